"""Import alias for the hyphen-named package directory
`hydrodynamic-limits-of-active-particle-systems-with-mean-field-interactions_b200/`."""
import importlib.util as _u
import os as _os
import sys as _sys

_real = _os.path.join(
    _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
    "hydrodynamic-limits-of-active-particle-systems-with-mean-field-interactions_b200",
)
_spec = _u.spec_from_file_location(
    "aps_b200", _os.path.join(_real, "__init__.py"), submodule_search_locations=[_real]
)
_mod = _u.module_from_spec(_spec)
_sys.modules["aps_b200"] = _mod
_spec.loader.exec_module(_mod)
