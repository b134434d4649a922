#!/usr/bin/env python
"""Install the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun).

The reference is 15 flat .py files with no setup.py / pyproject.toml, so the contract's command
    pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
fails with "Neither 'setup.py' nor 'pyproject.toml' found".  Per the contract's fallback the install runs from a copy
under /tmp to which ONE file is added: a setup.py that lists the reference's modules as `py_modules` (no source file is
touched).  The installed files are byte-identical to /root/reference (checked below).
Used by: `bench.py --impl reference` / `cpu_baseline` (kind "reference"), tests/test_unmodified_drivers.py, and the
plotting methods of the drop-in class, which delegate to the reference's own code.
"""
import filecmp
import glob
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def install(force=False):
    mods = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(REF, "*.py")))
    if not mods:
        raise SystemExit(f"{REF} is not available")
    if not force and all(os.path.exists(os.path.join(DST, m + ".py")) and
                         filecmp.cmp(os.path.join(REF, m + ".py"), os.path.join(DST, m + ".py"), shallow=False) for m in mods):
        return DST
    tmp = tempfile.mkdtemp(prefix="aps_ref_")
    src = os.path.join(tmp, "reference")
    shutil.copytree(REF, src)
    with open(os.path.join(src, "setup.py"), "w") as f:
        f.write("from setuptools import setup\n"
                f"setup(name='hydrodynamic-limits-reference', version='0', py_modules={mods!r})\n")
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "-q",
                           "--find-links", "/opt/wheelhouse", "--target", DST, src])
    for m in mods:
        assert filecmp.cmp(os.path.join(REF, m + ".py"), os.path.join(DST, m + ".py"), shallow=False), m
    shutil.rmtree(tmp, ignore_errors=True)
    return DST


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
