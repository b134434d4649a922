"""Stand-ins for matplotlib / vispy so that the reference's UNMODIFIED scripts (which import them at module top and
plot after every run) can be executed on a box without those packages.  Everything is a MagicMock except the calls
whose return value the scripts unpack (`plt.subplots`)."""
import sys
from unittest.mock import MagicMock

import numpy as np


def _subplots(nrows=1, ncols=1, *a, **k):
    fig = MagicMock(name="Figure")
    if nrows == 1 and ncols == 1:
        return fig, MagicMock(name="Axes")
    axes = np.empty((nrows, ncols), dtype=object)
    for i in range(nrows):
        for j in range(ncols):
            axes[i, j] = MagicMock(name=f"Axes[{i},{j}]")
    return fig, (axes if nrows > 1 and ncols > 1 else axes.reshape(-1))


class _Cmap:
    """Stands in for a matplotlib colormap: called with an array it returns one RGBA row per entry (the scripts zip the
    result with their data), called with a scalar one RGBA tuple."""

    def __call__(self, x, *a, **k):
        if np.ndim(x) == 0:
            return (0.0, 0.0, 0.0, 1.0)
        return np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (len(x), 1))


class _CmapRegistry(MagicMock):
    """`plt.cm.<name>` / `cm.<name>`: every attribute is a colormap; get_cmap(name, n) too."""

    def __getattr__(self, name):
        if name.startswith("_") or name in ("assert_called", "method_calls", "mock_calls", "call_args", "call_args_list",
                                            "called", "call_count", "return_value", "side_effect"):
            return super().__getattr__(name)
        if name == "get_cmap":
            return lambda *a, **k: _Cmap()
        return _Cmap()


def install():
    """No-op for packages that are really installed (then matplotlib runs headless on the Agg backend)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
    except ImportError:
        mpl, plt, cm = MagicMock(name="matplotlib"), MagicMock(name="matplotlib.pyplot"), _CmapRegistry(name="matplotlib.cm")
        plt.subplots.side_effect = _subplots
        plt.cm = cm
        plt.get_cmap = lambda *a, **k: _Cmap()
        mpl.pyplot, mpl.cm = plt, cm
        for name, mod in [("matplotlib", mpl), ("matplotlib.pyplot", plt), ("matplotlib.cm", cm),
                          ("matplotlib.colors", MagicMock()), ("matplotlib.ticker", MagicMock()),
                          ("matplotlib.gridspec", MagicMock()), ("matplotlib.animation", MagicMock()),
                          ("matplotlib.lines", MagicMock()), ("matplotlib.patches", MagicMock())]:
            sys.modules.setdefault(name, mod)
    try:
        import vispy  # noqa: F401
    except ImportError:
        v = MagicMock(name="vispy")
        for name in ["vispy", "vispy.app", "vispy.scene", "vispy.io"]:
            sys.modules.setdefault(name, v if name == "vispy" else getattr(v, name.split(".")[1]))
