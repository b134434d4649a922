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


def install():
    """No-op for packages that are really installed (then matplotlib runs headless on the Agg backend)."""
    try:
        import matplotlib
        matplotlib.use("Agg")
    except ImportError:
        mpl, plt, cm = MagicMock(name="matplotlib"), MagicMock(name="matplotlib.pyplot"), MagicMock(name="matplotlib.cm")
        plt.subplots.side_effect = _subplots
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
