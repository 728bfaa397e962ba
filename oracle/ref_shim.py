"""Import shim for the UNMODIFIED reference (test infrastructure only).

Only usable inside the build container where /root/reference exists; it is
what `tests/golden/make_golden.py` uses to generate the committed fixtures.
Nothing in the product path, the `-m gpu` tests, smoke() or bench.py imports
this module.  Recipe: SURVEY.md Appendix B (numpy>=2 moved a private helper the
reference imports, and matplotlib is absent from this image).
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SCPN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "scpn_fusion"))


def install() -> None:
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import numpy.lib.format as f
    import numpy.lib._format_impl as fi

    if not hasattr(f, "_read_array_header"):
        f._read_array_header = fi._read_array_header
    for m in (
        "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
        "matplotlib.cm", "matplotlib.gridspec", "matplotlib.animation", "matplotlib.figure",
        "matplotlib.axes", "matplotlib.collections", "matplotlib.lines", "matplotlib.ticker",
        "mpl_toolkits", "mpl_toolkits.mplot3d",
    ):
        sys.modules.setdefault(m, MagicMock())
    sys.dont_write_bytecode = True
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
