"""Overlay package: the hot-path modules (utils.utils hot functions, utils.loss) live here; everything else the
reference's scripts import from `utils` (utils.datasets: TensorDataset / collate_fn — host-side I/O, out of scope) is
resolved from the reference checkout's own `utils/` directory when one is found further down sys.path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.join(os.path.abspath(_p or "."), "utils")
    if _cand != _here and os.path.isfile(os.path.join(_cand, "datasets.py")) and _cand not in __path__:
        __path__.append(_cand)
