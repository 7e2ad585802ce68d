"""`util` package of the drop-in: only `util.model_utils` (the trunk factory / checkpoint layout) is mirrored here.

The reference keeps its training and rollout loops in `util/learn_utils.py` and its simulator dataset in
`util/data_utils.py`; those are the unchanged CALLERS of the accelerated path and must
keep resolving to the reference's own files when this directory sits ahead of the reference root on PYTHONPATH.
The reference's `util/` has no `__init__.py` (a namespace package), which a regular package of the same name
shadows completely -- so the other `util/` directories on sys.path are appended to this package's search path.
`util.model_utils` still resolves here first.
"""
import os
import sys

_here = os.path.abspath(os.path.dirname(__file__))
for _p in list(sys.path):
    _d = os.path.abspath(os.path.join(_p or ".", "util"))
    if _d != _here and os.path.isdir(_d) and _d not in __path__:
        __path__.append(_d)
