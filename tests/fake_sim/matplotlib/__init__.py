"""Stub (TEST INFRASTRUCTURE): the reference imports matplotlib.pyplot at util/learn_utils.py:14 and
util/model_utils.py:7 and only ever calls plt.close('all') on the hot path (util/learn_utils.py:516)."""
from . import pyplot  # noqa: F401
