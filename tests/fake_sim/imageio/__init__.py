"""Stub (TEST INFRASTRUCTURE): scripts/rollout.py:11 imports imageio for --record_video only."""


def get_writer(*args, **kwargs):
    raise RuntimeError("imageio stub: video recording is outside the accelerated path")
