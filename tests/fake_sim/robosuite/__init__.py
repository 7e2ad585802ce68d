"""A stand-in for robosuite (TEST INFRASTRUCTURE): just enough of its surface for the reference's UNCHANGED
scripts/train_model.py, scripts/rollout.py, util/learn_utils.py and util/data_utils.py to run offline.

The real simulator (robosuite + mujoco-py) is not installable here; this package sits LAST on PYTHONPATH, so a real
install would win.  `make()` returns a deterministic environment that renders pseudo-random uint8 256 x 256 frames
(robosuite's default camera size) and smooth pseudo-random poses; what the estimators are fed is therefore exactly
the kind of data the real loops feed them (uint8 HWC frames through the reference's own torchvision transform).
"""
import os

import numpy as np

from . import utils  # noqa: F401
from .utils import transform_utils  # noqa: F401


# scripts/train_model.py never seeds its RNGs, so two runs start from different random weights.  Both scripts import
# robosuite on their first line: with PE_FAKE_SIM_SEED set, the stand-in seeds torch and numpy right there, which makes
# the "reference" and "ours" arms of tests/dropin_runner.py start from identical weights and draw identical noise.
if os.environ.get("PE_FAKE_SIM_SEED"):
    import torch
    torch.manual_seed(int(os.environ["PE_FAKE_SIM_SEED"]))
    np.random.seed(int(os.environ["PE_FAKE_SIM_SEED"]))


def load_controller_config(default_controller=None, custom_fpath=None):
    return {"type": default_controller}


def _unit_quat(rng):
    q = rng.standard_normal(4)
    return q / np.linalg.norm(q)


class _FakeEnv:
    OBJECTS = ("cube", "hammer", "pot", "peg")

    def __init__(self, n_arms, horizon, camera_names, camera_depths, seed=12345):
        self.n_arms = n_arms
        self.horizon = horizon
        self.camera = camera_names if isinstance(camera_names, str) else camera_names[0]
        self.camera_depths = camera_depths
        self.rng = np.random.RandomState(seed)
        self.t = 0
        self.action_dim = 7 * n_arms
        self._poses = {}

    @property
    def action_spec(self):
        return -np.ones(self.action_dim), np.ones(self.action_dim)

    def _obs(self):
        obs = {self.camera + "_image": self.rng.randint(0, 256, size=(256, 256, 3)).astype(np.uint8)}
        if self.camera_depths:
            obs[self.camera + "_depth"] = self.rng.rand(256, 256, 1).astype(np.float32)
        for name in ["robot%d_eef" % i for i in range(max(self.n_arms, 2))] + list(self.OBJECTS):
            pos, quat = self._poses[name]
            pos = pos + 0.01 * self.rng.standard_normal(3)
            quat = quat + 0.02 * self.rng.standard_normal(4)
            quat = quat / np.linalg.norm(quat)
            self._poses[name] = (pos, quat)
            obs[name + "_pos"] = pos.copy()
            obs[name + "_quat"] = quat.copy()
        return obs

    def reset(self):
        self.t = 0
        for name in ["robot%d_eef" % i for i in range(max(self.n_arms, 2))] + list(self.OBJECTS):
            self._poses[name] = (self.rng.uniform(-0.5, 0.5, size=3), _unit_quat(self.rng))
        return self._obs()

    def step(self, action):
        assert len(action) == self.action_dim
        self.t += 1
        return self._obs(), 0.0, self.t >= self.horizon, {}

    def move_indicator(self, pos):
        pass

    def render(self):
        pass


class Lift(_FakeEnv):
    pass


class TwoArmLift(_FakeEnv):
    pass


def make(env_name, robots=None, horizon=100, camera_names="frontview", camera_depths=False, **kwargs):
    cls = {"Lift": Lift, "TwoArmLift": TwoArmLift}.get(env_name)
    if cls is None:
        cls = type(env_name, (_FakeEnv,), {})
    return cls(2 if "TwoArm" in env_name else 1, horizon, camera_names, camera_depths)
