"""The reference's UNCHANGED loops against the drop-in mirrors.

scripts/train_model.py and scripts/rollout.py -- and through them util/learn_utils.train (:128-184), rollout
(:366-455) and util/data_utils.MultiEpisodeDataset -- are executed as they are, from the git-ignored copy of the
reference (oracle/_ref), with PYTHONPATH ordered as INTEGRATION.md section 1 prescribes and a stand-in simulator
(tests/fake_sim).  The "reference" arm runs the reference's own nn.Modules on the host CPU; the "ours" arm runs the
mirrors on the B200.  Both start from the same seed-0 weights and see the same frames, so the numbers the loops print
(per-phase loss / position error / orientation error), the checkpoint train() saves and the positions rollout() writes
to model_outputs.npy must agree to TF32 tolerance.  tests/golden/dropin_<kind>.json holds the reference arm's result
from the build container (tests/dropin_runner.py make-golden).
"""
import json
import os
import subprocess
import sys

import pytest

import dropin_runner as dr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _fixture(kind):
    return json.load(open(os.path.join(GOLDEN, "dropin_%s.json" % kind)))


def _close(a, b, tol, what):
    assert abs(a - b) <= tol * max(abs(b), 0.05), (what, a, b)


def _compare(res, ref, tol_loss, tol_out):
    for phase in ("train", "val"):
        for key in ("loss", "pos_err"):
            _close(res["train"][phase][key], ref["train"][phase][key], tol_loss, (phase, key))
        # summed |rotation angle| over a handful of near-random quaternions: same tolerance, radians
        _close(res["train"][phase]["ori_err"], ref["train"][phase]["ori_err"], tol_loss, (phase, "ori_err"))
    _close(res["train"]["best_val_err"], ref["train"]["best_val_err"], tol_loss, "best_val_err")
    # checkpoint file name = class _ env _ horizon _ episodes _ timestamp (util/learn_utils.py:223-231)
    assert res["train"]["checkpoint_name"].rsplit("_", 2)[0] == ref["train"]["checkpoint_name"].rsplit("_", 2)[0]
    ck, ck_ref = res["checkpoint"], ref["checkpoint"]
    assert list(ck) == list(ck_ref), "state_dict keys / order differ"
    for k, (s, sq, shape) in ck_ref.items():
        assert ck[k][2] == shape, k
        if k.endswith("num_batches_tracked"):
            assert ck[k][0] == s, k                              # exact: one dummy forward + train steps
        elif "running_" in k:
            assert abs(ck[k][1] - sq) <= tol_loss * max(sq, 1e-3), (k, ck[k][1], sq)
        else:
            # two Adam steps at lr 1e-5 move every weight by at most ~2e-5 (Adam's first steps are +-lr whatever the
            # gradient's size, so elements whose gradient sign differs between the arms end up 4e-5 apart): this pins
            # the layout and the magnitude, with that much slack per element
            numel = 1
            for d in shape:
                numel *= d
            assert abs(ck[k][1] - sq) <= 1e-3 * sq + numel * (4e-5) ** 2 + 2 * 4e-5 * (numel * sq) ** 0.5, (k, ck[k][1], sq)
    if ref["rollout"] is not None:
        r, rr = res["rollout"], ref["rollout"]
        assert len(r["model_outputs"]) == len(rr["model_outputs"]) == 30
        worst = max(abs(a - b) for pa, pb in zip(r["model_outputs"], rr["model_outputs"]) for a, b in zip(pa, pb))
        scale = max(abs(b) for pb in rr["model_outputs"] for b in pb)
        assert worst <= tol_out * scale, (worst, scale)
        _close(r["pos_mean"], rr["pos_mean"], tol_loss, "rollout pos mean")
        _close(r["ori_mean"], rr["ori_mean"], tol_loss, "rollout ori mean")


def test_reference_callers_resolve_next_to_the_mirrors():
    """With the drop-in package AHEAD of the reference root on PYTHONPATH, `models.*` and `util.model_utils` are the
    mirrors while `util.learn_utils` / `util.data_utils` stay the reference's own files (no GPU needed to import)."""
    if not dr.have_reference():
        pytest.skip("no reference tree / oracle/_ref copy")
    code = ("import util.learn_utils as lu, util.data_utils as du, util.model_utils as mu, models.naive as mn, "
            "models.time_sensitive as mt, models.losses as ml; "
            "print(lu.__file__); print(du.__file__); print(mu.__file__); print(mn.__file__); print(mt.__file__); "
            "print(ml.__file__); print(lu.NaiveObjectStateEstimator is mn.NaiveObjectStateEstimator)")
    r = subprocess.run([sys.executable, "-c", code], env=dr._env("ours"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lu, du, mu, mn, mt, ml, same = r.stdout.strip().splitlines()[-7:]
    assert lu.startswith(dr.REF) and du.startswith(dr.REF), (lu, du)
    assert all(p.startswith(dr.PKG) for p in (mu, mn, mt, ml)), (mu, mn, mt, ml)
    assert same == "True"


def test_reference_arm_reproduces_fixture():
    """The reference's own modules, run here through the unchanged scripts, reproduce the committed fixture (pins
    the fixture and the oracle/_ref copy to each other; host-CPU arithmetic only)."""
    if not dr.have_reference():
        pytest.skip("no reference tree / oracle/_ref copy")
    res = dr.run_case("reference", "tdo")
    assert res["train"]["device"] == "cpu"
    _compare(res, _fixture("tdo"), 2e-3, 2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["tdo", "td"])
def test_unchanged_scripts_drive_the_mirrors(kind):
    """scripts/train_model.py (one epoch: train + val phases, Adam, checkpoint) and scripts/rollout.py (30 batch-1
    steps with carried LSTM state, host tensors in / numpy out) run unchanged on the B200 mirrors and agree with the
    reference's own modules.  Tolerance: full ResNet-50 depth on 2-4 frames per step, TF32 operands -> 2e-2 on the
    printed losses / errors and on the rollout positions."""
    if not dr.have_reference():
        pytest.skip("oracle/_ref was not shipped with the snapshot")
    res = dr.run_case("ours", kind)
    assert res["train"]["device"] == "cuda:0"
    _compare(res, _fixture(kind), 2e-2, 2e-2)


@pytest.mark.gpu
def test_live_reference_arm_on_this_host():
    """Same comparison against the reference arm executed on THIS machine's host cores (not only the fixture)."""
    if not dr.have_reference():
        pytest.skip("oracle/_ref was not shipped with the snapshot")
    ref = dr.run_case("reference", "tdo")
    _compare(ref, _fixture("tdo"), 2e-3, 2e-3)
