"""GPU parity tests of the estimators (through the reference-facing nn.Module API and the C ABI) against
the CPU oracle and the golden fixtures generated from the reference's own modules.

Tolerances (TF32 tensor-core operands, fp32 accumulate; see DESIGN.md "Numerics"):
  * full ResNet-50 depth, random init, 2-4 frames: outputs / loss within 2e-2 relative of the fp32 oracle.
    The yard-stick is torch's own cuDNN TF32 path, which lands 3e-3..5e-3 from the same oracle.
  * 4-block trunk (well conditioned): outputs / loss within 5e-3.
  * gradients: a random-init BN/ReLU stack amplifies rounding noise ~2000x (cuDNN *fp32* already differs
    from CPU fp32 by 2e-3..3e-2 in relative gradient norm, cuDNN TF32 by 0.1..0.8).  We therefore assert
    that our per-parameter gradient error is no worse than 1.5x torch-TF32's error (+2e-2) on the same
    problem, and pin the kernels individually to 1e-4 in test_kernels_gpu.py.
"""
import json
import os

import pytest
import torch

import model_checks as mc
from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _strict_fp32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    mc.SHALLOW[0] = False
    yield
    mc.SHALLOW[0] = False


def _split(rows):
    fwd = [(n, e, t) for n, e, t in rows if " grad " not in n and "conv-weight grad" not in n and "[oracle" not in n]
    grads = [(n, e) for n, e, t in rows if " grad " in n and "is None" not in n and "MISSING" not in n]
    struct = [(n, e, t) for n, e, t in rows if "is None" in n or "MISSING" in n]
    return fwd, grads, struct


@pytest.mark.parametrize("kind", ["no", "tdo", "td", "n", "tdo_v2"])
def test_shallow_trunk_parity(kind):
    mc.SHALLOW[0] = True
    rows = mc.check_train_step(kind, n=4, verbose=True)
    if kind in ("tdo", "td", "tdo_v2"):
        rows += mc.check_rollout(kind)
    fwd, grads, struct = _split(rows)
    bad = [(n, e) for n, e, t in fwd if not e <= max(t, 5e-3)] + [(n, e) for n, e, t in struct if e != 0.0]
    assert not bad, bad
    # gradient yard-stick: torch's TF32 path on the same problem
    if kind == "no":
        cal = dict((n, e) for n, e, _ in mc.calibrate(kind, n=4))
        yard = cal["calib torch-tf32 no n4 worst conv weight grad"]
        worst = max(e for n, e in grads)
        assert worst <= 1.5 * yard + 2e-2, (worst, yard)


# Full ResNet-50 depth against the PLAIN fp32 oracle (= the reference's own arithmetic): what TF32 operands cost.
# Calibration: torch's cuDNN TF32 path lands 4.7e-3 from the same oracle on the outputs and
# 1.3e-2 on running_var at 8 frames; eval-mode outputs sit higher than train-mode ones because nothing re-normalises
# the activations there -- running statistics after ONE update are ~(0.1 mean, 0.9 + 0.1 var), so operand rounding
# compounds through 53 un-normalised layers instead of being divided out by each layer's batch statistics.
# Measured here (B200, r2 logs under profiles/): outputs 4e-3 at 8 frames, 1.05e-2 at 4 frames (TDO), running_var
# 1.5e-2 / 1.6e-2 at 8 / 4 frames -- the SAME buffers agree with the oracle run at TF32 operand precision to 5e-7, and
# the outputs to 2e-4, so these are the price of TF32 operands on a handful of frames, not of the implementation.
FULL_DEPTH_OUT_TOL = {8: 6e-3, 4: 1.2e-2}
FULL_DEPTH_EVAL_TOL = 1.2e-2
FULL_DEPTH_STATS_TOL = 2e-2


def _full_depth_bad(rows, frames):
    fwd, _, struct = _split(rows)
    bad = []
    for n, e, t in fwd:
        if "[tf32-operand oracle]" in n:
            tol = t                      # the row's own tolerance (eval 3e-3, rollout 8e-3: model_checks.py)
        elif "running_" in n:
            tol = FULL_DEPTH_STATS_TOL
        elif "eval" in n or "rollout" in n:
            tol = FULL_DEPTH_EVAL_TOL
        else:
            tol = FULL_DEPTH_OUT_TOL[frames]
        if t == 0.0:
            tol = 0.0
        if not e <= tol:
            bad.append((n, e, tol))
    return bad + [(n, e, 0.0) for n, e, t in struct if e != 0.0]


@pytest.mark.parametrize("kind", ["no", "tdo"])
def test_full_depth_parity(kind):
    rows = mc.check_train_step(kind, n=4 if kind == "no" else 2)
    if kind == "tdo":
        rows += mc.check_rollout(kind, steps=2)
    bad = _full_depth_bad(rows, 4)
    assert not bad, bad


def test_config1_naive_object_cube_batch8():
    """BASELINE config 1 exactly: naive-object estimator, cube target, 8 frames, full depth, forward + backward.
    Outputs / loss / BatchNorm buffers against the plain fp32 oracle at the stated TF32 tolerance; every
    per-parameter gradient against the oracle at the same operand precision on the same ReLU masks (<= 1e-2)."""
    bad = _full_depth_bad(mc.check_train_step("no", n=8), 8)
    assert not bad, bad
    rows = mc.check_forced("no", n=8, verbose=True)
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad


@pytest.mark.parametrize("kind", ["no", "tdo"])
def test_full_depth_gradients_with_cta_pairs_forced(kind):
    """The production step (256 frames) runs nearly every conv launch on CTA pairs (tcgen05 cta_group::2); the small
    test batches never reach the automatic rule (one tile per SM).  Same full-depth teacher-forced gradient check with
    pairs forced on every launch that allows them."""
    from pe_b200 import native
    L = native.lib()
    L.pe_debug_cta_group(2)
    try:
        rows = mc.check_forced(kind, n=4 if kind == "no" else 2, verbose=True)
    finally:
        L.pe_debug_cta_group(0)
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad


@pytest.mark.parametrize("kind", ["no", "tdo", "td", "n", "tdo_v2"])
def test_full_depth_gradients_teacher_forced(kind):
    """Every per-parameter gradient of the FULL [3,4,6,3] trunk + head, all five estimators, within 1e-2
    (||g - g_ref|| / ||g_ref||) of the oracle run at the CUDA path's operand precision (TF32 operands,
    float64 accumulation) on the same ReLU masks -- the first model-level check of the identity-residual backward
    (lazy masked join gradient merged in the dgrad epilogue).  Outputs, loss and running statistics of the same run
    must agree to 1e-3 (north_star's example tolerance), every convolution's forward output to 5e-4.

    Why teacher forcing: a ReLU network's gradient is discontinuous in its activations; without common masks even
    fp32 vs fp64 accumulation on the CPU differ by 0.09 (median) in these gradients (oracle/pose_oracle.py)."""
    rows = mc.check_forced(kind, n=4 if kind in ("no", "n") else 2, verbose=True)
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad


@pytest.mark.parametrize("kind", ["no", "tdo", "td", "n", "tdo_v2"])
def test_against_reference_fixture(kind):
    """Outputs / loss / eval outputs of the CUDA path vs numbers produced by the reference's own modules
    (tests/golden/forward_<kind>.json)."""
    from models.losses import PoseDistanceLoss
    fx = json.load(open(os.path.join(GOLDEN, "forward_%s.json" % kind)))
    model = mc.build_model(kind).cuda().train()
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **fx["shapes"])
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    crit = PoseDistanceLoss(**fx["loss_cfg"])
    if kind in ("td", "tdo", "tdo_v2"):
        model.reset_initial_state(img.shape[1])
    out = model(img, None, x0)
    outs = list(out) if isinstance(out, tuple) else [out]
    loss = crit(outs[0], tgt) if kind in ("no", "tdo", "tdo_v2") else crit(outs[0], x0) + crit(outs[1], tgt)
    loss.backward()
    def close(a, b, tol=3e-2):
        b = torch.tensor(b)
        return float((a.detach().cpu() - b).abs().max()) <= tol * max(float(b.abs().max()), 0.1)

    for o, g in zip(outs, fx["outputs"]):
        assert close(o, g)
    if fx["loss"] != fx["loss"]:
        # the reference itself yields NaN here (all-zero ReLU'd quaternion, no epsilon: quirks Q2/Q7)
        assert torch.isnan(loss)
    else:
        assert abs(float(loss) - fx["loss"]) <= 3e-2 * abs(fx["loss"]), (float(loss), fx["loss"])
    named = dict(model.named_parameters())
    for n, gn in fx["grad_norms"].items():
        assert (named[n].grad is None) == (gn is None), n
    model.eval()
    with torch.no_grad():
        if kind in ("td", "tdo", "tdo_v2"):
            model.reset_initial_state(img.shape[1])
        oe = model(img, None, x0)
    oe = list(oe) if isinstance(oe, tuple) else [oe]
    for o, g in zip(oe, fx["eval_outputs"]):
        assert close(o, g)


def test_loss_module_known_answers():
    from models.losses import PoseDistanceLoss
    for c in json.load(open(os.path.join(GOLDEN, "loss_vectors.json"))):
        pred = torch.tensor(c["pred"], device="cuda", requires_grad=True)
        truth = torch.tensor(c["truth"], device="cuda")
        if c["mode"] == "val":
            pos, ang = PoseDistanceLoss(mode="val")(pred.detach(), truth)
            assert abs(float(pos) - c["pos"]) <= 1e-5 * max(1, abs(c["pos"]))
            assert abs(ang - c["angle"]) <= 1e-3 * max(1, abs(c["angle"]))
            continue
        crit = PoseDistanceLoss(distance_metric=c["metric"], scale_factor=c["scale_factor"], alpha=c["alpha"],
                                mode=c["mode"])
        loss = crit(pred, truth)
        loss.backward()
        assert abs(float(loss) - c["loss"]) <= 1e-5 * max(1, abs(c["loss"])), c
        assert torch.allclose(pred.grad.cpu(), torch.tensor(c["grad"]), rtol=1e-4, atol=1e-6), c


def test_fused_trainer_matches_autograd_path():
    """FusedTrainer (flat arenas + fused Adam) == model.forward / loss.backward() / torch.optim.Adam on the
    same kernels: identical parameters after two steps up to fp32 rounding of the update."""
    from models.losses import PoseDistanceLoss
    from pe_b200.trainer import FusedTrainer
    mc.SHALLOW[0] = True
    lk = mc.CONFIGS["no"]["loss"]
    img, x0, tgt = po.synthetic_batch("no", 4, seed=1)
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    a = mc.build_model("no").cuda().train()
    b = mc.build_model("no").cuda().train()
    tr = FusedTrainer(a, lr=1e-4, **lk)
    opt = torch.optim.Adam(b.parameters(), lr=1e-4)
    crit = PoseDistanceLoss(**lk)
    for _ in range(2):
        la = tr.step(img, x0, tgt)
        opt.zero_grad()
        lb = crit(b(img, None, x0), tgt)
        lb.backward()
        opt.step()
        assert abs(float(la) - float(lb)) <= 1e-4 * abs(float(lb))
    # Adam moves every weight by about +-lr per step whatever the gradient's size, so two runs that differ
    # only in atomic summation order may end up a few lr apart on weights whose gradient is ~0
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert float((pa - pb).abs().max()) <= 10 * 1e-4, n
    for (n, ba), (_, bb) in zip(a.named_buffers(), b.named_buffers()):
        assert mc.rel(ba.float(), bb.float()) <= 5e-3, n


@pytest.mark.parametrize("kind,name", [("no", "curve_no_lr1e-5.json"), ("tdo", "curve_tdo_lr1e-5.json")])
def test_loss_curve_vs_reference(kind, name):
    """100 Adam steps (lr 1e-5) against the loss curve of the reference's own modules + torch.optim.Adam.

    Stated tolerance: TF32 rounding noise (and the atomic summation order of the split-K wgrad) is amplified by
    training on a 4-frame batch, so the curves are required to agree to 5e-2 over the first 10 steps and, while the
    loss is still falling (steps 10-25), to stay inside the reference's +-2-step envelope widened by 10 % (twelve
    runs: always inside the un-widened envelope; plain pointwise deviation 0.17-0.30, dominated by sub-step lag).
    After that both curves sit on a noisy plateau
    (the reference's own values wander between 0.28 and 0.37), where a pointwise comparison is noise against
    noise: there every value must stay within a factor of two of the reference's plateau mean and the mean of the
    last 30 steps must agree to 20 %.  The reference's naive-object run dies (final-layer ReLU zeroes the
    quaternion -> NaN loss) around step 45; ours must die within 15 steps of it and track it to
    2e-2 until then."""
    import math
    path = os.path.join(GOLDEN, name)
    from pe_b200.trainer import FusedTrainer
    fx = json.load(open(path))
    model = mc.build_model(kind).cuda().train()
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **fx["shapes"])
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    tr = FusedTrainer(model, lr=fx["lr"], **fx["loss_cfg"])
    ref = fx["losses"]
    ours = [float(tr.step(img, x0, tgt)) for _ in ref]
    nan_ref = next((i for i, v in enumerate(ref) if math.isnan(v)), len(ref))
    nan_ours = next((i for i, v in enumerate(ours) if math.isnan(v)), len(ours))
    alive = min(nan_ref, nan_ours)
    dev = [abs(a - b) / abs(b) for a, b in zip(ours[:alive], ref[:alive])]
    assert max(dev[:10]) <= 5e-2, max(dev[:10])
    if kind == "no":
        assert abs(nan_ref - nan_ours) <= 15, (nan_ref, nan_ours)
        assert max(dev[:max(1, alive - 2)]) <= 2e-2, max(dev)
    else:
        # steep descent (loss falls 3 -> 0.4 in 15 steps): a one-step lag already moves the pointwise ratio by
        # 20-30 %, so the curve is compared with the reference's +-2-step envelope, widened by 10 %
        for i in range(10, 25):
            lo, hi = min(ref[i - 2:i + 3]), max(ref[i - 2:i + 3])
            assert 0.9 * lo <= ours[i] <= 1.1 * hi, (i, ours[i], lo, hi)
        plateau_ref = sum(ref[25:]) / len(ref[25:])
        assert all(0.5 * plateau_ref <= v <= 2.0 * plateau_ref for v in ours[25:]), (min(ours[25:]), max(ours[25:]))
        tail_o, tail_r = sum(ours[-30:]) / 30, sum(ref[-30:]) / 30
        assert abs(tail_o - tail_r) <= 0.2 * tail_r, (tail_o, tail_r)


def test_checkpoint_round_trip_through_gpu(tmp_path):
    """A checkpoint written from the CUDA model loads bit-identically into a fresh CPU copy and back."""
    m = mc.build_model("tdo").cuda()
    path = tmp_path / "ck.pth"
    torch.save(m.state_dict(), path)
    sd = torch.load(path, map_location="cpu")
    m2 = mc.build_model("tdo", seed=5)
    m2.load_state_dict(sd)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.cpu(), b), k


@pytest.mark.parametrize("kind", ["tdo", "no"])
def test_cuda_graph_rollout_matches_eager_and_oracle(kind):
    """Batch-1 streaming rollout under CUDA Graph capture == eager kernels == CPU oracle with carried state."""
    from pe_b200.rollout import StreamingEstimator
    mc.SHALLOW[0] = True
    model = mc.build_model(kind).cuda().eval()
    orc = mc.oracle_for(kind, model)
    orc.reset_state(1)
    g = StreamingEstimator(model, batch_size=1, use_graph=True)
    e = StreamingEstimator(model, batch_size=1, use_graph=False)
    g.reset()
    e.reset()
    for t in range(4):
        img, x0, _ = po.synthetic_batch(kind, 1, s=1, seed=20 + t) if kind == "tdo" else po.synthetic_batch(kind, 1, seed=20 + t)
        og = g.step(img.pin_memory(), x0.pin_memory()).clone()
        oe = e.step(img.cuda(), x0.cuda()).clone()
        ref = orc.forward(img, x0, training=False, rollout=True)
        assert torch.equal(og.reshape(-1), oe.reshape(-1)), t
        assert mc.rel(og.reshape(-1), ref.reshape(-1)) <= 5e-3, t


@pytest.mark.parametrize("kind", ["no", "tdo", "td", "n"])
def test_fused_head_matches_gemm_head(kind):
    """Rollout-sized inference: the one-launch fused head (pe_fused_head, fp32 FMA) and the tap-GEMM head (TF32
    tensor cores) agree on outputs and carried LSTM state."""
    from pe_b200 import estimators as est
    mc.SHALLOW[0] = True
    model = mc.build_model(kind).cuda().eval()
    seq = kind in ("td", "tdo")
    n = 3
    res = {}
    import models.naive as mn
    for fused in (True, False):
        est.FUSED_HEAD[0] = fused
        mn._drop_stream(model)          # the captured rollout step bakes the head variant in
        try:
            if seq:
                model.rollout = True
                model.reset_initial_state(n)
            outs = []
            with torch.no_grad():
                for t in range(2):
                    img, x0, _ = po.synthetic_batch(kind, n, s=1, seed=20 + t) if seq else po.synthetic_batch(kind, n, seed=20 + t)
                    o = model(img.cuda(), None, x0.cuda())
                    outs += [v.clone() for v in (o if isinstance(o, tuple) else (o,))]
            res[fused] = outs
        finally:
            est.FUSED_HEAD[0] = True
    for a, b in zip(res[True], res[False]):
        assert mc.rel(a, b) <= 5e-3, (kind, mc.rel(a, b))


def test_streaming_estimator_graph_with_fused_head():
    """CUDA-graph replay of the TDO streaming step == eager steps (state carried in place by the fused head)."""
    from pe_b200.rollout import StreamingEstimator
    mc.SHALLOW[0] = True
    model = mc.build_model("tdo").cuda().eval()
    outs = {}
    for use_graph in (True, False):
        est_ = StreamingEstimator(model, batch_size=2, use_graph=use_graph)
        est_.reset()
        seq = []
        for t in range(3):
            img, x0, _ = po.synthetic_batch("tdo", 2, s=1, seed=30 + t)
            seq.append(est_.step(img.cuda(), x0.cuda()).clone())
        outs[use_graph] = seq
    for a, b in zip(outs[True], outs[False]):
        assert mc.rel(a, b) <= 1e-5, mc.rel(a, b)


def test_space_to_depth_stem_matches_im2col_stem():
    """conv1 as a 4x4/1 convolution over the space-to-depth image read through an overlapping TMA view
    (pe_stem_conv_fwd / pe_stem_conv_wgrad, no im2col matrix) gives the same loss, gradients (conv1's weight included)
    and BatchNorm buffers as im2col + GEMM.  The two sum conv1's 147 products in a different order; the 1e-7 difference
    of the first layer flips a few ReLU masks downstream, and a flipped mask is an O(1) change of that element's
    gradient: the loss agrees to 1e-4, BatchNorm buffers to 1e-3 (the re-rounding of every activation to TF32
    amplifies the first layer's 1e-7 layer by layer: 1e-5 at layer3, 1.2e-4 at layer4 with 4 frames), eval outputs to 5e-3, the gradients of this shallow
    trunk only to 0.2 (sanity bound; measured up to 6e-2).  The tight gradient checks of the stem are the
    kernel check (fp64 reference, 1e-4) and the teacher-forced full-depth tests (common masks, 1e-2)."""
    from models.losses import PoseDistanceLoss
    from pe_b200 import engine
    mc.SHALLOW[0] = True
    img, x0, tgt = po.synthetic_batch("no", 4, seed=12)
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    crit = PoseDistanceLoss(distance_metric="l2", alpha=0.5, mode="pose")
    res = {}
    for s2d in (False, True):
        engine.STEM_S2D[0] = s2d
        try:
            model = mc.build_model("no").cuda().train()
            with torch.no_grad():
                getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
            loss = crit(model(img, None, x0), tgt)
            loss.backward()
            model.eval()
            with torch.no_grad():
                ev = model(img, None, x0).clone()
            res[s2d] = (float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                        {n: b.clone() for n, b in model.named_buffers()}, ev)
        finally:
            engine.STEM_S2D[0] = True
    assert abs(res[True][0] - res[False][0]) <= 1e-4 * abs(res[False][0])
    assert res[True][1].keys() == res[False][1].keys()
    for n, g in res[False][1].items():
        assert mc.relnorm(res[True][1][n], g) <= 0.2, (n, mc.relnorm(res[True][1][n], g))
    for n, b in res[False][2].items():
        assert mc.relnorm(res[True][2][n].float(), b.float()) <= 1e-3, n
    assert mc.rel(res[True][3], res[False][3]) <= 5e-3


def test_fused_stem_tail_matches_separate_kernels():
    """The one-pass stem tail (bn1 + ReLU + max pool + aux branch, pe_stem_post_train; bn1's output never written)
    gives the same loss, gradients and BN buffers as bn_train_apply + maxpool + aux as three kernels.  Tolerance
    1e-4: the split-K wgrad accumulates with unordered fp32 atomics, everything else is the same arithmetic."""
    from models.losses import PoseDistanceLoss
    from pe_b200 import engine
    mc.SHALLOW[0] = True
    img, x0, tgt = po.synthetic_batch("no", 4, seed=11)
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    crit = PoseDistanceLoss(distance_metric="l2", alpha=0.5, mode="pose")
    res = {}
    for fused in (False, True):
        engine.FUSED_STEM_TAIL[0] = fused
        try:
            model = mc.build_model("no").cuda().train()
            with torch.no_grad():      # keep the ReLU'd output alive so the loss is finite (quirk Q2)
                getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
            loss = crit(model(img, None, x0), tgt)
            loss.backward()
            res[fused] = (float(loss), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                          {n: b.clone() for n, b in model.named_buffers()})
        finally:
            engine.FUSED_STEM_TAIL[0] = True
    assert abs(res[True][0] - res[False][0]) <= 1e-5 * abs(res[False][0])
    assert res[True][1].keys() == res[False][1].keys()
    for n, g in res[False][1].items():
        assert mc.relnorm(res[True][1][n], g) <= 1e-4, (n, mc.relnorm(res[True][1][n], g))
    for n, b in res[False][2].items():
        assert mc.relnorm(res[True][2][n].float(), b.float()) <= 1e-6, n


def test_frozen_trunk_feature_extraction():
    """feature_extract semantics (util/model_utils.py:110-113): with every trunk parameter frozen except the
    replaced fc, the backward pass skips the convolutions; the head / fc / aux-conv gradients are the same as in
    the full backward and the frozen parameters get none."""
    from models.losses import PoseDistanceLoss
    mc.SHALLOW[0] = True
    img, x0, tgt = po.synthetic_batch("no", 3, seed=5)
    img, x0, tgt = img.cuda(), x0.cuda(), tgt.cuda()
    crit = PoseDistanceLoss(distance_metric="l2", alpha=0.5, mode="pose")
    grads = {}
    for frozen in (False, True):
        model = mc.build_model("no").cuda().train()
        with torch.no_grad():      # keep the ReLU'd output alive so the loss is finite (quirk Q2)
            getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        trunk = model.feature_net.module
        if frozen:
            for n, p in trunk.named_parameters():
                if not n.startswith("fc."):
                    p.requires_grad_(False)
        loss = crit(model(img, None, x0), tgt)
        loss.backward()
        grads[frozen] = {n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}
    for n, g in grads[True].items():
        is_trunk_conv = n.startswith("feature_net.module.") and not n.startswith("feature_net.module.fc.")
        if is_trunk_conv:
            assert g is None, n
        elif grads[False][n] is None:
            assert g is None, n
        else:
            assert mc.relnorm(g, grads[False][n]) <= 1e-3, (n, mc.relnorm(g, grads[False][n]))


@pytest.mark.parametrize("kind", ["no", "tdo"])
def test_use_depth_parity(kind):
    """use_depth=True: depth branch forward and the gradients it adds (InstanceNorm affine, aux conv,
    trunk through the aux path) against the oracle's autograd on the same weights."""
    import contextlib
    import io
    import models.naive as mn
    import models.time_sensitive as mt
    import util.model_utils as mu
    from models.losses import PoseDistanceLoss
    mu._RESNET_LAYERS[50] = [1, 1, 1, 1]
    try:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "no":
                model = mn.NaiveObjectStateEstimator("cube", [1024, 256, 64], 50, 512, False, (9,), True, False)
            else:
                model = mt.TemporallyDependentObjectStateEstimator("robot1_eef", 512, 50, 512, 20, feature_extract=False,
                                                                   use_depth=True, use_pretrained=False)
        with torch.no_grad():
            model.depth_nets[0].module[2].weight.fill_(0.7)
            model.depth_nets[0].module[2].bias.fill_(0.2)
            if kind == "no":
                getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        orc = po.OracleEstimator(kind, {k: v.detach().cpu() for k, v in model.state_dict().items()})
        shape = dict(n=3) if kind == "no" else dict(n=2, s=2)
        img, x0, tgt = po.synthetic_batch(kind, seed=6, **shape)
        depth = torch.rand(*img.shape[:-3], 1, 224, 224, generator=torch.Generator().manual_seed(9))
        lk = dict(distance_metric="l2", alpha=0.5, mode="pose")
        for k in orc.param_names:
            orc.sd[k].requires_grad_(True)
        ref_out = orc.forward(img, x0, training=True, depth=depth)
        ref_loss = po.pose_loss(ref_out, tgt, **lk)
        ref_loss.backward()
        model.cuda().train()
        if kind == "tdo":
            model.reset_initial_state(2)
        out = model(img.cuda(), depth.cuda(), x0.cuda())
        loss = PoseDistanceLoss(**lk)(out, tgt.cuda())
        loss.backward()
        assert mc.rel(out, ref_out.detach()) <= 5e-3, mc.rel(out, ref_out.detach())
        assert abs(float(loss) - float(ref_loss)) <= 5e-3 * abs(float(ref_loss))
        named = dict(model.named_parameters())
        for n in ("depth_nets.0.module.2.weight", "depth_nets.0.module.2.bias", "aux_nets.0.module.0.weight",
                  "aux_nets.0.module.0.bias", "feature_net.module.conv1.weight"):
            g, gr = named[n].grad, orc.sd[n].grad
            assert g is not None and gr is not None, n
            # head-side gradients are tight; the stem conv sits behind the whole TF32 trunk, where torch's own
            # cuDNN-TF32 path is ~0.1 away from the fp32 oracle on this 4-block net (DESIGN.md "Numerics")
            tol = 0.15 if n.startswith("feature_net") else 5e-2
            assert mc.relnorm(g, gr) <= tol, (n, mc.relnorm(g, gr))
        # inference path (fused head for <= 8 frames) with the depth branch
        model.eval()
        with torch.no_grad():
            if kind == "tdo":
                model.reset_initial_state(2)
            oe = model(img.cuda(), depth.cuda(), x0.cuda())
        for k in orc.param_names:
            orc.sd[k].requires_grad_(False)
        oe_ref = orc.forward(img, x0, training=False, depth=depth)
        assert mc.rel(oe, oe_ref) <= 5e-3, mc.rel(oe, oe_ref)
    finally:
        mu._RESNET_LAYERS[50] = [3, 4, 6, 3]


def test_streaming_estimator_raw_frames():
    """step_raw (uint8 HWC 256x256 frames, preprocessing inside the captured graph) == step on frames preprocessed
    with the reference's transform arithmetic."""
    from pe_b200.preprocess import IMAGENET_MEAN, IMAGENET_STD
    from pe_b200.rollout import StreamingEstimator
    mc.SHALLOW[0] = True
    model = mc.build_model("tdo").cuda().eval()
    raw_est = StreamingEstimator(model, batch_size=2, use_graph=True, raw_hw=256)
    ref_est = StreamingEstimator(model, batch_size=2, use_graph=False)
    raw_est.reset()
    ref_est.reset()
    g = torch.Generator().manual_seed(3)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    for t in range(3):
        raw = torch.randint(0, 256, (2, 256, 256, 3), dtype=torch.uint8, generator=g)
        x0 = torch.randn(2, 7, generator=g)
        img = (raw[:, 16:240, 16:240, :].permute(0, 3, 1, 2).float() / 255.0 - mean) / std
        a = raw_est.step_raw(raw.pin_memory(), x0.pin_memory()).clone()
        b = ref_est.step(img.cuda(), x0.cuda()).clone()
        assert mc.rel(a, b) <= 1e-5, (t, mc.rel(a, b))


@pytest.mark.parametrize("kind,raw", [("tdo", False), ("tdo", True), ("td", False), ("no", False)])
def test_pipelined_estimator_matches_one_shot_step(kind, raw):
    """PipelinedEstimator (chunks of the batch copied on a side stream while the previous chunk computes, one CUDA graph
    per chunk, LSTM state per chunk) == the one-shot StreamingEstimator step, row for row, over carried state."""
    from pe_b200.rollout import PipelinedEstimator, StreamingEstimator
    mc.SHALLOW[0] = True
    model = mc.build_model(kind).cuda().eval()
    hw = 256 if raw else None
    pipe = PipelinedEstimator(model, batch_size=12, chunk=4, raw_hw=hw)
    ref = StreamingEstimator(model, batch_size=12, use_graph=False, raw_hw=hw)
    pipe.reset()
    ref.reset()
    g = torch.Generator().manual_seed(5)
    for t in range(3):
        if raw:
            frames = torch.randint(0, 256, (12, 256, 256, 3), dtype=torch.uint8, generator=g).pin_memory()
        else:
            frames = torch.randn(12, 3, 224, 224, generator=g).pin_memory()
        x0 = torch.randn(12, 7, generator=g).pin_memory()
        a = pipe.step_raw(frames, x0) if raw else pipe.step(frames, x0)
        b = ref.step_raw(frames, x0) if raw else ref.step(frames, x0)
        a = a if isinstance(a, tuple) else (a,)
        b = b if isinstance(b, tuple) else (b,)
        for u, v in zip(a, b):
            assert u.shape == v.shape
            assert mc.rel(u.float().cpu(), v.float().cpu()) <= 2e-3, (kind, t, mc.rel(u.cpu(), v.cpu()))


@pytest.mark.parametrize("kind", ["no", "tdo"])
def test_no_proprioception_variant(kind):
    """no_proprioception=True (models/naive.py:147, models/time_sensitive.py:295): the 7 proprio columns vanish from
    the fusion rows (first head layer has 7 fewer inputs); forward, loss and head gradients against the oracle, in
    training (GEMM head) and rollout-sized eval (fused head) modes."""
    import contextlib
    import io
    import models.naive as mn
    import models.time_sensitive as mt
    import util.model_utils as mu
    from models.losses import PoseDistanceLoss
    mu._RESNET_LAYERS[50] = [1, 1, 1, 1]
    try:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            if kind == "no":
                model = mn.NaiveObjectStateEstimator("cube", [1024, 256, 64], 50, 512, False, (9,), False, False,
                                                     no_proprioception=True)
            else:
                model = mt.TemporallyDependentObjectStateEstimator("robot1_eef", 512, 50, 512, 20, feature_extract=False,
                                                                   use_pretrained=False, no_proprioception=True)
        with torch.no_grad():
            if kind == "no":
                getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "running_" not in k)
              for k, v in model.state_dict().items()}
        shape = dict(n=3) if kind == "no" else dict(n=2, s=2)
        img, x0, tgt = po.synthetic_batch(kind, seed=8, **shape)
        lk = dict(distance_metric="l2", alpha=0.5, mode="pose")
        if kind == "no":
            ref_out = po.naive_object_forward(sd, img, x0, True, model.n_fc, use_proprio=False)
        else:
            ref_out, _ = po.tdo_forward(sd, img, x0, True, None, use_proprio=False)
        ref_loss = po.pose_loss(ref_out, tgt, **lk)
        ref_loss.backward()
        model.cuda().train()
        if kind == "tdo":
            model.reset_initial_state(2)
        out = model(img.cuda(), None, x0.cuda())
        loss = PoseDistanceLoss(**lk)(out, tgt.cuda())
        loss.backward()
        assert mc.rel(out, ref_out.detach()) <= 5e-3, mc.rel(out, ref_out.detach())
        assert abs(float(loss) - float(ref_loss)) <= 5e-3 * abs(float(ref_loss))
        named = dict(model.named_parameters())
        first = "fc0.module.weight" if kind == "no" else "rnn.module.weight_ih_l0"
        assert named[first].shape[1] == 512 + 3136
        assert mc.relnorm(named[first].grad, sd[first].grad) <= 5e-2
        model.eval()
        with torch.no_grad():
            if kind == "tdo":
                model.reset_initial_state(2)
            oe = model(img.cuda(), None, x0.cuda())
        sd2 = {k: v.detach() for k, v in sd.items()}
        # eval uses the running statistics the training forward just updated on the GPU model
        sd2.update({k: v.detach().cpu() for k, v in model.state_dict().items() if "running_" in k})
        if kind == "no":
            oe_ref = po.naive_object_forward(sd2, img, x0, False, model.n_fc, use_proprio=False)
        else:
            oe_ref, _ = po.tdo_forward(sd2, img, x0, False, None, use_proprio=False)
        assert mc.rel(oe, oe_ref) <= 5e-3, mc.rel(oe, oe_ref)
    finally:
        mu._RESNET_LAYERS[50] = [3, 4, 6, 3]


@pytest.mark.parametrize("n", [1, 5])
def test_odd_batch_sizes(n):
    """Ragged sizes: 1 and 5 frames (tiles of 128 pixels never divide evenly; BatchNorm over a single frame)."""
    mc.SHALLOW[0] = True
    rows = mc.check_train_step("no", n=n)
    fwd, grads, struct = _split(rows)
    bad = [(name, e) for name, e, t in fwd if not e <= max(t, 5e-3)] + [(name, e) for name, e, t in struct if e != 0.0]
    assert not bad, bad


def test_device_prefetcher_order_and_values():
    """DevicePrefetcher yields every batch, in order, with the host values, while copying one batch ahead."""
    from pe_b200.loader import DevicePrefetcher
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(4, 3, 8, 8, generator=g).pin_memory(), (torch.randn(4, 7, generator=g).pin_memory(),
                                                                     torch.randn(4, 7, generator=g).pin_memory()))
               for _ in range(5)]
    seen = 0
    for k, (img, (a, b)) in enumerate(DevicePrefetcher(batches, dev)):
        # consume on the compute stream before the slot can be recycled
        assert torch.equal(img.cpu(), batches[k][0]) and torch.equal(a.cpu(), batches[k][1][0])
        assert torch.equal(b.cpu(), batches[k][1][1])
        seen += 1
    assert seen == 5


@pytest.mark.parametrize("kind,latent", [("no", 50), ("n", 50), ("n", 25), ("td", 25), ("tdo", 50)])
def test_latent_dims_that_are_not_multiples_of_four(kind, latent):
    """latent_dim = 50 is the DEFAULT of all five constructors (models/naive.py:19,143; models/time_sensitive.py:18,
    287): the trunk fc's dgrad / wgrad operands then need padded rows (16-byte TMA alignment).  latent 25 makes
    latent + 7 (and latent + 3136 + 7) a multiple of 32, the case where the fusion rows have no padding columns at
    all.  Forward, loss and the presence of every gradient against the oracle on the 4-block trunk."""
    mc.SHALLOW[0] = True
    rows = mc.check_train_step(kind, n=3, latent=latent, verbose=True)
    fwd, grads, struct = _split(rows)
    bad = [(n, e) for n, e, t in fwd if not e <= max(t, 5e-3)] + [(n, e) for n, e, t in struct if e != 0.0]
    assert not bad, bad
    # the trunk fc is the layer whose shadows needed the padding: a layout error there is an O(1) error.  Against the
    # PLAIN fp32 oracle its gradient carries the TF32 operand error of a 3-frame batch (measured 1.5e-2 .. 2.4e-2
    # depending on the summation order of the stem): 4e-2
    fc = [e for n, e in grads if ".fc.weight" in n or ".fc.bias" in n]
    assert fc and max(fc) <= 4e-2, fc


@pytest.mark.parametrize("kind", ["tdo", "no"])
def test_host_tensors_are_staged_like_the_rollout_loop_feeds_them(kind):
    """util/learn_utils.py:415-455 hands the model CPU tensors and reads the pose with .numpy(); scripts/rollout.py
    never calls model.cuda().  The mirrors move themselves to the current CUDA device on first use, stage host inputs
    and return host outputs; results equal the all-device call."""
    mc.SHALLOW[0] = True
    model = mc.build_model(kind).eval()          # parameters still on the host
    model.rollout = True
    seq = kind == "tdo"
    outs = {}
    for where in ("host", "device"):
        model.reset_initial_state(1)
        res = []
        for t in range(3):
            img, x0, _ = po.synthetic_batch(kind, 1, s=1, seed=40 + t) if seq else po.synthetic_batch(kind, 1, seed=40 + t)
            if where == "device":
                img, x0 = img.cuda(), x0.cuda()
            o = model(img, None, x0)             # autograd left on, as in the reference loop (quirk Q9)
            assert o.device.type == ("cpu" if where == "host" else "cuda")
            assert not o.requires_grad
            res.append(o.detach().cpu().numpy())
        outs[where] = res
    assert next(model.parameters()).is_cuda
    for a, b in zip(outs["host"], outs["device"]):
        assert abs(a - b).max() <= 1e-6 * max(abs(b).max(), 1e-3)
    from models.losses import PoseDistanceLoss
    pos, ang = PoseDistanceLoss(mode="val")(torch.tensor(outs["host"][0]).reshape(-1, 7), torch.tensor([[0., 0, 0, 0, 0, 0, 1]]))
    assert pos == pos and ang == ang


def test_two_forwards_before_one_backward():
    """torch allows loss(model(a)) + loss(model(b)) with a single backward: each taped forward keeps its own copy of
    the BatchNorm coefficient vectors, so the first forward's backward does not read the second one's statistics."""
    from models.losses import PoseDistanceLoss
    mc.SHALLOW[0] = True
    lk = dict(distance_metric="l2", alpha=0.5, mode="pose")
    crit = PoseDistanceLoss(**lk)
    a = po.synthetic_batch("tdo", 2, s=2, seed=3)
    b = po.synthetic_batch("tdo", 2, s=2, seed=4)
    grads = []
    for joint in (True, False):
        model = mc.build_model("tdo").cuda().train()
        if joint:
            la = crit(model(a[0].cuda(), None, a[1].cuda()), a[2].cuda())
            lb = crit(model(b[0].cuda(), None, b[1].cuda()), b[2].cuda())
            (la + lb).backward()
            grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
        else:
            # the same two forwards (BatchNorm buffers evolve identically), each followed by its own backward
            crit(model(a[0].cuda(), None, a[1].cuda()), a[2].cuda()).backward()
            crit(model(b[0].cuda(), None, b[1].cuda()), b[2].cuda()).backward()
            grads.append({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    for n, g in grads[1].items():
        assert mc.relnorm(grads[0][n], g) <= 1e-3, (n, mc.relnorm(grads[0][n], g))


def test_loss_curve_at_the_reference_learning_rate():
    """TDO, lr 1e-3 (scripts/train_model.py:26,228 default), 30 Adam steps on the fixture's 2 x 2 frames.

    At this learning rate the reference does not reproduce ITSELF: tests/golden/curve_tdo.json and
    curve_tdo_lr1e-3_threads{1,3}.json are the same reference modules, seed and data run with 8, 1 and 3 CPU threads
    (i.e. a different summation order inside oneDNN) -- they agree to 0.3 % for four steps and are 10-100 % apart from
    step 6 on (Adam moves every weight by ~lr per step whatever the gradient's size, and ReLU masks flip).  Stated
    tolerance, accordingly: the first four steps within 1.5e-2 of the reference, step 4 within 8e-2; from then on every
    value inside the envelope of the three reference realisations over a +-1-step window, widened by a factor of two
    (our three GPU runs in profiles/curve_r02.log stay inside a factor 1.5 except for one step), and the mean of the
    last ten steps within the realisations' range widened by 1.5."""
    from pe_b200.trainer import FusedTrainer
    fx = json.load(open(os.path.join(GOLDEN, "curve_tdo.json")))
    refs = [fx["losses"]] + [json.load(open(os.path.join(GOLDEN, "curve_tdo_lr1e-3_threads%d.json" % t)))["losses"]
                             for t in (1, 3)]
    assert fx["lr"] == 1e-3
    model = mc.build_model("tdo").cuda().train()
    img, x0, tgt = po.synthetic_batch("tdo", seed=1, **fx["shapes"])
    tr = FusedTrainer(model, lr=fx["lr"], **fx["loss_cfg"])
    ours = [float(tr.step(img.cuda(), x0.cuda(), tgt.cuda())) for _ in range(30)]
    for i in range(4):
        assert abs(ours[i] - refs[0][i]) <= 1.5e-2 * refs[0][i], (i, ours[i], refs[0][i])
    assert abs(ours[4] - refs[0][4]) <= 8e-2 * refs[0][4], (ours[4], refs[0][4])
    for i in range(5, 30):
        window = [r[j] for r in refs for j in range(max(0, i - 1), min(30, i + 2))]
        assert min(window) / 2 <= ours[i] <= max(window) * 2, (i, ours[i], min(window), max(window))
    tails = [sum(r[-10:]) / 10 for r in refs]
    tail = sum(ours[-10:]) / 10
    assert min(tails) / 1.5 <= tail <= max(tails) * 1.5, (tail, tails)


def test_resnet18_trunk():
    """num_resnet_layers = 18 (BasicBlock trunk, the other depth of the reference's option set torchvision can build):
    two 3x3 convs per block, the identity / downsample gradient merged in the 3x3 conv1 dgrad epilogue (or summed
    for the stride-2 stage transitions).  Outputs / loss vs the plain fp32 oracle, every per-parameter gradient vs the
    TF32-operand teacher-forced oracle."""
    rows = mc.check_train_step("no", n=4, layers=18)
    bad = _full_depth_bad(rows, 4)
    assert not bad, bad
    rows = mc.check_forced("no", n=4, layers=18, verbose=True) + mc.check_forced("tdo", n=2, layers=18)
    bad = [(n, e, t) for n, e, t in rows if not e <= t]
    assert not bad, bad
