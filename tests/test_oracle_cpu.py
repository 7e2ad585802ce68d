"""CPU checks of the oracle: against the golden fixtures (always) and against the reference's own
modules imported from /root/reference (build container only)."""
import json
import os

import pytest
import torch

from oracle import pose_oracle as po
from oracle import ref_shim

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def test_loss_known_answers():
    """Every (metric, mode) of PoseDistanceLoss: value and gradient equal the reference's."""
    for c in _golden("loss_vectors.json"):
        pred = torch.tensor(c["pred"], requires_grad=True)
        truth = torch.tensor(c["truth"])
        if c["mode"] == "val":
            pos, ang = po.pose_val_metrics(pred.detach(), truth)
            assert abs(pos - c["pos"]) <= 1e-5 * max(1, abs(c["pos"]))
            assert abs(ang - c["angle"]) <= 1e-4 * max(1, abs(c["angle"]))
            continue
        loss = po.pose_loss(pred, truth, c["metric"], c["scale_factor"], c["alpha"], 1e-4, c["mode"])
        loss.backward()
        assert abs(float(loss) - c["loss"]) <= 1e-6 * max(1, abs(c["loss"])), c
        assert torch.allclose(pred.grad, torch.tensor(c["grad"]), rtol=1e-5, atol=1e-6), c


def test_survey_appendix_c_values():
    """The survey's hand-recorded numbers for combined / pose."""
    pred = torch.tensor([[0.1, 0.2, 0.3, 0.5, 0.5, 0.5, 0.5], [-0.4, 0.25, 0.0, 0.1, -0.2, 0.3, -0.4]])
    truth = torch.tensor([[0., 0., 0., 0., 0., 0., 1.], [0.1, 0.25, -0.3, 0., 0.6, 0., 0.8]])
    want = {("l1", "position"): 1.40000010, ("l2", "pose"): 1.87496209, ("combined", "pose"): 4.07496214}
    for (metric, mode), v in want.items():
        assert abs(float(po.pose_loss(pred, truth, metric, 1.0, 0.5, 1e-4, mode)) - v) < 1e-5
    pos, ang = po.pose_val_metrics(pred, truth)
    assert abs(pos - 0.9574803) < 1e-5 and abs(ang - 3.3702678813908973) < 1e-4
    nan = po.pose_loss(torch.tensor([[0.1, 0, 0, 0., 0., 0., 0.]]), truth[:1], "l2", 1.0, 0.5, 1e-4, "pose")
    assert torch.isnan(nan)


@pytest.mark.parametrize("kind", ["no", "tdo", "td", "n", "tdo_v2"])
def test_forward_backward_golden(kind):
    """Oracle forward / loss / gradients / BN running stats / eval outputs == reference fixtures."""
    import model_checks as mc
    fx = _golden("forward_%s.json" % kind)
    model = mc.build_model(kind)
    orc = mc.oracle_for(kind, model)
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **fx["shapes"])
    lk = fx["loss_cfg"]
    for k in orc.param_names:
        orc.sd[k].requires_grad_(True)
    out = orc.forward(img, x0, training=True)
    outs = list(out) if isinstance(out, tuple) else [out]
    loss = po.pose_loss(outs[0], tgt, **lk) if kind in ("no", "tdo", "tdo_v2") else \
        po.pose_loss(outs[0], x0, **lk) + po.pose_loss(outs[1], tgt, **lk)
    loss.backward()
    for o, g in zip(outs, fx["outputs"]):
        assert torch.allclose(o.detach(), torch.tensor(g), rtol=1e-4, atol=1e-6)
    if fx["loss"] != fx["loss"]:
        # reference quirk Q2/Q7: the "n" model's ReLU'd pre-measurement head emits an all-zero quaternion at
        # init, which PoseDistanceLoss normalises without an epsilon -> NaN in the reference itself
        assert torch.isnan(loss)
    else:
        assert abs(float(loss) - fx["loss"]) <= 1e-5 * abs(fx["loss"])
    for n, gn in fx["grad_norms"].items():
        g = orc.sd[n].grad
        if gn is None:
            assert g is None, n
        else:
            assert abs(float(g.norm()) - gn) <= 2e-3 * max(gn, 1e-6), (n, float(g.norm()), gn)
    for k, v in fx["running_var_sum"].items():
        assert abs(float(orc.sd[k].double().sum()) - v) <= 1e-5 * abs(v), k
    for k in orc.param_names:
        orc.sd[k].requires_grad_(False)
    oe = orc.forward(img, x0, training=False)
    oe = list(oe) if isinstance(oe, tuple) else [oe]
    for o, g in zip(oe, fx["eval_outputs"]):
        assert torch.allclose(o, torch.tensor(g), rtol=1e-4, atol=1e-6)


def test_adam_loss_curve_golden():
    """First steps of the reference's Adam loss curve (torch.optim.Adam on the reference module)."""
    import model_checks as mc
    fx = _golden("curve_no.json")
    model = mc.build_model("no")
    orc = mc.oracle_for("no", model)
    img, x0, tgt = po.synthetic_batch("no", seed=1, **fx["shapes"])
    for i in range(4):
        _, loss = orc.train_step(img, x0, tgt, fx["loss_cfg"], lr=fx["lr"])
        assert abs(float(loss) - fx["losses"][i]) <= 2e-3 * abs(fx["losses"][i]), (i, float(loss), fx["losses"][i])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kind", ["no", "tdo", "tdo_v2"])
def test_oracle_vs_live_reference(kind):
    """Direct comparison with the unmodified reference modules (build container only)."""
    ref = ref_shim.load()
    m = ref_shim.build_reference_model(ref, kind)
    orc = po.OracleEstimator(kind, m.state_dict())
    shape = dict(n=2) if kind == "no" else dict(n=2, s=2)
    img, x0, tgt = po.synthetic_batch(kind, seed=3, **shape)
    m.train()
    if kind in ("tdo", "tdo_v2"):
        m.reset_initial_state(2)
    with ref_shim.quiet():
        want = m(img, None, x0)
    got = orc.forward(img, x0, training=True)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    sd = m.state_dict()
    for k in sd:
        assert torch.allclose(orc.sd[k].float(), sd[k].float(), rtol=1e-5, atol=1e-6), k


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("kind", ["no", "tdo"])
def test_oracle_depth_branch_vs_live_reference(kind):
    """use_depth=True: aux features gated by the pooled, instance-normalised depth map (models/naive.py:324-330)."""
    ref = ref_shim.load()
    m = ref_shim.build_reference_model(ref, kind, use_depth=True)
    with torch.no_grad():      # non-trivial affine parameters
        m.depth_nets[0].module[2].weight.fill_(0.7)
        m.depth_nets[0].module[2].bias.fill_(0.2)
    orc = po.OracleEstimator(kind, m.state_dict())
    shape = dict(n=2) if kind == "no" else dict(n=2, s=2)
    img, x0, tgt = po.synthetic_batch(kind, seed=4, **shape)
    depth = torch.rand(*img.shape[:-3], 1, 224, 224, generator=torch.Generator().manual_seed(9))
    m.train()
    if kind == "tdo":
        m.reset_initial_state(2)
    with ref_shim.quiet():
        want = m(img, depth, x0)
    got = orc.forward(img, x0, training=True, depth=depth)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_tf32_rounding_matches_cvt_rna():
    """round_tf32 = cvt.rna.tf32.f32: 10 explicit mantissa bits, round to nearest, ties away from zero."""
    x = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 - 2 ** -20, 1.0 + 2 ** -10, -1.0 - 2 ** -11, 3.14159, 0.0])
    want = torch.tensor([1.0, 1.0 + 2 ** -10, 1.0, 1.0 + 2 ** -10, -1.0 - 2 ** -10, 3.140625, 0.0])
    assert torch.equal(po.round_tf32(x), want)
    assert torch.equal(po.round_tf32(x.double()), want.double())


def test_teacher_forcing_puts_two_precisions_on_the_same_masks():
    """The machinery the GPU gradient tests rely on, exercised between two CPU realisations of the oracle (float32
    and float64 accumulation, both with TF32 operands) on the 4-block trunk: un-forced, their trunk gradients differ
    by several percent (ReLU masks flip); with the float32 run's conv outputs teacher-forced into the float64 run they
    agree to well under 1e-2, and every convolution's own output still matches to 5e-4."""
    import model_checks as mc
    mc.SHALLOW[0] = True
    try:
        model = mc.build_model("tdo")
    finally:
        mc.SHALLOW[0] = False
    img, x0, tgt = po.synthetic_batch("tdo", 2, s=2, seed=1)
    lk = mc.CONFIGS["tdo"]["loss"]

    def run(dt, force=None, record=False):
        orc = mc.oracle_for("tdo", model)
        orc.sd = {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in orc.sd.items()}
        po.RECORD[0] = [] if record else None
        try:
            with po.tf32_operands(True):
                if force is not None:
                    with po.forced_conv_outputs(force) as errs:
                        _, _, g = orc.loss_and_grads(img.to(dt), x0.to(dt), tgt.to(dt), lk)
                        assert len(errs) == 17 and max(errs) <= 5e-4, errs
                else:
                    _, _, g = orc.loss_and_grads(img.to(dt), x0.to(dt), tgt.to(dt), lk)
            return g, po.RECORD[0]
        finally:
            po.RECORD[0] = None

    g32, rec = run(torch.float32, record=True)
    g64_forced, _ = run(torch.float64, force=rec)
    g64_free, _ = run(torch.float64)
    forced = max(mc.relnorm(g32[k], g64_forced[k]) for k in g32 if g32[k] is not None)
    free = max(mc.relnorm(g32[k], g64_free[k]) for k in g32 if g32[k] is not None)
    assert forced <= 5e-3, forced
    assert free > 3 * forced, (free, forced)


def test_resnet18_trunk_vs_live_reference():
    """import_resnet(18, ...) (util/model_utils.py:130-136, BasicBlock trunk): the oracle's restatement and the mirror's
    seed-0 state_dict against the reference's own NaiveObjectStateEstimator built with num_resnet_layers=18."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not available")
    import model_checks as mc
    ref = ref_shim.load()
    m_ref = ref_shim.build_reference_model(ref, "no", num_resnet_layers=18)
    mirror = mc.build_model("no", layers=18)
    sd_ref, sd = m_ref.state_dict(), mirror.state_dict()
    assert list(sd_ref) == list(sd)
    for k in sd_ref:
        assert torch.equal(sd_ref[k], sd[k]), k
    img, x0, tgt = po.synthetic_batch("no", 2, seed=1)
    lk = mc.CONFIGS["no"]["loss"]
    with torch.no_grad():          # keep the ReLU'd quaternion alive (quirk Q2/Q7)
        getattr(m_ref, "fc%d" % (m_ref.n_fc - 1)).module.bias.fill_(0.5)
    orc = po.OracleEstimator("no", m_ref.state_dict())
    m_ref.train()
    with ref_shim.quiet():
        out = m_ref(img, None, x0)
    loss = ref.losses.PoseDistanceLoss(**lk)(out, tgt)
    loss.backward()
    out_o, loss_o, grads = orc.loss_and_grads(img, x0, tgt, lk)
    assert mc.rel(out_o, out) <= 1e-5
    assert abs(float(loss_o) - float(loss)) <= 1e-5 * abs(float(loss))
    named = dict(m_ref.named_parameters())
    for k in ("feature_net.module.conv1.weight", "feature_net.module.layer2.0.downsample.0.weight",
              "feature_net.module.layer4.1.conv2.weight", "fc0.module.weight"):
        assert mc.relnorm(grads[k], named[k].grad) <= 1e-3, k
