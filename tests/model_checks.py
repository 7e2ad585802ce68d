"""End-to-end parity of the CUDA estimators against the CPU oracle (GPU only).

`python tests/model_checks.py [kinds...]` prints per-output / per-parameter errors without stopping;
tests/test_models_gpu.py asserts on the same numbers.  Tolerances (TF32 operands, fp32 accumulate):
outputs and loss 1e-3 relative (north_star), gradients ||g-g_ref|| / ||g_ref|| <= 1e-2 per parameter.
"""
import contextlib
import io
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200"))
sys.path.insert(0, ROOT)

from oracle import pose_oracle as po  # noqa: E402

CONFIGS = {
    # kind: (ctor kwargs, loss kwargs)  -- hyper-parameters from scripts/train_no.sbatch / train_tdo.sbatch
    "no": dict(latent=512, hidden=[1024, 256, 64], loss=dict(distance_metric="combined", alpha=0.5, mode="pose")),
    "tdo": dict(latent=512, hidden=512, loss=dict(distance_metric="combined", alpha=0.5, mode="pose")),
    "tdo_v2": dict(latent=512, hidden=512, loss=dict(distance_metric="combined", alpha=0.5, mode="pose")),
    "td": dict(latent=1024, hidden=512, loss=dict(distance_metric="l2", alpha=0.5, mode="pose")),
    "n": dict(latent=1024, hidden=[512], loss=dict(distance_metric="l2", alpha=0.5, mode="pose")),
}


SHALLOW = [False]   # when set, Bottleneck stacks are [1,1,1,1] instead of ResNet-50's [3,4,6,3]


def build_model(kind, seed=0, latent=None, layers=50):
    import models.naive as mn
    import models.time_sensitive as mt
    import util.model_utils as mu
    cfg = dict(CONFIGS[kind])
    if latent is not None:
        cfg["latent"] = latent
    mu._RESNET_LAYERS[50] = [1, 1, 1, 1] if SHALLOW[0] else [3, 4, 6, 3]
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "no":
            return mn.NaiveObjectStateEstimator("cube", list(cfg["hidden"]), layers, cfg["latent"], False, (9,), False,
                                                False)
        if kind == "tdo":
            return mt.TemporallyDependentObjectStateEstimator("robot1_eef", cfg["hidden"], layers, cfg["latent"], 20,
                                                              feature_extract=False, use_pretrained=False)
        if kind == "tdo_v2":
            return mt.TemporallyDependentObjectStateEstimatorV2("robot1_eef", cfg["hidden"], 64, 50, cfg["latent"], 20,
                                                                feature_extract=False, use_pretrained=False)
        if kind == "td":
            return mt.TemporallyDependentStateEstimator(cfg["hidden"], cfg["hidden"], 50, cfg["latent"], 10,
                                                        feature_extract=False, use_pretrained=False)
        if kind == "n":
            orig = mn.import_resnet
            mn.import_resnet = lambda n, o, fe=True, use_pretrained=True: orig(n, o, fe, use_pretrained=False)
            try:
                return mn.NaiveEndEffectorStateEstimator(list(cfg["hidden"]), list(cfg["hidden"]), 50, cfg["latent"],
                                                         False)
            finally:
                mn.import_resnet = orig
    raise ValueError(kind)


def oracle_for(kind, model):
    extra = None
    if kind == "td":
        extra = {"aux_w": model.aux_nets[0][0].weight.detach().cpu(), "aux_b": model.aux_nets[0][0].bias.detach().cpu()}
    return po.OracleEstimator(kind, {k: v.detach().cpu() for k, v in model.state_dict().items()}, extra)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def relnorm(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def check_train_step(kind, n=2, s=2, seed=1, verbose=False, latent=None, layers=50):
    """forward (train mode) + loss + backward vs oracle.  Returns rows (name, err, tol)."""
    from models.losses import PoseDistanceLoss
    cfg = CONFIGS[kind]
    model = build_model(kind, latent=latent, layers=layers)
    orc = oracle_for(kind, model)
    if kind in ("no", "n"):
        img, x0, tgt = po.synthetic_batch(kind, n, seed=seed)
    else:
        img, x0, tgt = po.synthetic_batch(kind, n, s=s, seed=seed)
    lk = cfg["loss"]
    t0 = time.time()
    if kind in ("no", "tdo", "tdo_v2"):
        outs_ref, loss_ref, grads_ref = orc.loss_and_grads(img, x0, tgt, lk)
        outs_ref = (outs_ref,)
    else:
        # two-headed models train on loss(pre, x0) + loss(post, x1)  (util/learn_utils.py:166-172)
        for k in orc.param_names:
            orc.sd[k].requires_grad_(True)
            orc.sd[k].grad = None
        pre, post = orc.forward(img, x0, training=True)
        loss_ref = po.pose_loss(pre, x0, **lk) + po.pose_loss(post, tgt, **lk)
        loss_ref.backward()
        grads_ref = {k: orc.sd[k].grad for k in orc.param_names}
        for k in orc.param_names:
            orc.sd[k].requires_grad_(False)
        outs_ref = (pre.detach(), post.detach())
        loss_ref = loss_ref.detach()
    t_cpu = time.time() - t0

    model.cuda().train()
    crit = PoseDistanceLoss(distance_metric=lk["distance_metric"], alpha=lk["alpha"], mode=lk["mode"])
    if kind in ("td", "tdo", "tdo_v2"):
        model.reset_initial_state(n)
    out = model(img.cuda(), None, x0.cuda())
    outs = out if isinstance(out, tuple) else (out,)
    if kind in ("no", "tdo", "tdo_v2"):
        loss = crit(outs[0], tgt.cuda())
    else:
        loss = crit(outs[0], x0.cuda()) + crit(outs[1], tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    rows = []
    tag = "%s n%d" % (kind, n) + ("" if kind in ("no", "n") else " s%d" % s)
    for i, (a, b) in enumerate(zip(outs, outs_ref)):
        rows.append(("%s out%d" % (tag, i), rel(a, b), 1e-3))
    rows.append(("%s loss" % tag, rel(loss.reshape(1), loss_ref.reshape(1)), 1e-3))
    named = dict(model.named_parameters())
    worst = []
    for k in orc.param_names:
        g_ref = grads_ref[k]
        g = named[k].grad
        if g_ref is None:
            rows.append(("%s grad %s is None" % (tag, k), 0.0 if g is None else 1.0, 0.0))
            continue
        if g is None:
            rows.append(("%s grad %s MISSING" % (tag, k), 1.0, 0.0))
            continue
        worst.append((relnorm(g, g_ref), k))
    worst.sort(reverse=True)
    for e, k in (worst if verbose else worst[:6]):
        rows.append(("%s grad %s (|g_ref| %.2e)" % (tag, k, float(grads_ref[k].norm())), e, 1e-2))
    convs = [e for e, k in worst if ("conv" in k or "downsample.0" in k) and k.endswith("weight")]
    if convs:
        rows.append(("%s worst conv-weight grad" % tag, max(convs), 1e-2))
        rows.append(("%s median conv-weight grad" % tag, sorted(convs)[len(convs) // 2], 1e-2))
    # running statistics after one training forward
    sd = model.state_dict()
    e_rm = max(rel(sd[k], orc.sd[k]) for k in sd if k.endswith("running_mean"))
    e_rv = max(rel(sd[k], orc.sd[k]) for k in sd if k.endswith("running_var"))
    nbt_ok = all(int(sd[k]) == int(orc.sd[k]) for k in sd if k.endswith("num_batches_tracked"))
    rows.append(("%s running_mean (max over BNs)" % tag, e_rm, 2e-3))
    rows.append(("%s running_var (max over BNs)" % tag, e_rv, 2e-3))
    rows.append(("%s num_batches_tracked" % tag, 0.0 if nbt_ok else 1.0, 0.0))
    rows.append(("%s [oracle cpu seconds]" % tag, t_cpu, float("inf")))

    # eval-mode forward with the same (now updated) running statistics
    model.eval()
    with torch.no_grad():
        if kind in ("td", "tdo", "tdo_v2"):
            model.reset_initial_state(n)
        oe = model(img.cuda(), None, x0.cuda())
    oe = oe if isinstance(oe, tuple) else (oe,)
    ref_e = orc.forward(img, x0, training=False)
    ref_e = ref_e if isinstance(ref_e, tuple) else (ref_e,)
    for i, (a, b) in enumerate(zip(oe, ref_e)):
        rows.append(("%s eval out%d" % (tag, i), rel(a, b), 2e-3))
    # the same eval forward against the oracle at the CUDA path's operand precision (TF32 operands, float64
    # accumulation), on the running statistics the CUDA model now holds: what is left is accumulation order
    orc2 = oracle_for(kind, model)
    orc2.sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in orc2.sd.items()}
    orc2.extra = {k: v.double() for k, v in orc2.extra.items()}
    with po.tf32_operands(True), torch.no_grad():
        ref_t = orc2.forward(img.double(), x0.double(), training=False)
    ref_t = ref_t if isinstance(ref_t, tuple) else (ref_t,)
    for i, (a, b) in enumerate(zip(oe, ref_t)):
        # 3e-3: with TF32 operands the eval-mode network (no per-layer re-normalisation) is itself sensitive to the
        # accumulation order -- float32 vs float64 accumulation of the SAME TF32-operand oracle differ by 2-3e-3 on
        # the CPU (measured); measured against the CUDA path: 1.0e-3 .. 1.5e-3
        rows.append(("%s eval out%d [tf32-operand oracle]" % (tag, i), rel(a, b), 3e-3))
    return rows


def oracle_loss_and_grads(orc, kind, img, x0, tgt, lk):
    """(outputs tuple, loss, {name: grad}) of the training objective of util/learn_utils.py:160-172: the object-pose
    models train on loss(out, obj), the two-headed ones on loss(pre, x0) + loss(post, x1)."""
    if kind in ("no", "tdo", "tdo_v2"):
        outs, loss, grads = orc.loss_and_grads(img, x0, tgt, lk)
        return (outs,), loss, grads
    for k in orc.param_names:
        orc.sd[k].requires_grad_(True)
        orc.sd[k].grad = None
    pre, post = orc.forward(img, x0, training=True)
    loss = po.pose_loss(pre, x0, **lk) + po.pose_loss(post, tgt, **lk)
    loss.backward()
    grads = {k: orc.sd[k].grad for k in orc.param_names}
    for k in orc.param_names:
        orc.sd[k].requires_grad_(False)
    return (pre.detach(), post.detach()), loss.detach(), grads


def check_forced(kind, n=2, s=2, seed=1, verbose=False, layers=50):
    """Full-depth training step against the oracle at the SAME operand precision (TF32 operands, float64 accumulation)
    and on the SAME ReLU masks: the raw convolution outputs the CUDA path saved for its backward pass are teacher-forced
    into the oracle's forward (oracle/pose_oracle.py: tf32_operands, forced_conv_outputs).  Every convolution is still
    computed by the oracle and compared (per-layer forward rows); BatchNorm, ReLU, pooling, heads, loss and the whole
    backward pass are the oracle's own autograd.  Returns rows (name, err, tol)."""
    from models.losses import PoseDistanceLoss
    from pe_b200 import engine
    cfg = CONFIGS[kind]
    lk = cfg["loss"]
    model = build_model(kind, layers=layers)
    with torch.no_grad():      # finite loss for the models that ReLU their output (quirk Q2/Q7)
        if kind == "no":
            getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        elif kind == "n":
            getattr(model, "pre_fc%d" % (model.n_pre_hidden - 1)).bias.fill_(0.5)
            getattr(model, "post_fc%d" % (model.n_post_hidden - 1)).bias.fill_(0.5)
    orc = oracle_for(kind, model)
    orc.sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in orc.sd.items()}
    orc.extra = {k: v.double() for k, v in orc.extra.items()}
    if kind in ("no", "n"):
        img, x0, tgt = po.synthetic_batch(kind, n, seed=seed)
    else:
        img, x0, tgt = po.synthetic_batch(kind, n, s=s, seed=seed)

    model.cuda().train()
    crit = PoseDistanceLoss(distance_metric=lk["distance_metric"], alpha=lk["alpha"], mode=lk["mode"])
    if kind in ("td", "tdo", "tdo_v2"):
        model.reset_initial_state(n)
    from pe_b200 import estimators
    engine.CAPTURE_CONV_OUTPUTS[0] = []
    estimators.CAPTURE_HEAD_OUTPUTS[0] = []
    try:
        out = model(img.cuda(), None, x0.cuda())
        acts = engine.CAPTURE_CONV_OUTPUTS[0]
        heads = [h.detach().double().cpu() for h in estimators.CAPTURE_HEAD_OUTPUTS[0]]
    finally:
        engine.CAPTURE_CONV_OUTPUTS[0] = None
        estimators.CAPTURE_HEAD_OUTPUTS[0] = None
    outs = out if isinstance(out, tuple) else (out,)
    loss = crit(outs[0], tgt.cuda()) if kind in ("no", "tdo", "tdo_v2") else \
        crit(outs[0], x0.cuda()) + crit(outs[1], tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    ys = [a.t.view(a.B, a.H, a.W, a.C).permute(0, 3, 1, 2).double().cpu() for a in acts]
    del acts

    t0 = time.time()
    with po.tf32_operands(True), po.forced_conv_outputs(ys) as ferr, po.forced_head_outputs(heads or None):
        outs_ref, loss_ref, grads_ref = oracle_loss_and_grads(orc, kind, img.double(), x0.double(), tgt.double(), lk)
        ferr = list(ferr)
    t_cpu = time.time() - t0
    tag = "forced %s n%d" % (kind, n) + ("" if kind in ("no", "n") else " s%d" % s)
    n_conv = len(ys)
    rows = [("%s conv forward, worst of %d layers" % (tag, n_conv), max(ferr[:n_conv]), 5e-4)]
    if len(ferr) > n_conv:
        # the naive heads' layer outputs (incl. the final pose) are forced too: each layer's own result is the check
        rows.append(("%s head layers forward (incl. output), worst of %d" % (tag, len(ferr) - n_conv),
                     max(ferr[n_conv:]), 1e-3))
    else:
        for i, (a, b) in enumerate(zip(outs, outs_ref)):
            rows.append(("%s out%d" % (tag, i), rel(a, b), 1e-3))
    rows.append(("%s loss" % tag, rel(loss.reshape(1), loss_ref.reshape(1)), 1e-3))
    named = dict(model.named_parameters())
    errs = []
    for k in orc.param_names:
        g_ref, g = grads_ref[k], named[k].grad
        if g_ref is None or g is None:
            rows.append(("%s grad %s is None" % (tag, k), 0.0 if (g is None) == (g_ref is None) else 1.0, 0.0))
            continue
        if g_ref.numel() == 1 and k.endswith(".bias"):
            # a single scalar that is a signed sum of thousands of terms (the aux conv's bias: the sum of the aux
            # gradient over all pixels) can come out arbitrarily close to zero, which makes |dg| / |g| meaningless;
            # its error is measured against the scale of that sum -- the norm of the sibling weight's gradient, which
            # is the same sum weighted by O(1) activations, spread over its elements
            w_ref = grads_ref[k[:-len("bias")] + "weight"]
            scale = max(float(g_ref.abs()), float(w_ref.norm()) / w_ref.numel() ** 0.5)
            errs.append((float((g.detach().double().cpu() - g_ref.double()).abs()) / scale, k))
        else:
            errs.append((relnorm(g, g_ref), k))
    if verbose:
        for e, k in errs:                     # network order
            rows.append(("%s grad %s" % (tag, k), e, 1e-2))
    errs.sort(reverse=True)
    if not verbose:
        for e, k in errs[:5]:
            rows.append(("%s grad %s" % (tag, k), e, 1e-2))
    rows.append(("%s worst grad over %d parameters" % (tag, len(errs)), errs[0][0], 1e-2))
    rows.append(("%s median grad" % tag, errs[len(errs) // 2][0], 1e-2))
    sd = model.state_dict()
    rows.append(("%s running_mean (max over BNs)" % tag,
                 max(rel(sd[k], orc.sd[k]) for k in sd if k.endswith("running_mean")), 1e-3))
    rows.append(("%s running_var (max over BNs)" % tag,
                 max(rel(sd[k], orc.sd[k]) for k in sd if k.endswith("running_var")), 1e-3))
    rows.append(("%s num_batches_tracked" % tag,
                 0.0 if all(int(sd[k]) == int(orc.sd[k]) for k in sd if k.endswith("num_batches_tracked")) else 1.0, 0.0))
    rows.append(("%s [oracle cpu seconds]" % tag, t_cpu, float("inf")))
    return rows


def check_rollout(kind="tdo", steps=3):
    """Batch-1 streaming inference with carried LSTM state vs the plain fp32 oracle and vs the oracle at the CUDA
    path's operand precision (TF32 trunk operands, fp32 fused head; float64 accumulation)."""
    model = build_model(kind)
    orc = oracle_for(kind, model)
    orc_t = oracle_for(kind, model)
    orc_t.sd = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in orc_t.sd.items()}
    orc_t.extra = {k: v.double() for k, v in orc_t.extra.items()}
    model.cuda().eval()
    model.rollout = True
    model.reset_initial_state(1)
    orc.reset_state(1)
    orc_t.reset_state(1)
    if orc_t.state is not None:
        def dbl(t):
            return tuple(dbl(u) for u in t) if isinstance(t, tuple) else t.double()
        orc_t.state = dbl(orc_t.state)
    rows = []
    for t in range(steps):
        img, x0, _ = po.synthetic_batch(kind, 1, s=1, seed=10 + t)
        with torch.no_grad():
            o = model(img.cuda(), None, x0.cuda())
            r = orc.forward(img, x0, training=False, rollout=True)
            with po.tf32_operands(True):
                rt = orc_t.forward(img.double(), x0.double(), training=False, rollout=True)
        o = o if isinstance(o, tuple) else (o,)
        r = r if isinstance(r, tuple) else (r,)
        rt = rt if isinstance(rt, tuple) else (rt,)
        rows.append(("%s rollout step %d" % (kind, t), max(rel(a, b) for a, b in zip(o, r)), 2e-3))
        # 8e-3: a fresh model's running statistics (mean 0, var 0.9 after the constructor's dummy forward) leave the
        # eval network un-normalised, where float32 vs float64 accumulation of the same TF32-operand oracle already
        # differ by 2.3e-3 .. 3.1e-3 on the CPU; measured against the CUDA path: 2e-3 .. 5e-3
        rows.append(("%s rollout step %d [tf32-operand oracle]" % (kind, t), max(rel(a, b) for a, b in zip(o, rt)), 8e-3))
    return rows


def calibrate(kind="no", n=2, s=2, seed=1):
    """How far does torch's own cuDNN path (fp32 and TF32) land from the CPU fp32 oracle on the same
    problem?  Uses the oracle restatement on the GPU; this is the yard-stick for the TF32 tolerances."""
    cfg = CONFIGS[kind]
    model = build_model(kind)
    orc = oracle_for(kind, model)
    img, x0, tgt = po.synthetic_batch(kind, n, seed=seed) if kind in ("no", "n") else po.synthetic_batch(kind, n, s=s, seed=seed)
    outs_ref, loss_ref, grads_ref = orc.loss_and_grads(img, x0, tgt, cfg["loss"])
    rows = []
    for tf in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf
        torch.backends.cudnn.allow_tf32 = tf
        g = oracle_for(kind, model)
        g.sd = {k: v.cuda() for k, v in g.sd.items()}
        outs, loss, grads = g.loss_and_grads(img.cuda(), x0.cuda(), tgt.cuda(), cfg["loss"])
        tag = "calib torch-%s %s n%d" % ("tf32" if tf else "fp32", kind, n)
        rows.append((tag + " out", rel(outs, outs_ref), float("inf")))
        rows.append((tag + " loss", rel(loss.reshape(1), loss_ref.reshape(1)), float("inf")))
        worst = sorted(((relnorm(grads[k], grads_ref[k]), k) for k in grads_ref if grads_ref[k] is not None), reverse=True)
        for e, k in worst[:4]:
            rows.append((tag + " grad " + k, e, float("inf")))
        convs = [e for e, k in worst if k.endswith("conv1.weight") or k.endswith("conv2.weight") or k.endswith("conv3.weight")]
        rows.append((tag + " worst conv weight grad", max(convs), float("inf")))
        rm = max(rel(g.sd[k], orc.sd[k]) for k in g.sd if k.endswith("running_var"))
        rows.append((tag + " running_var", rm, float("inf")))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return rows


def main(argv):
    if "--shallow" in argv and "--calib" in argv:
        SHALLOW[0] = True
        argv = [a for a in argv if a != "--shallow"]
    if "--calib" in argv:
        argv = [a for a in argv if a != "--calib"]
        n = 2
        for a in list(argv):
            if a.startswith("n="):
                n = int(a[2:]); argv.remove(a)
        for name, err, tol in calibrate(argv[0] if argv else "no", n=n):
            print("%-70s %.3e" % (name, err), flush=True)
        return 0
    if "--forced" in argv:
        argv = [a for a in argv if a != "--forced"]
        nb = 2
        for a in list(argv):
            if a.startswith("n="):
                nb = int(a[2:]); argv.remove(a)
        nfail = 0
        for kind in [a for a in argv if not a.startswith("-")] or ["no", "tdo", "td", "n", "tdo_v2"]:
            for name, err, tol in check_forced(kind, n=nb, verbose=("-v" in argv)):
                ok = err <= tol
                nfail += (not ok)
                print("%-4s %-72s err %.3e tol %.1e" % ("ok" if ok else "FAIL", name, err, tol), flush=True)
        return nfail
    nb = 2
    if "--shallow" in argv:
        SHALLOW[0] = True
        argv = [a for a in argv if a != "--shallow"]
    for a in list(argv):
        if a.startswith("n="):
            nb = int(a[2:]); argv.remove(a)
    kinds = [a for a in argv if not a.startswith("-")] or ["no", "tdo", "td", "n"]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    nfail = 0
    from pe_b200 import native
    for kind in kinds:
        try:
            rows = check_train_step(kind, n=nb, verbose=("-v" in argv))
            if kind in ("tdo", "td"):
                rows += check_rollout(kind)
        except Exception as e:
            import traceback
            traceback.print_exc()
            rows = [("EXCEPTION %s: %r" % (kind, e), float("inf"), 0.0)]
        for name, err, tol in rows:
            ok = err <= tol
            nfail += (not ok)
            print("%-4s %-64s err %.3e tol %.1e" % ("ok" if ok else "FAIL", name, err, tol), flush=True)
    print("device error flag:", native.lib().pe_device_error())
    print("FAILED: %d" % nfail)
    return nfail


if __name__ == "__main__":
    sys.exit(1 if main([a for a in sys.argv[1:]]) else 0)
