"""Generate tests/golden/*.json from the UNMODIFIED reference modules (build container only).

TEST INFRASTRUCTURE.  Run as `python oracle/make_golden.py` where /root/reference exists.  The fixtures
pin (a) the oracle restatement (tests/test_oracle_cpu.py, CPU) and (b) the CUDA path (tests/test_models_gpu.py)
to what the reference's own code computes, on machines where the reference tree is not available.

Contents
  loss_vectors.json   PoseDistanceLoss known-answer vectors: every metric x mode, loss + gradient, 'val'
  state_dicts.json    key / shape / dtype manifest of each estimator's state_dict and parameter order,
                      plus checksums of the seed-0 random init (sum, sum of squares per tensor)
  forward_<kind>.json seed-0 weights, seed-1 synthetic batch: outputs, loss, per-parameter gradient norms,
                      BN running-stat checksums after one training forward, eval-mode outputs
  curve_<kind>.json   loss curve of N Adam steps (torch.optim.Adam, lr 1e-3) on a fixed synthetic batch
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pose_oracle as po  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LOSS_CFG = {"no": dict(distance_metric="combined", alpha=0.5, mode="pose"),
            "tdo": dict(distance_metric="combined", alpha=0.5, mode="pose"),
            "tdo_v2": dict(distance_metric="combined", alpha=0.5, mode="pose"),
            "td": dict(distance_metric="l2", alpha=0.5, mode="pose"),
            "n": dict(distance_metric="l2", alpha=0.5, mode="pose")}
SHAPES = {"no": dict(n=2), "n": dict(n=2), "tdo": dict(n=2, s=2), "td": dict(n=2, s=2), "tdo_v2": dict(n=2, s=2)}


def dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f)
    print("wrote", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


def loss_vectors(ref):
    g = torch.Generator().manual_seed(7)
    pred = torch.randn(5, 7, generator=g)
    truth = torch.randn(5, 7, generator=g)
    truth[:, 3:] = truth[:, 3:] / truth[:, 3:].norm(dim=1, keepdim=True)
    truth[:, 6] = truth[:, 6].abs()
    cases = [("random5", pred, truth),
             ("survey_appendix_c", torch.tensor([[0.1, 0.2, 0.3, 0.5, 0.5, 0.5, 0.5], [-0.4, 0.25, 0.0, 0.1, -0.2, 0.3, -0.4]]),
              torch.tensor([[0., 0., 0., 0., 0., 0., 1.], [0.1, 0.25, -0.3, 0., 0.6, 0., 0.8]])),
             ("edge_ties", torch.tensor([[0.3, 0.3, 0., 0., 0., 1., 0.]]), torch.tensor([[0., 0., 0., 0., 0., 1., 0.]])),
             ("seq_shape", pred[:4].reshape(2, 2, 7), truth[:4].reshape(2, 2, 7))]
    out = []
    for name, p, t in cases:
        for metric in ("l1", "l2", "linf", "combined"):
            for mode in ("position", "pose"):
                pp = p.clone().requires_grad_(True)
                crit = ref.losses.PoseDistanceLoss(distance_metric=metric, scale_factor=1.5, alpha=0.5, mode=mode)
                loss = crit(pp, t)
                loss.backward()
                out.append(dict(case=name, metric=metric, mode=mode, scale_factor=1.5, alpha=0.5, pred=p.tolist(),
                                truth=t.tolist(), loss=float(loss), grad=pp.grad.tolist()))
        crit = ref.losses.PoseDistanceLoss(mode="val")
        pos, ang = crit(p, t)
        out.append(dict(case=name, metric="l2", mode="val", pred=p.tolist(), truth=t.tolist(), pos=float(pos),
                        angle=float(ang)))
    dump("loss_vectors.json", out)


def manifest(ref):
    out = {}
    for kind in ("no", "n", "td", "tdo", "tdo_v2"):
        m = ref_shim.build_reference_model(ref, kind)
        sd = m.state_dict()
        out[kind] = dict(
            keys=[[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
            params=[n for n, _ in m.named_parameters()],
            checksum={k: [float(v.double().sum()), float((v.double() ** 2).sum())] for k, v in sd.items()},
        )
    dump("state_dicts.json", out)


def run_reference(ref, kind, img, x0, tgt, lk):
    m = ref_shim.build_reference_model(ref, kind)
    m.train()
    if kind in ("td", "tdo", "tdo_v2"):
        m.reset_initial_state(img.shape[1])
    crit = ref.losses.PoseDistanceLoss(**lk)
    out = m(img, None, x0)
    if kind in ("no", "tdo", "tdo_v2"):
        loss = crit(out, tgt)
        outs = [out]
    else:
        loss = crit(out[0], x0) + crit(out[1], tgt)
        outs = list(out)
    loss.backward()
    return m, outs, loss


def forward_fixture(ref, kind):
    lk = LOSS_CFG[kind]
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **SHAPES[kind])
    with ref_shim.quiet():
        m, outs, loss = run_reference(ref, kind, img, x0, tgt, lk)
    sd = m.state_dict()
    fx = dict(kind=kind, shapes=SHAPES[kind], loss_cfg=lk, outputs=[o.detach().tolist() for o in outs],
              loss=float(loss),
              grad_norms={n: (None if p.grad is None else float(p.grad.norm())) for n, p in m.named_parameters()},
              grad_head={n: p.grad.flatten()[:8].tolist() for n, p in m.named_parameters()
                         if p.grad is not None and ("fc" in n or "rnn" in n) and n.endswith("bias")},
              running_mean_sum={k: float(v.double().sum()) for k, v in sd.items() if k.endswith("running_mean")},
              running_var_sum={k: float(v.double().sum()) for k, v in sd.items() if k.endswith("running_var")})
    m.eval()
    with torch.no_grad(), ref_shim.quiet():
        if kind in ("td", "tdo", "tdo_v2"):
            m.reset_initial_state(img.shape[1])
        oe = m(img, None, x0)
    oe = list(oe) if isinstance(oe, tuple) else [oe]
    fx["eval_outputs"] = [o.tolist() for o in oe]
    dump("forward_%s.json" % kind, fx)


def curve_fixture(ref, kind, steps, lr=1e-3, name=None, threads=None):
    lk = LOSS_CFG[kind]
    shape = dict(n=4) if kind in ("no", "n") else dict(n=2, s=2)
    img, x0, tgt = po.synthetic_batch(kind, seed=1, **shape)
    if threads:
        torch.set_num_threads(threads)
    with ref_shim.quiet():
        m = ref_shim.build_reference_model(ref, kind)
    m.train()
    crit = ref.losses.PoseDistanceLoss(**lk)
    opt = torch.optim.Adam(m.parameters(), lr=lr)
    losses = []
    for i in range(steps):
        opt.zero_grad()
        if kind in ("td", "tdo"):
            m.reset_initial_state(img.shape[1])
        out = m(img, None, x0)
        loss = crit(out, tgt) if kind in ("no", "tdo") else crit(out[0], x0) + crit(out[1], tgt)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        if i % 10 == 0:
            print(kind, i, losses[-1], flush=True)
    torch.set_num_threads(os.cpu_count())
    dump(name or "curve_%s.json" % kind, dict(kind=kind, shapes=shape, loss_cfg=lk, lr=lr, losses=losses,
                                              threads=threads or os.cpu_count()))


def main():
    torch.set_num_threads(os.cpu_count())
    ref = ref_shim.load()
    what = sys.argv[1:] or ["loss", "manifest", "forward", "curve"]
    if "loss" in what:
        loss_vectors(ref)
    if "manifest" in what:
        manifest(ref)
    if "forward" in what:
        for kind in ("no", "tdo", "td", "n", "tdo_v2"):
            forward_fixture(ref, kind)
    if "forward_v2" in what:
        forward_fixture(ref, "tdo_v2")
    if "curve" in what:
        curve_fixture(ref, "no", 100)
        curve_fixture(ref, "tdo", 30)
    if "curve_self_noise" in what:
        # the reference against ITSELF at the scripts' default lr 1e-3: same modules, same seed, same data, only the
        # number of CPU threads (= the summation order inside oneDNN) differs.  The two realisations part ways after a
        # handful of steps, which is the yard-stick for what "the same loss curve" can mean at this learning rate.
        curve_fixture(ref, "tdo", 30, lr=1e-3, name="curve_tdo_lr1e-3_threads1.json", threads=1)
        curve_fixture(ref, "tdo", 30, lr=1e-3, name="curve_tdo_lr1e-3_threads3.json", threads=3)
    if "curve_small_lr" in what:
        # lr = 1e-5: the smooth regime, where a 100-step curve is a meaningful pointwise target (at the
        # scripts' default 1e-3 the tiny synthetic batch puts training in a chaotic regime after ~4 steps)
        curve_fixture(ref, "no", 100, lr=1e-5, name="curve_no_lr1e-5.json")
        curve_fixture(ref, "tdo", 100, lr=1e-5, name="curve_tdo_lr1e-5.json")


if __name__ == "__main__":
    main()
