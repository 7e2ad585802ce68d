#!/usr/bin/env python
"""Benchmark of the pose-estimator hot path (driver contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model no|tdo|td|n]

A "step" is one training step (forward, pose loss, backward, Adam) of the naive-object estimator on a
synthetic batch of 256 RGB frames per GPU -- BASELINE.json configs[1] ("naive model training on
1 B200, synthetic batch 256, hammer target").  `value` is whole-job samples/s with inputs resident in
HBM; `e2e` is the same step driven from pinned HOST buffers (H2D of the frames / proprio / targets and
a D2H read of the loss inside the timed region, every step; the region is run twice and the faster is
reported, both are listed).  Weak scaling: each rank owns its own 256 frames.

Beside that headline (`value`, `e2e`, `roofline`, `cpu_baseline`) the same JSON line carries the rest of
BASELINE.json's metric: `configs` = the TD (config 3) and TDO (config 4) training steps, each with its own roofline
fraction, and `rollout` = config 5 (batch-1 CUDA-graph latency p50 / p99 from host frames to host pose, and frames/s
for batches 1..1024).  Under torchrun (N > 1) `value` stays config 2 (so the driver's per-N scaling arithmetic
compares like with like) and `configs.tdo` is config 4 as north_star names it: TDO, S = 20, 32 episodes per GPU.

`--impl reference` times the reference's OWN nn.Modules (models/naive.py, models/time_sensitive.py, models/losses.py,
torch.optim.Adam) on all host cores: they are imported from /root/reference in the build container and from the
git-ignored copy oracle/_ref (oracle/build_ref.py) on the GPU box; only if neither exists does it fall back to the
oracle port.  `oracle/` is imported only by that leg and by the `cpu_baseline` leg; the GPU arm builds its models
and synthetic data on its own.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200"))
sys.path.insert(0, ROOT)

LOSS = dict(distance_metric="combined", alpha=0.5, mode="pose")          # scripts/train_no.sbatch:61-83
TRAIN_GFLOP_PER_FRAME = {"no": 24.32, "tdo": 24.35, "td": 24.42, "n": 24.30, "tdo_v2": 24.35}
SEQ_KINDS = ("td", "tdo", "tdo_v2")
TRAIN_MB_PER_FRAME = 223.0                                                # SURVEY 8(d) compulsory fp32 traffic


def resnet50_convs():
    """(H, Cin, Cout, k, stride, pad) of every convolution of torchvision's resnet50 at 224x224 input
    (util/model_utils.py:10-31 builds exactly this trunk)."""
    convs = [(224, 3, 64, 7, 2, 3)]
    H, cin = 56, 64
    for planes, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            convs += [(H, cin, planes, 1, 1, 0), (H, planes, planes, 3, s, 1), (H // s, planes, planes * 4, 1, 1, 0)]
            if b == 0:
                convs.append((H, cin, planes * 4, 1, s, 0))
            H, cin = H // s, planes * 4
    return convs


def layerwise_bound_ms(batch, hbm_gbs, tf32_tflops):
    """Sum over every conv pass (forward, dgrad, wgrad; no dgrad for the stem) of max(algorithmic bytes / HBM peak,
    flops / TF32 peak): the time the convolutions would take if each ran at whichever roofline bounds it.  The narrow
    1x1 convolutions (K or N = 64..128) are HBM-bound, the 3x3 ones tensor-bound, so neither peak alone describes the
    family.  Returns (bound ms, algorithmic GB, GFLOP)."""
    t = by_tot = fl_tot = 0.0
    for i, (H, ci, co, k, s, p) in enumerate(resnet50_convs()):
        Ho = (H + 2 * p - k) // s + 1
        by = 4.0 * (batch * H * H * ci + batch * Ho * Ho * co + co * ci * k * k)
        fl = 2.0 * batch * Ho * Ho * co * ci * k * k
        n_pass = 2 if i == 0 else 3
        t += n_pass * max(by / (hbm_gbs * 1e9), fl / (tf32_tflops * 1e12))
        by_tot += n_pass * by
        fl_tot += n_pass * fl
    return t * 1e3, by_tot / 1e9, fl_tot / 1e9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops_sustained"], which="measured")
    return dict(hbm=6650.0, bf16=1400.0, which="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:                      # let nvidia-smi finish tearing down NVML before the next (host-synchronous)
                self.proc.wait(5)     # timed region starts: its driver calls can stall kernel launches
            except Exception:
                pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# Model hyper-parameters of the reference's launchers (scripts/train_no.sbatch:61-83, train_tdo.sbatch:61-83,
# train_td.sbatch:61-62 -> scripts/train_model.py:22-42 defaults)
MODEL_CFG = {"no": dict(latent=512, hidden=[1024, 256, 64]), "tdo": dict(latent=512, hidden=512),
             "tdo_v2": dict(latent=512, hidden=512), "td": dict(latent=1024, hidden=512),
             "n": dict(latent=1024, hidden=[512])}


def build(kind, seed=0):
    """Reference constructor (the drop-in mirrors in rgb-proprioceptive-pose-estimator_b200/models) under
    torch.manual_seed(seed), random init.  The one-shot models put a ReLU after their LAST layer (reference quirk,
    models/naive.py:343-345), so a random-init network emits all-zero quaternions and the reference loss (no
    epsilon in the normalisation, models/losses.py:68-69) is NaN from step 0.  The bench keeps the arithmetic
    finite by starting the last layer's bias at +0.5; shapes, FLOPs and bytes are unchanged."""
    import contextlib
    import io
    import torch
    import models.naive as mn
    import models.time_sensitive as mt
    cfg = MODEL_CFG[kind]
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == "no":
            model = mn.NaiveObjectStateEstimator("hammer", list(cfg["hidden"]), 50, cfg["latent"], False, (9,), False,
                                                 False)
        elif kind == "tdo":
            model = mt.TemporallyDependentObjectStateEstimator("robot1_eef", cfg["hidden"], 50, cfg["latent"], 20,
                                                               feature_extract=False, use_pretrained=False)
        elif kind == "tdo_v2":
            model = mt.TemporallyDependentObjectStateEstimatorV2("robot1_eef", cfg["hidden"], 64, 50, cfg["latent"],
                                                                 20, feature_extract=False, use_pretrained=False)
        elif kind == "td":
            model = mt.TemporallyDependentStateEstimator(cfg["hidden"], cfg["hidden"], 50, cfg["latent"], 10,
                                                         feature_extract=False, use_pretrained=False)
        else:
            # the "n" constructor cannot pass use_pretrained (models/naive.py:42); there is no network here
            orig = mn.import_resnet
            mn.import_resnet = lambda n, o, fe=True, use_pretrained=True: orig(n, o, fe, use_pretrained=False)
            try:
                model = mn.NaiveEndEffectorStateEstimator(list(cfg["hidden"]), list(cfg["hidden"]), 50, cfg["latent"],
                                                          False)
            finally:
                mn.import_resnet = orig
    with torch.no_grad():
        if kind == "no":
            getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        elif kind == "n":
            getattr(model, "pre_fc%d" % (model.n_pre_hidden - 1)).bias.fill_(0.5)
            getattr(model, "post_fc%d" % (model.n_post_hidden - 1)).bias.fill_(0.5)
    return model


def synth(kind, n, s, seed):
    """Synthetic inputs of SURVEY 8(d): img ~ N(0,1) (normalised frames); positions U(-0.5,0.5)^3, unit quaternions
    with w >= 0 (util/data_utils.py:207-211).  Returns (img, x0bar, target) as CPU fp32 tensors."""
    import torch
    g = torch.Generator().manual_seed(seed)
    lead = (s, n) if kind in SEQ_KINDS else (n,)
    img = torch.randn(*lead, 3, 224, 224, generator=g)

    def pose():
        pos = torch.rand(*lead, 3, generator=g) - 0.5
        q = torch.randn(*lead, 4, generator=g)
        q = q / q.norm(dim=-1, keepdim=True)
        q[..., 3] = q[..., 3].abs()
        return torch.cat([pos, q], dim=-1)

    return img, pose(), pose()


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own modules on the host cores
# ------------------------------------------------------------------------------------------------
def _reference_model(kind):
    """(model, criterion, kind-of-baseline): the UNMODIFIED reference nn.Modules when their sources are reachable
    (/root/reference in the build container, oracle/_ref on the GPU box), else None."""
    from oracle import ref_shim
    if not ref_shim.available():
        return None
    import torch
    ref = ref_shim.load()
    cfg = MODEL_CFG[kind]
    kw = dict(latent_dim=cfg["latent"])
    if kind == "no":
        kw.update(object_name="hammer", hidden_dims=list(cfg["hidden"]))
    elif kind in ("tdo", "tdo_v2"):
        kw.update(object_name="robot1_eef", hidden_dim=cfg["hidden"], sequence_length=20)
    elif kind == "td":
        kw.update(hidden_dim=cfg["hidden"], sequence_length=10)
    else:
        kw.update(hidden_pre=list(cfg["hidden"]), hidden_post=list(cfg["hidden"]))
    model = ref_shim.build_reference_model(ref, kind, seed=0, **kw)
    with torch.no_grad():        # same finite-loss start as the GPU arm (see build())
        if kind == "no":
            getattr(model, "fc%d" % (model.n_fc - 1)).module.bias.fill_(0.5)
        elif kind == "n":
            getattr(model, "pre_fc%d" % (model.n_pre_hidden - 1)).bias.fill_(0.5)
            getattr(model, "post_fc%d" % (model.n_post_hidden - 1)).bias.fill_(0.5)
    return model, ref.losses.PoseDistanceLoss(**LOSS)


def cpu_reference_rate(kind, batch, seq, steps, warmup, lr=1e-4):
    """samples/s of the reference's training step (util/learn_utils.py:151-184: zero_grad, forward, loss, backward,
    Adam) on the host cores, bounded sample.  Returns (rate, seconds/step, frames/step, kind of baseline)."""
    import torch
    torch.set_num_threads(os.cpu_count())
    img, x0, tgt = synth(kind, batch, seq, 1)
    frames = img.shape[0] * (img.shape[1] if kind in SEQ_KINDS else 1)
    times = []
    got = _reference_model(kind)
    if got is not None:
        model, crit = got
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=lr)          # scripts/train_model.py:228
        which = "reference"

        def step():
            model.reset_initial_state(batch)
            opt.zero_grad()
            out = model(img, None, x0)
            loss = crit(out, tgt) if kind in ("no", "tdo", "tdo_v2") else crit(out[0], x0) + crit(out[1], tgt)
            loss.backward()
            opt.step()
            return float(loss.item())
    else:
        from oracle import pose_oracle as po
        model = build(kind)          # parameter container only (CPU); the arithmetic below is the oracle's
        extra = None
        if kind == "td":
            extra = {"aux_w": model.aux_nets[0][0].weight, "aux_b": model.aux_nets[0][0].bias}
        orc = po.OracleEstimator(kind, model.state_dict(), extra)
        which = "port"

        def step():
            if kind in ("no", "tdo", "tdo_v2"):
                return orc.train_step(img, x0, tgt, LOSS)
            return orc.train_step(img, x0, tgt, LOSS, which=-1)
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return frames / med, med, frames, which


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # a bounded sample of the GPU arm's workload: the same model, loss and optimizer on `cpu_batch` frames per step
    # (256 frames of ResNet-50 training take ~8 s per step on 16 host cores); the rate is per frame
    batch = args.cpu_batch
    rate, sec, frames, which = cpu_reference_rate(args.model, batch, args.seq, args.steps, args.warmup, args.lr)
    cores = os.cpu_count()
    sample = ("%s training step (forward, pose loss, backward, Adam) on %d-frame batches -- a bounded sample of the "
              "%d-frame GPU step -- %d warm-up + %d timed steps, median; torch CPU fp32, %d threads"
              % (args.model, frames, args.batch * (args.seq if args.model in SEQ_KINDS else 1), args.warmup,
                 args.steps, cores))
    cfg = workload_config(args)
    cfg["cpu_batch_frames"] = frames
    cfg["note"] = "per-frame rate measured on cpu_batch_frames-frame steps, not on per_gpu_batch-frame steps"
    out = {"impl": "reference", "metric": "train_samples_per_s", "value": rate, "unit": "samples/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": cfg,
           "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": which, "sample": sample},
           "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


def workload_config(args, kind=None, batch=None, seq=None):
    kind = kind or args.model
    batch = batch if batch is not None else args.batch
    seq = seq if seq is not None else args.seq
    names = {"no": "naive-object estimator (models/naive.py) training step, hammer target, latent 512, hidden 1024/256/64, combined pose loss, Adam",
             "tdo": "TDO estimator training step, robot1_eef target, latent 512, LSTM 512",
             "td": "TD estimator training step, latent 1024, LSTM 512",
             "tdo_v2": "TDO-v2 estimator training step (image LSTM 512 + proprio LSTM 64), robot1_eef target",
             "n": "naive end-effector estimator training step"}
    cfg = {"workload": names[kind], "per_gpu_batch": batch, "frame": "3x224x224 fp32",
           "cache": "inputs_larger_than_l2", "lr": args.lr}
    if kind in SEQ_KINDS:
        cfg["sequence_length"] = seq
        cfg["frames_per_gpu_step"] = batch * seq
    return cfg


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def measure_tf32_peak(dev, seconds=1.5):
    """Dense TF32 throughput of this box: torch.matmul (cuBLAS, allow_tf32) on 8192^3 fp32 operands, best of 10
    launches (burst) and back to back for `seconds` (sustained) -- the TF32 twin of the driver's bf16 figure in
    MEASURED_PEAKS.json.  The tap-GEMM is timed inside a long step, so the sustained number is its roofline."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        c = torch.empty(n, n, device=dev)
        for _ in range(3):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        flop = 2.0 * n ** 3
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        reps = max(10, int(seconds * 1e3 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        sustained = e0.elapsed_time(e1) / reps
        del a, b, c
        return {"burst": flop / best / 1e9, "sustained": flop / sustained / 1e9, "unit": "TFLOP/s",
                "how": "torch.matmul fp32 operands with allow_tf32 (cuBLAS TF32), 8192^3: best of 10 / %d back to back" % reps}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


GEMM_FAMS = ("pe_conv2d_fwd", "pe_conv2d_dgrad", "pe_conv2d_dgrad_bn", "pe_conv2d_wgrad", "pe_linear_fwd",
             "pe_linear_wgrad", "pe_stem_conv_fwd", "pe_stem_conv_wgrad")
HBM_FAMS = ("pe_bn_train_apply", "pe_bn_bwd_reduce", "pe_bn_bwd_apply", "pe_adam_step")


def ncu_profile():
    """Newest committed ncu launch list of the `no` / 256 step (profiles/launches_*_summary.json): the cross-check
    for the live byte estimates (ncu's per-launch times are cold and serialised; only bytes and shares are used)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "launches_r*_train_step_no_b256_summary.json")))
    if not files:
        return None, {}
    try:
        pj = json.load(open(files[-1]))
        prof = {k["kernel"]: k for k in pj["kernels"]}
        tg = [k for k in pj["kernels"] if k["kernel"].startswith("tapgemm")]
        if tg:
            prof["tapgemm"] = {"launches": sum(k["launches"] for k in tg),
                               "dram_read_GB": sum(k["dram_read_GB"] for k in tg),
                               "dram_write_GB": sum(k["dram_write_GB"] for k in tg)}
        return os.path.relpath(files[-1], ROOT), prof
    except Exception:
        return None, {}


class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.pg = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.pg = dist.group.WORLD
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())


def train_bench(h, kind, batch, seq, steps, warmup, pk, tf32_peak, main):
    """One training configuration: device-resident K-step region, end-to-end region(s) from pinned host buffers, and
    one instrumented step for the per-kernel-family split.  Returns the dict that becomes the JSON line (main) or an
    entry of `configs`."""
    torch = h.torch
    from pe_b200 import native
    from pe_b200.loader import DevicePrefetcher
    from pe_b200.trainer import FusedTrainer
    args, dev, world, rank = h.args, h.dev, h.world, h.rank
    L = native.lib()
    model = build(kind).to(dev).train()
    trainer = FusedTrainer(model, lr=args.lr, process_group=h.pg, **LOSS)
    img, x0, tgt = synth(kind, batch, seq, 1 + rank)
    frames = batch * (seq if kind in SEQ_KINDS else 1)
    targets_h = (x0, tgt) if kind in ("td", "n") else tgt
    img_h, x0_h = img.pin_memory(), x0.pin_memory()
    tg_h = tuple(t.pin_memory() for t in targets_h) if isinstance(targets_h, tuple) else targets_h.pin_memory()
    img_d, x0_d = img_h.to(dev), x0_h.to(dev)
    tg_d = tuple(t.to(dev) for t in tg_h) if isinstance(tg_h, tuple) else tg_h.to(dev)
    # raw renderer frames for the uint8 end-to-end variant: (.., 256, 256, 3) uint8, preprocessed on the device
    lead = img.shape[:-3]
    raw_h = torch.randint(0, 256, (*lead, 256, 256, 3), dtype=torch.uint8,
                          generator=torch.Generator().manual_seed(7 + rank)).pin_memory()
    last_loss = [None]
    loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
    loss_events = [torch.cuda.Event() for _ in range(2)]

    def step_resident():
        last_loss[0] = trainer.step(img_d, x0_d, tg_d)

    def nbytes(*ts):
        tot = 0
        for t in ts:
            for u in (t if isinstance(t, tuple) else (t,)):
                tot += u.numel() * u.element_size()
        return tot

    def run_e2e(frames_h, n):
        """every step's inputs come from pinned HOST memory (one H2D copy of the whole batch per step, issued by
        DevicePrefetcher on a side stream while the previous step computes) and every step's loss is read back"""
        feed = DevicePrefetcher(((frames_h, x0_h, tg_h) for _ in range(n)), dev)
        # every step's loss goes device -> pinned host buffer and is READ by the host inside the region, one step late
        # (the asynchronous-logging pattern: a blocking .item() per step idles the GPU for the ~1 ms the host needs to
        # enqueue the next step's launches); the last one is read before the region closes
        pending = None
        for k, (i, x, t) in enumerate(feed):
            loss = trainer.step(i, x, t)
            buf, ev = loss_host[k & 1], loss_events[k & 1]
            buf.copy_(loss.detach().reshape(1), non_blocking=True)
            ev.record()
            if pending is not None:
                pending[1].synchronize()
                last_loss[0] = float(pending[0][0])
            pending = (buf, ev)
        if pending is not None:
            pending[1].synchronize()
            last_loss[0] = float(pending[0][0])

    for _ in range(max(warmup, 3)):
        step_resident()
    L.check_device()
    sampler = ClockSampler(h.local) if (rank == 0 and main) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    n_calls0 = native.call_count()
    ms = h.timed(step_resident, steps)
    n_calls = native.call_count() - n_calls0
    clocks = sampler.stop() if sampler else None
    # the end-to-end regions depend on the host keeping up with the launches, so one stall of a shared host shows up in
    # full: the headline config runs two K-step regions (the faster is reported, both are listed)
    run_e2e(img_h, 3)
    e2e_regions = [h.timed(lambda: run_e2e(img_h, steps), 1) for _ in range(2 if main else 1)]
    ms_e2e = min(e2e_regions)
    run_e2e(raw_h, 2)
    ms_u8 = h.timed(lambda: run_e2e(raw_h, steps), 1)
    loss_val = last_loss[0]

    # ---- per-kernel-family device time and algorithmic bytes (one instrumented step, outside the timed regions) --
    native.enable_timing(True)
    step_resident()
    torch.cuda.synchronize()
    fam = native.timing_summary()
    fam_bytes = native.bytes_summary()
    native.enable_timing(False)
    L.check_device()
    # The production step folds BatchNorm-backward reductions into 32 of the dgrad launches (pe_conv2d_dgrad_bn): that
    # time is BN work done inside the GEMM kernel.  One more instrumented step with the fusion off gives the family's
    # time as plain GEMMs (reported next to the production figure, never instead of it).
    gemm_plain_ms = None
    if main:
        from pe_b200 import engine
        engine.FUSE_BN_REDUCE[0] = False
        try:
            step_resident()
            native.enable_timing(True)
            step_resident()
            torch.cuda.synchronize()
            fam_plain = native.timing_summary()
            native.enable_timing(False)
            gemm_plain_ms = sum(v for k, v in fam_plain.items() if k in GEMM_FAMS)
        finally:
            engine.FUSE_BN_REDUCE[0] = True
            native.enable_timing(False)

    value = world * frames * steps / (ms / 1e3)
    gemm_ms = sum(v for k, v in fam.items() if k in GEMM_FAMS)
    total_ms = sum(fam.values())
    ach = TRAIN_GFLOP_PER_FRAME[kind] * frames / 1e3 / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
    conv_bytes = sum(fam_bytes.get(k, 0) for k in GEMM_FAMS)
    n_gemm = None
    prof_file, prof = ncu_profile()
    traffic_ncu = None
    if kind == "no" and batch == 256 and "tapgemm" in prof:
        n_gemm = prof["tapgemm"]["launches"]
        traffic_ncu = (prof["tapgemm"]["dram_read_GB"] + prof["tapgemm"]["dram_write_GB"]) * 1e9 / n_gemm
    hbm_kernels = {}
    for f in HBM_FAMS:
        if fam.get(f) and fam_bytes.get(f):
            rate = fam_bytes[f] / 1e9 / (fam[f] / 1e3)
            hbm_kernels[f] = {"ms": round(fam[f], 3), "algorithmic_GB": round(fam_bytes[f] / 1e9, 3),
                              "GB_per_s": round(rate, 1), "frac_of_peak": round(rate / pk["hbm"], 3)}
    ncu_name = {"pe_bn_train_apply": "bn_train_apply_kernel", "pe_bn_bwd_reduce": "channel_reduce_kernel<1>",
                "pe_bn_bwd_apply": "bn_bwd_apply_kernel", "pe_adam_step": "adam_kernel"}
    if kind == "no" and batch == 256:
        for f, kn in ncu_name.items():
            if f in hbm_kernels and kn in prof:
                hbm_kernels[f]["ncu_dram_GB"] = round(prof[kn]["dram_read_GB"] + prof[kn]["dram_write_GB"], 3)
    lw_ms, lw_gb, _ = layerwise_bound_ms(frames, pk["hbm"], tf32_peak["sustained"])
    step_bytes = TRAIN_MB_PER_FRAME * frames / 1e3
    res = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms / steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic", "config": workload_config(args, kind, batch, seq),
        "e2e": {"value": world * frames * steps / (ms_e2e / 1e3), "unit": "samples/s",
                "h2d_bytes_per_step": nbytes(img_h, x0_h, tg_h), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps, "regions_ms": [round(v, 2) for v in e2e_regions],
                "note": "preprocessed fp32 frames (the reference loop's own tensor format) from pinned host memory "
                        "every step (DevicePrefetcher, one batch ahead) + every step's loss copied to the host and read "
                        "there inside the region (asynchronously, one step late; the last before the region closes); "
                        "faster of the listed regions"},
        "e2e_u8": {"value": world * frames * steps / (ms_u8 / 1e3), "unit": "samples/s",
                   "h2d_bytes_per_step": nbytes(raw_h, x0_h, tg_h), "d2h_bytes_per_step": 4,
                   "ms_per_step": ms_u8 / steps,
                   "note": "raw uint8 HWC 256x256 renderer frames from pinned host memory; CenterCrop(224) + /255 + "
                           "Normalize (util/data_utils.py:48-54) run on the device at the head of the step"},
        "gpu_launches": n_calls,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tapgemm_kernel (conv fwd/dgrad/wgrad + dense layers)",
                     "achieved": ach, "peak": tf32_peak["sustained"], "unit": "TFLOP/s",
                     "frac": ach / tf32_peak["sustained"] if tf32_peak["sustained"] else None,
                     "frac_of_burst": ach / tf32_peak["burst"] if tf32_peak["burst"] else None,
                     "traffic": traffic_ncu,
                     "traffic_note": ("ncu dram__bytes read+written, average per tapgemm launch over one step (%s)"
                                      % prof_file) if traffic_ncu else "no ncu capture for this configuration",
                     "traffic_algorithmic": conv_bytes / n_gemm if (n_gemm and conv_bytes) else None,
                     "algorithmic_GB_per_step": round(conv_bytes / 1e9, 2),
                     "algorithmic_flop_per_step": TRAIN_GFLOP_PER_FRAME[kind] * frames * 1e9,
                     "launches_per_step": n_gemm,
                     "peak_source": tf32_peak["how"] + " (sustained); burst %.0f" % tf32_peak["burst"],
                     "measured_ms": round(gemm_ms, 3),
                     "plain_gemm": None if not gemm_plain_ms else {
                         "measured_ms": round(gemm_plain_ms, 3),
                         "achieved": TRAIN_GFLOP_PER_FRAME[kind] * frames / 1e3 / (gemm_plain_ms / 1e3),
                         "frac": TRAIN_GFLOP_PER_FRAME[kind] * frames / 1e3 / (gemm_plain_ms / 1e3) / tf32_peak["sustained"],
                         "note": "same family with the BatchNorm-backward sums NOT folded into the dgrad epilogues "
                                 "(one extra instrumented step): the production figure above carries that BN work, "
                                 "which the step pays back in the pe_bn_bwd_reduce family"},
                     "share_of_step": gemm_ms / total_ms if total_ms else None,
                     "layerwise_bound": {"ms": round(lw_ms, 3), "measured_ms": round(gemm_ms, 3),
                                         "frac": round(lw_ms / gemm_ms, 3) if gemm_ms else None,
                                         "algorithmic_GB_per_step": round(lw_gb, 2),
                                         "note": "sum over conv passes of max(bytes/HBM peak, flops/TF32 peak): the "
                                                 "narrow 1x1 convs are HBM-bound, the 3x3 convs tensor-bound"}},
        "roofline_step": {"bound": "hbm", "achieved": step_bytes / (ms / steps / 1e3), "peak": pk["hbm"],
                          "unit": "GB/s", "frac": step_bytes / (ms / steps / 1e3) / pk["hbm"],
                          "note": "compulsory fp32 activation traffic (223 MB/frame) over the whole step"},
        "hbm_kernels": hbm_kernels,
        "kernel_ms": {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
        "loss": loss_val,
        "loss_finite": bool(loss_val == loss_val and abs(loss_val) != float("inf")),
    }
    del trainer, model, img_d, x0_d, tg_d
    torch.cuda.empty_cache()
    if not main:
        keep = ("value", "unit", "ms_per_step", "steps", "warmup", "config", "e2e", "e2e_u8", "gpu_launches", "loss")
        slim = {k: res[k] for k in keep}
        slim["roofline"] = {k: res["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "measured_ms",
                                                            "share_of_step")}
        slim["roofline_step_frac"] = res["roofline_step"]["frac"]
        slim["kernel_ms_top"] = dict(list(res["kernel_ms"].items())[:8])
        return slim
    return res


def rollout_bench(h, batches=(1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024)):
    """BASELINE config 5: the per-step forward of util/learn_utils.py:366-455 for the TDO estimator (eval, LSTM state
    carried) under CUDA Graph capture.  Timed on the HOST clock around the call a rollout loop makes: pinned host
    frame + measurement in, pose read back to the host, every step."""
    torch = h.torch
    from pe_b200.rollout import PipelinedEstimator, StreamingEstimator
    dev = h.dev
    model = build("tdo").to(dev).eval()
    out = {"workload": "TDO estimator rollout step (eval, state carried, CUDA graph), host frame in -> host pose out",
           "sweep": {}}
    g = torch.Generator().manual_seed(3)
    for N in batches:
        est = StreamingEstimator(model, batch_size=N, use_graph=True)
        est.reset()
        img_h = torch.randn(1, N, 3, 224, 224, generator=g).pin_memory()
        x0_h = torch.randn(1, N, 7, generator=g).pin_memory()
        for _ in range(5):
            est.step(img_h, x0_h)
        torch.cuda.synchronize()
        iters = 300 if N == 1 else (50 if N <= 64 else 12)
        lat = []
        for _ in range(iters):
            t0 = time.perf_counter()
            pose = est.step(img_h, x0_h).cpu()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        p50 = statistics.median(lat)
        row = {"p50_ms": round(p50, 4), "p99_ms": round(lat[min(len(lat) - 1, int(0.99 * len(lat)))], 4),
               "frames_per_s": round(N / (p50 / 1e3), 1)}
        out["sweep"][str(N)] = row
        if N >= 512:
            # same call, batch cut into 256-episode chunks whose H2D copy overlaps the previous chunk's trunk
            del est
            est = PipelinedEstimator(model, batch_size=N, chunk=256)
            est.reset()
            for _ in range(3):
                est.step(img_h, x0_h)
            torch.cuda.synchronize()
            lat = []
            for _ in range(iters):
                t0 = time.perf_counter()
                pose = est.step(img_h, x0_h).cpu()
                lat.append((time.perf_counter() - t0) * 1e3)
            p50 = statistics.median(lat)
            row["pipelined_p50_ms"] = round(p50, 4)
            row["pipelined_frames_per_s"] = round(N / (p50 / 1e3), 1)
        if N == 1:
            out.update({"batch1_p50_ms": row["p50_ms"], "batch1_p99_ms": row["p99_ms"], "iters": iters,
                        "h2d_bytes_per_step": img_h.numel() * 4 + 28, "d2h_bytes_per_step": 28,
                        "finite": bool(torch.isfinite(pose).all())})
        del est
    # the reference-facing call itself: model(img, depth, x0bar) with host tensors in rollout mode, as the unchanged
    # util/learn_utils.rollout() issues it (routed through the same captured step by the mirror)
    model.rollout = True
    model.reset_initial_state(1)
    img_c = torch.randn(1, 1, 3, 224, 224, generator=g)
    x0_c = torch.randn(1, 1, 7, generator=g)
    for _ in range(5):
        model(img_c, None, x0_c)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        pose = model(img_c, None, x0_c)
        _ = pose.squeeze().detach().numpy()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    out["module_api_batch1_p50_ms"] = round(statistics.median(lat), 4)
    out["module_api_batch1_p99_ms"] = round(lat[int(0.99 * len(lat))], 4)
    model.rollout = False
    # raw uint8 frames at the largest batch (preprocessing inside the graph)
    N = batches[-1]
    est = StreamingEstimator(model, batch_size=N, use_graph=True, raw_hw=256)
    est.reset()
    raw_h = torch.randint(0, 256, (1, N, 256, 256, 3), dtype=torch.uint8, generator=g).pin_memory()
    x0_h = torch.randn(1, N, 7, generator=g).pin_memory()
    for _ in range(3):
        est.step_raw(raw_h, x0_h)
    torch.cuda.synchronize()
    lat = []
    for _ in range(12):
        t0 = time.perf_counter()
        est.step_raw(raw_h, x0_h).cpu()
        lat.append((time.perf_counter() - t0) * 1e3)
    out["raw_u8_batch%d_frames_per_s" % N] = round(N / (statistics.median(lat) / 1e3), 1)
    del est
    est = PipelinedEstimator(model, batch_size=N, chunk=256, raw_hw=256)
    est.reset()
    for _ in range(3):
        est.step_raw(raw_h, x0_h)
    torch.cuda.synchronize()
    lat = []
    for _ in range(12):
        t0 = time.perf_counter()
        est.step_raw(raw_h, x0_h).cpu()
        lat.append((time.perf_counter() - t0) * 1e3)
    out["raw_u8_batch%d_pipelined_frames_per_s" % N] = round(N / (statistics.median(lat) / 1e3), 1)
    out["peak_frames_per_s"] = max(max(v["frames_per_s"], v.get("pipelined_frames_per_s", 0.0))
                                   for v in out["sweep"].values())
    del est, model
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    h = Harness(args)
    torch = h.torch
    from pe_b200 import native
    native.lib()
    pk = peaks()
    tf32_peak = measure_tf32_peak(h.dev)
    out = train_bench(h, args.model, args.batch, args.seq, args.steps, args.warmup, pk, tf32_peak, main=True)
    out["tf32_peak"] = {k: (round(v, 1) if isinstance(v, float) else v) for k, v in tf32_peak.items()}
    extra_steps = min(args.steps, 5)
    if not args.only_main and args.model == "no":
        configs = {}
        # config 4 (north_star's multi-GPU configuration): TDO, S = 20, 32 episodes per GPU, weak scaling
        configs["tdo"] = train_bench(h, "tdo", 32, 20, extra_steps, 3, pk, tf32_peak, main=False)
        if h.world == 1:
            # config 3: TD, S = 10, "batch 128" read as 128 episodes (1280 frames per step, the reference's own
            # sense of batch, util/learn_utils.py:118); 64 episodes if that does not fit next to the caching allocator
            for n_ep in (128, 64):
                try:
                    configs["td"] = train_bench(h, "td", n_ep, 10, extra_steps, 3, pk, tf32_peak, main=False)
                    break
                except torch.OutOfMemoryError:
                    torch.cuda.empty_cache()
        out["configs"] = configs
        if h.world == 1:
            out["rollout"] = rollout_bench(h)
    if h.rank == 0:
        if h.world == 1 and not args.no_cpu_baseline:
            try:
                # the reference's own modules in a CPU-ONLY child process (this bench's reference arm): with a GPU
                # visible the reference's nn.DataParallel wrappers would scatter to it (models/naive.py:224,253,274)
                n_cpu = 24
                child = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--model",
                                        args.model, "--steps", str(n_cpu), "--warmup", "2", "--cpu-batch",
                                        str(args.cpu_batch), "--lr", str(args.lr)],
                                       capture_output=True, text=True, timeout=900)
                line = [l for l in child.stdout.splitlines() if l.startswith("{")][-1]
                out["cpu_baseline"] = json.loads(line)["cpu_baseline"]
            except Exception as e:  # the GPU numbers stand on their own
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out), flush=True)
    if h.world > 1:
        h.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--model", default="no", choices=["no", "tdo", "td", "n", "tdo_v2"])
    ap.add_argument("--batch", type=int, default=None, help="frames (naive) or episodes (sequence models) per GPU")
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--cpu-batch", type=int, default=8, help="frames (episodes) per step of the host-CPU arms")
    # Adam moves every parameter by ~lr per step whatever the gradient: the synthetic job repeats ONE batch for a few
    # hundred steps, and at 1e-4 it drives some sample's ReLU'd quaternion to exactly zero around step 32 -> the
    # reference loss (no epsilon, models/losses.py:68-69) and then the parameters turn NaN (tests/probe_loss_traj.py).
    # 1e-6 keeps the whole run finite; the update is the same arithmetic at any lr.
    ap.add_argument("--lr", type=float, default=1e-6)
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: total frames / episodes over all ranks (per-rank batch = this / world size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-main", action="store_true",
                    help="skip the extra configurations (TD / TDO training, rollout sweep): profiling runs")
    args = ap.parse_args()
    if args.seq is None:
        args.seq = {"tdo": 20, "td": 10, "tdo_v2": 20}.get(args.model, 1)
    args.scaling = "weak"
    if args.global_batch is not None:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.global_batch % world:
            ap.error("--global-batch must be divisible by the number of ranks")
        args.batch = args.global_batch // world
        args.scaling = "strong"
    if args.batch is None:
        args.batch = {"no": 256, "n": 256, "tdo": 32, "td": 64, "tdo_v2": 32}[args.model]
    if args.impl == "reference":
        # host cores only: hide the GPUs before torch is imported, otherwise the reference's nn.DataParallel wrappers
        # (models/naive.py:224,253,274) would move the replicas to cuda:0
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
        if args.model in SEQ_KINDS and args.cpu_batch == 8:
            args.cpu_batch = 1          # episodes: 1 x S frames per step
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
