#!/usr/bin/env python
"""Benchmark of the pose-estimator hot path (driver contract: one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model no|tdo|td|n]

A "step" is one training step (forward, pose loss, backward, Adam) of the naive-object estimator on a
synthetic batch of 256 RGB frames per GPU -- BASELINE.json configs[1] ("naive model training on
1 B200, synthetic batch 256, hammer target").  `value` is whole-job samples/s with inputs resident in
HBM; `e2e` is the same step driven from pinned HOST buffers (H2D of the frames / proprio / targets and
a D2H read of the loss inside the timed region, every step; the region is run twice and the faster is
reported, both are listed).  Weak scaling: each rank owns its own 256 frames.

`--impl reference` times the reference algorithm's CPU path (the oracle restatement of the reference
modules; the reference tree itself cannot travel to the GPU box) on all host cores.  `oracle/` is imported only by
that leg and by the `cpu_baseline` leg; the GPU arm builds its models and synthetic data on its own.
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
def cpu_reference_rate(kind, batch, seq, steps, warmup):
    """samples/s of the reference algorithm on the host cores (oracle port), bounded sample."""
    import torch
    from oracle import pose_oracle as po
    torch.set_num_threads(os.cpu_count())
    model = build(kind)          # parameter container only (CPU); the arithmetic below is the oracle's
    extra = None
    if kind == "td":
        extra = {"aux_w": model.aux_nets[0][0].weight, "aux_b": model.aux_nets[0][0].bias}
    orc = po.OracleEstimator(kind, model.state_dict(), extra)
    img, x0, tgt = synth(kind, batch, seq, 1)
    frames = img.shape[0] * (img.shape[1] if kind in SEQ_KINDS else 1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        if kind in ("no", "tdo", "tdo_v2"):
            orc.train_step(img, x0, tgt, LOSS)
        else:
            orc.train_step(img, x0, tgt, LOSS, which=-1)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    med = statistics.median(times)
    return frames / med, med, frames


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    batch = args.cpu_batch
    rate, sec, frames = cpu_reference_rate(args.model, batch, args.seq, args.steps, min(args.warmup, 2))
    cores = os.cpu_count()
    sample = "%d-frame batches of the %s training step, %d timed steps (median)" % (frames, args.model, args.steps)
    out = {"impl": "reference", "metric": "train_samples_per_s", "value": rate, "unit": "samples/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args),
           "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)
    return 0


def workload_config(args):
    names = {"no": "naive-object estimator (models/naive.py) training step, hammer target, latent 512, hidden 1024/256/64, combined pose loss, Adam",
             "tdo": "TDO estimator training step, robot1_eef target, latent 512, LSTM 512",
             "td": "TD estimator training step, latent 1024, LSTM 512",
             "tdo_v2": "TDO-v2 estimator training step (image LSTM 512 + proprio LSTM 64), robot1_eef target",
             "n": "naive end-effector estimator training step"}
    cfg = {"workload": names[args.model], "per_gpu_batch": args.batch, "frame": "3x224x224 fp32",
           "cache": "inputs_larger_than_l2"}
    if args.model in SEQ_KINDS:
        cfg["sequence_length"] = args.seq
    return cfg


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from pe_b200 import native
    from pe_b200.trainer import FusedTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    L = native.lib()
    kind = args.model
    model = build(kind).to(dev).train()
    trainer = FusedTrainer(model, lr=args.lr, process_group=pg, **LOSS)
    seq = args.seq
    img, x0, tgt = synth(kind, args.batch, seq, 1 + rank)
    frames = args.batch * (seq if kind in SEQ_KINDS else 1)
    targets_h = (x0, tgt) if kind in ("td", "n") else tgt
    img_h, x0_h = img.pin_memory(), x0.pin_memory()
    tg_h = tuple(t.pin_memory() for t in targets_h) if isinstance(targets_h, tuple) else targets_h.pin_memory()
    img_d, x0_d = img_h.to(dev), x0_h.to(dev)
    tg_d = tuple(t.to(dev) for t in tg_h) if isinstance(tg_h, tuple) else tg_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    last_loss = [None]

    def step_resident():
        last_loss[0] = trainer.step(img_d, x0_d, tg_d)

    h2d = img_h.numel() * 4 + x0_h.numel() * 4 + sum(t.numel() * 4 for t in (tg_h if isinstance(tg_h, tuple) else (tg_h,)))

    # end to end: every step's inputs come from pinned HOST memory (one H2D copy of the whole batch per step, issued
    # by pe_b200.loader.DevicePrefetcher on a side stream while the previous step computes) and every step's loss is
    # read back to the host (a 4-byte D2H copy + sync per step)
    from pe_b200.loader import DevicePrefetcher

    def run_e2e(steps):
        t0 = time.perf_counter()
        marks = []
        feed = DevicePrefetcher(((img_h, x0_h, tg_h) for _ in range(steps)), dev)
        for i, x, t in feed:
            last_loss[0] = float(trainer.step(i, x, t).item())
            marks.append(round((time.perf_counter() - t0) * 1e3, 1))
        log("e2e host marks (ms since the region began):", marks)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    L.check_device()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    n_calls0 = native.call_count()
    ms = timed(step_resident, args.steps)
    n_calls = native.call_count() - n_calls0
    clocks = sampler.stop() if sampler else None
    run_e2e(3)
    # every step of this region synchronises with the host (.item()), so one stall of a shared host shows up in
    # full: two K-step regions, the faster one is reported, both are listed
    e2e_regions = [timed(lambda: run_e2e(args.steps), 1) for _ in range(2)]
    ms_e2e = min(e2e_regions)
    loss_val = last_loss[0]

    # ---- per-kernel-family device time (one instrumented step, outside the timed regions) ----------
    native.enable_timing(True)
    step_resident()
    torch.cuda.synchronize()
    fam = native.timing_summary()
    native.enable_timing(False)
    L.check_device()

    value = world * frames * args.steps / (ms / 1e3)
    e2e = world * frames * args.steps / (ms_e2e / 1e3)
    pk = peaks()
    gemm_ms = sum(v for k, v in fam.items() if k in ("pe_conv2d_fwd", "pe_conv2d_dgrad", "pe_conv2d_wgrad",
                                                      "pe_linear_fwd", "pe_linear_wgrad"))
    total_ms = sum(fam.values())
    tf32_peak = pk["bf16"] / 2.0
    ach_tflops = TRAIN_GFLOP_PER_FRAME[kind] * frames / 1e3 / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
    # DRAM bytes per kernel family from the committed ncu launch list of this same step (profiles/): live CUDA-event
    # times divided into ncu-measured bytes give the achieved HBM rate of the streaming kernels
    prof = {}
    try:
        pj = json.load(open(os.path.join(ROOT, "profiles", "launches_r01d_train_step_no_b256_summary.json")))
        prof = {k["kernel"]: k for k in pj["kernels"]}
        # tapgemm_kernel<2> / <4> (two / four epilogue groups) are one kernel family
        tg = [k for k in pj["kernels"] if k["kernel"].startswith("tapgemm_kernel")]
        if tg:
            prof["tapgemm_kernel"] = {"launches": sum(k["launches"] for k in tg),
                                      "dram_read_GB": sum(k["dram_read_GB"] for k in tg),
                                      "dram_write_GB": sum(k["dram_write_GB"] for k in tg)}
    except Exception:
        pass
    fam_of = {"bn_train_apply_kernel": "pe_bn_train_apply", "channel_reduce_kernel<1>": "pe_bn_bwd_reduce",
              "bn_bwd_apply_kernel": "pe_bn_bwd_apply", "adam_kernel": "pe_adam_step"}
    hbm_kernels = {}
    if kind == "no" and args.batch == 256:
        for kname, fname in fam_of.items():
            if kname in prof and fam.get(fname):
                gb = prof[kname]["dram_read_GB"] + prof[kname]["dram_write_GB"]
                rate = gb / (fam[fname] / 1e3)
                hbm_kernels[fname] = {"ms": round(fam[fname], 3), "dram_GB": round(gb, 3), "GB_per_s": round(rate, 1),
                                      "frac_of_peak": round(rate / pk["hbm"], 3)}
    lw_ms, lw_gb, _ = layerwise_bound_ms(frames, pk["hbm"], tf32_peak)
    gemm_launches = prof.get("tapgemm_kernel", {}).get("launches")
    gemm_traffic = None
    if kind == "no" and args.batch == 256 and gemm_launches:
        gemm_traffic = (prof["tapgemm_kernel"]["dram_read_GB"] + prof["tapgemm_kernel"]["dram_write_GB"]) * 1e9 / gemm_launches
    out = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps, "regions_ms": [round(v, 2) for v in e2e_regions],
                "note": "faster of two K-step regions; each step copies its inputs from pinned host memory "
                        "(DevicePrefetcher, one batch ahead) and reads the loss back"},
        "gpu_launches": n_calls,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "tapgemm_kernel (conv fwd/dgrad/wgrad + dense layers)",
                     "achieved": ach_tflops, "peak": tf32_peak, "unit": "TFLOP/s",
                     "frac": ach_tflops / tf32_peak if tf32_peak else None, "traffic": gemm_traffic,
                     "traffic_note": "ncu dram bytes read+written, average per tapgemm launch over one step "
                                     "(profiles/launches_r01d_train_step_no_b256_summary.json)",
                     "algorithmic_flop_per_step": TRAIN_GFLOP_PER_FRAME[kind] * frames * 1e9,
                     "launches_per_step": gemm_launches,
                     "peak_source": "%s bf16 sustained / 2 (TF32 runs at half the bf16 rate)" % pk["which"],
                     "share_of_step": gemm_ms / total_ms if total_ms else None,
                     "layerwise_bound": {"ms": round(lw_ms, 3), "measured_ms": round(gemm_ms, 3),
                                         "frac": round(lw_ms / gemm_ms, 3) if gemm_ms else None,
                                         "algorithmic_GB_per_step": round(lw_gb, 2),
                                         "note": "sum over conv passes of max(bytes/HBM peak, flops/TF32 peak): the "
                                                 "narrow 1x1 convs are HBM-bound, the 3x3 convs tensor-bound"}},
        "roofline_step": {"bound": "hbm", "achieved": TRAIN_MB_PER_FRAME * frames / 1e3 / (ms / args.steps / 1e3),
                          "peak": pk["hbm"], "unit": "GB/s",
                          "frac": TRAIN_MB_PER_FRAME * frames / 1e3 / (ms / args.steps / 1e3) / pk["hbm"],
                          "note": "compulsory fp32 activation traffic (223 MB/frame) over the whole step"},
        "hbm_kernels": hbm_kernels,
        "kernel_ms": {k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])},
        "loss": loss_val,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                n_cpu = 40
                rate, sec, fr = cpu_reference_rate(kind, args.cpu_batch, seq, n_cpu, 2)
                out["cpu_baseline"] = {"value": rate, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "%d-frame batches of the same training step, median of %d steps "
                                                 "(%.2f s/step, ~%.0f s of CPU work)" % (fr, n_cpu, sec, n_cpu * sec)}
            except Exception as e:  # the GPU numbers stand on their own
                out["cpu_baseline"] = {"error": repr(e)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--cpu-batch", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: total frames / episodes over all ranks (per-rank batch = this / world size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
