"""Probe (GPU): does a tensor-bound wgrad overlap the HBM-bound BN backward kernels when they run on two streams?
python tests/probe_overlap.py   -- prints serial vs two-stream times for a few pairings."""
import torch

import kernel_checks as kc
from pe_b200 import native

P = kc.P


def main():
    L = native.lib()
    B = 256
    dev = "cuda"
    side = torch.cuda.Stream()
    main_s = torch.cuda.current_stream()

    def wgrad_job(H, c, k):
        x = torch.randn(B, H, H, c, device=dev)
        dy = torch.randn(B, H, H, c, device=dev)
        dw = torch.empty(k * k, c, c, device=dev)
        return lambda st: L.pe_conv2d_wgrad(P(x), P(dy), P(dw), B, H, H, c, c, k, k, 1, (k - 1) // 2, st)

    def bn_job(Pn, C):
        d = torch.randn(Pn, C, device=dev)
        y = torch.randn(Pn, C, device=dev)
        dyo = torch.empty_like(y)
        mean, invstd, sc, sh, gamma = (torch.rand(C, device=dev) + 0.5 for _ in range(5))
        sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
        dg, db = torch.zeros(C, device=dev), torch.zeros(C, device=dev)

        def run(st):
            L.pe_bn_bwd_reduce(P(d), None, None, P(y), P(mean), P(invstd), P(sc), P(sh), None, P(sums), Pn, C, 1, st)
            L.pe_bn_bwd_apply(P(d), None, None, P(y), P(mean), P(invstd), P(gamma), P(sc), P(sh), None, P(sums), P(dyo),
                              None, 0, P(dg), P(db), 0, Pn, C, 1, 1, st)
        return run

    def timeit(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    for (H, c, k), (Pn, C) in [((14, 256, 3), (B * 14 * 14, 1024)), ((28, 128, 3), (B * 28 * 28, 512)),
                               ((56, 64, 3), (B * 56 * 56, 256)), ((7, 512, 3), (B * 7 * 7, 2048)),
                               ((14, 256, 3), (B * 56 * 56, 256))]:
        wg, bn = wgrad_job(H, c, k), bn_job(Pn, C)
        ms, ss = main_s.cuda_stream, side.cuda_stream

        def serial():
            wg(ms)
            bn(ms)

        def overlapped():
            side.wait_stream(main_s)
            wg(ss)
            bn(ms)
            main_s.wait_stream(side)

        t_w, t_b = timeit(lambda: wg(ms)), timeit(lambda: bn(ms))
        print("wgrad %dx%d %d k%d: %.0f us | bn bwd P=%d C=%d: %.0f us | serial %.0f us | two streams %.0f us"
              % (H, H, c, k, t_w, Pn, C, t_b, timeit(serial), timeit(overlapped)), flush=True)
    print("device error flag:", L.pe_device_error())


if __name__ == "__main__":
    main()
