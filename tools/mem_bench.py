#!/usr/bin/env python
"""Memory-bound operators of the NCSN++ step at the shapes of the 16 x 4 s workload: achieved algorithmic GB/s and the
fraction of the measured HBM copy rate (MEASURED_PEAKS.json).  CUDA events on the launching stream, inputs far larger
than L2 at the big shapes (537 MB per tensor), a 256 MB buffer rewritten between iterations at the small ones.
    python tools/mem_bench.py > gpurun_out/mem_bench.jsonl"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snr_aligned_diffse_b200 import ops  # noqa: E402
from snr_aligned_diffse_b200.sgmse.sdes import axpby  # noqa: E402

try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6448.4


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    g = torch.Generator().manual_seed(0)
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
    rows = []
    for (B, H, W, C) in [(16, 256, 512, 128), (16, 128, 256, 128), (16, 64, 128, 256), (16, 16, 32, 256)]:
        x = (torch.randn(B, H, W, C, generator=g) * 0.5).to(torch.bfloat16).cuda()
        gam, bet = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        el = x.numel() * 2
        small = el < (256 << 20)
        fl = flush if small else None
        cases = [
            ("groupnorm+silu (gn_stats + gn_finalize + gn_apply)", lambda: ops.groupnorm_nhwc(x, gam, bet), 3 * el),   # read, read, write
            ("fir_down2", lambda: ops.fir_nhwc(x, False), el + el // 4),
            ("fir_up2", lambda: ops.fir_nhwc(x, True), el + 4 * el) if H <= 128 else None,
        ]
        for c in cases:
            if c is None:
                continue
            name, fn, by = c
            ms = timeit(fn, flush=fl)
            rows.append(dict(op=name, shape=[B, H, W, C], ms=round(ms, 4), algorithmic_MB=round(by / 1e6, 1),
                             GBps=round(by / ms / 1e6, 1), frac_of_hbm=round(by / ms / 1e6 / HBM, 3)))
    # sampler state update (lincomb): x, y, s, z read + mean and x written, complex64 [16,1,256,512]
    S = [torch.view_as_complex(torch.randn(16, 1, 256, 512, 2, generator=g)).cuda() for _ in range(4)]
    by = 6 * S[0].numel() * 8
    ms = timeit(lambda: axpby(x=S[0], a=1.0, y=S[1], b=0.5, s=S[2], c=0.1, z=S[3], d=0.2, mean=True), flush=flush)
    rows.append(dict(op="lincomb (4 in, 2 out)", shape=list(S[0].shape), ms=round(ms, 4), algorithmic_MB=round(by / 1e6, 1),
                     GBps=round(by / ms / 1e6, 1), frac_of_hbm=round(by / ms / 1e6 / HBM, 3)))
    for r in rows:
        r["hbm_peak_GBps"] = HBM
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
