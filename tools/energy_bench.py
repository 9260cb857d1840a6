#!/usr/bin/env python
"""Energy per launch of the step's kernels in the power-capped steady state (NVML total-energy counter).

The graphed 16 x 4 s step runs at the board's power cap, where time per step ~ energy per step / cap: removing stall
cycles buys little there, removing Joules does.  Every case is captured in a CUDA graph of REP launches and replayed
back to back for ~0.6 s of warm-up and >= 1.2 s of measurement; reported: J per launch, average W, sustained TFLOP/s
(or GB/s), pJ per FLOP, mean SM clock.  cuBLAS bf16 8192^3 is the reference efficiency.
    python tools/energy_bench.py > gpurun_out/energy.jsonl"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pynvml  # noqa: E402

from snr_aligned_diffse_b200 import ops  # noqa: E402

pynvml.nvmlInit()
H = pynvml.nvmlDeviceGetHandleByIndex(0)


def energy_mj():
    return pynvml.nvmlDeviceGetTotalEnergyConsumption(H)


def measure(name, fn, flops=0.0, nbytes=0.0, rep=10, warm_s=0.6, meas_s=1.2):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rep):
            fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    per = max(e0.elapsed_time(e1) * 1e-3, 1e-5)
    for _ in range(max(1, int(warm_s / per))):
        g.replay()
    torch.cuda.synchronize()
    n = max(2, int(meas_s / per))
    clocks = []
    j0, t0 = energy_mj(), time.perf_counter()
    e0.record()
    for i in range(n):
        g.replay()
        if i % max(1, n // 8) == 0:
            clocks.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
    e1.record()
    torch.cuda.synchronize()
    j1, t1 = energy_mj(), time.perf_counter()
    sec = e0.elapsed_time(e1) * 1e-3
    launches = n * rep
    joule = (j1 - j0) * 1e-3
    row = dict(case=name, ms_per_launch=round(sec / launches * 1e3, 4), J_per_launch=round(joule / launches, 5),
               watts=round(joule / (t1 - t0), 1), sm_mhz=round(sum(clocks) / len(clocks)))
    if flops:
        row["TFLOPs"] = round(flops * launches / sec / 1e12, 1)
        row["pJ_per_FLOP"] = round(joule / (flops * launches) * 1e12, 4)
    if nbytes:
        row["GBps"] = round(nbytes * launches / sec / 1e9, 1)
        row["pJ_per_byte"] = round(joule / (nbytes * launches) * 1e12, 2)
    print(json.dumps(row), flush=True)
    return row


def main():
    gen = torch.Generator().manual_seed(0)
    # idle power
    torch.cuda.synchronize()
    j0, t0 = energy_mj(), time.perf_counter()
    time.sleep(1.0)
    print(json.dumps(dict(case="idle", watts=round((energy_mj() - j0) * 1e-3 / (time.perf_counter() - t0), 1))), flush=True)

    a = torch.randn(8192, 8192, generator=gen).to(torch.bfloat16).cuda()
    b = torch.randn(8192, 8192, generator=gen).to(torch.bfloat16).cuda()
    c = torch.empty(8192, 8192, dtype=torch.bfloat16, device="cuda")
    measure("cuBLAS bf16 8192^3", lambda: torch.matmul(a, b, out=c), flops=2.0 * 8192 ** 3, rep=4)
    del a, b, c

    def conv_cases(B, Hh, W, Ci, Co, Cs, tags):
        x = (torch.randn(B, Hh, W, Ci, generator=gen) * 0.5).to(torch.bfloat16).cuda()
        x1 = (torch.randn(B, Hh, W, Cs, generator=gen) * 0.5).to(torch.bfloat16).cuda() if Cs else None
        res = (torch.randn(B, Hh, W, Co, generator=gen) * 0.5).to(torch.bfloat16).cuda()
        wt = (torch.randn(Co, 9 * Ci + Cs, generator=gen) / math.sqrt(9 * Ci + Cs)).to(torch.bfloat16).cuda()
        bias = torch.randn(Co, generator=gen).cuda()
        gamma = (torch.rand(Ci, generator=gen) + 0.5).cuda()
        beta = (torch.randn(Ci, generator=gen) * 0.2).cuda()
        fl = 2.0 * B * Hh * W * Co * (9 * Ci + Cs)
        px = B * Hh * W
        shp = f"{Ci}->{Co}" + (f"+sc{Cs}" if Cs else "") + f" @{Hh}x{W}"
        if "plain" in tags:
            measure(f"conv {shp} plain", lambda: ops.conv_nhwc(x, wt, 9, x1=x1, bias=bias), flops=fl,
                    nbytes=2.0 * px * (Ci + Cs + Co))
        if "res" in tags:
            measure(f"conv {shp} +residual", lambda: ops.conv_nhwc(x, wt, 9, x1=x1, bias=bias, res=res), flops=fl,
                    nbytes=2.0 * px * (Ci + Cs + 2 * Co))
        if "stats" in tags:
            measure(f"conv {shp} +GroupNorm sums in the epilogue", lambda: ops.conv3x3_nhwc_stats(x, wt, x1=x1, bias=bias)[0],
                    flops=fl, nbytes=2.0 * px * (Ci + Cs + Co))
        if "gn" in tags:
            measure(f"gn_stats+finalize+conv {shp} GroupNorm+SiLU in flight",
                    lambda: ops.gn_silu_conv3x3_nhwc(x, gamma, beta, wt, x1=x1, bias=bias), flops=fl,
                    nbytes=2.0 * px * (2 * Ci + Cs + Co))
        if "gnpass" in tags:
            measure(f"groupnorm+silu three passes C={Ci} @{Hh}x{W}", lambda: ops.groupnorm_nhwc(x, gamma, beta),
                    nbytes=2.0 * px * Ci * 3)
        if "fir" in tags:
            measure(f"fir_down2 C={Ci} @{Hh}x{W}", lambda: ops.fir_nhwc(x, False), nbytes=2.0 * px * Ci * 1.25)

    conv_cases(16, 256, 512, 128, 128, 0, ("plain", "res", "gn", "gnpass", "fir"))
    conv_cases(16, 256, 512, 256, 128, 0, ("plain", "gn"))
    conv_cases(16, 256, 512, 128, 128, 256, ("plain", "gn"))
    conv_cases(16, 128, 256, 256, 256, 0, ("plain", "gn"))
    conv_cases(16, 64, 128, 256, 256, 0, ("plain", "gn"))
    conv_cases(16, 64, 128, 256, 256, 512, ("gn",))
    conv_cases(16, 32, 64, 256, 256, 0, ("gn",))
    conv_cases(16, 16, 32, 256, 256, 0, ("gn",))
    conv_cases(16, 8, 16, 256, 256, 0, ("gn",))

    # the whole step (one enhancer), as bench.py captures it
    import bench
    from snr_aligned_diffse_b200.pipeline import GraphedEnhancer
    dev = torch.device("cuda", 0)
    L = int(bench.SECONDS * bench.SR)
    model, _ = bench.build_models(dev, with_estimator=True)
    p = GraphedEnhancer(model, bench.BATCH, L, dev, oracle=False)
    p.y_dev.copy_(bench.synth_waves(bench.BATCH, L, seed=1000).to(dev))
    p.capture(warmup=2)
    torch.cuda.synchronize()
    with torch.cuda.stream(p.stream):
        for _ in range(40):
            p.graph.replay()
        torch.cuda.synchronize()
        clocks = []
        j0, t0 = energy_mj(), time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(p.stream)
        n = 100
        for i in range(n):
            p.graph.replay()
            if i % 10 == 0:
                clocks.append(pynvml.nvmlDeviceGetClockInfo(H, pynvml.NVML_CLOCK_SM))
        e1.record(p.stream)
        torch.cuda.synchronize()
        j1, t1 = energy_mj(), time.perf_counter()
    print(json.dumps(dict(case="whole 16 x 4 s step (one enhancer, graph)", ms_per_step=round(e0.elapsed_time(e1) / n, 3),
                          J_per_step=round((j1 - j0) * 1e-3 / n, 3), watts=round((j1 - j0) * 1e-3 / (t1 - t0), 1),
                          sm_mhz=round(sum(clocks) / len(clocks)), pJ_per_FLOP=round((j1 - j0) * 1e-3 / n / 16.994e12 * 1e12, 4))),
          flush=True)


if __name__ == "__main__":
    main()
