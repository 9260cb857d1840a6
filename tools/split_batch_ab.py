#!/usr/bin/env python
"""Does splitting the 16 x 4 s step into concurrent sub-batches (own graphs, own arenas, own streams) shorten the step?
The tails of one sub-batch (small-map convolutions, GroupNorm / FIR passes, SNR estimator, iSTFT) can overlap the other's
large convolutions, at the price of less parallelism per launch.  Prints ms per 16 utterances for 1 x 16, 2 x 8, 4 x 4,
interleaved round-robin (steady state).  Usage: python tools/split_batch_ab.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from snr_aligned_diffse_b200.pipeline import GraphedEnhancer  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L = int(bench.SECONDS * bench.SR)
y = bench.synth_waves(bench.BATCH, L, seed=1000).to(dev)
variants = {}
first = True
for parts in (1, 2, 4):
    b = bench.BATCH // parts
    pipes = []
    for k in range(parts):
        model, _ = bench.build_models(dev, with_estimator=first)
        first = False
        p = GraphedEnhancer(model, b, L, dev, oracle=False)
        p.y_dev.copy_(y[k * b:(k + 1) * b])
        p.capture(warmup=2)
        pipes.append(p)
    variants[parts] = pipes
torch.cuda.synchronize()
main = torch.cuda.Stream(device=dev)


def step(pipes):
    # fork from `main`, replay every sub-batch on its own stream, join
    ev = torch.cuda.Event()
    ev.record(main)
    for p in pipes:
        p.stream.wait_event(ev)
        with torch.cuda.stream(p.stream):
            p.graph.replay()
        main.wait_stream(p.stream)


res = {k: [] for k in variants}
for k, pipes in variants.items():
    for _ in range(10):
        step(pipes)
torch.cuda.synchronize()
for rnd in range(6):
    for k, pipes in variants.items():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for _ in range(20):
            step(pipes)
        e1.record(main)
        torch.cuda.synchronize()
        res[k].append(e0.elapsed_time(e1) / 20)
print(json.dumps({f"{k} x {bench.BATCH // k}": dict(rounds=[round(v, 3) for v in r], median=round(sorted(r)[len(r) // 2], 3))
                  for k, r in res.items()}))
