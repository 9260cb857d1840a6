#!/usr/bin/env python
"""BASELINE.json configs 3-5 on 1..8 GPUs (one process per GPU under torchrun, or a single process):

  --workload vbd    824 synthetic VoiceBank-DEMAND-test-shaped utterances (1.5-10 s), sharded by
                    LPT over the ranks, equal-Tpad batches of <= 16, SNR estimator in the loop (config 4 with
                    --fixed-snr 0.17783 / 0.31623 / 0.56234)
  --workload long   60 s utterances (config 5): --count per rank, batch 1
  --workload pc     the generic reverse loop: OUVE score model on the same NCSN++ network, predictor-
                    corrector sampler at the eval.py defaults (N=30, reverse diffusion + 1 annealed-Langevin step =
                    60 network evaluations) on a batch of 16 x 4 s per rank

Prints one JSON line on rank 0: whole-job enhanced audio-seconds per wall-second (slowest rank).  No data-path
collective; torch.distributed only gathers the per-utterance table and the timings.
    python tools/sweep_bench.py --workload vbd
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_bench.py --workload vbd
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from snr_aligned_diffse_b200.shard import synthetic_lengths  # noqa: E402
from snr_aligned_diffse_b200.sweep import enhance_sweep, gather_metrics  # noqa: E402


def synth_wave(length, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length) / bench.SR
    f0 = 100.0 + (seed % 37) * 5.0
    speech = sum(torch.sin(2 * torch.pi * f0 * (k + 1) * t + k) / (k + 1) for k in range(5))
    env = 0.5 + 0.5 * torch.sin(2 * torch.pi * (1.5 + 0.01 * (seed % 50)) * t)
    return (0.1 * speech * env + (0.01 + 0.0005 * (seed % 40)) * torch.randn(length, generator=g)).float()


def run_pc(args, world, rank, dev):
    """60-NFE predictor-corrector loop (bbed-style score head on the OUVE SDE), batch 16 x 4 s per rank."""
    import torch.distributed as dist
    from snr_aligned_diffse_b200 import ops
    from snr_aligned_diffse_b200.sgmse.model import ScoreModel
    from snr_aligned_diffse_b200.synth import synth_state_dict
    n_enh = max(1, args.enhancers) if args.graphs else 1
    B, L = bench.BATCH, int(bench.SECONDS * bench.SR)
    models, specs, peaks, samplers = [], [], [], []
    for e in range(n_enh):    # independent enhancers (own executor + activation arena) whose loops overlap on two streams
        model = ScoreModel(backbone="ncsnpp", sde="ouve", model_type="bbed", snr_conditioned="false", theta=1.5,
                           sigma_min=0.05, sigma_max=0.5, N=30, base_dir="")
        model._error_loading_ema = True
        model.load_state_dict(synth_state_dict({"dnn." + k: v for k, v in model.dnn.param_shapes().items()}, seed=0))
        model.eval(no_ema=True)
        y = bench.synth_waves(B, L, seed=2000 + rank + 100 * e).to(dev)
        peak = ops.absmax(y)
        Y = ops.stft(y, scale=peak, scale_is_divisor=True)[:, None]
        sampler = model.get_pc_sampler("reverse_diffusion", "ald", Y, N=30, corrector_steps=1, snr=0.5, graph=bool(args.graphs))
        sample, nfe = sampler()                       # warm-up: plans, weights, graph capture
        models.append(model); specs.append(Y); peaks.append(peak); samplers.append(sampler)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(1, args.repeat)
    main_s = torch.cuda.current_stream(dev)
    side = [torch.cuda.Stream(device=dev) for _ in samplers]   # one caller stream per enhancer: no cross-enhancer ordering
    e0.record()
    for st in side:
        st.wait_stream(main_s)
    for _ in range(reps):
        outs, x_hats = [], []
        for sm, st, pk in zip(samplers, side, peaks):   # graph mode: each call only enqueues; the loops run concurrently
            with torch.cuda.stream(st):
                o = sm()
                outs.append(o)
                x_hats.append(ops.istft(o[0][:, 0].contiguous(), L, scale=pk))
    for st in side:
        main_s.wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    nfe = outs[0][1]
    ms = torch.tensor([e0.elapsed_time(e1) / (reps * n_enh)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        print(json.dumps(dict(metric="enhanced audio-sec/sec (inverse RTF), PC sampler 60 NFE", workload="pc", unit=bench.UNIT,
                              value=world * B * bench.SECONDS / sec, n_gpus=world, nfe=int(nfe), ms_per_batch=round(sec * 1e3, 2),
                              ms_per_nfe=round(sec * 1e3 / nfe, 3), batch=B, enhancers_per_gpu=n_enh,
                              finite=all(bool(torch.isfinite(x).all()) for x in x_hats),
                              mode=(f"one CUDA graph per reverse step (corrector + predictor), {n_enh} independent batches "
                                    "in flight per GPU") if args.graphs else "host loop, eager launches",
                              scaling="weak")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="vbd", choices=["vbd", "long", "pc"])
    ap.add_argument("--count", type=int, default=0, help="utterances (vbd: total, default 824; long: per rank, default 2)")
    ap.add_argument("--fixed-snr", type=float, default=bench.FIXED_SNR)
    ap.add_argument("--max-batch", type=int, default=16)
    ap.add_argument("--repeat", type=int, default=2, help="passes over the list; the last one is timed")
    ap.add_argument("--enhancers", type=int, default=2, help="pc workload: independent batches in flight per GPU (graph mode)")
    ap.add_argument("--graphs", type=int, default=1, help="1: one CUDA graph per batch shape seen twice (default), 0: eager launches")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    bench.FIXED_SNR = args.fixed_snr
    if args.workload == "pc":
        return run_pc(args, world, rank, dev)
    model, est = bench.build_models(dev)
    if args.workload == "vbd":
        n = args.count or 824
        lengths = synthetic_lengths(n, seed=0)
        max_batch = args.max_batch
    else:
        n = (args.count or 2) * world
        lengths = [60 * bench.SR] * n
        max_batch = 1
    waves = [synth_wave(int(l), seed=i) for i, l in enumerate(lengths)]
    # one activation arena for all buckets: sized for the largest (batch, Tpad) this rank will see
    from snr_aligned_diffse_b200.shard import batch_shards
    need = max(model.dnn.engine.workspace_bytes(len(idx), 256, tpad)
               for tpad, idx in batch_shards([int(l) for l in lengths], world, max_batch)[rank])
    model.dnn._ensure_device_weights()
    model.dnn.engine.reserve(need)
    from snr_aligned_diffse_b200.pipeline import GraphedEnhancerCache
    fn = GraphedEnhancerCache(model, dev, min_uses=1, oracle=False) if args.graphs else (
        lambda y, lens: model.enhance_batch(y, lengths=lens, oracle=False))
    res = None
    for _ in range(max(1, args.repeat)):       # first pass: plans, workspaces, func attributes for every bucket shape
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        res = enhance_sweep(fn, waves, rank=rank, world=world, max_batch=max_batch, device=dev)
    allm = gather_metrics(res, world)
    if rank == 0:
        audio_s = sum(allm["samples"]) / bench.SR
        ok = all(c == c for c in allm["checksum"])            # no NaN anywhere
        print(json.dumps(dict(metric=bench.METRIC, workload=args.workload, value=audio_s / allm["job_seconds"], unit=bench.UNIT,
                              n_gpus=world, utterances=len(allm["ids"]), audio_seconds=round(audio_s, 1),
                              job_seconds=round(allm["job_seconds"], 4), batches_rank0=res["batches"],
                              fixed_snr=args.fixed_snr, max_batch=max_batch, finite=ok, mode="CUDA graph per batch shape" if args.graphs else "eager launches, no CUDA graph",
                              scaling="strong" if args.workload == "vbd" else "weak")), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
