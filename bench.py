#!/usr/bin/env python
"""Benchmark of the SNR-aligned diffusion enhancement hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): sebridge_v3 NCSN++ (65.6 M parameters, seeded synthetic
"de-degenerated" weights), a batch of 16 synthetic 4 s / 16 kHz utterances per GPU, the full
reference-exact sebridge_v3 pass per step: max|y| -> SNR estimator -> t snap / norm factor -> STFT +
exponent transform -> X_T = Y + sigma t Z -> preconditioned NCSN++ (1 NFE) -> inverse transform + iSTFT,
captured in one CUDA graph.  Metric: enhanced audio-seconds per wall-second (inverse RTF), whole job.

  value : inputs resident in HBM, CUDA-graph replay, CUDA-event timed, max over ranks.
  e2e   : the public API call with HOST buffers: pinned-host -> device copy of the waveforms and
          device -> pinned-host copy of the enhanced waveforms inside the timed region, every step.
  roofline : the implicit-GEMM convolution kernel (tensor bound): algorithmic FLOPs of all its launches in
          one step / their summed duration, measured with CUDA events on the launch stream.
  cpu_baseline / --impl reference : the CPU oracle port of the reference path (oracle/), all host threads,
          on a bounded sample (one utterance per step).
Multi-GPU (torchrun, one rank per GPU): utterances are independent, every rank processes its own batch of
16 (weak scaling), no data-path collective; NCCL only for the barrier and the max-over-ranks timing.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, SECONDS, SR = 16, 4.0, 16000
FIXED_SNR = 0.17783
METRIC = "enhanced audio-sec/sec (inverse RTF), sebridge_v3 1 NFE"
UNIT = "audio_s/s"


def synth_waves(batch, length, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length) / SR
    waves = []
    for b in range(batch):
        f0 = 110.0 + 17.0 * b
        speech = sum(torch.sin(2 * torch.pi * f0 * (k + 1) * t + k) / (k + 1) for k in range(6))
        env = 0.5 + 0.5 * torch.sin(2 * torch.pi * (2.0 + 0.1 * b) * t)
        noise = torch.randn(length, generator=g)
        waves.append(0.1 * speech * env + (0.01 + 0.004 * b) * noise)
    return torch.stack(waves).to(torch.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            half = sm[len(sm) // 2:]          # samples under load dominate the upper half of the region
            out = dict(sm_mhz=half[len(half) // 2] if half else sm[-1], sm_max_mhz=max(mx), reasons=sorted(reasons),
                       samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# the workload both arms (--impl b200 / --impl reference) are measured on
WORKLOAD = ("sebridge_v3 NCSN++ 65.6M (synthetic de-degenerated weights), 16 x 4 s @ 16 kHz per GPU (Tpad=512), 1 NFE, "
            "SNR estimator in the loop")


def cpu_reference_step(sd, snr_sd, wave, Z):
    """One utterance through the CPU oracle port of ScoreModel.enhance (sebridge_v3, estimator in the loop)."""
    import torch
    from oracle import sampler as o_sampler, snrnet as o_snrnet
    with torch.no_grad():
        ratio = float(o_snrnet.estimate_noise_over_clean(snr_sd, wave)[0, 0])
        return o_sampler.enhance_v3(sd, wave, Z, ratio, FIXED_SNR, sigma_max=1.0)["x_hat"]


def run_reference(args):
    """--impl reference: the reference's algorithm on the host CPU (oracle port; the Python reference cannot
    be shipped to the GPU box), all host threads, bounded sample = one utterance per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle.topology import NCSNppConfig, param_specs, snrnet_param_specs
    from snr_aligned_diffse_b200.synth import synth_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
    snr_sd = synth_state_dict(snrnet_param_specs(), seed=1)
    total = args.steps + args.warmup
    seconds = SECONDS if total <= 16 else (2.0 if total <= 40 else 1.0)
    L = int(seconds * SR)
    wave = synth_waves(1, L, seed=0)
    tpad = 64 * ((1 + L // 128 + 63) // 64)
    Z = torch.view_as_complex(torch.randn(1, 1, 256, tpad, 2, generator=torch.Generator().manual_seed(1)) * 0.5 ** 0.5)
    for _ in range(args.warmup):
        cpu_reference_step(sd, snr_sd, wave, Z)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(sd, snr_sd, wave, Z)
    dt = time.perf_counter() - t0
    value = args.steps * seconds / dt
    sample = f"1 synthetic {seconds:g} s utterance per step (of the 16 x 4 s batch), {args.steps} steps"
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * dt / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, seconds_per_utterance=SECONDS, nfe=1, cpu_model=_cpu_model(),
                            implementation="CPU port of the reference path (oracle/), fp32, all host threads"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def _cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build_models(device, with_estimator=True):
    """Score model + SNR estimator with seeded synthetic weights, packed on `device`.  with_estimator=False: another
    score model that shares the process-wide estimator already installed (sgmse.model.set_snr_model)."""
    from snr_aligned_diffse_b200.sgmse import model as sg_model
    from snr_aligned_diffse_b200.sgmse.model import ScoreModel
    from snr_aligned_diffse_b200.sgmse.snr_estimator import SNRModel
    from snr_aligned_diffse_b200.synth import synth_state_dict
    model = ScoreModel(backbone="ncsnpp", sde="ouve", model_type="sebridge_v3", snr_conditioned="true",
                       fixed_snr=FIXED_SNR, theta=1.5, sigma_min=0.05, sigma_max=1.0, base_dir="")
    model._error_loading_ema = True
    model.load_state_dict(synth_state_dict({"dnn." + k: v for k, v in model.dnn.param_shapes().items()}, seed=0))
    model.eval(no_ema=True)
    if not with_estimator:
        return model, sg_model.get_snr_model()
    est = SNRModel(base_dir="")
    est._error_loading_ema = True
    est.load_state_dict(synth_state_dict(est.dnn.engine.param_shapes(), seed=1))
    est.eval(no_ema=True)
    sg_model.set_snr_model(est)
    return model, est


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from snr_aligned_diffse_b200 import _lib
    lib = _lib.load()
    _lib.require_device()          # no CUDA extension / no B200 -> fail loudly, there is no fallback
    model, est = build_models(dev)

    L = int(SECONDS * SR)
    host_in = synth_waves(BATCH, L, seed=1000 + rank).pin_memory()
    host_out = torch.empty(BATCH, L, dtype=torch.float32).pin_memory()
    y_dev = host_in.to(dev)
    stream = torch.cuda.Stream(device=dev)

    from snr_aligned_diffse_b200.pipeline import GraphedEnhancer
    pipe = GraphedEnhancer(model, BATCH, L, dev, oracle=False, stream=stream)
    pipe.y_dev.copy_(y_dev)
    y_dev = pipe.y_dev
    # --streams 2: a second, independent enhancer (own network workspace, own streams); consecutive steps alternate
    # between the two so that the small kernels of one batch (grids below 148 CTAs on the 32x64 and smaller maps, the
    # front end, the SNR estimator, the iSTFT) overlap the other batch's work.  Every step is still one full pass over
    # one batch of 16 utterances.
    pipes = [pipe]
    for _ in range(args.streams - 1):
        model2, _ = build_models(dev, with_estimator=False)   # the estimator (weights, read-only) is shared
        pipe2 = GraphedEnhancer(model2, BATCH, L, dev, oracle=False)
        pipe2.y_dev.copy_(y_dev)
        with torch.cuda.stream(pipe2.stream):
            pipe2._step()
            pipe2.capture(warmup=1)
            pipe2.stream.synchronize()
        pipes.append(pipe2)

    with torch.cuda.stream(stream):
        pipe._step()                           # first eager pass: packs weights, builds the plan
        stream.synchronize()
        n0 = lib.snrse_launch_count()
        pipe.capture(warmup=1)
        launches_per_step = int(lib.snrse_launch_count() - n0) // 2     # one eager warm-up + the captured pass
        graph, out_dev = pipe.graph, pipe.out_dev
        for _ in range(max(args.warmup, 3)):
            for p in pipes:
                p.replay()
        torch.cuda.synchronize()

        def barrier():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(body):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for p in pipes[1:]:
                p.stream.wait_event(e0)                    # the other enhancer's stream starts inside the timed region
            for i in range(args.steps):
                body(i)
            for p in pipes[1:]:
                stream.wait_stream(p.stream)               # ... and ends inside it
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        clocks = ClockSampler(local) if rank == 0 else None
        time.sleep(0.3)
        ms_dev = timed(lambda i: pipes[i % len(pipes)].replay())
        time.sleep(0.3)                     # every timed region starts from the same idle state (the board is power-capped:
        ms_single = timed(lambda i: pipes[0].replay()) if len(pipes) > 1 else ms_dev   # the first ~0.1 s after idle run at boost)

        host_outs = [host_out] + [torch.empty_like(host_out).pin_memory() for _ in pipes[1:]]

        def e2e_body(i=0):
            # the package's host-buffer API: pinned host -> device copy of this step's inputs, graph replay, device ->
            # pinned host copy of this step's result; copies run on a second stream and overlap the neighbouring steps
            k = i % len(pipes)
            pipes[k].enhance_host(host_in, host_outs[k])

        for i in range(4):
            e2e_body(i)
        for p in pipes:
            p.flush()

        def timed_e2e():
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for p in pipes[1:]:
                p.stream.wait_event(e0)
            for i in range(args.steps):
                e2e_body(i)
            for p in pipes:
                p.flush()                                  # the last read-backs are inside the timed region
            for p in pipes[1:]:
                stream.wait_stream(p.stream)
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        time.sleep(0.3)
        ms_e2e = timed_e2e()
        clk = clocks.stop() if clocks else None

        # ---- roofline of the dominant kernel (implicit-GEMM conv), measured live with CUDA events
        roof = None
        if rank == 0:
            aux = model.enhance_batch(y_dev, oracle=False, return_aux=True)[1]
            eng = model.dnn.engine
            eng.profile_forward(aux["X_T"], aux["Y"], aux["t"], mode=1)             # warm
            time.sleep(0.3)                                                          # same idle start as the timed regions
            passes = [eng.profile_forward(aux["X_T"], aux["Y"], aux["t"], mode=1) for _ in range(5)]
            passes.sort(key=lambda pr: sum(q["ms"] for q in pr))
            prof = passes[len(passes) // 2]                                          # the pass with the median total time
            gemm = [p for p in prof if p["kind"] == 1]
            tot_ms = sum(p["ms"] for p in prof)
            g_ms = sum(p["ms"] for p in gemm)
            g_fl = sum(p["flops"] for p in gemm)
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            except Exception:
                pass
            peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
            achieved = g_fl / (g_ms * 1e-3) / 1e12
            by_kind = {}
            names = {0: "other", 1: "conv_gemm_tcgen05", 2: "groupnorm_silu", 3: "fir", 4: "attention", 5: "thin_conv", 6: "pack_temb_head"}
            for p in prof:
                d = by_kind.setdefault(names[p["kind"]], dict(ms=0.0, launches=0, bytes=0.0))
                d["ms"] += p["ms"]; d["launches"] += 1; d["bytes"] += p["bytes"]
            hbm = float(peaks.get("hbm_gbs", 6650.0))
            for k, d in by_kind.items():
                d["ms"] = round(d["ms"], 4)
                d["share"] = round(d["ms"] / tot_ms, 4)
                d["algo_GBps"] = round(d.pop("bytes") / (d["ms"] * 1e-3) / 1e9, 1) if d["ms"] > 0 else None
            traffic, traffic_note = None, None
            try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture (profiles/)
                cap = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_halo2_ncu.json")))
                c = cap["gn_silu_in_flight"]
                traffic = int(c["dram_bytes_read"] + c["dram_bytes_write"])
                traffic_note = (f"bytes per launch of {cap['shape']}; algorithmic bytes of that launch "
                                f"{cap['algorithmic_bytes']}; tensor pipe active {c['tensor_subpipe_hmma_cycles_active_pct']} %")
            except Exception:
                pass
            roof = dict(bound="tensor", kernel="conv_halo2_kernel (3x3, W>=8) + conv_gemm_kernel (1x1 / NIN / 4x8 maps)",
                        achieved=round(achieved, 2), peak=peak, unit="TFLOP/s",
                        frac=round(achieved / peak, 4), traffic=traffic, traffic_note=traffic_note,
                        peak_source="MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)",
                        launches_per_step=len(gemm), avg_launch_ms=round(g_ms / max(1, len(gemm)), 4),
                        algorithmic_tflop_per_step=round(g_fl / 1e12, 3), kernel_share_of_network=round(g_ms / tot_ms, 4),
                        hbm_peak_GBps=hbm, network_ms_eager=round(tot_ms, 3), by_kind=by_kind)

    audio_s = world * BATCH * SECONDS * args.steps
    value = audio_s / (ms_dev * 1e-3)
    e2e_value = audio_s / (ms_e2e * 1e-3)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle.topology import NCSNppConfig, param_specs, snrnet_param_specs
            from snr_aligned_diffse_b200.synth import synth_state_dict
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
            snr_sd = synth_state_dict(snrnet_param_specs(), seed=1)
            w1 = host_in[:1].clone()
            Z = torch.view_as_complex(torch.randn(1, 1, 256, 512, 2, generator=torch.Generator().manual_seed(1)) * 0.5 ** 0.5)
            cpu_reference_step(sd, snr_sd, w1[:, :SR], Z[..., :128])          # warm-up on 1 s
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < 12.0 and reps < 12):   # ~12 s of CPU work
                cpu_reference_step(sd, snr_sd, w1, Z)
                reps += 1
            dt = time.perf_counter() - t0
            cpu = dict(value=reps * SECONDS / dt, unit=UNIT, cores=cores, kind="port", cpu_model=_cpu_model(),
                       sample=f"utterance 0 of the batch (4 s) x {reps} runs after a 1 s warm-up, fp32, all host threads")
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                    data="synthetic",
                    config=dict(workload=WORKLOAD,
                                implementation="CUDA graph" + (f", {args.streams} alternating enhancers (streams)" if args.streams > 1 else ""),
                                global_batch=world * BATCH, seconds_per_utterance=SECONDS, nfe=1, parallelism=f"dp{world} (utterance-sharded, no collective)",
                                l2="per-step working set 5.4 GB >> 126 MB L2, no flush needed", accumulate="fp32", storage="bf16 activations"),
                    e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=ms_e2e / args.steps,
                             h2d_bytes_per_step=int(host_in.numel() * 4), d2h_bytes_per_step=int(host_out.numel() * 4)),
                    single_stream=dict(ms_per_step=ms_single / args.steps, value=audio_s / (ms_single * 1e-3), unit=UNIT,
                                       note="same K steps through ONE enhancer (steps strictly back to back); the headline runs "
                                            "two independent enhancers whose steps alternate on two streams"),
                    gpu_launches=launches_per_step * args.steps, launches_per_step=launches_per_step,
                    clocks=clk, roofline=roof, cpu_baseline=cpu, impl="b200")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2, 3, 4],
                    help="independent enhancers per GPU whose steps alternate (2: small kernels of one batch overlap the other batch)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
