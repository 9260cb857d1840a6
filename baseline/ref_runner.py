#!/usr/bin/env python
"""Run the UNMODIFIED reference (baseline/_ref/sgmse-bbed, a byte copy of /root/reference/sgmse-bbed) in its own
process and print one JSON line.  Used by `bench.py --impl reference` (the reference arm), by bench.py's
`gpu_eager_baseline` record and by the on-box parity tests; never by the product path.

    python baseline/ref_runner.py --task import
    python baseline/ref_runner.py --task enhance_cpu --steps K --warmup W --utts-per-step U --seconds S
    python baseline/ref_runner.py --task eager_gpu --batch B --seconds S --precision fp32|tf32|bf16 --reps R
    python baseline/ref_runner.py --task parity --batch B --seconds S --out file.npz [--device cuda|cpu]

How the reference is imported: `oracle/ref_shims` (stand-ins for six absent pure-Python
dependencies) and the reference tree go on sys.path; the SNR-estimator checkpoint the reference loads at import time
(sgmse/model.py:25-30) is written from the same seeded synthetic weights the product uses, into a temp cwd; the score
model gets the seeded "de-degenerated" weights of snr_aligned_diffse_b200.synth.  Nothing of the reference is edited.

  enhance_cpu : the reference's own `ScoreModel.enhance(x, y)` exactly as eval.py:94-132 drives it: model.cpu(), one
                utterance per call, SNR estimator + front-end STFT where the reference puts them (CUDA, model.py:716,
                742-743), NCSN++ on the host CPU (model.py:824) with all host threads.
  eager_gpu   : the same reference functions composed for a batch on the GPU (model.to('cuda')): stft ->
                _forward_transform -> pad_spec -> snr_model -> t snap / normfac -> X_T -> model(X_T, t, Y) -> to_audio.
                `fp32` = torch defaults (cuDNN convolutions may use TF32, matmul fp32); `tf32` additionally allows TF32
                matmuls; `fp16` / `bf16` run the network under torch.autocast (bf16 needs two upcast wrappers, see code).  This is the cuDNN / cuBLAS / cuFFT
                kernel set the B200-native path has to beat on the same box.
  parity      : strict fp32 (TF32 off) composed pass with explicit noise Z; writes x_hat / sample / t / norm_factor.
"""
import argparse
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref", "sgmse-bbed")
FIXED_SNR = 0.17783
SR = 16000
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _import_reference():
    """Returns (ref_model_module, ScoreModel, SpecsDataModule, pad_spec, pad_spec_16)."""
    import torch
    if not os.path.isdir(REF):
        raise RuntimeError(f"{REF} missing: run baseline/install_ref.py where /root/reference exists")
    sys.dont_write_bytecode = True
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", os.path.join(HERE, "_ref", "torch_extensions"))
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    for p in (ROOT, REF, os.path.join(ROOT, "oracle", "ref_shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    from oracle.topology import snrnet_param_specs
    from snr_aligned_diffse_b200.synth import synth_state_dict
    tmp = tempfile.mkdtemp(prefix="refrun_")
    os.chdir(tmp)
    os.makedirs("sgmse-bbed/sgmse", exist_ok=True)
    from sgmse.data_module import SpecsDataModule
    from sgmse.snr_estimator import SNRModel
    m = SNRModel(backbone="snrnet", data_module_cls=SpecsDataModule, base_dir="")
    m.load_state_dict(synth_state_dict(snrnet_param_specs(), seed=1))
    torch.save({"state_dict": m.state_dict(),
                "hyper_parameters": {"backbone": "snrnet", "data_module_cls": SpecsDataModule, "base_dir": ""}},
               "sgmse-bbed/sgmse/snr_estimator.ckpt")
    import sgmse.model as ref_model
    from sgmse.util.other import pad_spec, pad_spec_16
    return ref_model, ref_model.ScoreModel, SpecsDataModule, pad_spec, pad_spec_16


def _score_model(ScoreModel, SpecsDataModule):
    from oracle.topology import NCSNppConfig, param_specs
    from snr_aligned_diffse_b200.synth import synth_state_dict
    mod = ScoreModel(backbone="ncsnpp", sde="ouve", model_type="sebridge_v3", snr_conditioned="true",
                     fixed_snr=FIXED_SNR, data_module_cls=SpecsDataModule, base_dir="", theta=1.5, sigma_min=0.05,
                     sigma_max=1.0)
    mod.load_state_dict(synth_state_dict(param_specs(NCSNppConfig()), seed=0))
    mod.train(False, no_ema=True)
    return mod


def _cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def task_enhance_cpu(a):
    import torch
    from snr_aligned_diffse_b200.synth import synth_waves
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref_model, ScoreModel, DM, _, _ = _import_reference()
    model = _score_model(ScoreModel, DM)
    model.cpu()                                        # eval.py:101
    L = int(a.seconds * SR)
    waves = synth_waves(a.batch, L, seed=a.seed)
    torch.manual_seed(0)

    def step(i):
        for u in range(a.utts_per_step):
            y = waves[(i * a.utts_per_step + u) % a.batch][None]
            with torch.no_grad():
                out = model.enhance(y, y)              # eval.py:127-132 (oracle=False: estimator in the loop)
            assert out.shape == (L,)
        return out

    for i in range(a.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(a.steps):
        last = step(a.warmup + i)
    dt = time.perf_counter() - t0
    import numpy as np
    print(json.dumps(dict(task="enhance_cpu", seconds_total=dt, steps=a.steps, utts_per_step=a.utts_per_step,
                          audio_s_per_s=a.steps * a.utts_per_step * a.seconds / dt, cores=cores, cpu_model=_cpu_model(),
                          finite=bool(np.isfinite(last).all()), torch_threads=torch.get_num_threads())), flush=True)


def _compose_batch(ref_model, model, pad_spec, pad_spec_16, y, Z=None, autocast=None):
    """model.py:713-752,810-830 for a batch [B,L] on y.device, using only the reference's own functions.
    Returns (x_hat [B,L], aux) and per-stage CUDA-event times."""
    import numpy as np
    import torch
    dev = y.device
    cuda = dev.type == "cuda"
    marks = []

    def mark(name):
        if cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))
        else:
            marks.append((name, time.perf_counter()))

    B, L = y.shape
    mark("start")
    peak = y.abs().amax(dim=1, keepdim=True)                                            # :715,726
    S = torch.stft(y / peak, n_fft=510, hop_length=128, center='True',
                   window=torch.hann_window(510, periodic=True, device=dev), return_complex=True)   # :716
    feat = pad_spec_16(torch.view_as_real(S).permute(0, 3, 1, 2))                       # :717-719
    est_gt = ref_model.snr_model(feat)                                                  # :720
    est_snr = est_gt / (1 - est_gt)                                                     # :721
    mark("snr_estimator")
    t_raw = model.calculate_snr_direct(1, est_snr, model.fixed_snr).detach().cpu().numpy().reshape(B)   # :732-733
    idx = np.abs(ref_model.t_30[None, :] - t_raw[:, None]).argmin(axis=1)               # :734
    t_ = ref_model.t_30[idx]                                                            # :735
    est_snr_ = torch.FloatTensor(10 ** 0.25 * model.fixed_snr * t_).to(dev)             # :737-738
    normfac_ = model.calculate_normfac_direct(1, est_snr_, model.fixed_snr)             # :739
    norm_factor = peak[:, 0] * normfac_                                                 # :740
    mark("scalars_host_sync")
    yn = y / norm_factor[:, None]                                                       # :745
    Y = pad_spec(torch.unsqueeze(model._forward_transform(model._stft(yn)), 1))         # :749-751 (batch on dim 0)
    vec_t = torch.as_tensor(t_, dtype=torch.float32, device=dev)[:, None, None, None]   # :819-820
    if Z is None:
        Z = torch.randn_like(Y)
    X_T = Y + Z * model.sigma_max * vec_t                                               # :822-823
    mark("stft_transform_noise")
    if autocast is not None:
        with torch.autocast(dev.type, dtype=autocast):
            sample = model(X_T, vec_t, Y)
        sample = sample.to(torch.complex64)
    else:
        sample = model(X_T, vec_t, Y)                                                   # :824
    mark("ncsnpp_forward")
    x_hat = model.to_audio(sample.squeeze(1), L) * norm_factor[:, None]                 # :828-830
    mark("istft")
    if cuda:
        torch.cuda.synchronize()
        ms = {marks[i][0]: marks[i - 1][1].elapsed_time(marks[i][1]) for i in range(1, len(marks))}
    else:
        ms = {marks[i][0]: 1e3 * (marks[i][1] - marks[i - 1][1]) for i in range(1, len(marks))}
    return x_hat, dict(t=t_, idx=idx, norm_factor=norm_factor, sample=sample, Y=Y, ratio=est_snr), ms


def task_eager_gpu(a):
    import torch
    from snr_aligned_diffse_b200.synth import synth_waves
    t_imp = time.perf_counter()
    ref_model, ScoreModel, DM, pad_spec, pad_spec_16 = _import_reference()
    model = _score_model(ScoreModel, DM).to("cuda")
    t_imp = time.perf_counter() - t_imp
    L = int(a.seconds * SR)
    y = synth_waves(a.batch, L, seed=a.seed).cuda()
    results = {}
    for prec in a.precision.split(","):
        torch.backends.cudnn.allow_tf32 = True            # torch default
        torch.backends.cuda.matmul.allow_tf32 = prec == "tf32"
        torch.backends.cudnn.benchmark = True             # let cuDNN pick its best algorithm per shape
        ac = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(prec)
        if prec == "bf16":
            # The unmodified reference cannot run under bf16 autocast: its upfirdn2d CUDA op dispatches float / double /
            # half only (upfirdn2d_kernel.cu:311) and NCSNpp.forward ends in torch.view_as_complex, which rejects bf16
            # (ncsnpp.py:401).  Two wrappers in THIS process upcast at those two call sites (the reference files stay
            # byte-identical); fp16 autocast needs neither and is the reference's own reduced-precision mode.
            import sgmse.backbones.ncsnpp_utils.up_or_down_sampling as uds
            if not hasattr(uds, "_orig_upfirdn2d"):
                uds._orig_upfirdn2d = uds.upfirdn2d

                def _fir_fp32(x, *a_, **k_):
                    return uds._orig_upfirdn2d(x.float(), *a_, **k_).to(x.dtype)
                uds.upfirdn2d = _fir_fp32
                _vac = torch.view_as_complex
                torch.view_as_complex = lambda t_: _vac(t_.float() if t_.dtype == torch.bfloat16 else t_)
        try:
            with torch.no_grad():
                for _ in range(2):
                    _compose_batch(ref_model, model, pad_spec, pad_spec_16, y, autocast=ac)
                torch.cuda.synchronize()
                stage = {}
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    out, _, ms = _compose_batch(ref_model, model, pad_spec, pad_spec_16, y, autocast=ac)
                    for k, v in ms.items():
                        stage[k] = stage.get(k, 0.0) + v / a.reps
                e1.record()
                torch.cuda.synchronize()
                total = e0.elapsed_time(e1) / a.reps
                kernels = None
                try:
                    from torch.profiler import ProfilerActivity, profile
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        _compose_batch(ref_model, model, pad_spec, pad_spec_16, y, autocast=ac)
                        torch.cuda.synchronize()
                    kernels = sum(1 for ev in prof.events() if str(getattr(ev, "device_type", "")).endswith("CUDA"))
                except Exception:
                    pass
            results[prec] = dict(ms_per_step=round(total, 3), audio_s_per_s=round(a.batch * a.seconds / (total * 1e-3), 1),
                                 stage_ms={k: round(v, 3) for k, v in stage.items()}, kernel_launches=kernels,
                                 finite=bool(torch.isfinite(out).all().item()),
                                 peak_mem_GB=round(torch.cuda.max_memory_allocated() / 2 ** 30, 2))
        except Exception as ex:      # e.g. out of memory on the 60 s shape
            results[prec] = dict(unavailable=f"{type(ex).__name__}: {str(ex)[:200]}")
            torch.cuda.empty_cache()
    print(json.dumps(dict(task="eager_gpu", batch=a.batch, seconds=a.seconds, reps=a.reps, import_s=round(t_imp, 1),
                          torch=torch.__version__, cudnn=torch.backends.cudnn.version(), results=results)), flush=True)


def task_parity(a):
    import numpy as np
    import torch
    from snr_aligned_diffse_b200.synth import synth_noise, synth_waves
    torch.set_num_threads(os.cpu_count() or 1)
    ref_model, ScoreModel, DM, pad_spec, pad_spec_16 = _import_reference()
    dev = torch.device(a.device)
    model = _score_model(ScoreModel, DM).to(dev)
    ref_model.snr_model.to(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = False
    L = int(a.seconds * SR)
    tpad = 64 * ((1 + L // 128 + 63) // 64)
    y = synth_waves(a.batch, L, seed=a.seed)
    Z = synth_noise(a.batch, tpad, seed=a.seed + 1)
    t0 = time.perf_counter()
    with torch.no_grad():
        x_hat, aux, _ = _compose_batch(ref_model, model, pad_spec, pad_spec_16, y.to(dev), Z=Z.to(dev))
    dt = time.perf_counter() - t0
    np.savez(a.out, x_hat=x_hat.cpu().numpy(), t=np.asarray(aux["t"], dtype=np.float64), idx=np.asarray(aux["idx"]),
             norm_factor=aux["norm_factor"].cpu().numpy(), ratio=aux["ratio"].reshape(-1).cpu().numpy(),
             sample=aux["sample"][:, 0].cpu().numpy() if a.keep_sample else np.zeros(0))
    print(json.dumps(dict(task="parity", out=a.out, seconds_total=round(dt, 2), device=a.device)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", required=True, choices=["import", "enhance_cpu", "eager_gpu", "parity"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--utts-per-step", type=int, default=1)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--precision", default="fp32,fp16,bf16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--out", default="ref_parity.npz")
    ap.add_argument("--keep-sample", action="store_true")
    a = ap.parse_args()
    if a.task == "import":
        _import_reference()
        print(json.dumps(dict(task="import", ok=True)))
    elif a.task == "enhance_cpu":
        task_enhance_cpu(a)
    elif a.task == "eager_gpu":
        task_eager_gpu(a)
    else:
        task_parity(a)


if __name__ == "__main__":
    main()
