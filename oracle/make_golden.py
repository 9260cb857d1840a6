#!/usr/bin/env python
"""Generate `tests/golden/*` from the UNMODIFIED reference and pin the oracle against it.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # writes tests/golden/, asserts oracle == reference

What it does
  1. puts `oracle/ref_shims` (stand-ins for six absent pure-Python deps) and
     `/root/reference/sgmse-bbed` on sys.path, creates the `snr_estimator.ckpt` the reference loads
     at import time (sgmse/model.py:25-30) in a temp cwd, and imports the reference's own
     `ScoreModel`, `SNRNet`, `SpecsDataModule`, samplers and SDEs;
  2. loads the seeded synthetic weights (`snr_aligned_diffse_b200.synth`) into the reference modules;
  3. runs the reference on seeded inputs with every random tensor supplied explicitly
     (`torch.randn_like` is patched to pop from a prepared list);
  4. runs this repo's oracle (`oracle/*.py`) on the same inputs and asserts agreement;
  5. stores the REFERENCE outputs as small fixtures.

Nothing under tests/, bench.py or the product imports this script or /root/reference.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/sgmse-bbed"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "ref_shims"))

from oracle import frontend, ncsnpp as o_ncsnpp, sampler as o_sampler, snrnet as o_snrnet  # noqa: E402
from oracle.topology import NCSNppConfig, param_specs, snrnet_param_specs  # noqa: E402
from snr_aligned_diffse_b200.synth import synth_state_dict  # noqa: E402


def maxabs(a, b):
    return float((a - b).abs().max())


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    os.makedirs("sgmse-bbed/sgmse", exist_ok=True)

    # ---- reference imports (SNR estimator ckpt must exist before `sgmse.model` is imported)
    from sgmse.data_module import SpecsDataModule
    from sgmse.snr_estimator import SNRModel
    snr_sd = synth_state_dict(snrnet_param_specs(), seed=1)
    m = SNRModel(backbone="snrnet", data_module_cls=SpecsDataModule, base_dir="")
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_keys == snrnet_param_specs(), "SNRNet parameter inventory differs from the reference"
    m.load_state_dict(snr_sd)
    torch.save({"state_dict": m.state_dict(),
                "hyper_parameters": {"backbone": "snrnet", "data_module_cls": SpecsDataModule, "base_dir": ""}},
               "sgmse-bbed/sgmse/snr_estimator.ckpt")
    import sgmse.model as ref_model
    from sgmse.model import ScoreModel
    from sgmse.util.other import pad_spec, pad_spec_16
    from sgmse.backbones.ncsnpp_utils import up_or_down_sampling as ref_fir

    cfg = NCSNppConfig()
    specs = param_specs(cfg)
    sd = synth_state_dict(specs, seed=0)

    def make_model(model_type, snr_conditioned, sde, **kw):
        mod = ScoreModel(backbone="ncsnpp", sde=sde, model_type=model_type, snr_conditioned=snr_conditioned,
                         fixed_snr=0.17783, data_module_cls=SpecsDataModule, base_dir="", **kw)
        keys = [(k, tuple(v.shape)) for k, v in mod.state_dict().items()]
        assert keys == list(specs.items()), "NCSN++ parameter inventory / order differs from the reference"
        mod.load_state_dict(sd)
        mod.train(False, no_ema=True)
        return mod

    v3 = make_model("sebridge_v3", "true", "ouve", theta=1.5, sigma_min=0.05, sigma_max=1.0)
    with open(os.path.join(GOLD, "ncsnpp_param_specs.json"), "w") as f:
        json.dump({k: list(v) for k, v in specs.items()}, f)
    # EMA ordering: shadow params = requires_grad params in parameters() order
    ema_names = [n for n, p in v3.named_parameters() if p.requires_grad]
    with open(os.path.join(GOLD, "ncsnpp_ema_order.json"), "w") as f:
        json.dump(ema_names, f)

    report = {}
    g = torch.Generator().manual_seed(1234)

    # ---- (1) scalar tables: t_30, snap, normfac  (model.py:22-23, 627-634, 732-740)
    ratios = np.array([10 ** (-s / 20) for s in range(-5, 40, 5)] + [0.001, 0.05, 0.2, 1.0, 3.0, 50.0], dtype=np.float32)
    rows = []
    for fs in (0.17783, 0.31623, 0.56234):
        v3.fixed_snr = fs
        for r in ratios:
            est = torch.FloatTensor([float(r)])
            t_ = v3.calculate_snr_direct(1, est, fs).numpy()
            idx = int(np.abs(ref_model.t_30 - t_).argmin())
            t = ref_model.t_30[idx]
            est_ = torch.FloatTensor([10 ** 0.25 * fs * t])
            nfac = float(v3.calculate_normfac_direct(1, est_, fs).item())
            o_idx, o_t, o_nf = o_sampler.v3_scalars(float(r), fs, 1.0)
            assert (o_idx, o_t) == (idx, float(t)) and abs(o_nf - nfac) < 1e-7, (fs, r)
            rows.append([fs, float(r), idx, float(t), nfac])
    v3.fixed_snr = 0.17783
    assert np.array_equal(ref_model.t_30, o_sampler.T_30)
    np.savez(os.path.join(GOLD, "scalars.npz"), t_30=ref_model.t_30, rows=np.array(rows, dtype=np.float64))

    # ---- (2) front end: stft / spec_fwd / pad / spec_back / istft  (data_module.py:241-297)
    L = 8000 + 77
    wave = (torch.randn(2, L, generator=g) * 0.1)
    S = v3._stft(wave)
    Yf = v3._forward_transform(S)
    Yp = pad_spec(Yf.unsqueeze(1))
    Yb = v3._backward_transform(Yp.squeeze(1))
    back = v3._istft(Yb, L)
    snr_feat = pad_spec_16(torch.view_as_real(torch.stft(wave[:1] / wave[:1].abs().max().item(), n_fft=510, hop_length=128,
                                                           center='True', window=torch.hann_window(510, periodic=True),
                                                           return_complex=True)).permute(0, 3, 1, 2))
    report["stft"] = maxabs(torch.view_as_real(frontend.stft(wave)), torch.view_as_real(S))
    report["spec_fwd"] = maxabs(torch.view_as_real(frontend.spec_fwd(S)), torch.view_as_real(Yf))
    report["spec_back"] = maxabs(torch.view_as_real(frontend.spec_back(Yp.squeeze(1))), torch.view_as_real(Yb))
    report["istft"] = maxabs(frontend.istft(Yb, L), back)
    report["snr_feat"] = maxabs(o_snrnet.snr_features(wave[:1]), snr_feat)
    d = frontend.dft_stft(wave[0].numpy())
    report["dft_stft_vs_torch"] = float(np.abs(d - S[0].numpy()).max())
    di = frontend.dft_istft(Yb[0].numpy(), L)
    report["dft_istft_vs_torch"] = float(np.abs(di - back[0].numpy()).max())
    assert max(report["stft"], report["spec_fwd"], report["spec_back"], report["istft"], report["snr_feat"]) == 0.0
    assert report["dft_stft_vs_torch"] < 2e-4 and report["dft_istft_vs_torch"] < 2e-5, report
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), wave=wave.numpy(), stft=S.numpy(), spec=Yp.numpy(),
                        spec_back=Yb.numpy(), istft=back.numpy(), snr_feat=snr_feat.numpy())

    # ---- (3) FIR resampling (up_or_down_sampling.py:195-257 via upfirdn2d_native)
    xf = torch.randn(2, 3, 6, 10, generator=g)
    up, dn = ref_fir.upsample_2d(xf, [1, 3, 3, 1], factor=2), ref_fir.downsample_2d(xf, [1, 3, 3, 1], factor=2)
    report["fir_up"] = maxabs(o_ncsnpp.fir_upsample_2d(xf), up)
    report["fir_down"] = maxabs(o_ncsnpp.fir_downsample_2d(xf), dn)
    assert report["fir_up"] < 1e-6 and report["fir_down"] < 1e-6
    np.savez(os.path.join(GOLD, "fir.npz"), x=xf.numpy(), up=up.numpy(), down=dn.numpy())

    # ---- (4) NCSN++ forward (ncsnpp.py:247-404), B=2, 256x64, two different t
    B, T = 2, 64
    xin = torch.view_as_complex(torch.randn(B, 2, 256, T, 2, generator=g) * 0.3)
    tt = torch.tensor([0.3, float(ref_model.t_30[12])], dtype=torch.float32)
    with torch.no_grad():
        ref_out = v3.dnn(xin, tt)
        ora_out = o_ncsnpp.ncsnpp_forward(sd, xin, tt)
    report["ncsnpp_forward"] = maxabs(torch.view_as_real(ora_out), torch.view_as_real(ref_out))
    report["ncsnpp_forward_ref_absmean"] = float(ref_out.abs().mean())
    assert report["ncsnpp_forward"] <= 1e-4 * float(ref_out.abs().max()), report
    np.savez_compressed(os.path.join(GOLD, "ncsnpp_forward.npz"), x=xin.numpy(), t=tt.numpy(), out=ref_out.numpy())

    # ---- (5) sebridge_v3 enhance, composed exactly as model.py:726-830 on CPU with explicit Z
    Lw = 7000
    y_wave = torch.randn(1, Lw, generator=g) * 0.05
    y_wave = y_wave + 0.2 * torch.sin(torch.arange(Lw) * 0.05)[None]
    ratio = 0.35  # noise_rms / clean_rms  (oracle=True, model.py:723)
    Z = torch.view_as_complex(torch.randn(1, 1, 256, 64, 2, generator=g) * (0.5 ** 0.5))
    est_snr = torch.FloatTensor([ratio])
    norm_factor = y_wave.abs().max().item()
    t_ = v3.calculate_snr_direct(1, est_snr, v3.fixed_snr).detach().cpu().numpy()
    t_ = ref_model.t_30[np.abs(ref_model.t_30 - t_).argmin()]
    est_snr_ = torch.FloatTensor([10 ** 0.25 * v3.fixed_snr * t_])
    norm_factor = norm_factor * v3.calculate_normfac_direct(1, est_snr_, v3.fixed_snr)
    y = y_wave / norm_factor
    Y = pad_spec(torch.unsqueeze(v3._forward_transform(v3._stft(y)), 0))
    vec_t = (torch.ones(Y.shape[0]) * t_).unsqueeze(1).unsqueeze(2).unsqueeze(3)
    X_T = Y + Z * v3.sigma_max * t_
    with torch.no_grad():
        sample = v3(X_T, vec_t, Y)
    x_hat = (v3.to_audio(sample.squeeze(), Lw) * norm_factor).squeeze()
    o = o_sampler.enhance_v3(sd, y_wave, Z, ratio, 0.17783, sigma_max=1.0)
    report["enhance_v3_sample"] = maxabs(torch.view_as_real(o["sample"]), torch.view_as_real(sample))
    report["enhance_v3_wave"] = maxabs(o["x_hat"], x_hat)
    report["enhance_v3_wave_peak"] = float(x_hat.abs().max())
    assert abs(o["norm_factor"] - float(norm_factor)) < 1e-7 and o["t"] == float(t_)
    assert report["enhance_v3_wave"] <= 1e-4 * report["enhance_v3_wave_peak"], report
    np.savez_compressed(os.path.join(GOLD, "enhance_v3.npz"), y=y_wave.numpy(), Z=Z.numpy(), ratio=ratio, t=float(t_),
                        norm_factor=float(norm_factor), sample=sample.numpy(), x_hat=x_hat.numpy())

    # ---- (6) PC sampler (reverse_diffusion + ald) with the bbed score head on OUVE, N=2 -> 4 NFE
    bb = make_model("bbed", "false", "ouve", theta=1.5, sigma_min=0.05, sigma_max=0.5)
    Ypc = Y.clone()
    noise_list = [torch.view_as_complex(torch.randn(1, 1, 256, 64, 2, generator=g) * (0.5 ** 0.5)) for _ in range(1 + 2 * 2)]
    feed = iter(noise_list)
    orig_randn_like = torch.randn_like
    torch.randn_like = lambda x, *a, **k: next(feed).to(x.dtype)
    try:
        sampler = bb.get_pc_sampler("reverse_diffusion", "ald", Ypc, N=2, corrector_steps=1, snr=0.5, intermediate=False)
        ref_pc, nfe = sampler()
    finally:
        torch.randn_like = orig_randn_like
    ora_pc, o_nfe = o_sampler.pc_sample(sd, Ypc, o_sampler.OUVE(1.5, 0.05, 0.5, N=2), noise_list, N=2, eps=0.03, snr=0.5)
    report["pc_ouve"] = maxabs(torch.view_as_real(ora_pc), torch.view_as_real(ref_pc))
    report["pc_ouve_absmax"] = float(ref_pc.abs().max())
    assert nfe == o_nfe == 4 and report["pc_ouve"] <= 1e-4 * report["pc_ouve_absmax"], report
    np.savez_compressed(os.path.join(GOLD, "pc_ouve.npz"), Y=Ypc.numpy(), noises=torch.stack(noise_list).numpy(),
                        out=ref_pc.numpy(), nfe=nfe)

    # ---- (6c) Langevin corrector in the loop (correctors.py:38-57) with the reverse-diffusion predictor, OUVE, N=2
    noise_lv = [torch.view_as_complex(torch.randn(1, 1, 256, 64, 2, generator=g) * (0.5 ** 0.5)) for _ in range(1 + 2 * 2)]
    feed = iter(noise_lv)
    torch.randn_like = lambda x, *a, **k: next(feed).to(x.dtype)
    try:
        ref_lv, nfe_lv = bb.get_pc_sampler("reverse_diffusion", "langevin", Ypc, N=2, corrector_steps=1, snr=0.5,
                                           intermediate=False)()
    finally:
        torch.randn_like = orig_randn_like
    ora_lv, _ = o_sampler.pc_sample(sd, Ypc, o_sampler.OUVE(1.5, 0.05, 0.5, N=2), noise_lv, N=2, eps=0.03, snr=0.5,
                                    corrector="langevin")
    report["pc_langevin"] = maxabs(torch.view_as_real(ora_lv), torch.view_as_real(ref_lv))
    report["pc_langevin_absmax"] = float(ref_lv.abs().max())
    assert nfe_lv == 4 and report["pc_langevin"] <= 1e-4 * report["pc_langevin_absmax"], report
    np.savez_compressed(os.path.join(GOLD, "pc_langevin.npz"), Y=Ypc.numpy(), noises=torch.stack(noise_lv).numpy(),
                        out=ref_lv.numpy(), nfe=nfe_lv)

    # ---- (6d) Euler-Maruyama predictor (predictors.py:41-52), called directly as update_fn(x, t, y): through
    # pc_sampler the reference hands it a 4th positional argument (sampling/__init__.py:72) that reaches
    # OUVESDE.sde(x, t, y) (sdes.py:121,192) and raises TypeError -- pinned here as behaviour.
    from sgmse.sampling.predictors import EulerMaruyamaPredictor as RefEM
    sde_em = bb.sde.copy()
    sde_em.N = 30
    em = RefEM(sde_em, bb, probability_flow=False)
    x_em = Ypc + 0.3 * noise_lv[0]
    t_em = torch.tensor([0.7])
    feed = iter([noise_lv[1]])
    torch.randn_like = lambda x, *a, **k: next(feed).to(x.dtype)
    try:
        with torch.no_grad():
            em_x, em_mean = em.update_fn(x_em, t_em, Ypc)
    finally:
        torch.randn_like = orig_randn_like
    o_x, o_mean = o_sampler.em_step(sd, x_em, t_em, Ypc, o_sampler.OUVE(1.5, 0.05, 0.5, N=30), noise_lv[1], 30)
    report["em_step"] = max(maxabs(torch.view_as_real(o_x), torch.view_as_real(em_x)),
                            maxabs(torch.view_as_real(o_mean), torch.view_as_real(em_mean)))
    assert report["em_step"] <= 1e-4 * float(em_x.abs().max()), report
    em_in_loop = "no error"
    try:
        bb.get_pc_sampler("euler_maruyama", "none", Ypc, N=2)()
    except TypeError as ex:
        em_in_loop = "TypeError"
    report["em_in_pc_sampler"] = em_in_loop
    np.savez_compressed(os.path.join(GOLD, "em_step.npz"), Y=Ypc.numpy(), x=x_em.numpy(), t=t_em.numpy(),
                        z=noise_lv[1].numpy(), x_new=em_x.numpy(), x_mean=em_mean.numpy(), in_loop=em_in_loop)

    # ---- (6e) BBED full loop (sdes.py:240-307), reverse_diffusion + ald, T_sampling=0.5, N=2 (B=1: the reference's
    # drift broadcast `(y-x)/(Tc-t)` with t [B] is only well-formed for one utterance)
    bbed = make_model("bbed", "false", "bbed", T_sampling=0.5, k=2.6, theta=0.52, sigma_min=0.05, sigma_max=0.5)
    bbed.sde.logk, bbed.sde.Eilog = float(bbed.sde.logk), float(bbed.sde.Eilog)   # numpy-2 promotion workaround
    ref_copy = type(bbed.sde).copy

    def _copy(self):
        c = ref_copy(self)
        c.logk, c.Eilog = float(c.logk), float(c.Eilog)
        return c
    type(bbed.sde).copy = _copy
    noise_bb = [torch.view_as_complex(torch.randn(1, 1, 256, 64, 2, generator=g) * (0.5 ** 0.5)) for _ in range(1 + 2 * 2)]
    feed = iter(noise_bb)
    torch.randn_like = lambda x, *a, **k: next(feed).to(x.dtype)
    try:
        ref_bb, nfe_bb = bbed.get_pc_sampler("reverse_diffusion", "ald", Ypc, N=2, corrector_steps=1, snr=0.5,
                                             intermediate=False)()
    finally:
        torch.randn_like = orig_randn_like
        type(bbed.sde).copy = ref_copy
    ora_bb, _ = o_sampler.pc_sample(sd, Ypc, o_sampler.BBED(0.5, 2.6, 0.52, N=2), noise_bb, N=2, eps=0.03, snr=0.5)
    report["pc_bbed"] = maxabs(torch.view_as_real(ora_bb), torch.view_as_real(ref_bb))
    report["pc_bbed_absmax"] = float(ref_bb.abs().max())
    assert nfe_bb == 4 and ref_bb.dtype == torch.complex64 and report["pc_bbed"] <= 1e-4 * report["pc_bbed_absmax"], report
    np.savez_compressed(os.path.join(GOLD, "pc_bbed.npz"), Y=Ypc.numpy(), noises=torch.stack(noise_bb).numpy(),
                        out=ref_bb.numpy(), nfe=nfe_bb)

    # ---- (5b) the reference's own fixture: dataset/VBD_SNR-5/valid/noisy/p232_001.wav with the oracle ratio of
    # valid/active_rms.txt row 1 (eval.py:76-83,127-129: enhance(oracle=True, clean_rms, noise_rms)); composed as (5)
    import wave as _wave
    wf = _wave.open("/root/reference/dataset/VBD_SNR-5/valid/noisy/p232_001.wav", "rb")
    assert wf.getframerate() == 16000 and wf.getnchannels() == 1 and wf.getsampwidth() == 2
    pcm = np.frombuffer(wf.readframes(wf.getnframes()), dtype="<i2")
    wf.close()
    y_p = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]          # torchaudio.load normalisation
    row = open("/root/reference/dataset/VBD_SNR-5/valid/active_rms.txt").readline().split()
    assert row[0] == "p232_001.wav"
    clean_rms, noise_rms = float(row[1]), float(row[2])
    ratio_p = noise_rms / clean_rms
    Lp = y_p.shape[1]
    tp = 64 * ((1 + Lp // 128 + 63) // 64)
    Zp = torch.view_as_complex(torch.randn(1, 1, 256, tp, 2, generator=g) * (0.5 ** 0.5))
    est_snr = torch.FloatTensor([ratio_p])
    nf = y_p.abs().max().item()
    t_ = v3.calculate_snr_direct(1, est_snr, v3.fixed_snr).detach().cpu().numpy()
    idx_p = int(np.abs(ref_model.t_30 - t_).argmin())
    t_ = ref_model.t_30[idx_p]
    nf = nf * v3.calculate_normfac_direct(1, torch.FloatTensor([10 ** 0.25 * v3.fixed_snr * t_]), v3.fixed_snr)
    Yp_ = pad_spec(torch.unsqueeze(v3._forward_transform(v3._stft(y_p / nf)), 0))
    vt = (torch.ones(1) * t_)[:, None, None, None]
    with torch.no_grad():
        samp_p = v3(Yp_ + Zp * v3.sigma_max * t_, vt, Yp_)
    xh_p = (v3.to_audio(samp_p.squeeze(), Lp) * nf).squeeze()
    o_p = o_sampler.enhance_v3(sd, y_p, Zp, ratio_p, 0.17783, sigma_max=1.0)
    report["p232_wave"] = maxabs(o_p["x_hat"], xh_p)
    report["p232_wave_peak"] = float(xh_p.abs().max())
    assert o_p["t_index"] == idx_p and report["p232_wave"] <= 1e-4 * report["p232_wave_peak"], report
    np.savez_compressed(os.path.join(GOLD, "p232_001.npz"), y=pcm, Z_seed_note="Z drawn after all earlier draws of generator 1234",
                        Z=Zp.numpy().astype(np.complex64), ratio=ratio_p, clean_rms=clean_rms, noise_rms=noise_rms,
                        t_index=idx_p, t=float(t_), norm_factor=float(nf), x_hat=xh_p.numpy())

    # ---- (5c) the reference's two training-folder wavs through the ESTIMATOR path (enhance(oracle=False), model.py:713-721):
    # SNR-branch STFT of y / max|y| -> pad_spec_16 -> the import-time snr_model -> n/s -> t snap -> composed pass as (5).
    # The noise draw is snr_aligned_diffse_b200.synth.synth_noise(1, Tpad, seed) (stored as the seed, not the tensor).
    from snr_aligned_diffse_b200.synth import synth_noise
    extra = {}
    for tag, path, seed in (("p226_001", "/root/reference/dataset/VBD_SNR-5/train/noisy/p226_001.wav", 226),
                            ("p286_001", "/root/reference/dataset/VBD_SNR-5/train2/noisy/p286_001.wav", 286)):
        wf = _wave.open(path, "rb")
        assert wf.getframerate() == 16000 and wf.getnchannels() == 1 and wf.getsampwidth() == 2
        pcm_e = np.frombuffer(wf.readframes(wf.getnframes()), dtype="<i2")
        wf.close()
        y_e = torch.from_numpy(pcm_e.astype(np.float32) / 32768.0)[None]
        Le = y_e.shape[1]
        tpe = 64 * ((1 + Le // 128 + 63) // 64)
        Ze = synth_noise(1, tpe, seed)
        y_chk = y_e / y_e.abs().max().item()
        feat_e = pad_spec_16(torch.view_as_real(torch.stft(y_chk, n_fft=510, hop_length=128, center='True',
                                                           window=torch.hann_window(510, periodic=True),
                                                           return_complex=True)).permute(0, 3, 1, 2))
        with torch.no_grad():
            est_gt = ref_model.snr_model(feat_e)
        est_snr_e = est_gt / (1 - est_gt)
        nf_e = y_e.abs().max().item()
        t_e = v3.calculate_snr_direct(1, est_snr_e, v3.fixed_snr).detach().cpu().numpy()
        idx_e = int(np.abs(ref_model.t_30 - t_e).argmin())
        t_e = ref_model.t_30[idx_e]
        nf_e = nf_e * v3.calculate_normfac_direct(1, torch.FloatTensor([10 ** 0.25 * v3.fixed_snr * t_e]), v3.fixed_snr)
        Ye = pad_spec(torch.unsqueeze(v3._forward_transform(v3._stft(y_e / nf_e)), 0))
        with torch.no_grad():
            samp_e = v3(Ye + Ze * v3.sigma_max * t_e, (torch.ones(1) * t_e)[:, None, None, None], Ye)
        xh_e = (v3.to_audio(samp_e.squeeze(), Le) * nf_e).squeeze()
        ratio_o = float(o_snrnet.estimate_noise_over_clean(snr_sd, y_e)[0, 0])
        o_e = o_sampler.enhance_v3(sd, y_e, Ze, ratio_o, 0.17783, sigma_max=1.0)
        report[tag + "_ratio"] = abs(ratio_o - float(est_snr_e))
        report[tag + "_wave"] = maxabs(o_e["x_hat"], xh_e)
        assert o_e["t_index"] == idx_e and report[tag + "_wave"] <= 1e-4 * float(xh_e.abs().max()), (tag, report)
        extra.update({tag + "_y": pcm_e, tag + "_seed": seed, tag + "_ratio": float(est_snr_e), tag + "_t_index": idx_e,
                      tag + "_t": float(t_e), tag + "_norm_factor": float(nf_e), tag + "_x_hat": xh_e.numpy()})
    np.savez_compressed(os.path.join(GOLD, "train_wavs.npz"), **extra)

    # ---- (6b) BBED scalar functions (sdes.py:275-293) under the reference's pinned-numpy semantics
    from sgmse.sdes import BBED as RefBBED
    rb = RefBBED(0.999, 2.6, 0.52, N=30)
    rb.logk, rb.Eilog = float(rb.logk), float(rb.Eilog)  # numpy-2 promotion workaround
    tsb = torch.tensor([0.03, 0.25, 0.5, 0.9, 0.999])
    ob = o_sampler.BBED(0.999, 2.6, 0.52, N=30)
    report["bbed_std"] = maxabs(ob.std(tsb), rb._std(tsb))
    assert report["bbed_std"] < 1e-7
    np.savez(os.path.join(GOLD, "bbed.npz"), t=tsb.numpy(), std=rb._std(tsb).numpy())

    # ---- (7) SNR estimator (snrnet.py:47-97) + n/s conversion (model.py:720-721)
    feat = torch.cat([snr_feat, snr_feat.flip(3) * 0.7], dim=0)
    with torch.no_grad():
        ref_g = ref_model.snr_model(feat)
        ora_g = o_snrnet.snrnet_forward(snr_sd, feat)
    report["snrnet"] = maxabs(ora_g, ref_g)
    assert report["snrnet"] < 1e-5, report
    np.savez_compressed(os.path.join(GOLD, "snrnet.npz"), feat=feat.numpy(), out=ref_g.numpy())

    with open(os.path.join(GOLD, "oracle_vs_reference.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
