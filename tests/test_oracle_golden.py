"""The oracle (CPU restatement) against fixtures produced by the unmodified reference
(`oracle/make_golden.py`, run in the build container).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import frontend, ncsnpp as o_ncsnpp, sampler as o_sampler, snrnet as o_snrnet
from oracle.topology import NCSNppConfig, param_specs, snrnet_param_specs
from snr_aligned_diffse_b200.synth import synth_state_dict


def _c(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_param_inventory_matches_reference(golden_dir):
    ref = json.load(open(os.path.join(golden_dir, "ncsnpp_param_specs.json")))
    mine = param_specs(NCSNppConfig())
    assert list(mine.keys()) == list(ref.keys())
    assert all(tuple(ref[k]) == tuple(v) for k, v in mine.items())
    assert sum(int(np.prod(v)) for v in mine.values()) == 65590822
    ema = json.load(open(os.path.join(golden_dir, "ncsnpp_ema_order.json")))
    # EMA shadow list == requires_grad parameters in registration order (Fourier W is frozen)
    assert ema == [k for k in mine if k != "dnn.all_modules.0.W"]


def test_scalars(golden_dir):
    z = np.load(os.path.join(golden_dir, "scalars.npz"))
    assert np.array_equal(z["t_30"], o_sampler.T_30)
    for fs, r, idx, t, nfac in z["rows"]:
        o_idx, o_t, o_nf = o_sampler.v3_scalars(float(np.float32(r)), float(fs), 1.0)
        assert o_idx == int(idx) and o_t == t and abs(o_nf - nfac) < 1e-7


def test_frontend(golden_dir):
    z = np.load(os.path.join(golden_dir, "frontend.npz"))
    wave = _c(z["wave"])
    L = wave.shape[1]
    S = frontend.stft(wave)
    assert S.shape == (2, 256, frontend.n_frames(L))
    assert torch.equal(torch.view_as_real(S), torch.view_as_real(_c(z["stft"])))
    Y = frontend.pad_spec(frontend.spec_fwd(S).unsqueeze(1))
    assert Y.shape[-1] == frontend.padded_frames(L) == 64
    assert torch.equal(torch.view_as_real(Y), torch.view_as_real(_c(z["spec"])))
    back = frontend.spec_back(Y.squeeze(1))
    assert torch.allclose(torch.view_as_real(back), torch.view_as_real(_c(z["spec_back"])), atol=0, rtol=0)
    assert torch.equal(frontend.istft(back, L), _c(z["istft"]))
    # independent direct-DFT statement (explicit reflect / OLA indices) agrees with torch.stft/istft
    d = frontend.dft_stft(z["wave"][0])
    assert np.abs(d - z["stft"][0]).max() < 2e-6
    di = frontend.dft_istft(z["spec_back"][0], L)
    assert np.abs(di - z["istft"][0]).max() < 1e-6
    assert torch.equal(o_snrnet.snr_features(wave[:1]), _c(z["snr_feat"]))


def test_frontend_edge_lengths():
    # ragged / boundary lengths: frame count, padded frame count, exact zero padding, round trip
    for L in (256, 257, 383, 384, 8191, 8192, 8193):  # reflect padding needs L > n_fft//2
        w = torch.randn(1, L, generator=torch.Generator().manual_seed(L))
        S = frontend.stft(w)
        assert S.shape[-1] == 1 + L // 128
        Y = frontend.pad_spec(S.unsqueeze(1))
        assert Y.shape[-1] % 64 == 0 and Y.shape[-1] - S.shape[-1] < 64
        assert Y[..., S.shape[-1]:].abs().sum() == 0
        rt = frontend.istft(frontend.spec_back(frontend.spec_fwd(S)), L)
        assert (rt - w).abs().max() < 1e-4
        d = frontend.dft_stft(w[0].numpy())
        assert np.abs(d - S[0].numpy()).max() < 1e-4


def test_fir(golden_dir):
    z = np.load(os.path.join(golden_dir, "fir.npz"))
    x = _c(z["x"])
    assert torch.allclose(o_ncsnpp.fir_upsample_2d(x), _c(z["up"]), atol=1e-6)
    assert torch.allclose(o_ncsnpp.fir_downsample_2d(x), _c(z["down"]), atol=1e-6)
    # separable closed forms on an odd-sized map
    x = torch.randn(1, 1, 3, 5)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1))
    up = o_ncsnpp.fir_upsample_2d(x)
    r = torch.zeros(1, 1, 6, 5)
    r[:, :, 0::2] = (xp[:, :, 0:3, 1:6] + 3 * xp[:, :, 1:4, 1:6]) / 4
    r[:, :, 1::2] = (3 * xp[:, :, 1:4, 1:6] + xp[:, :, 2:5, 1:6]) / 4
    rp = torch.nn.functional.pad(r, (1, 1, 0, 0))
    e = torch.zeros(1, 1, 6, 10)
    e[..., 0::2] = (rp[..., 0:5] + 3 * rp[..., 1:6]) / 4
    e[..., 1::2] = (3 * rp[..., 1:6] + rp[..., 2:7]) / 4
    assert torch.allclose(up, e, atol=1e-6)


def test_snrnet(golden_dir):
    z = np.load(os.path.join(golden_dir, "snrnet.npz"))
    sd = synth_state_dict(snrnet_param_specs(), seed=1)
    out = o_snrnet.snrnet_forward(sd, _c(z["feat"]))
    assert torch.allclose(out, _c(z["out"]), atol=1e-6)


def test_ncsnpp_forward(golden_dir):
    z = np.load(os.path.join(golden_dir, "ncsnpp_forward.npz"))
    sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
    with torch.no_grad():
        out = o_ncsnpp.ncsnpp_forward(sd, _c(z["x"]), _c(z["t"]))
    ref = _c(z["out"])
    assert out.shape == ref.shape == (2, 1, 256, 64)
    assert (out - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_enhance_v3(golden_dir):
    z = np.load(os.path.join(golden_dir, "enhance_v3.npz"))
    sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
    o = o_sampler.enhance_v3(sd, _c(z["y"]), _c(z["Z"]), float(z["ratio"]), 0.17783, sigma_max=1.0)
    assert o["t"] == float(z["t"]) and abs(o["norm_factor"] - float(z["norm_factor"])) < 1e-7
    ref = _c(z["x_hat"])
    assert o["x_hat"].shape == ref.shape
    assert (o["x_hat"] - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_pc_sampler(golden_dir):
    z = np.load(os.path.join(golden_dir, "pc_ouve.npz"))
    sd = synth_state_dict(param_specs(NCSNppConfig()), seed=0)
    noises = [n for n in _c(z["noises"])]
    out, nfe = o_sampler.pc_sample(sd, _c(z["Y"]), o_sampler.OUVE(1.5, 0.05, 0.5, N=2), noises, N=2, eps=0.03, snr=0.5)
    ref = _c(z["out"])
    assert nfe == int(z["nfe"]) == 4
    assert (out - ref).abs().max() <= 1e-4 * ref.abs().max()
    b = np.load(os.path.join(golden_dir, "bbed.npz"))
    assert torch.allclose(o_sampler.BBED(0.999, 2.6, 0.52).std(_c(b["t"])), _c(b["std"]), atol=1e-7)


def test_upfirdn2d_general_oracle_matches_reference_vectors(golden_dir):
    """General upfirdn2d restatement vs `upfirdn2d_native` outputs of the reference (oracle/make_golden_upfirdn2d.py)."""
    z = np.load(os.path.join(golden_dir, "upfirdn2d.npz"))
    for i, c in enumerate(z["cases"]):
        c = [int(v) for v in c]
        got = o_ncsnpp.upfirdn2d_general(z[f"x{i}"], z[f"k{i}"], *c[6:])
        ref = _c(z[f"y{i}"])
        assert tuple(got.shape) == tuple(ref.shape)                                  # output size: exact
        assert (got - ref).abs().max() <= 2e-6 * max(1.0, float(ref.abs().max()))
    # the two NCSN++ configurations are special cases of the general operator
    x = _c(np.load(os.path.join(golden_dir, "fir.npz"))["x"])
    k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]) / 64.0
    fz = np.load(os.path.join(golden_dir, "fir.npz"))
    assert (o_ncsnpp.upfirdn2d_general(x, k * 4, 2, 2, 1, 1, 2, 1, 2, 1) - _c(fz["up"])).abs().max() < 1e-6
    assert (o_ncsnpp.upfirdn2d_general(x, k, 1, 1, 2, 2, 1, 1, 1, 1) - _c(fz["down"])).abs().max() < 1e-6


# ----------------------------------------------------------------------------------------------- r02 fixtures
import pytest  # noqa: E402


@pytest.fixture(scope="module")
def sd():
    return synth_state_dict(param_specs(NCSNppConfig()), seed=0)


def test_oracle_sampler_variants_match_reference_fixtures(sd, golden_dir):
    """Langevin corrector loop, Euler-Maruyama step and the BBED loop (T_sampling 0.5) against the unmodified reference's
    outputs (oracle/make_golden.py sections 6c-6e)."""
    z = np.load(os.path.join(golden_dir, "pc_langevin.npz"))
    out, nfe = o_sampler.pc_sample(sd, _c(z["Y"]), o_sampler.OUVE(1.5, 0.05, 0.5, N=2), [n for n in _c(z["noises"])], N=2,
                                   eps=0.03, snr=0.5, corrector="langevin")
    assert nfe == int(z["nfe"]) == 4 and (out - _c(z["out"])).abs().max() <= 1e-4 * _c(z["out"]).abs().max()
    z = np.load(os.path.join(golden_dir, "em_step.npz"))
    x_new, x_mean = o_sampler.em_step(sd, _c(z["x"]), _c(z["t"]), _c(z["Y"]), o_sampler.OUVE(1.5, 0.05, 0.5, N=30), _c(z["z"]), 30)
    assert (x_new - _c(z["x_new"])).abs().max() <= 1e-4 * _c(z["x_new"]).abs().max()
    assert (x_mean - _c(z["x_mean"])).abs().max() <= 1e-4 * _c(z["x_mean"]).abs().max()
    assert str(z["in_loop"]) == "TypeError"          # what the reference does with this predictor inside pc_sampler
    z = np.load(os.path.join(golden_dir, "pc_bbed.npz"))
    out, nfe = o_sampler.pc_sample(sd, _c(z["Y"]), o_sampler.BBED(0.5, 2.6, 0.52, N=2), [n for n in _c(z["noises"])], N=2,
                                   eps=0.03, snr=0.5)
    assert nfe == int(z["nfe"]) == 4 and (out - _c(z["out"])).abs().max() <= 1e-4 * _c(z["out"]).abs().max()


def test_oracle_matches_reference_on_its_own_wav_fixtures(sd, golden_dir):
    """dataset/VBD_SNR-5 valid/p232_001.wav with the active_rms.txt oracle ratio, and train/p226_001.wav,
    train2/p286_001.wav through the estimator path (make_golden.py sections 5b, 5c)."""
    from oracle.topology import snrnet_param_specs
    from snr_aligned_diffse_b200.synth import synth_noise, synth_state_dict
    z = np.load(os.path.join(golden_dir, "p232_001.npz"))
    y = torch.from_numpy(z["y"].astype(np.float32) / 32768.0)[None]
    o = o_sampler.enhance_v3(sd, y, _c(z["Z"]), float(z["ratio"]), 0.17783, sigma_max=1.0)
    assert o["t_index"] == int(z["t_index"]) and abs(o["norm_factor"] - float(z["norm_factor"])) < 1e-7
    assert (o["x_hat"] - _c(z["x_hat"])).abs().max() <= 1e-4 * np.abs(z["x_hat"]).max()
    snr_sd = synth_state_dict(snrnet_param_specs(), seed=1)
    w = np.load(os.path.join(golden_dir, "train_wavs.npz"))
    for tag in ("p226_001", "p286_001"):
        y = torch.from_numpy(w[tag + "_y"].astype(np.float32) / 32768.0)[None]
        ratio = float(o_snrnet.estimate_noise_over_clean(snr_sd, y)[0, 0])
        assert abs(ratio / float(w[tag + "_ratio"]) - 1) <= 1e-5
        tpad = 64 * ((1 + y.shape[1] // 128 + 63) // 64)
        o = o_sampler.enhance_v3(sd, y, synth_noise(1, tpad, int(w[tag + "_seed"])), ratio, 0.17783, sigma_max=1.0)
        assert o["t_index"] == int(w[tag + "_t_index"])
        assert (o["x_hat"] - _c(w[tag + "_x_hat"])).abs().max() <= 1e-4 * np.abs(w[tag + "_x_hat"]).max()
