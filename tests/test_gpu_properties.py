"""Size-independent properties at BASELINE.json's FULL sizes (the oracle comparisons at these sizes are in
tests/test_gpu_configs.py; these checks need no reference output):

  * STFT + exponent transform -> inverse transform + iSTFT is the identity on the waveform (16 x 4 s and 1 x 60 s), and
    the padded frames [n_frames, Tpad) of the spectrogram are exactly zero;
  * the front end is linear before the transform: STFT(a x + b y) = a STFT(x) + b STFT(y);
  * batch invariance at the bench shape: utterance 5 of the 16 x 4 s batch, enhanced alone, gives the same bits;
  * run-to-run determinism at the bench shape (no floating-point atomics: GroupNorm sums are 64-bit fixed point);
  * a sweep over the utterances of the bench batch returns, per utterance, the bits of the batched call, and its
    checksum of checksums does not depend on the rank count.
"""
import pytest
import torch

from oracle.topology import NCSNppConfig, param_specs
from snr_aligned_diffse_b200.synth import synth_noise, synth_state_dict, synth_waves

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def v3():
    from snr_aligned_diffse_b200.sgmse.model import ScoreModel
    m = ScoreModel.from_state_dict(synth_state_dict(param_specs(NCSNppConfig()), seed=0), backbone="ncsnpp", sde="ouve",
                                   model_type="sebridge_v3", snr_conditioned="true", fixed_snr=0.17783, theta=1.5,
                                   sigma_min=0.05, sigma_max=1.0, base_dir="")
    return m.eval(no_ema=True)


@pytest.mark.parametrize("batch,seconds", [(16, 4.0), (1, 60.0)], ids=["16x4s", "1x60s"])
def test_stft_istft_round_trip_and_padding_at_full_size(batch, seconds):
    from snr_aligned_diffse_b200 import ops
    L = int(seconds * 16000)
    y = synth_waves(batch, L, seed=7).cuda()
    S = ops.stft(y)                                            # exponent transform on, Tpad = multiple of 64
    nf = 1 + L // 128
    assert S.shape[-1] % 64 == 0 and S.shape[-1] >= nf
    assert not S[..., nf:].abs().any()                         # pad_spec region: exact zeros
    back = ops.istft(S, L)
    assert back.shape == y.shape
    err = (back - y)[:, :L - 300].abs().max() / y.abs().max()  # the last ~255 samples see the zero padded frames
    assert float(err) <= 1e-4
    # linearity of the raw STFT (transform off)
    a, b = 0.37, -1.9
    x2 = synth_waves(batch, L, seed=8).cuda()
    lhs = ops.stft(a * y + b * x2, transform=False)
    rhs = a * ops.stft(y, transform=False) + b * ops.stft(x2, transform=False)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) <= 1e-5


def test_bench_shape_batch_invariance_and_determinism(v3):
    L, B = 64000, 16
    y = synth_waves(B, L, seed=1000)
    Z = synth_noise(B, 512, seed=5)
    ratios = [0.05 + 0.1 * b for b in range(B)]                # several t_30 indices in one batch
    out1, aux = v3.enhance_batch(y, oracle=True, noise_over_clean=ratios, noise=Z, return_aux=True)
    out1 = out1.clone()
    out2 = v3.enhance_batch(y, oracle=True, noise_over_clean=ratios, noise=Z)
    assert torch.equal(out1, out2)                             # run-to-run: same bits
    assert len(set(aux["t_index"].tolist())) >= 4
    alone = v3.enhance_batch(y[5:6], oracle=True, noise_over_clean=ratios[5:6], noise=Z[5:6])
    assert torch.equal(alone[0], out1[5])                      # item of a batch == the item alone
    assert torch.isfinite(out1).all()


def test_sweep_over_the_bench_utterances_matches_the_batched_call(v3):
    from snr_aligned_diffse_b200.sweep import enhance_sweep
    L, B = 64000, 16
    y = synth_waves(B, L, seed=1000)
    waves = [y[b] for b in range(B)]

    def fn(yb, n):          # noise fixed per utterance by its (unique) peak-independent index: derive it from the waveform
        idx = [int(torch.nonzero((y[:, :100].cuda() == yb[r, :100]).all(1))[0]) for r in range(yb.shape[0])]
        Z = torch.stack([synth_noise(1, 512, seed=100 + i)[0] for i in idx])
        return v3.enhance_batch(yb, lengths=n, oracle=True, noise_over_clean=[0.3] * yb.shape[0], noise=Z)

    ref = fn(torch.cat([y, torch.zeros(B, 128 * 512 - 1 - L)], 1).cuda(), torch.full((B,), L, dtype=torch.int32).cuda())
    sums = {}
    for world in (1, 2):
        tot = 0.0
        for rank in range(world):
            r = enhance_sweep(fn, waves, rank=rank, world=world, max_batch=8, device="cuda", keep_audio=True)   # two batches of eight
            for i, a in r["audio"].items():
                assert torch.equal(a, ref[i, :L].cpu()), i
            tot += sum(r["checksum"])
        sums[world] = tot
    assert abs(sums[1] - sums[2]) <= 1e-9 * max(1.0, abs(sums[1]))   # checksum of checksums: independent of the sharding
