"""CPU restatement of the signal front/back end (oracle; TEST INFRASTRUCTURE ONLY).

Restates `SpecsDataModule.{stft,istft,spec_fwd,spec_back}` (sgmse-bbed/sgmse/data_module.py:241-297),
`get_window` (:13-19) and `pad_spec` / `pad_spec_16` (sgmse-bbed/sgmse/util/other.py:83-99).
`torch.stft` / `torch.istft` are the third-party arithmetic the reference itself calls.
`dft_stft` / `dft_istft` are an independent float64 numpy statement of the same transform (direct
DFT, explicit framing indices) used to pin framing / overlap-add index conventions bit-exactly.
"""
import numpy as np
import torch

N_FFT = 510
HOP = 128
N_BINS = N_FFT // 2 + 1  # 256
SPEC_FACTOR = 0.15        # data_module.py:190
SPEC_ABS_EXPONENT = 0.5   # data_module.py:191


def hann_window(n_fft: int = N_FFT) -> torch.Tensor:
    return torch.hann_window(n_fft, periodic=True)  # data_module.py:16-17


def stft(sig: torch.Tensor, n_fft: int = N_FFT, hop: int = HOP) -> torch.Tensor:
    """data_module.py:291-293 (center=True -> reflect pad n_fft//2, onesided, unnormalised)."""
    return torch.stft(sig, n_fft=n_fft, hop_length=hop, window=hann_window(n_fft).to(sig.device),
                      center=True, return_complex=True)


def istft(spec: torch.Tensor, length=None, n_fft: int = N_FFT, hop: int = HOP) -> torch.Tensor:
    """data_module.py:295-297."""
    return torch.istft(spec, n_fft=n_fft, hop_length=hop, window=hann_window(n_fft).to(spec.device),
                       center=True, length=length)


def spec_fwd(spec, e: float = SPEC_ABS_EXPONENT, factor: float = SPEC_FACTOR, transform_type="exponent"):
    """data_module.py:241-254."""
    if transform_type == "exponent":
        if e != 1:
            spec = spec.abs() ** e * torch.exp(1j * spec.angle())
        return spec * factor
    if transform_type == "log":
        return torch.log(1 + spec.abs()) * torch.exp(1j * spec.angle()) * factor
    return spec


def spec_back(spec, e: float = SPEC_ABS_EXPONENT, factor: float = SPEC_FACTOR, transform_type="exponent"):
    """data_module.py:256-267 (divide by the factor BEFORE the power)."""
    if transform_type == "exponent":
        spec = spec / factor
        if e != 1:
            spec = spec.abs() ** (1 / e) * torch.exp(1j * spec.angle())
        return spec
    if transform_type == "log":
        spec = spec / factor
        return (torch.exp(spec.abs()) - 1) * torch.exp(1j * spec.angle())
    return spec


def pad_spec(Y: torch.Tensor, multiple: int = 64) -> torch.Tensor:
    """util/other.py:83-99: zero-pad the last (time) axis on the right to a multiple of 64 (or 16)."""
    T = Y.size(3)
    num_pad = (multiple - T % multiple) % multiple
    return torch.nn.functional.pad(Y, (0, num_pad, 0, 0))


def n_frames(length: int, hop: int = HOP) -> int:
    return 1 + length // hop


def padded_frames(length: int, multiple: int = 64, hop: int = HOP) -> int:
    nf = n_frames(length, hop)
    return multiple * ((nf + multiple - 1) // multiple)


# ---------------------------------------------------------------------------------------------
# Independent float64 statement with explicit indices (pins framing / OLA conventions)
# ---------------------------------------------------------------------------------------------
def reflect_index(i: int, length: int) -> int:
    """Index into the original signal for padded position i-(n_fft//2) (torch 'reflect' padding)."""
    if i < 0:
        return -i
    if i >= length:
        return 2 * (length - 1) - i
    return i


def frame_sample_index(frame: int, n: int, length: int, n_fft: int = N_FFT, hop: int = HOP) -> int:
    """Source sample read by tap n of STFT frame `frame` (center=True, reflect)."""
    return reflect_index(frame * hop + n - n_fft // 2, length)


def dft_stft(sig: np.ndarray, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """[L] float -> [n_fft//2+1, n_frames] complex128 by direct DFT."""
    L = sig.shape[-1]
    nf = n_frames(L, hop)
    n = np.arange(n_fft)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)
    idx = np.array([[frame_sample_index(f, k, L, n_fft, hop) for k in range(n_fft)] for f in range(nf)])
    frames = sig.astype(np.float64)[idx] * win[None, :]
    kk = np.arange(n_fft // 2 + 1)
    tw = np.exp(-2j * np.pi * np.outer(kk, n) / n_fft)
    return tw @ frames.T


def dft_istft(spec: np.ndarray, length: int, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """[n_fft//2+1, T] complex -> [length] float64: irDFT * window, overlap-add, / sum(w^2), trim."""
    K, T = spec.shape
    n = np.arange(n_fft)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)
    kk = np.arange(K)
    wgt = np.full(K, 2.0)
    wgt[0] = 1.0
    wgt[-1] = 1.0  # n_fft even -> last bin is Nyquist; imaginary parts of bins 0 and Nyquist are ignored
    ang = 2 * np.pi * np.outer(n, kk) / n_fft
    re, im = spec.real.astype(np.float64), spec.imag.astype(np.float64)
    im = im.copy()
    im[0] = 0.0
    im[-1] = 0.0
    frames = (np.cos(ang) * wgt[None]) @ re - (np.sin(ang) * wgt[None]) @ im  # [n_fft, T]
    frames = frames / n_fft * win[:, None]
    total = n_fft + hop * (T - 1)
    y = np.zeros(total)
    env = np.zeros(total)
    for f in range(T):
        y[f * hop:f * hop + n_fft] += frames[:, f]
        env[f * hop:f * hop + n_fft] += win ** 2
    start = n_fft // 2
    out = np.zeros(length)
    avail = min(length, total - start)
    out[:avail] = y[start:start + avail] / env[start:start + avail]
    return out
