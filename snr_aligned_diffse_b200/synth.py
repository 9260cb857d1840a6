"""Seeded synthetic ("de-degenerated") weights for benchmarks and parity tests.

The reference's default initialisation makes 56 tensors ~1e-10 (`init_scale=0.`,
sgmse-bbed/sgmse/backbones/ncsnpp.py:61, ncsnpp_utils/layers.py:88-91), so a freshly constructed
network outputs a constant and a parity test on it would prove nothing.  There is no network
access for the authors' checkpoints, so every measurement and parity check in this repo uses the
state dict produced here: every matrix/filter drawn with the reference's own DDPM rule at scale 1
(uniform, variance 1/fan_avg, layers.py:54-91), every bias and GroupNorm affine perturbed, the
Fourier frequencies ~ N(0, 16^2) (layerspp.py:37).  Each tensor has its own CRC-derived seed, so
the result does not depend on iteration order and is identical on every host with this torch.
"""
import math
import zlib
from typing import Dict, Tuple

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 63))
    return g


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    g = _gen(seed, name)
    shape = tuple(shape)
    if len(shape) >= 2:
        recept = 1
        for s in shape[2:]:
            recept *= s
        fan_avg = 0.5 * (shape[0] + shape[1]) * recept
        bound = math.sqrt(3.0 / fan_avg)
        return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound
    if name.endswith(".W"):                      # GaussianFourierProjection frequencies
        return torch.randn(shape, generator=g) * 16.0
    if name.endswith(".weight"):                 # 1-D weight == GroupNorm gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.05 * torch.randn(shape, generator=g)  # biases, GroupNorm beta, NIN b, LSTM biases


def synth_state_dict(specs: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: synth_tensor(k, v, seed) for k, v in specs.items()}


def synth_waves(batch: int, length: int, seed: int, sr: int = 16000) -> torch.Tensor:
    """Synthetic VoiceBank-DEMAND-shaped noisy speech [batch, length] float32: six harmonics of a per-utterance pitch
    under a slow envelope plus white noise whose level grows with the utterance index (so the SNR estimator sees
    different inputs).  The benchmark, the reference arm and the parity tests all draw their inputs here."""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(length) / sr
    waves = []
    for b in range(batch):
        f0 = 110.0 + 17.0 * b
        speech = sum(torch.sin(2 * torch.pi * f0 * (k + 1) * t + k) / (k + 1) for k in range(6))
        env = 0.5 + 0.5 * torch.sin(2 * torch.pi * (2.0 + 0.1 * b) * t)
        noise = torch.randn(length, generator=g)
        waves.append(0.1 * speech * env + (0.01 + 0.004 * b) * noise)
    return torch.stack(waves).to(torch.float32)


def synth_noise(batch: int, tpad: int, seed: int) -> torch.Tensor:
    """Explicit unit complex normal draw Z [batch,1,256,tpad] complex64 (Var re = Var im = 1/2, as randn_like on a
    complex tensor); parity runs feed the same Z to the reference / oracle and to the CUDA path."""
    g = torch.Generator().manual_seed(seed)
    return torch.view_as_complex(torch.randn(batch, 1, 256, tpad, 2, generator=g) * (0.5 ** 0.5))
