"""`pad_spec`, `pad_spec_16`, `si_sdr` (mirror of sgmse-bbed/sgmse/util/other.py:71-99)."""
import numpy as np
import torch


def _pad_time(Y, multiple):
    T = Y.size(3)
    num_pad = (multiple - T % multiple) % multiple
    return torch.nn.functional.pad(Y, (0, num_pad, 0, 0))   # zero fill on the right of the time axis


def pad_spec(Y):
    return _pad_time(Y, 64)


def pad_spec_16(Y):
    return _pad_time(Y, 16)


def si_sdr(s, s_hat):
    alpha = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    return 10 * np.log10(np.linalg.norm(alpha * s) ** 2 / np.linalg.norm(alpha * s - s_hat) ** 2)


def snr_dB(s, n):
    s_power = 1 / len(s) * np.sum(np.abs(s) ** 2)
    n_power = 1 / len(n) * np.sum(np.abs(n) ** 2)
    return 10 * np.log10(s_power / n_power)
