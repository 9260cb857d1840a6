"""Corrector half of the predictor-corrector samplers.

Public surface kept from the reference (sgmse-bbed/sgmse/sampling/correctors.py:9-94): `CorrectorRegistry` with the
keys 'langevin', 'ald', 'none'; classes constructed as `cls(sde, score_fn, snr, n_steps)` and called as
`update_fn(x, t, *args) -> (x, x_mean)`.  Each Langevin move

    x_mean = x + eps * score(x, t, y),      x = x_mean + sqrt(2 eps) * z,      z ~ CN(0, 1)

is one fused `lincomb` launch on the GPU; only how eps is chosen differs between the two correctors.
"""
import abc

import torch

from ..sdes import axpby
from ..util.registry import Registry

CorrectorRegistry = Registry("Corrector")


def _batch_mean_l2(v):
    """Mean over the batch of the per-sample Euclidean norm of a complex tensor."""
    return torch.linalg.vector_norm(torch.view_as_real(v).flatten(1), dim=-1).mean()


def _langevin_move(x, score, z, eps):
    """(x_mean, x_new) for one Langevin move with per-sample (or shared) step size `eps`."""
    return axpby(x=x, a=1.0, s=score, c=eps, z=z, d=torch.sqrt(2 * eps), mean=True)


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        self.sde = sde
        self.rsde = sde.reverse(score_fn)
        self.score_fn, self.snr, self.n_steps = score_fn, snr, n_steps

    @abc.abstractmethod
    def step_size(self, x, t, score, z, *args):
        """eps of the next Langevin move."""

    def update_fn(self, x, t, *args):
        x_mean = self._initial_mean(x)
        for _ in range(self.n_steps):
            score = self.score_fn(x, t, *args)
            z = torch.randn_like(x)
            x_mean, x = _langevin_move(x, score, z, self.step_size(x, t, score, z, *args))
        return x, x_mean

    @staticmethod
    def _initial_mean(x):
        return x


class LangevinCorrector(Corrector):
    """eps from the ratio of the noise and score norms, shared by the whole batch (correctors.py:37-56)."""

    def step_size(self, x, t, score, z, *args):
        return (2 * (self.snr * _batch_mean_l2(z) / _batch_mean_l2(score)) ** 2).reshape(1)


class AnnealedLangevinDynamics(Corrector):
    """eps = 2 (snr * std(t))^2 with the perturbation kernel's std, per sample (correctors.py:59-81)."""

    def update_fn(self, x, t, y):
        self._std = self.sde._std(t)          # fixed over the inner moves, as in the reference
        return super().update_fn(x, t, y)

    def step_size(self, x, t, score, z, *args):
        return 2 * (self.snr * self._std) ** 2

    @staticmethod
    def _initial_mean(x):
        return 0                              # what the reference returns as x_mean when n_steps == 0


class NoneCorrector(Corrector):
    """Identity (correctors.py:84-94): constructed with any arguments, performs no network evaluation."""

    def __init__(self, *args, **kwargs):
        self.snr, self.n_steps = 0, 0

    def step_size(self, *a):
        raise AssertionError("NoneCorrector never moves")

    def update_fn(self, x, t, *args):
        return x, x


for _key, _cls in (("langevin", LangevinCorrector), ("ald", AnnealedLangevinDynamics), ("none", NoneCorrector)):
    CorrectorRegistry.register(_key)(_cls)
