"""Correctors (mirror of sgmse-bbed/sgmse/sampling/correctors.py:9-94)."""
import abc

import torch

from ..sdes import axpby
from ..util.registry import Registry

CorrectorRegistry = Registry("Corrector")


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__()
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.snr = snr
        self.n_steps = n_steps

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        pass


@CorrectorRegistry.register(name='langevin')
class LangevinCorrector(Corrector):
    def update_fn(self, x, t, *args):
        x_mean = x
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, *args)
            noise = torch.randn_like(x)
            grad_norm = torch.linalg.vector_norm(torch.view_as_real(grad).reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = torch.linalg.vector_norm(torch.view_as_real(noise).reshape(noise.shape[0], -1), dim=-1).mean()
            step_size = ((self.snr * noise_norm / grad_norm) ** 2 * 2).reshape(1)
            x_mean, x = axpby(x=x, a=1.0, s=grad, c=step_size, z=noise, d=torch.sqrt(step_size * 2), mean=True)
        return x, x_mean


@CorrectorRegistry.register(name='ald')
class AnnealedLangevinDynamics(Corrector):
    def __init__(self, sde, score_fn, snr, n_steps):
        super().__init__(sde, score_fn, snr, n_steps)
        self.sde = sde

    def update_fn(self, x, t, y):
        x_mean = 0
        std = self.sde._std(t)
        for _ in range(self.n_steps):
            grad = self.score_fn(x, t, y)
            noise = torch.randn_like(x)
            step_size = (self.snr * std) ** 2 * 2
            x_mean, x = axpby(x=x, a=1.0, s=grad, c=step_size, z=noise, d=torch.sqrt(step_size * 2), mean=True)
        return x, x_mean


@CorrectorRegistry.register(name='none')
class NoneCorrector(Corrector):
    def __init__(self, *args, **kwargs):
        self.snr = 0
        self.n_steps = 0

    def update_fn(self, x, t, *args):
        return x, x
