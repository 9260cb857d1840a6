"""Predictors (mirror of sgmse-bbed/sgmse/sampling/predictors.py:10-94)."""
import abc

import numpy as np
import torch

from ..sdes import axpby
from ..util.registry import Registry

PredictorRegistry = Registry("Predictor")


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        super().__init__()
        self.sde = sde
        self.rsde = sde.reverse(score_fn)
        self.score_fn = score_fn
        self.probability_flow = probability_flow

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        pass

    def debug_update_fn(self, x, t, *args):
        raise NotImplementedError(f"Debug update function not implemented for predictor {self}.")


@PredictorRegistry.register('euler_maruyama')
class EulerMaruyamaPredictor(Predictor):
    def update_fn(self, x, t, *args):
        dt = -1. / self.rsde.N
        z = torch.randn_like(x)
        f, g = self.rsde.sde(x, t, *args)
        return tuple(reversed(axpby(x=x, a=1.0, y=f, b=dt, z=z, d=g * float(np.sqrt(-dt)), mean=True)))


@PredictorRegistry.register('reverse_diffusion')
class ReverseDiffusionPredictor(Predictor):
    def update_fn(self, x, t, y, stepsize):
        f, g = self.rsde.discretize(x, t, y, stepsize)
        z = torch.randn_like(x)
        x_mean, x_new = axpby(x=x, a=1.0, y=f, b=-1.0, z=z, d=g, mean=True)   # x_mean = x - f; x = x_mean + g z
        return x_new, x_mean


@PredictorRegistry.register('none')
class NonePredictor(Predictor):
    def __init__(self, *args, **kwargs):
        pass

    def update_fn(self, x, t, *args):
        return x, x
