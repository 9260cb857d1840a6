"""Predictor half of the predictor-corrector samplers.

Public surface kept from the reference (sgmse-bbed/sgmse/sampling/predictors.py:10-94): `PredictorRegistry` with the
keys 'euler_maruyama', 'reverse_diffusion', 'none'; `cls(sde, score_fn, probability_flow=False)`;
`update_fn(x, t, *args) -> (x, x_mean)`; `debug_update_fn` raising NotImplementedError.  A predictor move is

    x_mean = x + drift_term,        x = x_mean + noise_scale * z,        z ~ CN(0, 1)

evaluated as one fused `lincomb` launch; the two predictors differ in how the reverse SDE supplies the two terms.
"""
import abc
import math

import torch

from ..sdes import axpby
from ..util.registry import Registry

PredictorRegistry = Registry("Predictor")


def _move(x, drift, drift_coef, z, noise_scale):
    """(x_new, x_mean) with x_mean = x + drift_coef * drift and x_new = x_mean + noise_scale * z."""
    x_mean, x_new = axpby(x=x, a=1.0, y=drift, b=drift_coef, z=z, d=noise_scale, mean=True)
    return x_new, x_mean


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde, self.score_fn, self.probability_flow = sde, score_fn, probability_flow
        self.rsde = sde.reverse(score_fn)

    @abc.abstractmethod
    def update_fn(self, x, t, *args):
        """-> (next state, next state without the injected noise)."""

    def debug_update_fn(self, x, t, *args):
        raise NotImplementedError(f"Debug update function not implemented for predictor {self}.")


class EulerMaruyamaPredictor(Predictor):
    """Fixed step dt = -1/N of the reverse SDE's continuous drift and diffusion (predictors.py:41-52)."""

    def update_fn(self, x, t, *args):
        if len(args) != 1:
            # pc_sampler calls update_fn(x, t, y, stepsize) (sampling/__init__.py:72); in the reference the extra argument
            # reaches OUVESDE.sde / BBED.sde(x, t, y) through RSDE.rsde_parts (sdes.py:121) and raises exactly this.
            # Pinned by tests/golden/em_step.npz ("in_loop": TypeError): the predictor only works called as (x, t, y).
            raise TypeError(f"sde() takes 4 positional arguments but {3 + len(args)} were given")
        dt = -1.0 / self.rsde.N
        z = torch.randn_like(x)
        drift, g = self.rsde.sde(x, t, *args)
        return _move(x, drift, dt, z, g * math.sqrt(-dt))


class ReverseDiffusionPredictor(Predictor):
    """Ancestral step with the discretised reverse SDE: x_mean = x - f, x = x_mean + G z (predictors.py:70-83)."""

    def update_fn(self, x, t, y, stepsize):
        f, G = self.rsde.discretize(x, t, y, stepsize)
        return _move(x, f, -1.0, torch.randn_like(x), G)


class NonePredictor(Predictor):
    """Identity (predictors.py:86-94): constructed with any arguments."""

    def __init__(self, *args, **kwargs):
        pass

    def update_fn(self, x, t, *args):
        return x, x


for _key, _cls in (("euler_maruyama", EulerMaruyamaPredictor), ("reverse_diffusion", ReverseDiffusionPredictor),
                   ("none", NonePredictor)):
    PredictorRegistry.register(_key)(_cls)
