"""SDE registry, OU-VE and BBED processes and the reverse-SDE factory
(mirror of sgmse-bbed/sgmse/sdes.py:17-307).

Per-batch scalars (sigma(t), std(t), step coefficients) are tiny [B] tensors evaluated on the host
side exactly as the reference does (BBED._std keeps the scipy `expi` call, sdes.py:287-293); every
full-size complex tensor update is one fused `lincomb` launch  out = a*x + b*y + c*s (+ d*z).
"""
import abc
import warnings

import numpy as np
import scipy.special as sc
import torch

from .. import ops
from .util.registry import Registry

SDERegistry = Registry("SDE")


def _vec(v, like, B):
    """[B] float32 cuda tensor from a python scalar / 0-d / [B] tensor."""
    if torch.is_tensor(v):
        v = v.to(device=like.device, dtype=torch.float32).reshape(-1)
        return v.expand(B).contiguous() if v.numel() == 1 else v.contiguous()
    return torch.full((B,), float(v), dtype=torch.float32, device=like.device)


def axpby(x=None, a=None, y=None, b=None, s=None, c=None, z=None, d=None, mean=False):
    """a*x + b*y + c*s (+ d*z) on complex64 tensors [B,...]; returns (mean, full) or full."""
    ref = next(v for v in (x, y, s, z) if v is not None)
    B = ref.shape[0]
    va = _vec(a, ref, B) if x is not None else None
    vb = _vec(b, ref, B) if y is not None else None
    vc = _vec(c, ref, B) if s is not None else None
    vd = _vec(d, ref, B) if z is not None else None
    m, f = ops.lincomb(x, y, s, z, va, vb, vc, vd, want_mean=mean, want_x=True)
    return (m, f) if mean else f


class SDE(abc.ABC):
    def __init__(self, N):
        super().__init__()
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        pass

    @abc.abstractmethod
    def sde(self, x, t, *args):
        pass

    @abc.abstractmethod
    def drift_coeffs(self, t):
        """(a, b) with drift = a*x + b*y as [B] tensors."""

    @abc.abstractmethod
    def diffusion(self, t):
        """g(t) as a [B] tensor."""

    @abc.abstractmethod
    def marginal_prob(self, x, t, *args):
        pass

    @abc.abstractmethod
    def prior_sampling(self, shape, *args):
        pass

    def prior_logp(self, z):
        raise NotImplementedError("prior_logp is not implemented")

    def discretize(self, x, t, y, stepsize):
        """f = drift*dt, G = diffusion*sqrt(dt)  (sdes.py:73-91)."""
        dt = float(stepsize)
        drift, diffusion = self.sde(x, t, y)
        f = axpby(x=drift, a=dt)
        G = diffusion * float(np.sqrt(np.float32(dt)))
        return f, G

    def reverse(oself, score_model, probability_flow=False):
        """Reverse-time SDE / probability-flow ODE (sdes.py:93-142)."""
        N, T = oself.N, oself.T
        scale = 0.5 if probability_flow else 1.0

        class RSDE(oself.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, t, *args):
                parts = self.rsde_parts(x, t, *args)
                return parts["total_drift"], parts["diffusion"]

            def rsde_parts(self, x, t, *args):
                y = args[0]
                a, b = oself.drift_coeffs(t)
                g = oself.diffusion(t)
                score = score_model(x, t, *args)
                sde_drift = axpby(x=x, a=a, y=y, b=b)
                score_drift = axpby(x=score, a=-(g ** 2) * scale)
                total = axpby(x=x, a=a, y=y, b=b, s=score, c=-(g ** 2) * scale)
                diffusion = torch.zeros_like(g) if probability_flow else g
                return {'total_drift': total, 'diffusion': diffusion, 'sde_drift': sde_drift,
                        'sde_diffusion': g, 'score_drift': score_drift, 'score': score}

            def discretize(self, x, t, y, stepsize):
                """rev_f = f - G^2 * score, rev_G = G  (sdes.py:132-140), one fused launch."""
                dt = float(stepsize)
                a, b = oself.drift_coeffs(t)
                G = oself.diffusion(t) * float(np.sqrt(np.float32(dt)))
                score = score_model(x, t, y)
                rev_f = axpby(x=x, a=a * dt, y=y, b=b * dt, s=score, c=-(G ** 2) * scale)
                rev_G = torch.zeros_like(G) if probability_flow else G
                return rev_f, rev_G

        return RSDE()

    @abc.abstractmethod
    def copy(self):
        pass


@SDERegistry.register("ouve")
class OUVESDE(SDE):
    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--sde-n", type=int, default=1000)
        parser.add_argument("--theta", type=float, default=1.5)
        parser.add_argument("--sigma-min", type=float, default=0.05)
        parser.add_argument("--sigma-max", type=float, default=0.5)
        return parser

    def __init__(self, theta, sigma_min, sigma_max, N=1000, **ignored_kwargs):
        super().__init__(N)
        self.theta, self.sigma_min, self.sigma_max = theta, sigma_min, sigma_max
        self.logsig = np.log(self.sigma_max / self.sigma_min)
        self.N = N
        self._T = 1

    def copy(self):
        return OUVESDE(self.theta, self.sigma_min, self.sigma_max, N=self.N)

    @property
    def T(self):
        return self._T

    def drift_coeffs(self, t):
        th = torch.full_like(t, float(self.theta), dtype=torch.float32)
        return -th, th

    def diffusion(self, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return (sigma * np.sqrt(2 * self.logsig)).to(torch.float32)

    def sde(self, x, t, y):
        a, b = self.drift_coeffs(t)
        return axpby(x=x, a=a, y=y, b=b), self.diffusion(t)

    def _mean(self, x0, t, y):
        e = torch.exp(-self.theta * t)
        return axpby(x=x0, a=e, y=y, b=1 - e)

    def _std(self, t):
        sigma_min, theta, logsig = self.sigma_min, self.theta, self.logsig
        return torch.sqrt((sigma_min ** 2 * torch.exp(-2 * theta * t) * (torch.exp(2 * (theta + logsig) * t) - 1) * logsig)
                          / (theta + logsig))

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_sampling(self, shape, y):
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        std = self._std(torch.ones((y.shape[0],), device=y.device))
        z = torch.randn_like(y)
        return axpby(y=y, b=1.0, z=z, d=std), z


@SDERegistry.register("bbed")
class BBED(SDE):
    @staticmethod
    def add_argparse_args(parser):
        parser.add_argument("--sde-n", type=int, default=30)
        parser.add_argument("--T_sampling", type=float, default=0.999)
        parser.add_argument("--k", type=float, default=2.6)
        parser.add_argument("--theta", type=float, default=0.52)
        return parser

    def __init__(self, T_sampling, k, theta, N=1000, **kwargs):
        super().__init__(N)
        self.k = k
        self.logk = float(np.log(self.k))
        self.theta = theta
        self.N = N
        self.Eilog = float(sc.expi(-2 * self.logk))
        self._Tval = T_sampling
        self.Tc = 1

    def copy(self):
        return BBED(self.T, self.k, self.theta, N=self.N)

    @property
    def T(self):
        return self._Tval

    @T.setter
    def T(self, v):          # eval.py pokes `model.sde.T = reverse_starting_point` (eval.py:108)
        self._Tval = v

    def drift_coeffs(self, t):
        inv = (1.0 / (self.Tc - t)).to(torch.float32)   # per-sample 1/(Tc-t): the reference's B==1 semantics
        return -inv, inv

    def diffusion(self, t):
        return ((self.k) ** t * np.sqrt(self.theta)).to(torch.float32)

    def sde(self, x, t, y):
        a, b = self.drift_coeffs(t)
        return axpby(x=x, a=a, y=y, b=b), self.diffusion(t)

    def _mean(self, x0, t, y):
        time = t / self.Tc
        return axpby(x=x0, a=1 - time, y=y, b=time)

    def _std(self, t):
        hv = getattr(t, "_host_value", None)
        if hv is not None:
            # the sampler loops tag their uniform time vector with its host value: same float32 -> float64 expi
            # arithmetic as below, but no device->host read (one sync per step less; CUDA-graph capturable)
            t1 = np.full((1,), np.float32(hv))
            Eis = sc.expi(2 * (t1 - 1) * self.logk) - self.Eilog
            var = (self.k ** (2 * t1) - 1 + t1) + 2 * self.k ** 2 * self.logk * (1 - t1) * Eis
            var = torch.full_like(t, float(np.float32(var[0])), dtype=torch.float32) * (1 - t) * self.theta
            return torch.sqrt(var)
        t_np = t.detach().cpu().numpy()
        Eis = sc.expi(2 * (t_np - 1) * self.logk) - self.Eilog
        h = 2 * self.k ** 2 * self.logk
        var = (self.k ** (2 * t_np) - 1 + t_np) + h * (1 - t_np) * Eis
        var = torch.tensor(var).to(device=t.device, dtype=torch.float32) * (1 - t) * self.theta
        return torch.sqrt(var)

    def marginal_prob(self, x0, t, y):
        return self._mean(x0, t, y), self._std(t)

    def prior_sampling(self, shape, y):
        if shape != y.shape:
            warnings.warn(f"Target shape {shape} does not match shape of y {y.shape}! Ignoring target shape.")
        std = self._std(self.T * torch.ones((y.shape[0],), device=y.device))
        z = torch.randn_like(y)
        return axpby(y=y, b=1.0, z=z, d=std), z
