"""CPU restatement of the SNR-conditioned sampler logic (oracle; TEST INFRASTRUCTURE ONLY).

Restates, with file:line citations into /root/reference/sgmse-bbed/sgmse/:
  * the t_30 grid, timestep snapping and normalisation factor (model.py:22-23,627-634,726-740,811-817)
  * the sebridge_v3 preconditioned network call (model.py:481-484,536-543)
  * the bbed score head (model.py:488-489)
  * OUVESDE / BBED (sdes.py:149-307), SDE.discretize / reverse (sdes.py:73-142)
  * the predictor-corrector loop (sampling/__init__.py:54-91, predictors.py:70-83, correctors.py:59-81)
Random numbers are never drawn here: every function takes the noise explicitly.
"""
import math

import numpy as np
import scipy.special as sc
import torch

from . import frontend
from .ncsnpp import ncsnpp_forward

# model.py:22-23 (float64 numpy)
I_30 = np.arange(1, 30 + 1)
T_30 = (0.001 ** (1 / 7) + (I_30 - 1) / (30 - 1) * (1 ** (1 / 7) - 0.001 ** (1 / 7))) ** 7


def snr_to_t_raw(noise_over_clean, fixed_snr):
    """calculate_snr_direct(1, est_snr, fixed_snr) (model.py:627-629)."""
    return noise_over_clean / (10 ** 0.25 * fixed_snr)


def snap_t(t_raw: float):
    """model.py:733-735: nearest grid point of t_30 (argmin of |t_30 - t|, first on ties)."""
    idx = int(np.abs(T_30 - t_raw).argmin())
    return idx, float(T_30[idx])


def normfac(noise_over_clean, fixed_snr):
    """calculate_normfac_direct(1, n, fixed_snr) (model.py:631-634)."""
    return 2.040166 * (0.240253 + 0.759747 * fixed_snr ** 2) ** 0.5 / ((1 + noise_over_clean ** 2) ** 0.5)


def v3_scalars(noise_over_clean: float, fixed_snr: float, peak: float):
    """model.py:726-740: returns (t_index, t, norm_factor) with float32 rounding as the reference
    (est_snr_ and normfac_ are computed on float32 tensors)."""
    t_raw = float(np.float32(snr_to_t_raw(np.float32(noise_over_clean), fixed_snr)))
    idx, t = snap_t(t_raw)
    est_snr_ = torch.FloatTensor([10 ** 0.25 * fixed_snr * t])
    nf = normfac(est_snr_, fixed_snr)  # float32 tensor arithmetic, as in the reference
    return idx, t, float((peak * nf).item())


def precond_v3(t: torch.Tensor):
    """c_skip, c_out of model.py:537-539 (sigma_data=0.5, eps=0.001); t any shape, float32."""
    eps, sigma_data = 0.001, 0.5
    c_skip = sigma_data ** 2 / ((t - eps) ** 2 + sigma_data ** 2)
    c_out = (sigma_data * (t - eps)) / ((sigma_data ** 2 + t ** 2) ** 0.5)
    return c_skip, c_out


def score_forward(sd, x, t, y, model_type="sebridge_v3", cfg=None):
    """ScoreModel.forward (model.py:481-543) for the two heads the path uses.
    x, y: [B,1,F,T] c64; t: [B,1,1,1] (sebridge_v3) or [B] (bbed)."""
    kw = {} if cfg is None else {"cfg": cfg}
    dnn_input = torch.cat([x, y], dim=1)
    if model_type == "bbed":
        return -ncsnpp_forward(sd, dnn_input, t, **kw)
    if model_type in ("sebridge_v3", "sebridge", "sebridge_v2"):
        c_skip, c_out = precond_v3(t)
        tt = t.squeeze(3).squeeze(2).squeeze(1)
        return c_skip * x + c_out * ncsnpp_forward(sd, dnn_input, tt, **kw)
    raise ValueError(model_type)


def enhance_v3(sd, y_wave, Z, noise_over_clean, fixed_snr, sigma_max=1.0, cfg=None):
    """The sebridge_v3 / snr_conditioned='true' branch of ScoreModel.enhance (model.py:702-839)
    composed on CPU, with the SNR given (oracle=True path, :723-724) and the noise Z explicit.

    y_wave: [1,L] float32; Z: [1,1,256,Tpad] complex64 (unit complex normal).
    Returns dict(x_hat [L] float32, t_index, t, norm_factor, Y, X_T, sample)."""
    L = y_wave.size(1)
    peak = y_wave.abs().max().item()
    idx, t, norm_factor = v3_scalars(noise_over_clean, fixed_snr, peak)
    y = y_wave / norm_factor
    Y = frontend.pad_spec(frontend.spec_fwd(frontend.stft(y)).unsqueeze(0))  # :749-751
    vec_t = (torch.ones(Y.shape[0]) * t)[:, None, None, None]                 # :819-820
    X_T = Y + Z * sigma_max * t                                              # :822-823
    with torch.no_grad():
        sample = score_forward(sd, X_T, vec_t, Y, "sebridge_v3", cfg)         # :824
    x_hat = frontend.istft(frontend.spec_back(sample.squeeze()), L) * norm_factor  # :828-830
    return dict(x_hat=x_hat.squeeze(), t_index=idx, t=t, norm_factor=norm_factor, Y=Y, X_T=X_T, sample=sample)


# ---------------------------------------------------------------------------------------------
# SDEs (sdes.py)
# ---------------------------------------------------------------------------------------------
class OUVE:
    """OUVESDE (sdes.py:149-235)."""

    def __init__(self, theta=1.5, sigma_min=0.05, sigma_max=0.5, N=30, T=1.0):
        self.theta, self.sigma_min, self.sigma_max, self.N, self.T = theta, sigma_min, sigma_max, N, T
        self.logsig = float(np.log(sigma_max / sigma_min))

    def sde(self, x, t, y):
        drift = self.theta * (y - x)
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return drift, sigma * np.sqrt(2 * self.logsig)

    def std(self, t):
        s, th, ls = self.sigma_min, self.theta, self.logsig
        return torch.sqrt((s ** 2 * torch.exp(-2 * th * t) * (torch.exp(2 * (th + ls) * t) - 1) * ls) / (th + ls))

    def prior_std(self, batch):
        return self.std(torch.ones(batch))  # prior_sampling uses t=1, not T (sdes.py:227)


class BBED:
    """BBED (sdes.py:240-307); logk / Eilog held as Python floats (numpy-1.22 promotion semantics
    of the reference's pinned environment: float32 tensors stay float32)."""

    def __init__(self, T_sampling=0.999, k=2.6, theta=0.52, N=30):
        self.k, self.theta, self.N, self.T, self.Tc = k, theta, N, T_sampling, 1
        self.logk = float(np.log(k))
        self.Eilog = float(sc.expi(-2 * self.logk))

    def sde(self, x, t, y):
        # NOTE: t is [B]; the reference divides [B,1,F,T] by [B] (trailing-axis broadcast), which is
        # only well-formed for B == 1 (or B == T).  Restated as is (sdes.py:276).
        drift = (y - x) / (self.Tc - t)
        sigma = self.k ** t
        return drift, sigma * np.sqrt(self.theta)

    def std(self, t):
        t_np = t.cpu().numpy()
        Eis = sc.expi(2 * (t_np - 1) * self.logk) - self.Eilog
        h = 2 * self.k ** 2 * self.logk
        var = (self.k ** (2 * t_np) - 1 + t_np) + h * (1 - t_np) * Eis
        var = torch.tensor(var).to(torch.float32) * (1 - t) * self.theta
        return torch.sqrt(var)

    def prior_std(self, batch):
        return self.std(self.T * torch.ones(batch))


def em_step(sd, x, t, Y, sde, z, N, cfg=None, score_fn=None, probability_flow=False):
    """EulerMaruyamaPredictor.update_fn(x, t, y) (predictors.py:41-52) on the reverse SDE of `sde`
    (sdes.py:115-131): dt = -1/N, x_mean = x + (f - g^2 score [/2]) dt, x = x_mean + g sqrt(-dt) z.
    (Through pc_sampler the reference passes a fourth argument and raises TypeError; only the direct call is defined.)"""
    if score_fn is None:
        score_fn = lambda x_, t_, y_: score_forward(sd, x_, t_, y_, "bbed", cfg)
    with torch.no_grad():
        dt = -1.0 / N
        drift, g = sde.sde(x, t, Y)
        g4 = g[:, None, None, None] if torch.is_tensor(g) else g
        total = drift - g4 ** 2 * score_fn(x, t, Y) * (0.5 if probability_flow else 1.0)
        x_mean = x + total * dt
        gd = torch.zeros_like(g) if probability_flow else g
        return x_mean + gd[:, None, None, None] * np.sqrt(-dt) * z, x_mean


def pc_sample(sd, Y, sde, noises, N=30, eps=0.03, snr=0.5, corrector_steps=1, cfg=None,
              score_fn=None, trace=None, corrector="ald"):
    """get_pc_sampler(...)(), reverse_diffusion predictor + ald (default) or langevin corrector
    (sampling/__init__.py:54-75, predictors.py:75-80, correctors.py:38-57,69-81, sdes.py:73-91,132-140).

    `noises`: iterator over complex64 tensors shaped like Y, consumed in the order the reference
    calls `torch.randn_like` (prior, then per step: corrector noise(s), predictor noise)."""
    noises = iter(noises)
    if score_fn is None:
        score_fn = lambda x, t, y: score_forward(sd, x, t, y, "bbed", cfg)
    B = Y.shape[0]
    with torch.no_grad():
        xt = Y + next(noises) * sde.prior_std(B)[:, None, None, None]
        timesteps = torch.linspace(sde.T, eps, N)
        xt_mean = xt
        for i in range(N):
            t = timesteps[i]
            stepsize = t - timesteps[i + 1] if i != N - 1 else timesteps[-1]
            vec_t = torch.ones(B) * t
            # corrector: annealed Langevin dynamics
            std = sde.std(vec_t)
            for _ in range(corrector_steps):
                grad = score_fn(xt, vec_t, Y)
                noise = next(noises)
                if corrector == "langevin":      # correctors.py:46-52: one step size for the whole batch
                    grad_norm = torch.norm(grad.reshape(B, -1), dim=-1).mean()
                    noise_norm = torch.norm(noise.reshape(B, -1), dim=-1).mean()
                    step_size = ((snr * noise_norm / grad_norm) ** 2 * 2).unsqueeze(0)
                else:
                    step_size = (snr * std) ** 2 * 2
                xt_mean = xt + step_size[:, None, None, None] * grad
                xt = xt_mean + noise * torch.sqrt(step_size * 2)[:, None, None, None]
            # predictor: reverse diffusion
            drift, diffusion = sde.sde(xt, vec_t, Y)
            f = drift * stepsize
            G = diffusion * torch.sqrt(torch.tensor(float(stepsize)))
            rev_f = f - G[:, None, None, None] ** 2 * score_fn(xt, vec_t, Y)
            z = next(noises)
            xt_mean = xt - rev_f
            xt = xt_mean + G[:, None, None, None] * z
            if trace is not None:
                trace.append(xt.clone())
        return xt_mean, N * (corrector_steps + 1)


def si_sdr(s: np.ndarray, s_hat: np.ndarray) -> float:
    """util/other.py:71-75."""
    alpha = np.dot(s_hat, s) / np.linalg.norm(s) ** 2
    return float(10 * np.log10(np.linalg.norm(alpha * s) ** 2 / np.linalg.norm(alpha * s - s_hat) ** 2))
