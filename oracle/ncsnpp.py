"""CPU restatement of the reference NCSN++ forward pass (oracle; TEST INFRASTRUCTURE ONLY).

Functional fp32 PyTorch, driven directly by a reference-format `state_dict`
(`dnn.all_modules.<i>.<sub>` / `dnn.output_layer.*`).  Each function cites the reference lines
it restates (paths relative to /root/reference/sgmse-bbed/sgmse/backbones/).
"""
import math

import torch
import torch.nn.functional as F

from .topology import NCSNppConfig, build_modules

SQRT2 = math.sqrt(2.0)


# ---------------------------------------------------------------------------------------------
# FIR resampling (ncsnpp_utils/up_or_down_sampling.py:181-257, op/upfirdn2d.py:159-200)
# ---------------------------------------------------------------------------------------------
_FIR_1D = torch.tensor([1.0, 3.0, 3.0, 1.0])


def _fir_kernel_2d(gain: float) -> torch.Tensor:
    # _setup_kernel: outer([1,3,3,1]) normalised to sum 1 (up_or_down_sampling.py:181-188)
    k = torch.outer(_FIR_1D, _FIR_1D)
    return k / k.sum() * gain


def fir_upsample_2d(x: torch.Tensor) -> torch.Tensor:
    """upsample_2d(x, [1,3,3,1], factor=2): zero-insert x2, pad (2,1), 4x4 FIR with gain 4
    (up_or_down_sampling.py:195-224)."""
    n, c, h, w = x.shape
    k = _fir_kernel_2d(4.0).to(x)
    z = x.new_zeros(n, c, h, 2, w, 2)
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(n * c, 1, 2 * h, 2 * w)
    z = F.pad(z, (2, 1, 2, 1))
    out = F.conv2d(z, torch.flip(k, (0, 1))[None, None])
    return out.reshape(n, c, 2 * h, 2 * w)


def fir_downsample_2d(x: torch.Tensor) -> torch.Tensor:
    """downsample_2d(x, [1,3,3,1], factor=2): pad (1,1), 4x4 FIR, keep every 2nd sample
    (up_or_down_sampling.py:227-257)."""
    n, c, h, w = x.shape
    k = _fir_kernel_2d(1.0).to(x)
    z = F.pad(x.reshape(n * c, 1, h, w), (1, 1, 1, 1))
    out = F.conv2d(z, torch.flip(k, (0, 1))[None, None], stride=2)
    return out.reshape(n, c, out.shape[-2], out.shape[-1])


# ---------------------------------------------------------------------------------------------
# Blocks
# ---------------------------------------------------------------------------------------------
def _gn(sd, key, x):
    c = x.shape[1]
    # nn.GroupNorm(num_groups=min(C//4, 32), eps=1e-6)  (layerspp.py:221,233,69; ncsnpp.py:210)
    return F.group_norm(x, min(c // 4, 32), sd[key + ".weight"], sd[key + ".bias"], eps=1e-6)


def _nin(sd, key, x):
    # NIN: channel contraction x[b,c,h,w] W[c,o] + b[o]  (layers.py:537-555)
    return torch.einsum("bchw,co->bohw", x, sd[key + ".W"]) + sd[key + ".b"][None, :, None, None]


def resblock(sd, p, m, x, temb):
    """ResnetBlockBigGANpp.forward (layerspp.py:244-276)."""
    h = F.silu(_gn(sd, p + "GroupNorm_0", x))
    if m["up"]:
        h, x = fir_upsample_2d(h), fir_upsample_2d(x)
    elif m["down"]:
        h, x = fir_downsample_2d(h), fir_downsample_2d(x)
    h = F.conv2d(h, sd[p + "Conv_0.weight"], sd[p + "Conv_0.bias"], padding=1)
    h = h + F.linear(F.silu(temb), sd[p + "Dense_0.weight"], sd[p + "Dense_0.bias"])[:, :, None, None]
    h = F.silu(_gn(sd, p + "GroupNorm_1", h))
    h = F.conv2d(h, sd[p + "Conv_1.weight"], sd[p + "Conv_1.bias"], padding=1)
    if (p + "Conv_2.weight") in sd:
        x = F.conv2d(x, sd[p + "Conv_2.weight"], sd[p + "Conv_2.bias"])
    return (x + h) / SQRT2


def attnblock(sd, p, x):
    """AttnBlockpp.forward (layerspp.py:78-93): single head over all H*W positions."""
    b, c, hh, ww = x.shape
    h = _gn(sd, p + "GroupNorm_0", x)
    q = _nin(sd, p + "NIN_0", h).reshape(b, c, hh * ww)
    k = _nin(sd, p + "NIN_1", h).reshape(b, c, hh * ww)
    v = _nin(sd, p + "NIN_2", h).reshape(b, c, hh * ww)
    w = torch.einsum("bci,bcj->bij", q, k) * (int(c) ** (-0.5))
    w = torch.softmax(w, dim=-1)
    h = torch.einsum("bij,bcj->bci", w, v).reshape(b, c, hh, ww)
    h = _nin(sd, p + "NIN_3", h)
    return (x + h) / SQRT2


def time_embedding(sd, prefix, t):
    """Fourier features of log(t) + 2-layer MLP (ncsnpp.py:256-275, layerspp.py:32-43)."""
    w = sd[prefix + "all_modules.0.W"]
    proj = torch.log(t)[:, None] * w[None, :] * 2 * math.pi
    temb = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)
    temb = F.linear(temb, sd[prefix + "all_modules.1.weight"], sd[prefix + "all_modules.1.bias"])
    temb = F.linear(F.silu(temb), sd[prefix + "all_modules.2.weight"], sd[prefix + "all_modules.2.bias"])
    return temb


def ncsnpp_forward(sd, x, t, cfg: NCSNppConfig = NCSNppConfig(), prefix: str = "dnn.", taps=None):
    """NCSNpp.forward (ncsnpp.py:247-404).

    x: [B,2,F,T] complex64 (channel 0 = current state, channel 1 = noisy spectrogram y)
    t: [B] float32.   Returns [B,1,F,T] complex64.
    `taps`, if a dict, receives intermediate activations (NCHW fp32) keyed by module index.
    """
    mods = build_modules(cfg)
    L = cfg.num_resolutions

    def P(i):
        return f"{prefix}all_modules.{i}."

    x4 = torch.cat((x[:, [0]].real, x[:, [0]].imag, x[:, [1]].real, x[:, [1]].imag), dim=1)  # :253
    temb = time_embedding(sd, prefix, t)
    mi = 3
    input_pyramid = x4
    h = F.conv2d(x4, sd[P(mi) + "weight"], sd[P(mi) + "bias"], padding=1)  # :285
    if taps is not None:
        taps[mi] = h
    hs = [h]
    mi += 1
    for i_level in range(L):
        for _ in range(cfg.num_res_blocks):
            h = resblock(sd, P(mi), mods[mi], hs[-1], temb)
            if taps is not None:
                taps[mi] = h
            mi += 1
            if h.shape[-2] in cfg.attn_resolutions:  # :295 (frequency axis)
                h = attnblock(sd, P(mi), h)
                if taps is not None:
                    taps[mi] = h
                mi += 1
            hs.append(h)
        if i_level != L - 1:
            h = resblock(sd, P(mi), mods[mi], hs[-1], temb)  # DOWN block :306
            if taps is not None:
                taps[mi] = h
            mi += 1
            input_pyramid = fir_downsample_2d(input_pyramid)  # :310
            h = F.conv2d(input_pyramid, sd[P(mi) + "Conv_0.weight"], sd[P(mi) + "Conv_0.bias"]) + h  # Combine 'sum'
            if taps is not None:
                taps[mi] = h
            mi += 1
            hs.append(h)
    h = hs[-1]
    h = resblock(sd, P(mi), mods[mi], h, temb); mi += 1  # :325
    h = attnblock(sd, P(mi), h); mi += 1
    h = resblock(sd, P(mi), mods[mi], h, temb); mi += 1
    if taps is not None:
        taps[mi - 1] = h
    pyramid = None
    for i_level in reversed(range(L)):
        for _ in range(cfg.num_res_blocks + 1):
            h = resblock(sd, P(mi), mods[mi], torch.cat([h, hs.pop()], dim=1), temb)  # :337
            if taps is not None:
                taps[mi] = h
            mi += 1
        if h.shape[-2] in cfg.attn_resolutions:  # :341
            h = attnblock(sd, P(mi), h)
            if taps is not None:
                taps[mi] = h
            mi += 1
        ph = F.silu(_gn(sd, P(mi)[:-1], h)); mi += 1                         # :348 / :362
        ph = F.conv2d(ph, sd[P(mi) + "weight"], sd[P(mi) + "bias"], padding=1); mi += 1
        pyramid = ph if pyramid is None else fir_upsample_2d(pyramid) + ph  # :361-366
        if taps is not None:
            taps[mi - 1] = pyramid
        if i_level != 0:
            h = resblock(sd, P(mi), mods[mi], h, temb)  # UP block :384
            if taps is not None:
                taps[mi] = h
            mi += 1
    assert not hs and mi == len(mods)
    h = pyramid / t[:, None, None, None]  # :398
    h = F.conv2d(h, sd[prefix + "output_layer.weight"], sd[prefix + "output_layer.bias"])  # :401
    h = h.permute(0, 2, 3, 1).contiguous()
    return torch.view_as_complex(h)[:, None]


def upfirdn2d_general(x, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """General upfirdn2d restated in numpy float64 loops over taps (ncsnpp_utils/op/upfirdn2d.py:159-200):
    zero-insert by (up_x, up_y), pad (negative = crop), correlate with the flipped kernel, keep every
    (down_y, down_x)-th sample.  x [N, C, H, W], kernel [kh, kw] -> float32 tensor [N, C, out_h, out_w]."""
    import numpy as np
    xn = np.asarray(x, dtype=np.float64)
    kn = np.asarray(kernel, dtype=np.float64)
    N, C, H, W = xn.shape
    kh, kw = kn.shape
    U = np.zeros((N, C, H * up_y, W * up_x))
    U[:, :, ::up_y, ::up_x] = xn
    P = np.pad(U, ((0, 0), (0, 0), (max(pad_y0, 0), max(pad_y1, 0)), (max(pad_x0, 0), max(pad_x1, 0))))
    P = P[:, :, max(-pad_y0, 0):P.shape[2] - max(-pad_y1, 0), max(-pad_x0, 0):P.shape[3] - max(-pad_x1, 0)]
    fh, fw = P.shape[2] - kh + 1, P.shape[3] - kw + 1
    full = np.zeros((N, C, fh, fw))
    for ky in range(kh):
        for kx in range(kw):
            full += kn[kh - 1 - ky, kw - 1 - kx] * P[:, :, ky:ky + fh, kx:kx + fw]
    return torch.from_numpy(full[:, :, ::down_y, ::down_x].astype(np.float32))
