"""Module / parameter inventory of the reference NCSN++ score network (oracle side, test infrastructure).

Follows the constructor of `NCSNpp` (sgmse-bbed/sgmse/backbones/ncsnpp.py:45-245) for the only
configuration the reference instantiates (biggan resblocks, FIR resampling, input_skip /
output_skip pyramids with 'sum' combiner, Fourier embedding, conditional=True).

`build_modules(cfg)` returns the `all_modules` list as plain dicts; `param_specs(cfg)` returns the
`state_dict` entries (name -> shape) in registration order (`dnn.output_layer.*` first, because the
reference assigns `self.output_layer` before `self.all_modules`, ncsnpp.py:97,245).
"""
from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class NCSNppConfig:
    nf: int = 128
    ch_mult: Tuple[int, ...] = (1, 1, 2, 2, 2, 2, 2)
    num_res_blocks: int = 2
    attn_resolutions: Tuple[int, ...] = (16,)
    image_size: int = 256
    fourier_scale: float = 16.0
    num_channels: int = 4  # x.real, x.imag, y.real, y.imag (ncsnpp.py:96)

    @property
    def num_resolutions(self):
        return len(self.ch_mult)

    @property
    def all_resolutions(self):
        return [self.image_size // (2 ** i) for i in range(self.num_resolutions)]


def build_modules(cfg: NCSNppConfig) -> List[dict]:
    """`all_modules` in order (ncsnpp.py:99-245)."""
    nf, L = cfg.nf, cfg.num_resolutions
    mods: List[dict] = []
    mods.append(dict(kind="fourier", size=nf))                      # :103
    mods.append(dict(kind="linear", cin=2 * nf, cout=4 * nf))        # :113
    mods.append(dict(kind="linear", cin=4 * nf, cout=4 * nf))        # :116
    mods.append(dict(kind="conv3", cin=cfg.num_channels, cout=nf))  # :159
    hs_c = [nf]
    in_ch = nf
    for i_level in range(L):                                         # :163
        for _ in range(cfg.num_res_blocks):
            out_ch = nf * cfg.ch_mult[i_level]
            mods.append(dict(kind="res", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if cfg.all_resolutions[i_level] in cfg.attn_resolutions:
                mods.append(dict(kind="attn", c=in_ch))
            hs_c.append(in_ch)
        if i_level != L - 1:
            mods.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=True))  # :178
            mods.append(dict(kind="combine", cin=cfg.num_channels, cout=in_ch))        # :181
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    mods.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False))  # :192
    mods.append(dict(kind="attn", c=in_ch))
    mods.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False))
    for i_level in reversed(range(L)):                               # :198
        for _ in range(cfg.num_res_blocks + 1):
            out_ch = nf * cfg.ch_mult[i_level]
            mods.append(dict(kind="res", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if cfg.all_resolutions[i_level] in cfg.attn_resolutions:
            mods.append(dict(kind="attn", c=in_ch))
        mods.append(dict(kind="gn", c=in_ch))                                # :210 / :222
        mods.append(dict(kind="conv3", cin=in_ch, cout=cfg.num_channels))    # :212 / :224
        if i_level != 0:
            mods.append(dict(kind="res", cin=in_ch, cout=in_ch, up=True, down=False))  # :236
    assert not hs_c
    for i, m in enumerate(mods):
        m["idx"] = i
    return mods


def _gn(prefix, c):
    return [(prefix + ".weight", (c,)), (prefix + ".bias", (c,))]


def module_param_specs(m: dict, temb_dim: int) -> List[Tuple[str, tuple]]:
    """Parameter (sub-name, shape) list of one module in registration order."""
    k = m["kind"]
    if k == "fourier":
        return [("W", (m["size"],))]
    if k == "linear":
        return [("weight", (m["cout"], m["cin"])), ("bias", (m["cout"],))]
    if k == "conv3":
        return [("weight", (m["cout"], m["cin"], 3, 3)), ("bias", (m["cout"],))]
    if k == "gn":
        return [("weight", (m["c"],)), ("bias", (m["c"],))]
    if k == "combine":
        return [("Conv_0.weight", (m["cout"], m["cin"], 1, 1)), ("Conv_0.bias", (m["cout"],))]
    if k == "attn":
        c = m["c"]
        out = _gn("GroupNorm_0", c)
        for i in range(4):
            out += [(f"NIN_{i}.W", (c, c)), (f"NIN_{i}.b", (c,))]
        return out
    if k == "res":  # layerspp.py:214-242
        ci, co = m["cin"], m["cout"]
        out = _gn("GroupNorm_0", ci)
        out += [("Conv_0.weight", (co, ci, 3, 3)), ("Conv_0.bias", (co,))]
        out += [("Dense_0.weight", (co, temb_dim)), ("Dense_0.bias", (co,))]
        out += _gn("GroupNorm_1", co)
        out += [("Conv_1.weight", (co, co, 3, 3)), ("Conv_1.bias", (co,))]
        if ci != co or m["up"] or m["down"]:
            out += [("Conv_2.weight", (co, ci, 1, 1)), ("Conv_2.bias", (co,))]
        return out
    raise ValueError(k)


def param_specs(cfg: NCSNppConfig, prefix: str = "dnn.") -> Dict[str, tuple]:
    specs: Dict[str, tuple] = {}
    specs[prefix + "output_layer.weight"] = (2, cfg.num_channels, 1, 1)
    specs[prefix + "output_layer.bias"] = (2,)
    for m in build_modules(cfg):
        for sub, shape in module_param_specs(m, 4 * cfg.nf):
            specs[f"{prefix}all_modules.{m['idx']}.{sub}"] = shape
    return specs


SNRNET_SPECS = {  # sgmse-bbed/sgmse/backbones/snrnet.py:15-44
    "conv5x5_1.weight": (32, 2, 5, 5), "conv5x5_1.bias": (32,),
    "conv3x3_1.weight": (32, 32, 3, 3), "conv3x3_1.bias": (32,),
    "convt_1.weight": (32, 32, 64, 1), "convt_1.bias": (32,),
    "convt_2.weight": (32, 32, 64, 2), "convt_2.bias": (32,),
    "convt_3.weight": (32, 32, 64, 4), "convt_3.bias": (32,),
    "convt_4.weight": (32, 32, 64, 8), "convt_4.bias": (32,),
    "blstm.weight_ih_l0": (512, 128), "blstm.weight_hh_l0": (512, 128),
    "blstm.bias_ih_l0": (512,), "blstm.bias_hh_l0": (512,),
    "blstm.weight_ih_l0_reverse": (512, 128), "blstm.weight_hh_l0_reverse": (512, 128),
    "blstm.bias_ih_l0_reverse": (512,), "blstm.bias_hh_l0_reverse": (512,),
    "fc.weight": (1, 1024), "fc.bias": (1,),
}


def snrnet_param_specs(prefix: str = "dnn.") -> Dict[str, tuple]:
    return {prefix + k: v for k, v in SNRNET_SPECS.items()}
