"""`NCSNpp` backbone (mirror of sgmse-bbed/sgmse/backbones/ncsnpp.py:36-404).

Holds the reference-format fp32 parameters on the host and a native executor
(`snr_aligned_diffse_b200.engine.NCSNppEngine`) with the packed bf16/fp32 copy on the GPU.
`forward(x, time_cond)` keeps the reference contract: x [B,2,F,T] complex64, time_cond [B] -> [B,1,F,T].
"""
import torch

from ...engine import MODE_RAW, NCSNppEngine
from .shared import BackboneRegistry


@BackboneRegistry.register("ncsnpp")
class NCSNpp:
    @staticmethod
    def add_argparse_args(parser):
        return parser

    def __init__(self, scale_by_sigma=True, nonlinearity='swish', nf=128, ch_mult=(1, 1, 2, 2, 2, 2, 2),
                 num_res_blocks=2, attn_resolutions=(16,), resamp_with_conv=True, conditional=True, fir=True,
                 fir_kernel='song', skip_rescale=True, resblock_type='biggan', progressive='output_skip',
                 progressive_input='input_skip', progressive_combine='sum', init_scale=0., fourier_scale=16,
                 image_size=256, embedding_type='fourier', dropout=.0, **unused_kwargs):
        fixed = dict(nonlinearity=(nonlinearity, 'swish'), conditional=(conditional, True), fir=(fir, True),
                     skip_rescale=(skip_rescale, True), resblock_type=(resblock_type.lower(), 'biggan'),
                     progressive=(progressive.lower(), 'output_skip'),
                     progressive_input=(progressive_input.lower(), 'input_skip'),
                     progressive_combine=(progressive_combine.lower(), 'sum'),
                     embedding_type=(embedding_type.lower(), 'fourier'))
        for k, (got, want) in fixed.items():
            if got != want:
                raise NotImplementedError(f"NCSNpp({k}={got!r}) is not implemented on the B200 path (only {want!r})")
        if dropout != 0:
            raise NotImplementedError("dropout is a training-time option; inference uses p=0")
        self.nf, self.ch_mult, self.num_res_blocks = nf, tuple(ch_mult), num_res_blocks
        self.attn_resolutions, self.image_size = tuple(attn_resolutions), image_size
        self.engine = NCSNppEngine(nf, self.ch_mult, num_res_blocks, self.attn_resolutions, image_size)
        self._shapes = {k[len("dnn."):]: v for k, v in self.engine.param_shapes().items()}
        self._sd = None
        self._dirty = True

    # ---- parameters (reference state-dict names without the `dnn.` prefix)
    def param_shapes(self):
        return dict(self._shapes)

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._shapes if k not in sd]
        if missing and strict:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, shp in self._shapes.items():
            if k in sd and tuple(sd[k].shape) != tuple(shp):
                raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {tuple(shp)}")
        self._sd = {k: sd[k].detach().to("cpu", torch.float32).clone() for k in self._shapes if k in sd}
        self._dirty = True

    def state_dict(self):
        return dict(self._sd or {})

    def _ensure_device_weights(self):
        if self._sd is None:
            raise RuntimeError("NCSNpp has no weights: load a checkpoint / state dict first")
        if self._dirty:
            self.engine.load_state_dict({"dnn." + k: v for k, v in self._sd.items()}, "cuda")
            self._dirty = False

    def forward(self, x, time_cond, mode=MODE_RAW, state=None):
        """x [B,2,F,T] complex64 (channel 0 state, channel 1 noisy); time_cond [B]."""
        self._ensure_device_weights()
        dev = x.device
        xg = x if x.is_cuda else x.cuda()
        out = self.engine.forward(xg[:, 0], xg[:, 1], time_cond.to("cuda", torch.float32), mode=mode)
        return out[:, None].to(dev)

    __call__ = forward
