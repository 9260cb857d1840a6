"""Transform half of `SpecsDataModule` (mirror of sgmse-bbed/sgmse/data_module.py:178-297).

stft / istft / spec_fwd / spec_back keep the reference signatures and accept tensors on any device;
the arithmetic runs in the CUDA library (inputs on the host are moved to the GPU and the result is
returned on the caller's device).  Dataset / dataloader members are out of scope.
"""
import torch

from .. import ops


def get_window(window_type, window_length):
    if window_type == 'hann':
        return torch.hann_window(window_length, periodic=True)
    if window_type == 'sqrthann':
        return torch.sqrt(torch.hann_window(window_length, periodic=True))
    raise NotImplementedError(f"Window type {window_type} not implemented!")


def _to_gpu(t):
    return t if t.is_cuda else t.cuda()


class SpecsDataModule:
    def __init__(self, base_dir="", format='default', batch_size=8, n_fft=510, hop_length=128, num_frames=256,
                 window='hann', num_workers=4, dummy=False, spec_factor=0.15, spec_abs_exponent=0.5, gpu=True,
                 normalize='noisy', transform_type="exponent", fixed_snr=1, **kwargs):
        if n_fft != 510 or hop_length != 128 or window != 'hann':
            raise NotImplementedError("the B200 front end implements the reference geometry only: "
                                      "n_fft=510, hop_length=128, periodic Hann (data_module.py:184-187)")
        if transform_type not in ("exponent", "log", "none"):
            raise NotImplementedError(f"transform_type {transform_type!r} is not one of the reference's "
                                      "'exponent' / 'log' / 'none' (data_module.py:241-254)")
        self.base_dir, self.format, self.batch_size = base_dir, format, batch_size
        self.n_fft, self.hop_length, self.num_frames = n_fft, hop_length, num_frames
        self.window = get_window(window, n_fft)
        self.windows = {}
        self.num_workers, self.dummy = num_workers, dummy
        self.spec_factor, self.spec_abs_exponent = spec_factor, spec_abs_exponent
        self.gpu, self.normalize, self.transform_type, self.fixed_snr = gpu, normalize, transform_type, fixed_snr
        self.kwargs = kwargs

    # ---- reference properties
    @property
    def stft_kwargs(self):
        return {**self.istft_kwargs, "return_complex": True}

    @property
    def istft_kwargs(self):
        return dict(n_fft=self.n_fft, hop_length=self.hop_length, window=self.window, center=True)

    def _params(self):
        if self.transform_type == "none":
            return 1.0, 1.0
        return float(self.spec_abs_exponent), float(self.spec_factor)

    @property
    def transform_code(self):
        """Kernel-side code of the spectrogram transform: 0 none, 1 exponent, 2 log (include/snrse_b200.h)."""
        return {"none": 0, "exponent": 1, "log": 2}[self.transform_type]

    # ---- transforms
    def stft(self, sig):
        """[..., L] float32 -> [..., 256, 1 + L//128] complex64 (data_module.py:291-293)."""
        lead, dev = sig.shape[:-1], sig.device
        x = _to_gpu(sig).reshape(-1, sig.shape[-1]).to(torch.float32)
        out = ops.stft(x, transform=False, tpad=ops.n_frames(x.shape[-1]))
        return out.reshape(*lead, out.shape[-2], out.shape[-1]).to(dev)

    def istft(self, spec, length=None):
        """[..., 256, T] complex64 -> [..., length] float32 (data_module.py:295-297)."""
        lead, dev = spec.shape[:-2], spec.device
        s = _to_gpu(spec).reshape(-1, spec.shape[-2], spec.shape[-1])
        if length is None:
            length = self.hop_length * (s.shape[-1] - 1)
        out = ops.istft(s, int(length), transform=False)
        return out.reshape(*lead, out.shape[-1]).to(dev)

    def _transform(self, spec, inverse):
        from .. import _lib
        alpha, beta = self._params()
        dev = spec.device
        s = _to_gpu(spec).to(torch.complex64).contiguous()
        out = torch.empty_like(s)
        lib = _lib.load()
        _lib.require_device()
        _lib.check(lib.snrse_spec_transform(_lib.ptr(s), _lib.ptr(out), s.numel(), int(inverse), self.transform_code,
                                            alpha, beta, _lib.stream_ptr()), "spec_transform")
        return out.to(dev)

    def spec_fwd(self, spec):
        if self.transform_type == "none":
            return spec
        return self._transform(spec, False)

    def spec_back(self, spec):
        if self.transform_type == "none":
            return spec
        return self._transform(spec, True)
