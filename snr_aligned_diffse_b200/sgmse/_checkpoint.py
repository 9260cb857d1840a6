"""Lightning-format checkpoint plumbing shared by ScoreModel and SNRModel (no Lightning dependency).

Layout read: `state_dict` (reference names, `dnn.` prefix), `hyper_parameters`
(constructor kwargs, contains a pickled class reference `sgmse.data_module.SpecsDataModule`),
`ema` (torch-ema 0.3: decay, num_updates, shadow_params = list over the requires_grad parameters in
`parameters()` order, collected_params).
"""
import warnings

import torch

from . import install_alias


def load_checkpoint_file(path, map_location="cpu"):
    """torch.load of a Lightning checkpoint whose pickled class references (`sgmse.data_module.SpecsDataModule` in
    `hyper_parameters`, model.py:93) must resolve.  The `sgmse.*` names are pointed at this package only for the duration
    of the load and only where no other `sgmse` (e.g. the real reference, imported by the golden tooling in the same
    process) already owns them; whatever was in sys.modules before is restored afterwards.
    weights_only=False is required by that pickled class reference: load checkpoints from trusted sources only."""
    import sys
    before = {k: v for k, v in sys.modules.items() if k == "sgmse" or k.startswith("sgmse.")}
    foreign = {k: v for k, v in before.items() if not getattr(v, "__name__", "").startswith(__package__)}
    try:
        install_alias()
        return torch.load(path, map_location=map_location, weights_only=False)
    finally:
        if foreign:     # another package owns `sgmse`: put its modules back, drop the aliases we added
            for k in [k for k in sys.modules if k == "sgmse" or k.startswith("sgmse.")]:
                if k not in before:
                    del sys.modules[k]
            sys.modules.update(before)


class EMAState:
    """Just enough of torch_ema.ExponentialMovingAverage for inference: the shadow parameters."""

    def __init__(self, decay):
        self.decay = decay
        self.num_updates = 0
        self.shadow_params = None
        self.collected_params = None

    def load_state_dict(self, sd):
        self.decay = sd.get("decay", self.decay)
        self.num_updates = sd.get("num_updates", 0)
        self.shadow_params = [p.detach().to("cpu", torch.float32) for p in sd["shadow_params"]]
        self.collected_params = sd.get("collected_params", None)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params,
                    collected_params=self.collected_params)

    def to(self, *a, **k):
        return self


class CheckpointedModule:
    """Mixin: state-dict / EMA / eval-train swapping with the reference's semantics
    (model.py:109-134, snr_estimator.py:54-79)."""

    frozen_params = ()   # state-dict names that are not requires_grad (absent from the EMA list)

    def _init_ckpt(self, ema_decay):
        self.ema = EMAState(ema_decay)
        self._error_loading_ema = False
        self._raw_sd = None       # the optimiser's weights ("collected" when EMA weights are live)
        self._ema_live = False
        self.training = True

    def _dnn_names(self):
        raise NotImplementedError

    def load_state_dict(self, sd, strict=True):
        sub = {k[len("dnn."):]: v for k, v in sd.items() if k.startswith("dnn.")}
        self.dnn.load_state_dict(sub, strict=strict)
        self._raw_sd = self.dnn.state_dict()
        self._ema_live = False

    def state_dict(self):
        return {"dnn." + k: v for k, v in self.dnn.state_dict().items()}

    def on_load_checkpoint(self, checkpoint):
        ema = checkpoint.get('ema', None)
        if ema is not None:
            self.ema.load_state_dict(ema)
        else:
            self._error_loading_ema = True
            warnings.warn("EMA state_dict not found in checkpoint!")

    def on_save_checkpoint(self, checkpoint):
        checkpoint['ema'] = self.ema.state_dict()

    def _ema_state_dict(self):
        names = [n for n in self._dnn_names() if n not in self.frozen_params]
        sh = self.ema.shadow_params
        if sh is None or len(sh) != len(names):
            raise RuntimeError(f"EMA holds {0 if sh is None else len(sh)} tensors, model has {len(names)} trainable ones")
        sd = dict(self._raw_sd)
        for n, p in zip(names, sh):
            if tuple(p.shape) != tuple(sd[n].shape):
                raise RuntimeError(f"EMA tensor for {n} has shape {tuple(p.shape)}, expected {tuple(sd[n].shape)}")
            sd[n] = p
        return sd

    def train(self, mode=True, no_ema=False):
        self.training = bool(mode)
        if not self._error_loading_ema and self.ema.shadow_params is not None and self._raw_sd is not None:
            if mode is False and not no_ema:
                if not self._ema_live:
                    self.dnn.load_state_dict(self._ema_state_dict())   # ema.store + ema.copy_to
                    self._ema_live = True
            elif self._ema_live:
                self.dnn.load_state_dict(self._raw_sd)                  # ema.restore
                self._ema_live = False
        return self

    def eval(self, no_ema=False):
        return self.train(False, no_ema=no_ema)

    # device management.  The reference's eval.py calls `model.cpu()` (eval.py:101) and then runs NCSN++ on the host;
    # this package has no CPU path: weights and compute always live on the B200.  `.cpu()` / `.to("cpu")` are accepted
    # so that call sites keep working, do NOT move anything, and say so once; `forward` / `enhance` take host or device
    # tensors and return results on the device the inputs came from.
    def to(self, *args, **kwargs):
        dev = kwargs.get("device", args[0] if args else None)
        if isinstance(dev, (str, torch.device)) and torch.device(dev).type == "cpu":
            return self.cpu()
        return self

    def cpu(self):
        if not getattr(self, "_warned_cpu", False):
            warnings.warn("snr_aligned_diffse_b200: .cpu() / .to('cpu') is a no-op -- there is no CPU path; the network "
                          "keeps running on the B200 (the reference would run NCSN++ on the host here, eval.py:101)")
            self._warned_cpu = True
        return self

    def cuda(self, device=None):
        return self

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        ckpt = load_checkpoint_file(checkpoint_path)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        model = cls(**hp)
        model.on_load_checkpoint(ckpt)
        model.load_state_dict(ckpt["state_dict"])
        return model
