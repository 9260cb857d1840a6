"""Minimal stand-in for pytorch_lightning (absent in this image); golden generation only."""
import torch
import torch.nn as nn


class LightningModule(nn.Module):
    def save_hyperparameters(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass

    def to(self, *args, **kwargs):
        if not torch.cuda.is_available():
            args = tuple(a for a in args if not (isinstance(a, str) and a.startswith("cuda")))
            if kwargs.get("device", None) is not None and str(kwargs["device"]).startswith("cuda"):
                kwargs.pop("device")
            if not args and not kwargs:
                return self
        return super().to(*args, **kwargs)

    @classmethod
    def load_from_checkpoint(cls, path, map_location="cpu", **overrides):
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(overrides)
        obj = cls(**hp)
        obj.on_load_checkpoint(ckpt)
        obj.load_state_dict(ckpt["state_dict"])
        return obj

    def on_load_checkpoint(self, checkpoint):
        pass


class LightningDataModule:
    def __init__(self, *a, **k):
        pass


class Trainer:
    def __init__(self, *a, **k):
        raise RuntimeError("training is out of scope for the oracle shim")
