"""`SNRModel` (inference half; mirror of sgmse-bbed/sgmse/snr_estimator.py:19-79,143-146)."""
from ._checkpoint import CheckpointedModule
from .backbones.snrnet import SNRNet
from .data_module import SpecsDataModule


class SNRModel(CheckpointedModule):
    def __init__(self, backbone="snrnet", lr=1e-4, ema_decay=0.999, num_eval_files=10, loss_type='mse',
                 data_module_cls=None, **kwargs):
        self.dnn = SNRNet()
        self.lr, self.ema_decay, self.loss_type, self.num_eval_files = lr, ema_decay, loss_type, num_eval_files
        self._init_ckpt(ema_decay)
        data_module_cls = data_module_cls or SpecsDataModule
        self.data_module = data_module_cls(**kwargs, gpu=kwargs.get('gpus', 0) > 0)

    def _dnn_names(self):
        return list(self.dnn.state_dict().keys())

    def forward(self, y):
        return self.dnn(y)

    __call__ = forward

    def calculate_normfac_direct(self, s, n, fixed_snr):
        return (2.040166) * (0.240253 + 0.759747 * fixed_snr ** 2) ** 0.5 / ((1 + (n / s) ** 2) ** 0.5)
