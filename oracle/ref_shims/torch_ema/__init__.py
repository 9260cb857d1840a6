"""Minimal stand-in for torch_ema==0.3 ExponentialMovingAverage; golden generation only."""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params = [p.clone().detach() for p in parameters if p.requires_grad]
        self.collected_params = None

    def update(self, parameters):
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        with torch.no_grad():
            ps = [p for p in parameters if p.requires_grad]
            for s, p in zip(self.shadow_params, ps):
                s.sub_((1.0 - decay) * (s - p))

    def copy_to(self, parameters):
        ps = [p for p in parameters if p.requires_grad]
        for s, p in zip(self.shadow_params, ps):
            p.data.copy_(s.data)

    def store(self, parameters):
        self.collected_params = [p.clone() for p in parameters if p.requires_grad]

    def restore(self, parameters):
        ps = [p for p in parameters if p.requires_grad]
        for c, p in zip(self.collected_params, ps):
            p.data.copy_(c.data)

    def to(self, *args, **kwargs):
        if not torch.cuda.is_available():
            return
        self.shadow_params = [p.to(*args, **kwargs) for p in self.shadow_params]

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates,
                    shadow_params=self.shadow_params, collected_params=self.collected_params)

    def load_state_dict(self, sd):
        self.decay = sd["decay"]
        self.num_updates = sd["num_updates"]
        self.shadow_params = [p.clone() for p in sd["shadow_params"]]
        self.collected_params = sd.get("collected_params", None)
