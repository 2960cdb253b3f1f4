"""ModelMixin (import sites: unet_1d_condition.py:22, transformer_1d.py:10). `.dtype` is read at
unet_1d_condition.py:627."""
import torch


class ModelMixin(torch.nn.Module):
    _supports_gradient_checkpointing = False

    @property
    def dtype(self):
        for p in self.parameters():
            if p.is_floating_point():
                return p.dtype
        return torch.float32

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")
