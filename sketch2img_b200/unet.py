"""``UNet2DConditionModel`` stand-in backed by the CUDA engine.

Keeps what the reference touches on ``pipe.unet``: ``unet(sample, t, encoder_hidden_states=...).sample``
(/root/reference/modules/pipeline.py:96), ``.config.sample_size`` (:40), ``.in_channels`` (:64), ``.device``
(:38), ``.dtype``.  Weights come from a diffusers-named state dict (``from_state_dict``) and are packed once for
the tcgen05 kernels.
"""
from types import SimpleNamespace

import torch

from .engine import UNetEngine

SD15_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=8, cross_attention_dim=768, use_linear_projection=False, upcast_attention=False)
SD21_CONFIG = dict(
    sample_size=96, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True, upcast_attention=True)


class UNetOutput(dict):
    def __init__(self, sample):
        super().__init__(sample=sample)
        self.sample = sample


class UNet2DConditionModel:
    def __init__(self, config, state_dict, device=None):
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        self.config = SimpleNamespace(**cfg)
        self.engine = UNetEngine(cfg, state_dict, device=device)
        self.in_channels = int(cfg.get("in_channels", 4))
        self.device = self.engine.device
        self.dtype = torch.float32      # public tensors are fp32; GEMM operands are fp16 with fp32 accumulation

    @classmethod
    def from_state_dict(cls, config, state_dict, device=None):
        return cls(config, state_dict, device)

    def to(self, *args, **kwargs):
        return self

    def __call__(self, sample, timestep, encoder_hidden_states, return_dict=True, save_for_backward=False):
        t = float(timestep.item()) if torch.is_tensor(timestep) else float(timestep)
        out = self.engine.forward(sample, t, encoder_hidden_states, save_for_backward=save_for_backward)
        return UNetOutput(out) if return_dict else (out,)
