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


class _LinearShape:
    def __init__(self, in_features, out_features):
        self.in_features, self.out_features = in_features, out_features


class _AttentionShape:
    """What sketch_guided_attn.py:51-60 reads off ``base_layer.attn1``."""

    def __init__(self, dim, heads, upcast_attention):
        self.to_q = _LinearShape(dim, dim)
        self.heads = heads
        self.upcast_attention = upcast_attention


class BasicTransformerBlock:
    """Handle of one transformer block of the engine: the object ``unet.named_modules()`` yields where diffusers yields its
    ``BasicTransformerBlock`` (the reference's SatMixin selects modules by that class name, sketch_guided_attn.py:15-16, and
    AttnModule reads ``attn1.to_q.in_features / out_features``, ``attn1.heads``, ``attn1.upcast_attention``)."""

    def __init__(self, unet, block_path, dim, heads, upcast_attention):
        self._unet, self._path = unet, block_path
        self.attn1 = _AttentionShape(dim, heads, upcast_attention)


def transformer_block_paths(cfg):
    """Module paths of the UNet's Transformer2DModels in diffusers' ``named_modules`` order: down_blocks, up_blocks,
    mid_block (registration order: SURVEY A.1) -- what the reference's ``down + up[::-1] + mid`` assignment relies on."""
    layers = int(cfg.get("layers_per_block", 2))
    paths = [f"down_blocks.{i}.attentions.{j}" for i in range(3) for j in range(layers)]
    paths += [f"up_blocks.{i}.attentions.{j}" for i in range(1, 4) for j in range(layers + 1)]
    paths.append("mid_block.attentions.0")
    return paths


def transformer_block_handles(owner, cfg):
    """[(module name, BasicTransformerBlock handle)] in ``named_modules`` order for the UNet-like ``owner`` (needs ``.engine``)."""
    boc = list(cfg["block_out_channels"])
    heads = cfg.get("attention_head_dim", 8)
    heads = list(heads) if isinstance(heads, (tuple, list)) else [int(heads)] * len(boc)
    out = []
    for path in transformer_block_paths(cfg):
        lvl = int(path.split(".")[1]) if path.startswith("down") else 3 - int(path.split(".")[1]) if path.startswith("up") else 3
        out.append((path + ".transformer_blocks.0",
                    BasicTransformerBlock(owner, path, boc[lvl], heads[lvl], bool(cfg.get("upcast_attention", False)))))
    return out


class UNet2DConditionModel:
    def __init__(self, config, state_dict, device=None):
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        self.config = SimpleNamespace(**cfg)
        self.engine = UNetEngine(cfg, state_dict, device=device)
        self.in_channels = int(cfg.get("in_channels", 4))
        self.device = self.engine.device
        self.dtype = torch.float32      # public tensors are fp32; GEMM operands are fp16 with fp32 accumulation
        self._blocks = transformer_block_handles(self, cfg)

    @classmethod
    def from_state_dict(cls, config, state_dict, device=None):
        return cls(config, state_dict, device)

    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, device=None, **kwargs):
        """A diffusers model directory on local disk: ``config.json`` + ``diffusion_pytorch_model.safetensors`` (or ``.bin``),
        i.e. what ``from_pretrained`` of the reference's pipeline reads for ``pipe.unet`` (app.py:32-38).  fp16 checkpoints
        (``torch_dtype=torch.float16``) are fine: weights are packed to fp16 GEMM operands either way."""
        import json
        import os
        root = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(root, "config.json")) as f:
            cfg = json.load(f)
        st = os.path.join(root, "diffusion_pytorch_model.safetensors")
        if os.path.exists(st):
            from safetensors.torch import load_file
            sd = load_file(st)
        else:
            sd = torch.load(os.path.join(root, "diffusion_pytorch_model.bin"), map_location="cpu")
        return cls(cfg, sd, device=device)

    def named_modules(self):
        """(name, handle) of the transformer blocks, diffusers order -- enough for the reference's SatMixin loop."""
        yield "", self
        for name, blk in self._blocks:
            yield name, blk

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        """app.py:43: attention is always the fused tcgen05 kernel here."""
        return self

    def to(self, *args, **kwargs):
        return self

    def half(self):
        return self

    def eval(self):
        return self

    def __call__(self, sample, timestep, encoder_hidden_states, return_dict=True, save_for_backward=False):
        t = float(timestep.item()) if torch.is_tensor(timestep) else float(timestep)
        out = self.engine.forward(sample, t, encoder_hidden_states, save_for_backward=save_for_backward)
        return UNetOutput(out) if return_dict else (out,)
