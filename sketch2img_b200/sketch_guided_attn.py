"""Drop-in for /root/reference/modules/sketch_guided_attn.py (``SatMixin`` / ``AttnModule``) on the CUDA engine.

Same surface as the reference (sketch_guided_attn.py:8-85): ``SatMixin(unet)`` creates one ``AttnModule`` per
``BasicTransformerBlock`` of the UNet (16 in SD1.x / SD2.x), registered as sub-modules named
``sketch_attn_<module path with '.' -> '_'>`` in ``named_modules`` order (down, up, mid), each holding
``sketch_norm`` (LayerNorm), ``sketch_attn`` (bias-free q/k/v + biased out projection) and ``sketch_conv`` (Conv1d 1x1)
with the reference's parameter names, so a state dict trained with the reference loads unchanged;
``set_res_samples(res_samples)`` distributes the sketch encoder's per-down-block feature tuples exactly like
:29-40 and ``set_scale`` like :42-44.  The arithmetic of the injected block (:120-132: LayerNorm -> cross-attention to
the feature tokens -> 1x1 conv -> scaled residual) runs inside the engine's transformer blocks
(``s2i_unet_load_sat`` / ``s2i_unet_set_sat_feature`` / ``s2i_unet_set_sat_scale``); there is no CPU fallback.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class _SketchCrossAttention(nn.Module):
    """Parameter container with diffusers' CrossAttention names (to_q / to_k / to_v without bias, to_out.0 with)."""

    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])


class AttnModule(nn.Module):
    def __init__(self, sat_name, base_layer):
        """sketch_guided_attn.py:48-79.  ``base_layer``: the transformer block the module is injected into -- here the engine's
        ``BasicTransformerBlock`` handle from ``unet.named_modules()``; width, heads and head width come from its ``attn1``
        exactly like :51-60."""
        super().__init__()
        self.name = sat_name
        attn1 = base_layer.attn1
        dim = attn1.to_q.in_features
        heads = attn1.heads
        dim_head = attn1.to_q.out_features // heads
        self._unet, self._path = base_layer._unet, base_layer._path
        self.sketch_norm = nn.LayerNorm(dim)
        self.sketch_attn = _SketchCrossAttention(dim, heads, dim_head)
        self.sketch_conv = nn.Conv1d(dim, dim, 1)
        self.sketch_scale = 1.0
        self.res_sample = None

    def set_res_sample(self, res_sample):
        """sketch_guided_attn.py:81-82: the feature [b, c, h, w] becomes this block's key/value tokens."""
        self.res_sample = res_sample
        eng = self._unet.engine
        if res_sample is None:
            _lib.check(eng.lib.s2i_unet_set_sat_feature(eng._h, self._path.encode(), None, 0, 0, 0, 0, _lib.stream_ptr()))
            return
        f = res_sample.to(eng.device, torch.float32).contiguous()
        b, c, h, w = f.shape
        _lib.check(eng.lib.s2i_unet_set_sat_feature(eng._h, self._path.encode(), f.data_ptr(), b, c, h, w, _lib.stream_ptr()))

    def set_scale(self, scale):
        self.sketch_scale = scale


class SatMixin(nn.Module):
    def __init__(self, unet):
        super().__init__()
        self._unet = unet
        self.blocks = []
        prefix = "sketch_attn"
        for name, module in unet.named_modules():                        # sketch_guided_attn.py:15-21
            if module.__class__.__name__ == "BasicTransformerBlock":
                module_name = (prefix + "." + name).replace(".", "_")
                self.blocks.append(AttnModule(module_name, module))
        names = set()
        for block in self.blocks:                                        # :23-27
            assert block.name not in names, f"duplicated module name: {block.name}"
            names.add(block.name)
            self.add_module(block.name, block)
        self._pushed = None
        self._scale = None

    # ------------------------------------------------------------------ engine plumbing
    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def sync(self):
        """Upload the weights (again) when parameters changed (load_state_dict, optimiser step, .to ...)."""
        key = self._params_key()
        if key == self._pushed:
            return
        eng = self._unet.engine
        keep, names, ptrs, ndims, shapes = [], [], [], [], []
        for k, v in self.state_dict().items():
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            ndims.append(min(t.dim(), 4))
            shapes += (list(t.shape) + [1] * 4)[:4]
        n = len(names)
        with torch.cuda.device(eng.device):
            _lib.check(eng.lib.s2i_unet_load_sat(eng._h, n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs),
                                                 (C.c_int * n)(*ndims), (C.c_longlong * (4 * n))(*shapes)))
        self._pushed = key
        self._scale = None

    # ------------------------------------------------------------------ reference surface
    def set_res_samples(self, res_samples):
        """sketch_guided_attn.py:29-40."""
        self.sync()
        if self._scale is None:
            self.set_scale(self.blocks[0].sketch_scale)
        down_blocks, up_blocks = (), ()
        mid_block = (res_samples[-1][-1],)
        for mid_layers in res_samples:
            if len(mid_layers) == 3:
                down_blocks += (mid_layers[0], mid_layers[1])
                up_blocks += (mid_layers[0], mid_layers[1], mid_layers[1])
        total_blocks = down_blocks + up_blocks[::-1] + mid_block
        for idx in range(len(self.blocks)):
            self.blocks[idx].set_res_sample(total_blocks[idx])

    def set_scale(self, scale):
        """sketch_guided_attn.py:42-44."""
        self.sync()
        for block in self.blocks:
            block.set_scale(scale)
        eng = self._unet.engine
        _lib.check(eng.lib.s2i_unet_set_sat_scale(eng._h, float(scale), _lib.stream_ptr()))
        self._scale = float(scale)
