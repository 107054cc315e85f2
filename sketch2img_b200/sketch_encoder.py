"""Drop-in for /root/reference/modules/sketch_encoder.py (``SketchEncoder``) on the CUDA engine.

The reference's class is a ``UNet2DConditionModel`` whose ``forward`` stops after the down blocks and returns
``UNet2DConditionOutput(sample=down_block_res_samples)`` -- one tuple of feature maps per down block (its resnet outputs,
then the downsampled map for all but the last block) -- which is exactly what ``SatMixin.set_res_samples`` consumes
(sketch_guided_attn.py:29-40).  Its forward calls every down block WITHOUT ``encoder_hidden_states`` (sketch_encoder.py:93-95),
which diffusers only executes for attention-free blocks, so the encoder is built with ``down_block_types =
("DownBlock2D",) * 4``: conv_in, the sinusoidal time embedding + 2 Linear, and 4 x (2 ResnetBlock2D [+ stride-2 conv]).
All of it runs in libs2i (``s2i_sketch_encoder_*``: the UNet engine's conv3x3 / GroupNorm kernels); no CPU fallback.
"""
import ctypes as C

import torch

from . import _lib
from .engine import UNetEngine, device_view, unet_config_from
from .unet import UNetOutput


class _EncoderEngine(UNetEngine):
    def __init__(self, config, state_dict, device=None):
        if not torch.cuda.is_available():
            raise _lib.S2IError("sketch2img_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.cfg = unet_config_from(config)
        self._h = C.c_void_p()
        _lib.check(self.lib.s2i_sketch_encoder_create(C.byref(self.cfg), C.byref(self._h)))
        self._load({k: v for k, v in state_dict.items()
                    if k.startswith(("conv_in.", "time_embedding.", "down_blocks."))})
        self.in_channels = self.cfg.in_channels
        self._last = None

    def encode(self, x, t):
        x = x.to(self.device, torch.float32).contiguous()
        B, _, H, W = x.shape
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_sketch_encoder_forward(self._h, x.data_ptr(), B, H, W, float(t), _lib.stream_ptr()))
        n = self.lib.s2i_sketch_encoder_num_res_samples(self._h)
        maps = []
        for k in range(n):
            p, ld = C.c_void_p(), C.c_longlong()
            b, h, w, c = C.c_int(), C.c_int(), C.c_int(), C.c_int()
            _lib.check(self.lib.s2i_sketch_encoder_res_sample(self._h, k, C.byref(p), C.byref(ld), C.byref(b), C.byref(h),
                                                              C.byref(w), C.byref(c)))
            v = device_view(p.value, (b.value, h.value, w.value, c.value),
                            (h.value * w.value * ld.value, w.value * ld.value, ld.value, 1))
            maps.append(v.permute(0, 3, 1, 2).contiguous())        # NCHW copies: the views die with the next forward
        return maps


class SketchEncoder:
    def __init__(self, config, state_dict, device=None):
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        types = cfg.get("down_block_types")
        if types is not None and any(t != "DownBlock2D" for t in types):
            raise ValueError("SketchEncoder.forward calls its down blocks without encoder_hidden_states (sketch_encoder.py:93-95): "
                             "only attention-free DownBlock2D blocks can run; got %r" % (types,))
        from types import SimpleNamespace
        self.config = SimpleNamespace(**cfg)
        self.engine = _EncoderEngine(cfg, state_dict, device=device)
        self.device = self.engine.device
        self.dtype = torch.float32
        self.layers = int(cfg.get("layers_per_block", 2))

    def to(self, *a, **k):
        return self

    def forward(self, sample, timestep, encoder_hidden_states=None, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None):
        """sketch_encoder.py:13-98: ``.sample`` = [tuple of this down block's res_samples, ...] (NCHW fp32 tensors)."""
        t = float(timestep.item()) if torch.is_tensor(timestep) else float(timestep)
        maps = self.engine.encode(sample, t)
        out, k = [], 0
        for i in range(4):
            n = self.layers + (1 if i < 3 else 0)
            out.append(tuple(maps[k:k + n]))
            k += n
        return UNetOutput(out)

    __call__ = forward
