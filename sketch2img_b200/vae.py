"""``AutoencoderKL`` stand-in backed by the CUDA engine (libs2i ``s2i_vae_*``): the step either side of the sampling loop.

Keeps what the reference touches on ``vae``:
  * ``vae.encode(img).latent_dist.sample() * 0.18215``  -- /root/reference/app.py:107-109, the sketch target of
    ``AntiGradientPipeline.__call__(sketch_image=...)``;
  * ``vae.decode(latents / 0.18215).sample``            -- /root/reference/modules/pipeline.py:118 (decode_latents) and
    :163-174 (decode_latents_L);
  * ``.device``, ``.dtype``, ``.to(...)``, ``from_pretrained(dir, subfolder="vae")`` for local diffusers directories.
Weights come from a diffusers-named state dict (``encoder.*``, ``decoder.*``, ``quant_conv.*``, ``post_quant_conv.*``).
There is no CPU fallback: without a CUDA device / libs2i.so the constructor raises.
"""
import ctypes as C
from types import SimpleNamespace

import torch

from . import _lib

SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                     latent_channels=4, norm_num_groups=32, sample_size=512)


class DiagonalGaussianDistribution:
    """diffusers' posterior object: ``parameters`` = (mean | logvar) on the channel axis."""

    def __init__(self, parameters):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.parameters.device, dtype=self.parameters.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


class _Out(dict):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.__dict__.update(kw)


class AutoencoderKL:
    def __init__(self, config, state_dict, device=None):
        if not torch.cuda.is_available():
            raise _lib.S2IError("sketch2img_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        if int(cfg.get("norm_num_groups", 32)) != 32 or len(cfg["block_out_channels"]) != 4:
            raise ValueError("the s2i VAE engine supports the 4-level SD AutoencoderKL with 32 GroupNorm groups")
        self.config = SimpleNamespace(**cfg)
        self.lib = _lib.lib()
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.dtype = torch.float32
        c = _lib.VaeConfig()
        c.in_channels, c.out_channels = int(cfg.get("in_channels", 3)), int(cfg.get("out_channels", 3))
        c.latent_channels = int(cfg.get("latent_channels", 4))
        c.layers_per_block = int(cfg.get("layers_per_block", 2))
        for i in range(4):
            c.block_out_channels[i] = int(cfg["block_out_channels"][i])
        self._cfg = c
        self._h = C.c_void_p()
        _lib.check(self.lib.s2i_vae_create(C.byref(c), C.byref(self._h)))
        keep, names, ptrs, ndims, shapes = [], [], [], [], []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point or v.dim() > 4:
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            ndims.append(t.dim())
            shapes += list(t.shape) + [1] * (4 - t.dim())
        n = len(names)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_vae_load(self._h, n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs), (C.c_int * n)(*ndims),
                                             (C.c_longlong * (4 * n))(*shapes)))

    @classmethod
    def from_pretrained(cls, path, subfolder=None, torch_dtype=None, device=None, **kwargs):
        """A diffusers model directory on local disk (``config.json`` + ``diffusion_pytorch_model.safetensors`` / ``.bin``)."""
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

    def __del__(self):
        try:
            if self._h:
                self.lib.s2i_vae_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def encode(self, x, return_dict=True):
        x = x.to(self.device, torch.float32).contiguous()
        B, _, H, W = x.shape
        moments = torch.empty(B, 2 * self._cfg.latent_channels, H // 8, W // 8, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_vae_encode(self._h, x.data_ptr(), B, H, W, moments.data_ptr(), _lib.stream_ptr()))
        dist = DiagonalGaussianDistribution(moments)
        return _Out(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z, return_dict=True):
        z = z.to(self.device, torch.float32).contiguous()
        B, _, h, w = z.shape
        img = torch.empty(B, self._cfg.out_channels, 8 * h, 8 * w, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_vae_decode(self._h, z.data_ptr(), B, h, w, img.data_ptr(), _lib.stream_ptr()))
        return _Out(sample=img) if return_dict else (img,)

    def arena_bytes(self):
        return int(self.lib.s2i_vae_arena_bytes(self._h))
