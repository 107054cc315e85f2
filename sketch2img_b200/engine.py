"""Host-side handle of the CUDA UNet engine (libs2i `s2i_unet_*`).

Plays the role of ``self.unet`` + ``hook_unet`` in the reference loop
(/root/reference/modules/pipeline.py:96, /root/reference/modules/latent_predictor.py:47-81): one call runs the
SD UNet forward and leaves the 9 LGP feature taps resident; ``backward`` is the UNet part of
``torch.autograd.grad(loss, latents_prev)`` (pipeline.py:159).
"""
import ctypes as C

import torch

from . import _lib


class _DevView:
    """Zero-copy torch view of engine-owned device memory via __cuda_array_interface__."""

    def __init__(self, ptr, shape, strides_elems=None, itemsize=4, typestr="<f4"):
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False), "version": 3,
            "strides": None if strides_elems is None else tuple(int(s) * itemsize for s in strides_elems),
        }


def device_view(ptr, shape, strides_elems=None):
    return torch.as_tensor(_DevView(ptr, shape, strides_elems), device="cuda")


def unet_config_from(cfg):
    """cfg: a diffusers-style ``unet.config`` object or dict."""
    get = (lambda k, d=None: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d=None: getattr(cfg, k, d))
    boc = tuple(get("block_out_channels"))
    heads = get("attention_head_dim")
    heads = tuple(heads) if isinstance(heads, (tuple, list)) else (int(heads),) * len(boc)
    if len(boc) != 4:
        raise ValueError("the s2i engine supports the 4-level SD UNet topology")
    c = _lib.UNetConfig()
    c.in_channels = int(get("in_channels", 4))
    c.out_channels = int(get("out_channels", 4))
    for i in range(4):
        c.block_out_channels[i] = int(boc[i])
        c.num_heads[i] = int(heads[i])
    c.layers_per_block = int(get("layers_per_block", 2))
    c.cross_attention_dim = int(get("cross_attention_dim"))
    c.sample_size = int(get("sample_size", 64))
    c.ctx_len = int(get("ctx_len", 77))
    return c


class UNetEngine:
    def __init__(self, config, state_dict, device=None):
        if not torch.cuda.is_available():
            raise _lib.S2IError("sketch2img_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device or "cuda:%d" % torch.cuda.current_device())
        self.cfg = unet_config_from(config)
        self._h = C.c_void_p()
        _lib.check(self.lib.s2i_unet_create(C.byref(self.cfg), C.byref(self._h)))
        self._load(state_dict)
        self.in_channels = self.cfg.in_channels
        self._last = None

    def _load(self, state_dict):
        keep, names, ptrs, ndims, shapes = [], [], [], [], []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            if t.dim() > 4:
                continue
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            ndims.append(t.dim())
            shapes += list(t.shape) + [1] * (4 - t.dim())
        n = len(names)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_unet_load(
                self._h, n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs), (C.c_int * n)(*ndims),
                (C.c_longlong * (4 * n))(*shapes)))

    def __del__(self):
        try:
            if self._h:
                self.lib.s2i_unet_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def forward(self, x, t, ctx, save_for_backward=False):
        """x [B,4,H,W] fp32 cuda (NCHW), t scalar, ctx [B,77,D] fp32 cuda -> eps [B,4,H,W] fp32."""
        x = x.to(self.device, torch.float32).contiguous()
        ctx = ctx.to(self.device, torch.float32).contiguous()
        B, _, H, W = x.shape
        eps = torch.empty(B, self.cfg.out_channels, H, W, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.s2i_unet_forward(self._h, x.data_ptr(), B, H, W, float(t), ctx.data_ptr(),
                                                 eps.data_ptr(), int(bool(save_for_backward)), _lib.stream_ptr()))
        self._last = (x, ctx)
        return eps

    def tap(self, k):
        """NHWC fp32 view [B,H,W,C] of tap k (hook order of latent_predictor.py:63-80)."""
        p, B, H, W, Cc = C.c_void_p(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.s2i_unet_tap(self._h, k, C.byref(p), C.byref(B), C.byref(H), C.byref(W), C.byref(Cc)))
        ld = C.c_longlong()
        _lib.check(self.lib.s2i_unet_tap_stride(self._h, k, C.byref(ld)))
        return device_view(p.value, (B.value, H.value, W.value, Cc.value),
                           (H.value * W.value * ld.value, W.value * ld.value, ld.value, 1))

    def taps(self):
        return [self.tap(k) for k in range(9)]

    def backward(self, tap_grads, samples=None):
        """tap_grads: 9 NHWC fp32 cuda tensors (or None) -> dx [B,4,H,W] fp32 (NCHW).
        samples = (b0, nb): walk only the samples [b0, b0 + nb) of the forward's batch; tap_grads and dx hold nb samples."""
        x, _ = self._last
        gs = [None if g is None else g.to(self.device, torch.float32).contiguous() for g in tap_grads]
        arr = (C.c_void_p * 9)(*[None if g is None else g.data_ptr() for g in gs])
        with torch.cuda.device(self.device):
            if samples is None:
                dx = torch.empty_like(x)
                _lib.check(self.lib.s2i_unet_backward(self._h, arr, dx.data_ptr(), _lib.stream_ptr()))
            else:
                b0, nb = samples
                dx = torch.empty((nb,) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
                _lib.check(self.lib.s2i_unet_backward_samples(self._h, arr, dx.data_ptr(), int(b0), int(nb), _lib.stream_ptr()))
        return dx

    def set_debug(self, on=True):
        _lib.check(self.lib.s2i_unet_debug(self._h, int(on)))

    def debug_tensor(self, name):
        p, ld = C.c_void_p(), C.c_longlong()
        B, H, W, Cc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(self.lib.s2i_unet_debug_get(self._h, name.encode(), C.byref(p), C.byref(ld), C.byref(B), C.byref(H),
                                               C.byref(W), C.byref(Cc)))
        return device_view(p.value, (B.value, H.value, W.value, Cc.value),
                           (H.value * W.value * ld.value, W.value * ld.value, ld.value, 1))

    def arena_bytes(self):
        return int(self.lib.s2i_unet_arena_bytes(self._h))
