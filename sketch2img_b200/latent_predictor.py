"""Drop-in for /root/reference/modules/latent_predictor.py on the CUDA engine.

``LatentEdgePredictor`` keeps the reference constructor, parameter names and state-dict keys
(latent_predictor.py:9-35) so ``edge_predictor.pt`` loads unchanged; its arithmetic runs in libs2i
(``s2i_lgp_*``).  ``hook_unet`` returns the same 9 tap handles, in the same order, each exposing ``.output``
(latent_predictor.py:47-81) -- here as views of the engine-resident taps instead of forward hooks.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

_HIDDEN = (512, 256, 128, 64)


class LatentEdgePredictor(nn.Module):
    def __init__(self, input_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.input_dim, self.output_dim = input_dim, output_dim
        mods, prev = [], input_dim
        for w in _HIDDEN:
            mods += [nn.Linear(prev, w), nn.ReLU(), nn.BatchNorm1d(num_features=w)]
            prev = w
        mods.append(nn.Linear(prev, output_dim))
        self.layers = nn.Sequential(*mods)
        for m in self.layers:          # latent_predictor.py:32-35
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        self._engine = None
        self._engine_key = None

    # ------------------------------------------------------------------ engine plumbing
    def _params_key(self):
        return tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def engine(self):
        """(Re)build the device engine when parameters changed (load_state_dict, .to, .half ...)."""
        key = self._params_key()
        if self._engine is None or key != self._engine_key:
            self._engine = LGPEngine(self.input_dim, self.output_dim, self.num_layers, self.state_dict())
            self._engine_key = key
        return self._engine

    def pull_from_engine(self):
        """After ``trainer.training_step``: copy the engine's updated fp32 masters back into this module's parameters (so
        ``state_dict()`` / ``torch.save`` see them) without triggering a re-upload."""
        eng = self._engine
        if eng is None:
            return self
        with torch.no_grad():
            for name, p in self.named_parameters():
                p.copy_(eng.get_param(name, tuple(p.shape)).to(p.device, p.dtype))
        self._engine_key = self._params_key()
        return self

    def forward(self, x, t):
        """x [b, input_dim-4-4P, h, w] (resized, concatenated taps), t [b,4,h,w] -> fp16 [(b w h), output_dim]
        (latent_predictor.py:37-45).  b must hold (uncond, cond) pairs: BatchNorm statistics are per pair
         in train mode, the running ones in eval mode."""
        if not x.is_cuda:
            raise _lib.S2IError("LatentEdgePredictor.forward runs on the CUDA engine only (no CPU fallback)")
        if x.shape[2] != x.shape[3]:
            raise RuntimeError("LatentEdgePredictor engine expects square latents")
        eng = self.engine()
        b, _, h, _ = x.shape
        out = eng.forward_nchw(x.float().contiguous(), t.float().contiguous(), b, h, self.training)
        return out.to(torch.float16)


class LGPEngine:
    def __init__(self, input_dim, output_dim, num_layers, state_dict):
        self.lib = _lib.lib()
        self.input_dim, self.output_dim = input_dim, output_dim
        self._h = C.c_void_p()
        _lib.check(self.lib.s2i_lgp_create(input_dim, output_dim, num_layers, C.byref(self._h)))
        keep, names, ptrs, ndims, shapes = [], [], [], [], []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            names.append(k.encode())
            ptrs.append(t.data_ptr())
            ndims.append(t.dim())
            shapes += list(t.shape) + [1] * (4 - t.dim())
        n = len(names)
        _lib.check(self.lib.s2i_lgp_load(self._h, n, (C.c_char_p * n)(*names), (C.c_void_p * n)(*ptrs),
                                         (C.c_int * n)(*ndims), (C.c_longlong * (4 * n))(*shapes)))

    def __del__(self):
        try:
            if self._h:
                self.lib.s2i_lgp_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_grad_rounding(self, emulate_fp16=True):
        """True (default): mimic the reference's unscaled fp16 autograd rounding; False: loss-scaled gradients."""
        _lib.check(self.lib.s2i_lgp_set_grad_rounding(self._h, int(bool(emulate_fp16))))

    def forward_nchw(self, x, t, B, L, train):
        _lib.check(self.lib.s2i_lgp_forward_nchw(self._h, x.data_ptr(), t.data_ptr(), B, L, int(train), _lib.stream_ptr()))
        return self.output(B, L, x.device)

    def forward_taps(self, taps, B, L, noise, sigma, train):
        """taps: 9 NHWC fp32 cuda tensors; noise NCHW [B/2,4,L,L]."""
        taps = [t.contiguous() for t in taps]      # engine taps may be strided slices of the UNet's concat buffers
        self._taps_keepalive = taps
        ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in taps])
        sizes = (C.c_int * 9)(*[t.shape[1] for t in taps])
        chans = (C.c_int * 9)(*[t.shape[3] for t in taps])
        _lib.check(self.lib.s2i_lgp_forward_taps(self._h, ptrs, sizes, chans, B, L, noise.data_ptr(), float(sigma),
                                                 int(train), _lib.stream_ptr()))

    def forward_taps_batch(self, taps, B, L, noise_level):
        """LatentEdgePredictor.forward as the trainer calls it (trainer.py:245): B latents, BatchNorm statistics over all rows;
        taps: 9 NHWC fp32 cuda tensors [B, S, S, C]; noise_level NCHW [B, 4, L, L]."""
        taps = [t.contiguous() for t in taps]
        self._taps_keepalive = taps
        ptrs = (C.c_void_p * 9)(*[t.data_ptr() for t in taps])
        sizes = (C.c_int * 9)(*[t.shape[1] for t in taps])
        chans = (C.c_int * 9)(*[t.shape[3] for t in taps])
        _lib.check(self.lib.s2i_lgp_forward_taps_batch(self._h, ptrs, sizes, chans, B, L, noise_level.data_ptr(), _lib.stream_ptr()))

    def train_step(self, target, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, step=1):
        """One AdamW step on the batch of the last ``forward_taps_batch``; returns the loss (python float)."""
        loss = torch.empty(1, device=target.device, dtype=torch.float32)
        _lib.check(self.lib.s2i_lgp_train_step(self._h, target.data_ptr(), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                               float(weight_decay), int(step), loss.data_ptr(), _lib.stream_ptr()))
        return loss.item()

    def get_param(self, name, shape):
        out = torch.empty(shape, dtype=torch.float32)
        _lib.check(self.lib.s2i_lgp_get_param(self._h, name.encode(), out.data_ptr(), out.numel()))
        return out

    def output(self, B, L, device):
        out = torch.empty(B * L * L, self.output_dim, device=device, dtype=torch.float32)
        _lib.check(self.lib.s2i_lgp_output(self._h, out.data_ptr(), _lib.stream_ptr()))
        return out

    def loss_backward(self, target, taps, cond_only=False):
        """-> (loss [B/2], tap grads (9 NHWC fp32 tensors, scaled), grad_scale).  cond_only: the gradients of the B/2 cond
        samples only (batch entries 1, 3, ...), which is all pipeline.py:159 keeps."""
        S = target.shape[0]
        grads = [torch.empty((t.shape[0] // 2 if cond_only else t.shape[0],) + tuple(t.shape[1:]), device=t.device,
                             dtype=torch.float32) for t in taps]
        loss = torch.empty(S, device=target.device, dtype=torch.float32)
        scale = C.c_float()
        fn = self.lib.s2i_lgp_loss_backward_cond if cond_only else self.lib.s2i_lgp_loss_backward
        _lib.check(fn(self._h, target.data_ptr(), (C.c_void_p * 9)(*[g.data_ptr() for g in grads]),
                      loss.data_ptr(), C.byref(scale), _lib.stream_ptr()))
        return loss, grads, scale.value


class _Tap:
    """One tapped UNet sub-module (latent_predictor.py:65-79): ``.output`` is the fp32 NCHW feature of the last
    forward, like the reference's hook-stored attribute."""

    def __init__(self, unet, index, name):
        self._unet, self.index, self.name = unet, index, name

    @property
    def output(self):
        return self._unet.engine.tap(self.index).permute(0, 3, 1, 2)

    def __repr__(self):
        return f"<s2i tap {self.index}: {self.name}>"


TAP_NAMES = ("down_blocks.0", "down_blocks.1", "down_blocks.2", "mid_block.attentions.0", "mid_block.resnets.0",
             "mid_block.resnets.1", "up_blocks.0", "up_blocks.1", "up_blocks.2")


def hook_unet(unet):
    """Same return contract as latent_predictor.py:47-81: the 9 feature blocks in hook order."""
    return [_Tap(unet, i, n) for i, n in enumerate(TAP_NAMES)]
