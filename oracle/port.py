"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by sketch2img_b200/ (the shipped CUDA path).

Plain-PyTorch CPU restatement ("port") of the reference's sketch-guided sampling loop, written so it
runs WITHOUT /root/reference (which does not exist on the GPU box).  Each function cites the
reference lines it follows.  It is pinned in THIS container by tests/test_oracle_vs_reference.py,
which imports the reference's own modules/pipeline.py + modules/latent_predictor.py unmodified over
oracle/diffusers_shim and requires bit-identical latents, and by the committed fixtures under
tests/golden/ (made by oracle/make_golden.py from the unmodified reference files).

Parity status: the reference ships no tests / golden vectors of its own, and the
arithmetic of diffusers' UNet/DDIM lives in an un-vendored dependency restated in
oracle/diffusers_shim from SURVEY.md Appendix A.  => pinned against the reference's own Python files
run here; "parity unpinned" with respect to genuine diffusers + real checkpoints.
"""
import inspect
import math
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "diffusers_shim")


def add_shim_to_path():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)


add_shim_to_path()
from diffusers import DDIMScheduler, DPMSolverMultistepScheduler, UNet2DConditionModel  # noqa: E402  (the shim)
from diffusers.models.unet_2d_condition import SD15_CONFIG, SD21_CONFIG, TINY_CONFIG  # noqa: E402,F401

LGP_HIDDEN = (512, 256, 128, 64)
NUM_POS_LAYERS = 9


# --------------------------------------------------------------------------------------------------
# Latent guidance predictor  (/root/reference/modules/latent_predictor.py:9-45)
# --------------------------------------------------------------------------------------------------
class LatentEdgePredictorOracle(nn.Module):
    """Linear->ReLU->BatchNorm1d x4 -> Linear; state-dict keys ``layers.{0,3,6,9,12}`` (Linear) and
    ``layers.{2,5,8,11}`` (BN) as in latent_predictor.py:15-29; kaiming-uniform weights, zero bias
    (:32-35).  BatchNorm stays in TRAIN mode at inference in the reference."""

    def __init__(self, input_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        mods, prev = [], input_dim
        for width in LGP_HIDDEN:
            mods += [nn.Linear(prev, width), nn.ReLU(), nn.BatchNorm1d(width)]
            prev = width
        mods.append(nn.Linear(prev, output_dim))
        self.layers = nn.Sequential(*mods)
        for m in self.layers:
            if isinstance(m, nn.Linear):
                nn.init.kaiming_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x, t):
        # latent_predictor.py:39-45 -- sin(2*pi*t*2^-l) for l in range(num_layers), concat on C,
        # rows ordered (b, w, h), hard cast to fp16.
        pos = torch.cat([torch.sin(2 * math.pi * t * (2 ** -l)) for l in range(self.num_layers)], dim=1)
        z = torch.cat((x, t, pos), dim=1)
        b, c, h, w = z.shape
        z = z.permute(0, 3, 2, 1).reshape(b * w * h, c).to(torch.float16)
        return self.layers(z)


class LatentEdgePredictorOracle32(nn.Module):
    """NOT the reference: the same MLP evaluated in fp32 end to end (no ``.to(float16)`` cast, fp32 weights copied
    from an fp16 oracle LGP).  Used by the tests to separate implementation error from the reference's own fp16
    rounding noise (its unscaled fp16 gradients sit in the subnormal range)."""

    def __init__(self, lgp16):
        super().__init__()
        import copy
        self.num_layers = lgp16.num_layers
        self.layers = copy.deepcopy(lgp16.layers).float()

    def forward(self, x, t):
        pos = torch.cat([torch.sin(2 * math.pi * t * (2 ** -l)) for l in range(self.num_layers)], dim=1)
        z = torch.cat((x, t, pos), dim=1)
        b, c, h, w = z.shape
        return self.layers(z.permute(0, 3, 2, 1).reshape(b * w * h, c))


def tap_modules(unet):
    """The 9 tapped sub-modules in hook order (latent_predictor.py:63-80): down_blocks[0..2],
    mid attentions then mid resnets, up_blocks[0..2]."""
    mods = [blk for i, blk in enumerate(unet.down_blocks) if i in (0, 1, 2)]
    mods += list(unet.mid_block.attentions) + list(unet.mid_block.resnets)
    mods += [blk for i, blk in enumerate(unet.up_blocks) if i in (0, 1, 2)]
    return mods


def register_taps(unet):
    """Forward hooks storing ``module.output = feature.float()`` (latent_predictor.py:50-62):
    tuple -> [0]; dict-like (Transformer2DModelOutput) -> .sample; tensor as is."""
    def _store(module, _inp, out):
        if isinstance(out, tuple):
            out = out[0]
        if isinstance(out, dict):
            out = out.sample
        module.output = out.float()

    mods = tap_modules(unet)
    handles = [m.register_forward_hook(_store) for m in mods]
    return mods, handles


def lgp_input_dim(unet):
    boc = unet.config.block_out_channels
    # taps: down0,down1,down2 | mid x3 | up0,up1,up2  
    ch = [boc[0], boc[1], boc[2], boc[3], boc[3], boc[3], boc[3], boc[2], boc[1]]
    return sum(ch) + 4 + 4 * NUM_POS_LAYERS


# --------------------------------------------------------------------------------------------------
# Guidance update  (/root/reference/modules/pipeline.py:132-161)
# --------------------------------------------------------------------------------------------------
def noise_level(scheduler, noise, t):
    """pipeline.py:132-139: sqrt(1 - alpha_bar_t) * noise, factor broadcast as fp32 [1,1,1,1]."""
    s = ((1 - scheduler.alphas_cumprod[t]) ** 0.5).flatten()
    while s.dim() < noise.dim():
        s = s.unsqueeze(-1)
    return s.to(noise.device) * noise


def lgp_features(taps, size):
    """pipeline.py:145-151: bilinear (align_corners=False) resize of each tap to size x size, concat on C."""
    return torch.cat([F.interpolate(m.output, size=size, mode="bilinear") for m in taps], dim=1)


def anti_gradient(lgp, scheduler, taps, x_in, latents, noise, t, target, beta, record=None):
    """pipeline.py:141-161.  x_in is the CFG-doubled, grad-enabled UNet input [2,4,h,w]; the loss is the
    MSE between the sketch target and the LGP prediction on the cond half; the step is
    alpha = ||x_in - latents||_F / ||g_cond||_F * beta along g_cond = -dLoss/dx_in (cond half).
    record (tests / fixtures only): dict that receives the loss, the cond gradient and alpha of this call."""
    if target is None:
        return latents
    feats = lgp_features(taps, latents.shape[2])
    for m in taps:
        del m.output
    lvl = noise_level(scheduler, noise, t)
    out = lgp(feats, torch.cat([lvl] * 2))
    b, _, h, w = x_in.shape
    out = out.reshape(b, w, h, -1).permute(0, 3, 2, 1)          # "(b w h) c -> b c h w"
    cond = out.chunk(2)[1]
    loss = F.mse_loss(target.float(), cond.float(), reduction="mean")
    g = (-torch.autograd.grad(loss, x_in)[0]).chunk(2)[1]
    alpha = torch.linalg.norm(x_in - latents) / torch.linalg.norm(g) * beta
    if record is not None:
        record.update(loss=float(loss), g=g.detach().clone(), alpha=float(alpha))
    return latents + alpha * g


def guided_step(unet, lgp, scheduler, text_emb, latents, noise, t, target, guided, taps, guidance_scale=7.5, beta=1.6,
                record=None):
    """ONE iteration of the loop body pipeline.py:83-110 from explicit state (``scheduler.set_timesteps`` already called).
    record (tests / fixtures only): receives the pre-guidance scheduler output ``x_ddim`` and anti_gradient's record."""
    with torch.no_grad():
        x_in = torch.cat([latents] * 2)                                     # :85
        x_in = scheduler.scale_model_input(x_in, t).requires_grad_(True)    # :86-87
        with torch.enable_grad() if guided else torch.no_grad():
            eps = unet(x_in, t, encoder_hidden_states=text_emb).sample      # :96
        eps_u, eps_c = eps.chunk(2)                                         # :100
        eps = eps_u + guidance_scale * (eps_c - eps_u)                      # :101
        extra = {"eta": 0.0} if "eta" in inspect.signature(scheduler.step).parameters else {}           # :78
        latents = scheduler.step(eps, t, latents, **extra).prev_sample      # :104
        if record is not None:
            record["x_ddim"] = latents.detach().clone()
        if guided:
            with torch.enable_grad():
                latents = anti_gradient(lgp, scheduler, taps, x_in, latents, noise, t, target, beta, record)  # :109
    return latents


@torch.no_grad()
def guided_sample(unet, lgp, scheduler, text_emb, latents, target, num_steps=50, guidance_scale=7.5,
                  beta=1.6, stop_frac=0.5, taps=None, callback=None):
    """pipeline.py:59-115, CFG on, batch 1.  text_emb is [2,77,D] = [uncond, cond].
    Returns the final latent; ``callback(i, t, latents)`` fires every step like :112-115."""
    if taps is None:
        taps, _ = register_taps(unet)
    scheduler.set_timesteps(num_steps)
    ts = scheduler.timesteps
    latents = latents * scheduler.init_noise_sigma
    noise = latents.detach().clone()                                        # :75
    stop = stop_frac * len(ts)                                              # :90
    for i, t in enumerate(ts):
        guided = i <= stop                                                  # :89-92,108 (Q4)
        latents = guided_step(unet, lgp, scheduler, text_emb, latents, noise, t, target, guided, taps, guidance_scale, beta)
        if callback is not None:
            callback(i, t, latents)
    return latents


# --------------------------------------------------------------------------------------------------
# Seeded synthetic models / inputs shared by the oracle, the tests and bench.py
# --------------------------------------------------------------------------------------------------
# SD2.1-shaped small topology (linear projections, d_head = 64, upcast attention) for the injected-attention path
TINY21_CONFIG = dict(TINY_CONFIG, attention_head_dim=(1, 2, 4, 4), use_linear_projection=True, upcast_attention=True)
CONFIGS = {"sd15": SD15_CONFIG, "sd21": SD21_CONFIG, "tiny": TINY_CONFIG, "tiny21": TINY21_CONFIG}
WEIGHT_SEED = 1138
SAMPLE_SEED = 1139


def make_unet(name="sd15", seed=WEIGHT_SEED):
    """PyTorch-default-initialised UNet of the named topology under a fixed CPU seed (no checkpoints offline)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(**CONFIGS[name])
    torch.random.set_rng_state(g)
    return unet.eval()


def make_lgp(unet, seed=WEIGHT_SEED + 1, bn_affine_jitter=True):
    """fp16 LGP (reference: app.py:67-69).  BN affine is (1,0) by default init; a small seeded jitter
    makes the BN-affine code path observable in parity tests."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    lgp = LatentEdgePredictorOracle(lgp_input_dim(unet), 4, NUM_POS_LAYERS)
    if bn_affine_jitter:
        for m in lgp.layers:
            if isinstance(m, nn.BatchNorm1d):
                m.weight.data.add_(0.1 * torch.randn_like(m.weight))
                m.bias.data.add_(0.1 * torch.randn_like(m.bias))
    torch.random.set_rng_state(g)
    return lgp.half()       # stays in train mode: SURVEY.md Q2


def make_inputs(unet, seed=SAMPLE_SEED):
    """CPU-seeded initial latents [1,4,L,L], prompt embeddings [2,77,D] ([uncond, cond]) and sketch
    target [1,4,L,L]."""
    gen = torch.Generator().manual_seed(seed)
    L = unet.config.sample_size
    latents = torch.randn(1, 4, L, L, generator=gen)
    emb = torch.randn(2, 77, unet.config.cross_attention_dim, generator=gen)
    target = torch.randn(1, 4, L, L, generator=gen)
    return latents, emb, target


def make_scheduler(prediction_type="epsilon", kind="ddim"):
    """kind "ddim": BASELINE.json's scheduler; "dpmpp": the demo's DPM-Solver++(2M) exactly as /root/reference/app.py:14-25
    constructs it."""
    if kind == "dpmpp":
        return DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                           num_train_timesteps=1000, trained_betas=None, prediction_type=prediction_type,
                                           thresholding=False, algorithm_type="dpmsolver++", solver_type="midpoint",
                                           lower_order_final=True)
    return DDIMScheduler(prediction_type=prediction_type)


# --------------------------------------------------------------------------------------------------
# Injected sketch attention  (/root/reference/modules/sketch_guided_attn.py:8-161)
# --------------------------------------------------------------------------------------------------
class SatAttnOracle(nn.Module):
    """AttnModule restated (sketch_guided_attn.py:46-161): same sub-module / parameter names; rebinding the wrapped
    BasicTransformerBlock's forward like :74-79."""

    def __init__(self, sat_name, base_layer):
        super().__init__()
        from diffusers.models.attention import CrossAttention
        self.name = sat_name
        attn1 = base_layer.attn1
        dim = attn1.to_q.in_features
        heads = attn1.heads
        dim_head = attn1.to_q.out_features // heads
        self.sketch_norm = nn.LayerNorm(dim)
        self.sketch_attn = CrossAttention(query_dim=dim, heads=heads, dim_head=dim_head, dropout=0.0, bias=False,
                                          upcast_attention=attn1.upcast_attention)
        self.sketch_conv = nn.Conv1d(dim, dim, 1)
        self.sketch_scale = 1.0
        self.res_sample = None
        outer = self

        def forward(blk, hidden_states, encoder_hidden_states=None, timestep=None, attention_mask=None,
                    cross_attention_kwargs=None, class_labels=None):
            return outer.block_forward(blk, hidden_states, encoder_hidden_states, attention_mask)

        base_layer.forward = forward.__get__(base_layer, type(base_layer))

    def set_res_sample(self, res_sample):
        b, c, h, w = res_sample.shape
        self.res_sample = res_sample.permute(0, 2, 3, 1).reshape(b, h * w, c)          # "b c h w -> b (h w) c" (:82)

    def set_scale(self, scale):
        self.sketch_scale = scale

    def block_forward(self, blk, hidden_states, encoder_hidden_states, attention_mask):
        # :99-118 self-attention + residual
        attn_output = blk.attn1(blk.norm1(hidden_states),
                                encoder_hidden_states=encoder_hidden_states if blk.only_cross_attention else None,
                                attention_mask=attention_mask)
        hidden_states = attn_output + hidden_states
        if self.res_sample is not None:
            # :126-132 LayerNorm -> cross-attention to the sketch tokens -> slice (no-op) -> Conv1d 1x1 -> scale -> residual
            a = self.sketch_attn(self.sketch_norm(hidden_states), encoder_hidden_states=self.res_sample)
            a = a[:, :attn_output.shape[1], :attn_output.shape[2]].permute(0, 2, 1)
            a = self.sketch_scale * self.sketch_conv(a)
            hidden_states = a.permute(0, 2, 1) + hidden_states
        # :134-159 text cross-attention and feed-forward
        hidden_states = blk.attn2(blk.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states,
                                  attention_mask=attention_mask) + hidden_states
        return blk.ff(blk.norm3(hidden_states)) + hidden_states


class SatMixinOracle(nn.Module):
    """SatMixin restated (sketch_guided_attn.py:8-44)."""

    def __init__(self, unet):
        super().__init__()
        self.blocks = []
        for name, module in unet.named_modules():
            if module.__class__.__name__ == "BasicTransformerBlock":
                blk = SatAttnOracle(("sketch_attn." + name).replace(".", "_"), module)
                self.blocks.append(blk)
        for blk in self.blocks:
            self.add_module(blk.name, blk)

    def set_res_samples(self, res_samples):
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
        for blk in self.blocks:
            blk.set_scale(scale)


def make_sat(unet, cls=SatMixinOracle, seed=WEIGHT_SEED + 2):
    """Seeded SatMixin over `unet` (rebinds its transformer blocks' forward!).  cls: the port's restatement or the
    reference's own class (same constructor)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    sat = cls(unet)
    torch.random.set_rng_state(g)
    return sat


# --------------------------------------------------------------------------------------------------
# Sketch feature encoder  (/root/reference/modules/sketch_encoder.py:11-98)
# --------------------------------------------------------------------------------------------------
ENCODER_BLOCKS = ("DownBlock2D",) * 4


def make_sketch_encoder(name="tiny", seed=WEIGHT_SEED + 3, cls=None):
    """Seeded SketchEncoder-shaped model: the named UNet topology with attention-free down blocks -- the only form whose
    forward the reference can execute (sketch_encoder.py:93-95 calls the blocks without encoder_hidden_states).
    cls: the reference's own ``SketchEncoder`` class (same constructor) or None for the shim UNet."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    enc = (cls or UNet2DConditionModel)(**dict(CONFIGS[name], down_block_types=ENCODER_BLOCKS))
    torch.random.set_rng_state(g)
    return enc.eval()


@torch.no_grad()
def sketch_encoder_forward(enc, sample, timestep):
    """sketch_encoder.py:50-98 restated: time embedding, conv_in, the down blocks; returns the list of per-block res_samples
    tuples (the reference wraps it in UNet2DConditionOutput(sample=...))."""
    t = timestep
    if not torch.is_tensor(t):
        t = torch.tensor([t], dtype=torch.long, device=sample.device)      # :59-66
    elif t.dim() == 0:
        t = t[None].to(sample.device)
    t = t.expand(sample.shape[0])                                          # :71
    emb = enc.time_embedding(enc.time_proj(t).to(dtype=enc.dtype))         # :73-79
    sample = enc.conv_in(sample)                                           # :92
    out = []
    for blk in enc.down_blocks:                                            # :95-98
        sample, res = blk(hidden_states=sample, temb=emb)
        out.append(res)
    return out


# --------------------------------------------------------------------------------------------------
# VAE either side of the loop  (/root/reference/app.py:107-109, modules/pipeline.py:118, :163-174)
# --------------------------------------------------------------------------------------------------
def make_vae(name="tiny", seed=WEIGHT_SEED + 4):
    """Seeded AutoencoderKL (oracle/diffusers_shim): "sd" = the SD VAE topology, "tiny" = the same structure, narrow."""
    from diffusers.models.vae import SD_VAE_CONFIG, TINY_VAE_CONFIG, AutoencoderKL
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    vae = AutoencoderKL(**(SD_VAE_CONFIG if name == "sd" else TINY_VAE_CONFIG))
    torch.random.set_rng_state(g)
    return vae.eval()


def decode_latents_L(vae, latents):
    """pipeline.py:163-174 restated: decode, map to [0, 1], zero everything below 0.5, uint8 HWC."""
    import numpy as np
    image = vae.decode(1 / 0.18215 * latents).sample
    image = (image / 2 + 0.5).clamp(0, 1)
    image = image.detach().cpu().permute(0, 2, 3, 1).float().numpy()
    image[image < 0.5] = 0
    image = image.squeeze(0) * 255
    return image.astype(np.uint8)


def make_res_samples(unet, batch, seed=SAMPLE_SEED + 7, size=None):
    """Synthetic SketchEncoder output (modules/sketch_encoder.py:93-98): one tuple of feature maps per down block --
    (resnet/attn out) x layers_per_block (+ the downsampled map for all but the last block)."""
    gen = torch.Generator().manual_seed(seed)
    boc = unet.config.block_out_channels
    L = size or unet.config.sample_size
    out = []
    for i, c in enumerate(boc):
        maps = [torch.randn(batch, c, L, L, generator=gen) for _ in range(unet.config.layers_per_block)]
        if i < len(boc) - 1:
            L //= 2
            maps.append(torch.randn(batch, c, L, L, generator=gen))
        out.append(tuple(maps))
    return out
