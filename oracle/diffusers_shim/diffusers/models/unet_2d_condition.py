"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped CUDA path.

CPU restatement of diffusers' ``UNet2DConditionModel`` (SD1.x / SD2.x topology) as used by the
reference at /root/reference/modules/pipeline.py:96 (``self.unet(x, t, encoder_hidden_states=...)``),
/root/reference/modules/latent_predictor.py:47-81 (the 9 hook sites) and
/root/reference/modules/sketch_encoder.py:13-98.  diffusers is an un-vendored, unpinned dependency
(requirements.txt:3, API window ~v0.12-0.13); topology and arithmetic follow SURVEY.md Appendix
A.1-A.5.  Parameter names equal real diffusers names.
"""
import math
from collections import OrderedDict
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .attention import BasicTransformerBlock


class BaseOutput(OrderedDict):
    """dict subclass with attribute access (hook_unet's ``isinstance(output, dict)`` branch,
    latent_predictor.py:57-59)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class UNet2DConditionOutput(BaseOutput):
    def __init__(self, sample=None):
        super().__init__(sample=sample)


class Transformer2DModelOutput(BaseOutput):
    def __init__(self, sample=None):
        super().__init__(sample=sample)


def sinusoidal_timestep_embedding(t, dim, flip_sin_to_cos=True, freq_shift=0):
    """A.2 step 1: [cos | sin] halves after the flip."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device)
    exponent = exponent / (half - freq_shift)
    arg = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class Timesteps(nn.Module):
    def __init__(self, dim, flip_sin_to_cos, freq_shift):
        super().__init__()
        self.dim, self.flip, self.shift = dim, flip_sin_to_cos, freq_shift

    def forward(self, t):
        return sinusoidal_timestep_embedding(t, self.dim, self.flip, self.shift)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    """A.3."""

    def __init__(self, cin, cout, temb, groups, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / 1.0


class Transformer2DModel(nn.Module):
    """A.4."""

    def __init__(self, heads, dim_head, channels, cross_dim, groups, use_linear, upcast_attention):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        if use_linear:
            self.proj_in = nn.Linear(channels, inner)
        else:
            self.proj_in = nn.Conv2d(channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim=cross_dim,
                                  upcast_attention=upcast_attention)])
        if use_linear:
            self.proj_out = nn.Linear(inner, channels)
        else:
            self.proj_out = nn.Conv2d(inner, channels, 1)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, class_labels=None,
                cross_attention_kwargs=None, return_dict=True):
        b, c, hh, ww = hidden_states.shape
        res = hidden_states
        x = self.norm(hidden_states)
        if not self.use_linear_projection:
            x = self.proj_in(x)
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, x.shape[1])
        else:
            x = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
            x = self.proj_in(x)
        for blk in self.transformer_blocks:
            x = blk(x, encoder_hidden_states=encoder_hidden_states, timestep=timestep,
                    cross_attention_kwargs=cross_attention_kwargs, class_labels=class_labels)
        if not self.use_linear_projection:
            x = x.reshape(b, hh, ww, x.shape[-1]).permute(0, 3, 1, 2).contiguous()
            x = self.proj_out(x)
        else:
            x = self.proj_out(x)
            x = x.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
        out = x + res
        if not return_dict:
            return (out,)
        return Transformer2DModelOutput(sample=out)


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, groups, layers, add_down, attn_cfg=None):
        super().__init__()
        self.has_cross_attention = attn_cfg is not None
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups) for i in range(layers)])
        if attn_cfg is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(channels=cout, **attn_cfg) for _ in range(layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, cross_attention_kwargs=None):
        outs = ()
        for i, res in enumerate(self.resnets):
            hidden_states = res(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[i](
                    hidden_states, encoder_hidden_states=encoder_hidden_states,
                    cross_attention_kwargs=cross_attention_kwargs).sample
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlock2D(_DownBlock):
    pass


class DownBlock2D(_DownBlock):
    pass


class _UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, groups, layers, add_up, attn_cfg=None):
        super().__init__()
        self.has_cross_attention = attn_cfg is not None
        res = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            first = cprev if i == 0 else cout
            res.append(ResnetBlock2D(first + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(res)
        if attn_cfg is not None:
            self.attentions = nn.ModuleList([Transformer2DModel(channels=cout, **attn_cfg) for _ in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, encoder_hidden_states=None,
                cross_attention_kwargs=None, upsample_size=None):
        for i, res in enumerate(self.resnets):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, skip], dim=1)
            hidden_states = res(hidden_states, temb)
            if self.has_cross_attention:
                hidden_states = self.attentions[i](
                    hidden_states, encoder_hidden_states=encoder_hidden_states,
                    cross_attention_kwargs=cross_attention_kwargs).sample
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlock2D(_UpBlock):
    pass


class UpBlock2D(_UpBlock):
    pass


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, c, temb, groups, attn_cfg):
        super().__init__()
        self.has_cross_attention = True
        self.attentions = nn.ModuleList([Transformer2DModel(channels=c, **attn_cfg)])
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, groups), ResnetBlock2D(c, c, temb, groups)])

    def forward(self, hidden_states, temb=None, encoder_hidden_states=None, cross_attention_kwargs=None):
        hidden_states = self.resnets[0](hidden_states, temb)
        hidden_states = self.attentions[0](
            hidden_states, encoder_hidden_states=encoder_hidden_states,
            cross_attention_kwargs=cross_attention_kwargs).sample
        return self.resnets[1](hidden_states, temb)


SD15_CONFIG = dict(
    sample_size=64, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, use_linear_projection=False,
    upcast_attention=False, norm_num_groups=32)
SD21_CONFIG = dict(
    sample_size=96, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280),
    layers_per_block=2, attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
    use_linear_projection=True, upcast_attention=True, norm_num_groups=32)
# Small topology with the same block structure; used by fast tests and golden fixtures.
TINY_CONFIG = dict(
    sample_size=16, in_channels=4, out_channels=4, block_out_channels=(64, 128, 256, 256),
    layers_per_block=2, attention_head_dim=4, cross_attention_dim=64, use_linear_projection=False,
    upcast_attention=False, norm_num_groups=32)


class UNet2DConditionModel(nn.Module):
    """down x4 (cross-attn x3) -> mid -> up x4 (cross-attn x3); ``attention_head_dim`` is the number
    of heads (the diffusers naming quirk).  Registration order: conv_in, time_proj,
    time_embedding, down_blocks, up_blocks, mid_block, conv_norm_out, conv_act, conv_out."""

    def __init__(self, sample_size=64, in_channels=4, out_channels=4,
                 block_out_channels=(320, 640, 1280, 1280), layers_per_block=2, attention_head_dim=8,
                 cross_attention_dim=768, use_linear_projection=False, upcast_attention=False,
                 norm_num_groups=32, flip_sin_to_cos=True, freq_shift=0, down_block_types=None):
        super().__init__()
        boc = tuple(block_out_channels)
        nb = len(boc)
        heads = attention_head_dim if isinstance(attention_head_dim, (tuple, list)) else (attention_head_dim,) * nb
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            block_out_channels=boc, layers_per_block=layers_per_block, attention_head_dim=attention_head_dim,
            cross_attention_dim=cross_attention_dim, use_linear_projection=use_linear_projection,
            upcast_attention=upcast_attention, norm_num_groups=norm_num_groups,
            flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift, down_block_types=down_block_types,
            center_input_sample=False, class_embed_type=None)
        # read by modules/sketch_encoder.py:39,85 (SketchEncoder.forward)
        self.num_upsamplers = len(boc) - 1
        self.class_embedding = None
        self.in_channels = in_channels
        self.sample_size = sample_size
        temb = boc[0] * 4
        g = norm_num_groups

        def attn_cfg(i):
            return dict(heads=heads[i], dim_head=boc[i] // heads[i], cross_dim=cross_attention_dim, groups=g,
                        use_linear=use_linear_projection, upcast_attention=upcast_attention)

        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], flip_sin_to_cos, freq_shift)
        self.time_embedding = TimestepEmbedding(boc[0], temb)

        self.down_blocks = nn.ModuleList()
        out_c = boc[0]
        for i in range(nb):
            in_c, out_c = out_c, boc[i]
            last = i == nb - 1
            # diffusers' down_block_types: default (CrossAttnDownBlock2D x3, DownBlock2D); all-DownBlock2D is the only form the
            # reference's SketchEncoder.forward can execute (it calls the blocks without encoder_hidden_states)
            plain = last if down_block_types is None else down_block_types[i] == "DownBlock2D"
            cls = DownBlock2D if plain else CrossAttnDownBlock2D
            self.down_blocks.append(cls(in_c, out_c, temb, g, layers_per_block, not last,
                                        None if plain else attn_cfg(i)))

        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_c = rev[0]
        for i in range(nb):
            prev = out_c
            out_c = rev[i]
            in_c = rev[min(i + 1, nb - 1)]
            last = i == nb - 1
            first = i == 0
            cls = UpBlock2D if first else CrossAttnUpBlock2D
            self.up_blocks.append(cls(in_c, out_c, prev, temb, g, layers_per_block + 1, not last,
                                      None if first else attn_cfg(nb - 1 - i)))

        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb, g, attn_cfg(nb - 1))
        self.conv_norm_out = nn.GroupNorm(g, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, attention_mask=None,
                cross_attention_kwargs=None, return_dict=True):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.dim() == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(t).to(dtype=self.dtype))

        sample = self.conv_in(sample)
        skips = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, outs = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                   cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample, outs = blk(hidden_states=sample, temb=emb)
            skips += outs

        sample = self.mid_block(sample, emb, encoder_hidden_states=encoder_hidden_states,
                                cross_attention_kwargs=cross_attention_kwargs)

        for blk in self.up_blocks:
            n = len(blk.resnets)
            res = skips[-n:]
            skips = skips[:-n]
            if blk.has_cross_attention:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res,
                             encoder_hidden_states=encoder_hidden_states,
                             cross_attention_kwargs=cross_attention_kwargs)
            else:
                sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=res)

        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        if not return_dict:
            return (sample,)
        return UNet2DConditionOutput(sample=sample)
