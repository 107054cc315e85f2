"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped CUDA path.

CPU restatement of diffusers' ``AutoencoderKL`` (the SD VAE) as the reference uses it either side of the sampling loop:
/root/reference/app.py:107-109 (``vae.encode(img).latent_dist.sample() * 0.18215`` -> the sketch target) and
/root/reference/modules/pipeline.py:118, :163-174 (``vae.decode(latents / 0.18215).sample`` -> the image).
diffusers is an un-vendored, unpinned dependency (API window ~v0.12-0.13); the topology below is restated from
that version: Encoder / Decoder of DownEncoderBlock2D / UpDecoderBlock2D (ResnetBlock2D without time embedding, GroupNorm
eps 1e-6), a single-head AttentionBlock in the mid block, quant_conv / post_quant_conv 1x1, DiagonalGaussianDistribution.
Parameter names equal the real diffusers names (``encoder.down_blocks.0.resnets.0.norm1.weight`` ...,
``mid_block.attentions.0.{group_norm,query,key,value,proj_attn}``) so genuine checkpoints of that era load.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet_2d_condition import BaseOutput


class _Resnet(nn.Module):
    """ResnetBlock2D(temb_channels=None, eps=1e-6)."""

    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return (x + h) / 1.0


class AttentionBlock(nn.Module):
    """diffusers <= 0.14 ``AttentionBlock`` with num_head_channels=None (one head): softmax((q s)(k s)^T) v with
    s = 1 / sqrt(sqrt(C)), scores and softmax in fp32, then proj_attn and the residual."""

    def __init__(self, channels, groups):
        super().__init__()
        self.channels = channels
        self.num_heads = 1
        self.group_norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.query = nn.Linear(channels, channels)
        self.key = nn.Linear(channels, channels)
        self.value = nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)

    def forward(self, x):
        b, c, h, w = x.shape
        res = x
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.query(t), self.key(t), self.value(t)
        scale = 1 / math.sqrt(math.sqrt(self.channels / self.num_heads))
        scores = torch.matmul(q * scale, (k * scale).transpose(-1, -2))
        probs = torch.softmax(scores.float(), dim=-1).type(scores.dtype)
        t = self.proj_attn(torch.matmul(probs, v))
        return (t.transpose(-1, -2).reshape(b, c, h, w) + res) / 1.0


class _Downsample(nn.Module):
    """Downsample2D(use_conv=True, padding=0): pad right / bottom by one, then 3x3 stride 2."""

    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class _Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _DownEncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.downsamplers = nn.ModuleList([_Downsample(cout)]) if add_down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                x = d(x)
        return x


class _UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([_Resnet(cin if i == 0 else cout, cout, groups) for i in range(layers)])
        self.upsamplers = nn.ModuleList([_Upsample(cout)]) if add_up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                x = u(x)
        return x


class _MidBlock(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.attentions = nn.ModuleList([AttentionBlock(c, groups)])
        self.resnets = nn.ModuleList([_Resnet(c, c, groups), _Resnet(c, c, groups)])

    def forward(self, x):
        x = self.resnets[0](x)
        x = self.attentions[0](x)
        return self.resnets[1](x)


class Encoder(nn.Module):
    def __init__(self, in_channels, out_channels, boc, layers, groups):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(_DownEncoderBlock(c, co, layers, groups, i < len(boc) - 1))
            c = co
        self.mid_block = _MidBlock(boc[-1], groups)
        self.conv_norm_out = nn.GroupNorm(groups, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class Decoder(nn.Module):
    def __init__(self, in_channels, out_channels, boc, layers, groups):
        super().__init__()
        rev = list(reversed(boc))
        self.conv_in = nn.Conv2d(in_channels, rev[0], 3, padding=1)
        self.mid_block = _MidBlock(rev[0], groups)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(_UpDecoderBlock(c, co, layers + 1, groups, i < len(rev) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(self.conv_act(self.conv_norm_out(x)))


class DiagonalGaussianDistribution:
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


class AutoencoderKLOutput(BaseOutput):
    def __init__(self, latent_dist=None):
        super().__init__(latent_dist=latent_dist)


class DecoderOutput(BaseOutput):
    def __init__(self, sample=None):
        super().__init__(sample=sample)


SD_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                     latent_channels=4, norm_num_groups=32, sample_size=512)
# small topology with the same structure (fast tests / fixtures)
TINY_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(64, 128, 128, 128), layers_per_block=2,
                       latent_channels=4, norm_num_groups=32, sample_size=64)


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                 latent_channels=4, norm_num_groups=32, sample_size=512):
        super().__init__()
        boc = tuple(block_out_channels)
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, latent_channels=latent_channels,
                                      norm_num_groups=norm_num_groups, sample_size=sample_size)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block, norm_num_groups)
        self.decoder = Decoder(latent_channels, out_channels, boc, layers_per_block, norm_num_groups)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode(self, x, return_dict=True):
        moments = self.quant_conv(self.encoder(x))
        return AutoencoderKLOutput(latent_dist=DiagonalGaussianDistribution(moments))

    def decode(self, z, return_dict=True):
        return DecoderOutput(sample=self.decoder(self.post_quant_conv(z)))
