"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped CUDA path.

CPU restatement of the diffusers (~v0.12-0.13) attention building blocks the reference imports:
``from diffusers.models.attention import CrossAttention, BasicTransformerBlock``
(/root/reference/modules/sketch_guided_attn.py:5).  diffusers itself is absent from the image and is
an un-vendored, unpinned dependency of the reference (requirements.txt:3); the arithmetic below is
restated from its published algorithm.  Parameter names equal the real
diffusers names so genuine checkpoints load.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CrossAttention(nn.Module):
    """softmax(q k^T * d^-0.5) v with bias-free q/k/v projections and a biased output projection."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0,
                 bias=False, upcast_attention=False, upcast_softmax=False):
        super().__init__()
        inner = heads * dim_head
        kv_dim = query_dim if cross_attention_dim is None else cross_attention_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def _split(self, t):
        b, n, c = t.shape
        h = self.heads
        return t.reshape(b, n, h, c // h).permute(0, 2, 1, 3).reshape(b * h, n, c // h)

    def _merge(self, t):
        bh, n, d = t.shape
        h = self.heads
        return t.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self._split(self.to_q(hidden_states))
        k = self._split(self.to_k(ctx))
        v = self._split(self.to_v(ctx))
        dtype = q.dtype
        if self.upcast_attention:
            q, k = q.float(), k.float()
        scores = torch.baddbmm(
            torch.empty(q.shape[0], q.shape[1], k.shape[1], dtype=q.dtype, device=q.device),
            q, k.transpose(-1, -2), beta=0, alpha=self.scale)
        if attention_mask is not None:
            scores = scores + attention_mask
        if self.upcast_softmax:
            scores = scores.float()
        probs = scores.softmax(dim=-1).to(dtype)
        out = self._merge(torch.bmm(probs, v))
        out = self.to_out[0](out)
        return self.to_out[1](out)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        a, gate = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(dropout), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """LN -> self-attn -> LN -> cross-attn -> LN -> GEGLU FF, each with a residual add."""

    def __init__(self, dim, num_attention_heads, attention_head_dim, cross_attention_dim=None,
                 upcast_attention=False, only_cross_attention=False):
        super().__init__()
        self.only_cross_attention = only_cross_attention
        self.use_ada_layer_norm = False
        self.use_ada_layer_norm_zero = False
        self.attn1 = CrossAttention(
            query_dim=dim, heads=num_attention_heads, dim_head=attention_head_dim,
            cross_attention_dim=cross_attention_dim if only_cross_attention else None,
            upcast_attention=upcast_attention)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(
            query_dim=dim, cross_attention_dim=cross_attention_dim, heads=num_attention_heads,
            dim_head=attention_head_dim, upcast_attention=upcast_attention)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, hidden_states, encoder_hidden_states=None, timestep=None, attention_mask=None,
                cross_attention_kwargs=None, class_labels=None):
        kw = cross_attention_kwargs or {}
        hidden_states = self.attn1(
            self.norm1(hidden_states),
            encoder_hidden_states=encoder_hidden_states if self.only_cross_attention else None,
            attention_mask=attention_mask, **kw) + hidden_states
        hidden_states = self.attn2(
            self.norm2(hidden_states), encoder_hidden_states=encoder_hidden_states,
            attention_mask=attention_mask, **kw) + hidden_states
        return self.ff(self.norm3(hidden_states)) + hidden_states
