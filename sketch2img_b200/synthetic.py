"""Seeded synthetic weights / inputs for the sampling path (bench.py, smoke): there are no SD checkpoints offline.

``unet_state_dict`` enumerates the diffusers parameter names of the 4-level SD ``UNet2DConditionModel`` topology
(the names the reference's ``pipe.unet`` carries when loaded by ``from_pretrained``,
/root/reference/app.py:32-38) and fills them with PyTorch-default-style initial values (uniform +-1/sqrt(fan_in);
norm affine = (1, 0)).  ``count_macs`` walks the same topology and returns the algorithmic multiply-accumulate
counts bench.py's roofline uses.
"""
import math

import torch


def _uniform(gen, shape, fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2 - 1) * b


def _topology(cfg):
    """Yield (kind, name, dims) for every parametrised op in forward order.
    kinds: conv3 (Cout, Cin), lin (N, K, bias, as_conv1x1), norm (C)."""
    boc = list(cfg["block_out_channels"])
    layers = int(cfg.get("layers_per_block", 2))
    D = int(cfg["cross_attention_dim"])
    use_lin = bool(cfg.get("use_linear_projection", False))
    temb = boc[0] * 4
    ops = []

    def res(pre, cin, cout):
        ops.append(("norm", pre + ".norm1", (cin,)))
        ops.append(("conv3", pre + ".conv1", (cout, cin)))
        ops.append(("lin", pre + ".time_emb_proj", (cout, temb, True, False)))
        ops.append(("norm", pre + ".norm2", (cout,)))
        ops.append(("conv3", pre + ".conv2", (cout, cout)))
        if cin != cout:
            ops.append(("lin", pre + ".conv_shortcut", (cout, cin, True, True)))

    def tfm(pre, C):
        tb = pre + ".transformer_blocks.0"
        ops.append(("norm", pre + ".norm", (C,)))
        ops.append(("lin", pre + ".proj_in", (C, C, True, not use_lin)))
        for n in ("norm1", "norm2", "norm3"):
            ops.append(("norm", tb + "." + n, (C,)))
        for a, kd in (("attn1", C), ("attn2", D)):
            ops.append(("lin", f"{tb}.{a}.to_q", (C, C, False, False)))
            ops.append(("lin", f"{tb}.{a}.to_k", (C, kd, False, False)))
            ops.append(("lin", f"{tb}.{a}.to_v", (C, kd, False, False)))
            ops.append(("lin", f"{tb}.{a}.to_out.0", (C, C, True, False)))
        ops.append(("lin", tb + ".ff.net.0.proj", (8 * C, C, True, False)))
        ops.append(("lin", tb + ".ff.net.2", (C, 4 * C, True, False)))
        ops.append(("lin", pre + ".proj_out", (C, C, True, not use_lin)))

    ops.append(("conv3", "conv_in", (boc[0], int(cfg.get("in_channels", 4)))))
    ops.append(("lin", "time_embedding.linear_1", (temb, boc[0], True, False)))
    ops.append(("lin", "time_embedding.linear_2", (temb, temb, True, False)))
    ch = boc[0]
    for i in range(4):
        pre = f"down_blocks.{i}"
        for j in range(layers):
            res(f"{pre}.resnets.{j}", ch if j == 0 else boc[i], boc[i])
            if i < 3:
                tfm(f"{pre}.attentions.{j}", boc[i])
        ch = boc[i]
        if i < 3:
            ops.append(("conv3", f"{pre}.downsamplers.0.conv", (ch, ch)))
    res("mid_block.resnets.0", boc[3], boc[3])
    tfm("mid_block.attentions.0", boc[3])
    res("mid_block.resnets.1", boc[3], boc[3])
    rev = boc[::-1]
    out_c = rev[0]
    for i in range(4):
        pre = f"up_blocks.{i}"
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, 3)]
        for j in range(layers + 1):
            skip = in_c if j == layers else out_c
            first = prev if j == 0 else out_c
            res(f"{pre}.resnets.{j}", first + skip, out_c)
            if i > 0:
                tfm(f"{pre}.attentions.{j}", out_c)
        if i < 3:
            ops.append(("conv3", f"{pre}.upsamplers.0.conv", (out_c, out_c)))
    ops.append(("norm", "conv_norm_out", (boc[0],)))
    ops.append(("conv3", "conv_out", (int(cfg.get("out_channels", 4)), boc[0])))
    return ops


def unet_param_shapes(cfg):
    """name -> shape of every parameter ``unet_state_dict`` produces (no values)."""
    out = {}
    for kind, name, dims in _topology(cfg):
        if kind == "norm":
            out[name + ".weight"] = out[name + ".bias"] = (dims[0],)
        elif kind == "conv3":
            out[name + ".weight"] = (dims[0], dims[1], 3, 3)
            out[name + ".bias"] = (dims[0],)
        else:
            n, k, bias, as_conv = dims
            out[name + ".weight"] = (n, k, 1, 1) if as_conv else (n, k)
            if bias:
                out[name + ".bias"] = (n,)
    return out


def unet_state_dict(cfg, seed=1138):
    """diffusers-named fp32 CPU state dict of a randomly initialised SD UNet."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for kind, name, dims in _topology(cfg):
        if kind == "norm":
            sd[name + ".weight"] = torch.ones(dims[0])
            sd[name + ".bias"] = torch.zeros(dims[0])
        elif kind == "conv3":
            co, ci = dims
            sd[name + ".weight"] = _uniform(gen, (co, ci, 3, 3), ci * 9)
            sd[name + ".bias"] = _uniform(gen, (co,), ci * 9)
        else:
            n, k, bias, as_conv = dims
            sd[name + ".weight"] = _uniform(gen, (n, k, 1, 1) if as_conv else (n, k), k)
            if bias:
                sd[name + ".bias"] = _uniform(gen, (n,), k)
    return sd


def sketch_encoder_state_dict(cfg, seed=1140):
    """Randomly initialised SketchEncoder (modules/sketch_encoder.py): conv_in, time embedding and the down path of the
    topology with attention-free down blocks."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for kind, name, dims in _topology(cfg):
        if not name.startswith(("conv_in", "time_embedding", "down_blocks")) or ".attentions." in name:
            continue
        if kind == "norm":
            sd[name + ".weight"] = torch.ones(dims[0])
            sd[name + ".bias"] = torch.zeros(dims[0])
        elif kind == "conv3":
            co, ci = dims
            sd[name + ".weight"] = _uniform(gen, (co, ci, 3, 3), ci * 9)
            sd[name + ".bias"] = _uniform(gen, (co,), ci * 9)
        else:
            n, k, bias, as_conv = dims
            sd[name + ".weight"] = _uniform(gen, (n, k, 1, 1) if as_conv else (n, k), k)
            if bias:
                sd[name + ".bias"] = _uniform(gen, (n,), k)
    return sd


def lgp_input_dim(cfg, num_pos_layers=9):
    boc = cfg["block_out_channels"]
    return boc[0] + boc[1] + boc[2] + 3 * boc[3] + boc[3] + boc[2] + boc[1] + 4 + 4 * num_pos_layers


def sample_inputs(cfg, n, seed=1139, pin=False):
    """CPU-seeded per-image inputs: latents [n,4,L,L], prompt embeddings [2n,77,D] ([uncond..., cond...]),
    sketch targets [n,4,L,L]."""
    L, D = int(cfg["sample_size"]), int(cfg["cross_attention_dim"])
    lat, emb_u, emb_c, tgt = [], [], [], []
    for k in range(n):
        g = torch.Generator().manual_seed(seed + k)
        lat.append(torch.randn(1, 4, L, L, generator=g))
        e = torch.randn(2, 77, D, generator=g)
        emb_u.append(e[:1])
        emb_c.append(e[1:])
        tgt.append(torch.randn(1, 4, L, L, generator=g))
    out = [torch.cat(lat), torch.cat(emb_u + emb_c), torch.cat(tgt)]
    if pin and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return out


def count_macs(cfg, L=None, ctx_len=77):
    """Algorithmic MACs of ONE sample-forward, by class, plus the share upstream of the last LGP tap
    (up_blocks[2] output) that the guided backward traverses.  Attention counts QK^T + PV at the true head dim."""
    boc = list(cfg["block_out_channels"])
    L = int(L or cfg["sample_size"])
    layers = int(cfg.get("layers_per_block", 2))
    D = int(cfg["cross_attention_dim"])
    tot = {"conv3x3": 0, "linear": 0, "conv1x1": 0, "attn": 0}
    on = dict(tot)           # on the backward path
    on_self_attn = 0

    def add(kind, macs, onpath):
        tot[kind] += macs
        if onpath:
            on[kind] += macs

    def res(cin, cout, hw, onpath):
        add("conv3x3", hw * 9 * cin * cout + hw * 9 * cout * cout, onpath)
        if cin != cout:
            add("conv1x1", hw * cin * cout, onpath)

    def tfm(C, hw, onpath):
        nonlocal on_self_attn
        add("conv1x1", 2 * hw * C * C, onpath)                      # proj_in / proj_out
        add("linear", hw * C * C * 4, onpath)                       # q, k, v, out (self)
        add("linear", hw * C * C * 2 + 2 * ctx_len * D * C, onpath)  # q, out (cross) + k, v of the text
        add("linear", hw * C * 8 * C + hw * 4 * C * C, onpath)      # GEGLU + ff out
        add("attn", 2 * hw * hw * C + 2 * hw * ctx_len * C, onpath)
        if onpath:
            on_self_attn += 2 * hw * hw * C

    side = L
    add("conv3x3", side * side * 9 * int(cfg.get("in_channels", 4)) * boc[0], True)
    ch = boc[0]
    for i in range(4):
        for j in range(layers):
            res(ch if j == 0 else boc[i], boc[i], side * side, True)
            if i < 3:
                tfm(boc[i], side * side, True)
        ch = boc[i]
        if i < 3:
            side //= 2
            add("conv3x3", side * side * 9 * ch * ch, True)
    res(boc[3], boc[3], side * side, True)
    tfm(boc[3], side * side, True)
    res(boc[3], boc[3], side * side, True)
    rev = boc[::-1]
    out_c = rev[0]
    for i in range(4):
        prev, out_c = out_c, rev[i]
        in_c = rev[min(i + 1, 3)]
        onpath = i < 3
        for j in range(layers + 1):
            skip = in_c if j == layers else out_c
            first = prev if j == 0 else out_c
            res(first + skip, out_c, side * side, onpath)
            if i > 0:
                tfm(out_c, side * side, onpath)
        if i < 3:
            side *= 2
            add("conv3x3", side * side * 9 * out_c * out_c, onpath)
    add("conv3x3", side * side * 9 * boc[0] * int(cfg.get("out_channels", 4)), False)
    fwd = sum(tot.values())
    on_dense = on["conv3x3"] + on["linear"] + on["conv1x1"]
    return {"forward": fwd, "by_class": tot, "on_path_dense": on_dense, "on_path_attn": on["attn"]}


def flops_per_image(cfg, steps=50, guided_steps=26, L=None, lgp_hidden=(512, 256, 128, 64), lgp_out=4):
    """Algorithmic FLOPs of one sketch-guided image (CFG pair per step), following SURVEY 8d: unguided step =
    2 sample-forwards; guided step adds the dX backward (dense dgrad at 1x forward MACs, attention at 2x) and the LGP
    forward + dX backward over 2*L*L rows."""
    L = int(L or cfg["sample_size"])
    m = count_macs(cfg, L)
    fwd = 2 * 2 * m["forward"]
    bwd = 2 * 2 * (m["on_path_dense"] + 2 * m["on_path_attn"])
    widths = (lgp_input_dim(cfg),) + tuple(lgp_hidden) + (lgp_out,)
    lgp = 2 * (2 * L * L) * sum(a * b for a, b in zip(widths[:-1], widths[1:]))
    guided = fwd + bwd + 2 * lgp
    return {"unguided_step": fwd, "guided_step": guided, "lgp_fwd": lgp, "unet_bwd": bwd,
            "image": guided_steps * guided + (steps - guided_steps) * fwd}
