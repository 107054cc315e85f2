"""Bisecting aid (GPU box): per-tap / per-block relative error of the CUDA UNet against the CPU oracle."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from sketch2img_b200.engine import UNetEngine  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def main(name="tiny", t=981):
    torch.set_num_threads(os.cpu_count())
    unet = port.make_unet(name)
    lat, emb, _ = port.make_inputs(unet)
    x = torch.cat([lat] * 2)
    eng = UNetEngine(unet.config.__dict__, unet.state_dict())
    eng.set_debug(True)
    taps, _ = port.register_taps(unet)
    blocks = {}
    for i, b in enumerate(unet.down_blocks):
        b.register_forward_hook(lambda m, i_, o, k="down%d" % i: blocks.__setitem__(k, o[0].detach()))
    for i, b in enumerate(unet.up_blocks):
        b.register_forward_hook(lambda m, i_, o, k="up%d" % i: blocks.__setitem__(k, o.detach()))
    unet.mid_block.register_forward_hook(lambda m, i_, o: blocks.__setitem__("mid", o.detach()))
    unet.conv_in.register_forward_hook(lambda m, i_, o: blocks.__setitem__("conv_in", o.detach()))

    xg = x.clone().requires_grad_(True)
    t0 = time.time()
    with torch.enable_grad():
        eps_ref = unet(xg, torch.tensor(t), encoder_hidden_states=emb).sample
        tap_ref = [m.output for m in taps]
        gen = torch.Generator().manual_seed(7)
        G = [torch.randn(tr.shape, generator=gen) for tr in tap_ref]
        loss = sum((g * tr).sum() for g, tr in zip(G, tap_ref))
        dx_ref = torch.autograd.grad(loss, xg)[0]
    print(f"[{name}] oracle fwd+bwd {time.time() - t0:.1f}s")

    eps = eng.forward(x.cuda(), t, emb.cuda(), save_for_backward=True)
    torch.cuda.synchronize()
    print("eps rel err      %.3e" % rel(eps.cpu(), eps_ref))
    for k in ["conv_in", "down0", "down1", "down2", "down3", "mid", "up0", "up1", "up2", "up3"]:
        d = eng.debug_tensor(k).permute(0, 3, 1, 2).cpu()
        print("block %-8s rel err %.3e" % (k, rel(d, blocks[k])))
    for k in range(9):
        print("tap %d rel err %.3e" % (k, rel(eng.tap(k).permute(0, 3, 1, 2).cpu(), tap_ref[k])))
    dx = eng.backward([g.permute(0, 2, 3, 1).contiguous().cuda() for g in G])
    torch.cuda.synchronize()
    print("dx rel err       %.3e   (|dx_ref| %.3e)" % (rel(dx.cpu(), dx_ref), dx_ref.norm().item()))
    print("arena GB %.2f" % (eng.arena_bytes() / 1e9))
    # timing
    for save in (False, True):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.forward(x.cuda(), t, emb.cuda(), save_for_backward=save)
            if save:
                eng.backward([g.permute(0, 2, 3, 1).contiguous().cuda() for g in G])
        e1.record()
        torch.cuda.synchronize()
        print("save=%s  ms/iter %.2f" % (save, e0.elapsed_time(e1) / 3))


if __name__ == "__main__":
    main(*sys.argv[1:2])
