"""GPU box bisecting aid: ONE guided denoising step of the CUDA path against the CPU oracle, stage by stage
(eps, DDIM latent, LGP output/loss, tap gradients with the taps as detached leaves, UNet input gradient, updated
latent), plus a run-to-run determinism check of every UNet block output."""
import copy
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from sketch2img_b200 import _lib  # noqa: E402
from sketch2img_b200.engine import UNetEngine  # noqa: E402
from sketch2img_b200.latent_predictor import LGPEngine  # noqa: E402


def rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-30)).item()


def main(name="tiny", t=981, B=2):
    torch.set_num_threads(os.cpu_count())
    t, B = int(t), int(B)
    unet = port.make_unet(name)
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    sch = port.make_scheduler()
    sch.set_timesteps(50)
    L = lat.shape[2]
    eng = UNetEngine(unet.config.__dict__, unet.state_dict())
    leng = LGPEngine(port.lgp_input_dim(unet), 4, port.NUM_POS_LAYERS, copy.deepcopy(lgp).float().state_dict())

    # ---------------- determinism of the forward, block by block
    g = torch.Generator().manual_seed(11)
    xb = torch.randn(B, 4, L, L, generator=g).cuda()
    eb = torch.randn(B, 77, emb.shape[2], generator=g).cuda()
    eng.set_debug(True)
    names = ["conv_in", "r0.h1", "r0.out", "t0.t0", "t0.t1", "t0.t2", "t0.out", "r1.h1", "r1.out", "t1.out",
             "down0", "down1", "down2", "down3", "mid", "up0", "up1", "up2", "up3"]
    runs = []
    for _ in range(3):
        eps = eng.forward(xb, t, eb, save_for_backward=False).clone()
        runs.append(([eng.debug_tensor(k).clone() for k in names], eps))
    for r in (1, 2):
        print("run %d vs run 0:" % r, " ".join("%s %.1e" % (k, rel(a, b)) for k, a, b in zip(names, runs[r][0], runs[0][0])),
              "eps %.1e" % rel(runs[r][1], runs[0][1]))
    eng.set_debug(False)

    # ---------------- oracle: one guided step with every stage kept
    taps, handles = port.register_taps(unet)
    x_in = torch.cat([lat] * 2).requires_grad_(True)
    tt = torch.tensor(t)
    with torch.enable_grad():
        eps_ref = unet(x_in, tt, encoder_hidden_states=emb).sample
        tap_ref = [m.output for m in taps]
    eu, ec = eps_ref.detach().chunk(2)
    e = eu + 7.5 * (ec - eu)
    x_ddim = sch.step(e, tt, lat, eta=0.0).prev_sample
    # LGP on detached taps (direct partial gradients), then the chain through the UNet
    leaves = [tp.detach().clone().requires_grad_(True) for tp in tap_ref]
    with torch.enable_grad():
        feats = torch.cat([F.interpolate(lf, size=L, mode="bilinear") for lf in leaves], dim=1)
        lvl = port.noise_level(sch, lat, tt)
        out = lgp(feats, torch.cat([lvl] * 2))
        o4 = out.reshape(2, L, L, -1).permute(0, 3, 2, 1)
        loss = F.mse_loss(tgt.float(), o4.chunk(2)[1].float())
        g_tap = torch.autograd.grad(loss, leaves)
        dx_ref = torch.autograd.grad(tap_ref, x_in, grad_outputs=[gt.to(tp.dtype) for gt, tp in zip(g_tap, tap_ref)],
                                     retain_graph=True)[0]
    # the same chain with an fp32 LGP (no fp16 rounding anywhere): the "exact math" the fp16 reference approximates
    lgp32 = port.LatentEdgePredictorOracle32(lgp)
    leaves32 = [tp.detach().clone().requires_grad_(True) for tp in tap_ref]
    with torch.enable_grad():
        feats32 = torch.cat([F.interpolate(lf, size=L, mode="bilinear") for lf in leaves32], dim=1)
        out32 = lgp32(feats32, torch.cat([lvl] * 2))
        o432 = out32.reshape(2, L, L, -1).permute(0, 3, 2, 1)
        loss32 = F.mse_loss(tgt.float(), o432.chunk(2)[1].float())
        g_tap32 = torch.autograd.grad(loss32, leaves32)
        dx_ref32 = torch.autograd.grad(tap_ref, x_in, grad_outputs=list(g_tap32))[0]
    cos = lambda a, b: F.cosine_similarity(a.flatten().double().cpu(), b.flatten().double().cpu(), dim=0).item()
    print("oracle fp16-LGP vs fp32-LGP: out %.3e  tap grads %s  dx %.3e cos %.6f" % (
        rel(out.float(), out32), " ".join("%.2e" % rel(a, b) for a, b in zip(g_tap, g_tap32)), rel(dx_ref, dx_ref32),
        cos(dx_ref, dx_ref32)))
    gq = (-dx_ref).chunk(2)[1]
    alpha = torch.linalg.norm(x_in.detach() - x_ddim) / torch.linalg.norm(gq) * 1.6
    x_new_ref = x_ddim + alpha * gq
    print("oracle: loss %.6f alpha %.4e |dx| %.4e |x_ddim| %.4f |x_new| %.4f" % (loss.item(), alpha.item(), dx_ref.norm().item(),
                                                                              x_ddim.norm().item(), x_new_ref.norm().item()))

    # ---------------- CUDA path, same stages
    xin_d = x_in.detach().cuda()
    eps = eng.forward(xin_d, t, emb.cuda(), save_for_backward=True)
    print("eps rel err            %.3e" % rel(eps, eps_ref))
    tps = eng.taps()
    for k in range(9):
        print("  tap %d rel err        %.3e" % (k, rel(tps[k].permute(0, 3, 1, 2), tap_ref[k])))
    sigma = float((1 - sch.alphas_cumprod[t]) ** 0.5)
    leng.forward_taps(tps, 2, L, lat.cuda().contiguous(), sigma, True)
    mine = leng.output(2, L, "cuda")
    print("LGP output rel err     %.3e" % rel(mine, out.float()))
    l, grads, scale = leng.loss_backward(tgt.cuda().contiguous(), tps)
    print("loss mine %.6f (oracle %.6f) scale %g" % (l.item(), loss.item(), scale))
    for k in range(9):
        gm = grads[k].permute(0, 3, 1, 2) / scale
        print("  tap grad %d rel err   %.3e  |g_ref| %.3e cos %.6f" % (
            k, rel(gm, g_tap[k]), g_tap[k].norm(), F.cosine_similarity(gm.flatten().double().cpu(), g_tap[k].flatten().double(), dim=0)))
    # (a) UNet backward fed with the ORACLE's tap gradients: isolates the UNet adjoint
    eng_dx_a = eng.backward([gt.permute(0, 2, 3, 1).contiguous().cuda() * scale for gt in g_tap]) / scale
    print("dx (oracle tap grads)  %.3e  cos %.6f" % (rel(eng_dx_a, dx_ref), F.cosine_similarity(
        eng_dx_a.flatten().double().cpu(), dx_ref.flatten().double(), dim=0)))
    # (b) full chain
    eng.forward(xin_d, t, emb.cuda(), save_for_backward=True)
    tps = eng.taps()
    leng.forward_taps(tps, 2, L, lat.cuda().contiguous(), sigma, True)
    l, grads, scale = leng.loss_backward(tgt.cuda().contiguous(), tps)
    dx = eng.backward(grads) / scale
    print("dx (full chain)        %.3e  cos %.6f" % (rel(dx, dx_ref), F.cosine_similarity(
        dx.flatten().double().cpu(), dx_ref.flatten().double(), dim=0)))
    # exact-math mode: loss-scaled gradients without the fp16 rounding emulation, against the fp32-LGP oracle
    leng.set_grad_rounding(False)
    eng.forward(xin_d, t, emb.cuda(), save_for_backward=True)
    tps = eng.taps()
    leng.forward_taps(tps, 2, L, lat.cuda().contiguous(), sigma, True)
    l, grads, scale = leng.loss_backward(tgt.cuda().contiguous(), tps)
    for k in range(9):
        gm = grads[k].permute(0, 3, 1, 2) / scale
        print("  [exact] tap grad %d vs fp32 oracle %.3e" % (k, rel(gm, g_tap32[k])))
    dx_e = eng.backward(grads) / scale
    print("dx [exact] vs fp32-LGP oracle %.3e cos %.6f ; vs fp16-LGP oracle %.3e cos %.6f" % (
        rel(dx_e, dx_ref32), cos(dx_e, dx_ref32), rel(dx_e, dx_ref), cos(dx_e, dx_ref)))
    leng.set_grad_rounding(True)
    # LGP alone on IDENTICAL inputs (the oracle's taps): implementation error without input sensitivity
    otaps = [tp.detach().permute(0, 2, 3, 1).contiguous().cuda() for tp in tap_ref]
    for emul, gref_, oref_, nm in ((True, g_tap, out.float(), "emulated vs fp16 oracle"), (False, g_tap32, out32, "exact vs fp32 oracle")):
        leng.set_grad_rounding(emul)
        leng.forward_taps(otaps, 2, L, lat.cuda().contiguous(), sigma, True)
        mo = leng.output(2, L, "cuda")
        l, grads, scale = leng.loss_backward(tgt.cuda().contiguous(), otaps)
        print("LGP on oracle taps [%s]: out %.3e  tap grads %s" % (nm, rel(mo, oref_), " ".join(
            "%.2e" % rel(grads[k].permute(0, 3, 1, 2) / scale, gref_[k]) for k in range(9))))
    leng.set_grad_rounding(True)
    # the ORACLE's own sensitivity: fp32 LGP fed with the engine's taps (1e-3 away from its own)
    etaps = [tp.permute(0, 3, 1, 2).cpu().clone().requires_grad_(True) for tp in eng.taps()]
    with torch.enable_grad():
        fe = torch.cat([F.interpolate(lf, size=L, mode="bilinear") for lf in etaps], dim=1)
        oe = lgp32(fe, torch.cat([lvl] * 2))
        le = F.mse_loss(tgt.float(), oe.reshape(2, L, L, -1).permute(0, 3, 2, 1).chunk(2)[1].float())
        ge = torch.autograd.grad(le, etaps)
    print("fp32 oracle LGP, engine taps vs oracle taps: out %.3e  tap grads %s" % (rel(oe, out32), " ".join(
        "%.2e" % rel(a, b) for a, b in zip(ge, g_tap32))))
    gq_m = (-dx).chunk(2)[1].cpu()
    alpha_m = torch.linalg.norm(x_in.detach() - x_ddim) / torch.linalg.norm(gq_m) * 1.6
    print("alpha mine %.4e oracle %.4e" % (alpha_m.item(), alpha.item()))
    print("x_new rel err (chain)  %.3e" % rel(x_ddim + alpha_m * gq_m, x_new_ref))


if __name__ == "__main__":
    main(*sys.argv[1:4])
