"""GPU box: LGP engine (features, MLP with train-mode BN, edge loss, backward to the taps) against the CPU oracle,
both fed the ORACLE's taps so that only the LGP path is compared."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402
from sketch2img_b200.latent_predictor import LGPEngine  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def main(name="tiny", t=981):
    torch.set_num_threads(os.cpu_count())
    unet = port.make_unet(name)
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    sch = port.make_scheduler()
    sch.set_timesteps(50)
    taps, _ = port.register_taps(unet)
    x = torch.cat([lat] * 2).requires_grad_(True)
    L = lat.shape[2]
    with torch.enable_grad():
        unet(x, torch.tensor(t), encoder_hidden_states=emb)
        tap_out = [m.output for m in taps]
        feats = port.lgp_features(taps, L)
        lvl = port.noise_level(sch, lat, torch.tensor(t))
        out = lgp(feats, torch.cat([lvl] * 2))                       # [(b w h), 4] fp16
        o4 = out.reshape(2, L, L, -1).permute(0, 3, 2, 1)
        loss = F.mse_loss(tgt.float(), o4.chunk(2)[1].float())
        g_ref = torch.autograd.grad(loss, tap_out)
    print(f"[{name}] oracle loss {loss.item():.6f}")

    eng = LGPEngine(port.lgp_input_dim(unet), 4, port.NUM_POS_LAYERS, lgp.float().state_dict())
    taps_nhwc = [tp.detach().permute(0, 2, 3, 1).contiguous().cuda() for tp in tap_out]
    sigma = float((1 - sch.alphas_cumprod[t]) ** 0.5)
    eng.forward_taps(taps_nhwc, 2, L, lat.cuda().contiguous(), sigma, True)
    mine = eng.output(2, L, "cuda").cpu()                            # rows in (b w h) order
    print("LGP output rel err %.3e   (|out| %.3f)" % (rel(mine, out.float()), out.float().norm()))
    l, grads, scale = eng.loss_backward(tgt.cuda().contiguous(), taps_nhwc)
    torch.cuda.synchronize()
    print("loss mine %.6f  scale %g" % (l.item(), scale))
    for k in range(9):
        gm = grads[k].permute(0, 3, 1, 2).cpu() / scale
        print("tap grad %d rel err %.3e  |g_ref| %.3e  cos %.6f" % (
            k, rel(gm, g_ref[k]), g_ref[k].norm(),
            F.cosine_similarity(gm.flatten().double(), g_ref[k].flatten().double(), dim=0)))
    # forward() surface on concatenated features (b=2 -> statistics over both halves, same as the pipeline)
    out2 = eng.forward_nchw(feats.detach().cuda().contiguous(), torch.cat([lvl] * 2).cuda().contiguous(), 2, L, True).cpu()
    print("forward_nchw rel err %.3e" % rel(out2, out.float()))


if __name__ == "__main__":
    main(*sys.argv[1:2])
