"""GPU box: run the drop-in pipeline on the seeded synthetic inputs and compare per-step latents with the
committed golden fixtures (tests/golden/*.pt, produced by the unmodified reference files on CPU)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402  (seeded weights / inputs only)
from sketch2img_b200.latent_predictor import LatentEdgePredictor  # noqa: E402
from sketch2img_b200.pipeline import AntiGradientPipeline  # noqa: E402
from sketch2img_b200.scheduler import DDIMScheduler  # noqa: E402
from sketch2img_b200.unet import UNet2DConditionModel  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def build(name):
    o_unet = port.make_unet(name)
    o_lgp = port.make_lgp(o_unet)
    lat, emb, tgt = port.make_inputs(o_unet)
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    lgp = LatentEdgePredictor(port.lgp_input_dim(o_unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(o_lgp.float().state_dict())
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    return pipe, lat, emb, tgt


def main(name="tiny", steps=4):
    steps = int(steps)
    pipe, lat, emb, tgt = build(name)
    gold = torch.load(os.path.join(ROOT, "tests", "golden", f"{name}_{steps}step.pt"))
    got = {}
    torch.cuda.synchronize()
    t0 = time.time()
    out = pipe("synthetic", num_inference_steps=steps, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
               prompt_embeds=emb.cuda(), output_type="latent",
               callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"[{name} {steps} steps] wall {dt:.2f}s (first call incl. arena setup); reference CPU {gold['cpu_seconds']:.1f}s")
    for i in sorted(gold["latents"]):
        print("step %3d  rel err %.3e   |x| gold %.4f  got %.4f" % (i, rel(got[i], gold["latents"][i]),
                                                                   gold["latents"][i].norm(), got[i].norm()))
    print("FINAL rel L2 err %.3e" % rel(out.cpu(), gold["latents"][steps - 1]))
    torch.cuda.synchronize()
    t0 = time.time()
    pipe("synthetic", num_inference_steps=steps, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
         prompt_embeds=emb.cuda(), output_type="latent")
    torch.cuda.synchronize()
    print(f"second call wall {time.time() - t0:.3f}s  -> {(time.time() - t0) / steps * 1e3:.1f} ms/step")


if __name__ == "__main__":
    main(*sys.argv[1:3])
