"""GPU box, under ncu: a few denoising steps of the SD1.5 512x512 job (synthetic weights) with the profiler range
opened only around them (`ncu --profile-from-start off`).  usage: one_step.py [steps]  (3 -> 2 guided + 1 unguided; 1 -> 1 guided)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import _lib, synthetic  # noqa: E402
from sketch2img_b200.latent_predictor import LatentEdgePredictor  # noqa: E402
from sketch2img_b200.pipeline import AntiGradientPipeline  # noqa: E402
from sketch2img_b200.scheduler import DDIMScheduler  # noqa: E402
from sketch2img_b200.unet import SD15_CONFIG, UNet2DConditionModel  # noqa: E402


def main(steps=3):
    steps = int(steps)
    cfg = dict(SD15_CONFIG)
    unet = UNet2DConditionModel(cfg, synthetic.unet_state_dict(cfg))
    lgp = LatentEdgePredictor(synthetic.lgp_input_dim(cfg), 4, 9)
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    lat, emb, tgt = (t.cuda() for t in synthetic.sample_inputs(cfg, 1))
    # warm-up (arena sizing for both step kinds) outside the profiled range: 4 steps = 3 guided + 1 unguided
    pipe("x", num_inference_steps=4, latents=lat, sketch_image=tgt, prompt_embeds=emb, output_type="latent")
    torch.cuda.synchronize()
    guided = sum(1 for i in range(steps) if i <= 0.5 * steps)        # pipeline.py:90
    torch.cuda.profiler.start()
    n0 = _lib.launch_count()
    pipe("x", num_inference_steps=steps, latents=lat, sketch_image=tgt, prompt_embeds=emb, output_type="latent")
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled %d steps (%d guided, %d unguided): %d libs2i launches" % (steps, guided, steps - guided,
                                                                              _lib.launch_count() - n0))


if __name__ == "__main__":
    main(*sys.argv[1:2])
