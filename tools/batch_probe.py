"""GPU box: images/s of the sketch-guided SD1.5 512x512 50-step job at several images per call (BASELINE.json configs[2] runs
4 per GPU): the batch is [uncond_0.., cond_0..] inside the engine and the input-gradient walk covers the cond samples only.
usage: python tools/batch_probe.py [S ...]      (default 1 2 4)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sketch2img_b200 import synthetic  # noqa: E402
from sketch2img_b200.latent_predictor import LatentEdgePredictor  # noqa: E402
from sketch2img_b200.pipeline import AntiGradientPipeline  # noqa: E402
from sketch2img_b200.scheduler import DDIMScheduler  # noqa: E402
from sketch2img_b200.unet import SD15_CONFIG, UNet2DConditionModel  # noqa: E402


def main(sizes):
    cfg = dict(SD15_CONFIG)
    unet = UNet2DConditionModel(cfg, synthetic.unet_state_dict(cfg))
    lgp = LatentEdgePredictor(synthetic.lgp_input_dim(cfg), 4, 9)
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    for S in sizes:
        pipe.max_samples_per_launch = S
        lat, emb, tgt = (t.cuda() for t in synthetic.sample_inputs(cfg, S))
        prompts = ["x"] * S
        run = lambda n: pipe(prompts, num_inference_steps=n, latents=lat, sketch_image=tgt, prompt_embeds=emb, output_type="latent")
        run(4)            # arena sizing for both step kinds
        run(50)           # graph capture
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run(50)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        assert torch.isfinite(out).all()
        print("S = %d images per call: %.1f ms per call, %.1f ms per image, %.2f images/s, arena %.1f GB" %
              (S, ms, ms / S, 1e3 * S / ms, unet.engine.arena_bytes() / 1e9), flush=True)


if __name__ == "__main__":
    main([int(a) for a in sys.argv[1:]] or [1, 2, 4])
