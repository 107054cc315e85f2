"""GPU box: device time of one guided / one unguided denoising step of the SD1.5 512x512 job, launched the normal way
(host enqueues ~1000 kernels) and replayed from a CUDA graph captured around the same s2i_sampler_step call."""
import ctypes as C
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


def main():
    cfg = dict(SD15_CONFIG)
    unet = UNet2DConditionModel(cfg, synthetic.unet_state_dict(cfg))
    lgp = LatentEdgePredictor(synthetic.lgp_input_dim(cfg), 4, 9)
    sch = DDIMScheduler()
    pipe = AntiGradientPipeline(unet=unet, scheduler=sch)
    pipe.setup_lgp(lgp)
    lat, emb, tgt = (t.cuda() for t in synthetic.sample_inputs(cfg, 1))
    pipe("x", num_inference_steps=4, latents=lat, sketch_image=tgt, prompt_embeds=emb, output_type="latent")   # warm-up
    torch.cuda.synchronize()
    lib = _lib.lib()
    sampler = pipe._get_sampler()
    sch.set_timesteps(50)
    latents = lat.clone().float().contiguous()
    noise = latents.clone()
    ctx = emb.float().contiguous()
    target = tgt.float().contiguous()
    loss = torch.zeros(1, device="cuda")
    L = latents.shape[2]
    side = torch.cuda.Stream()

    def step(guided, stream):
        ti = 981
        sa_t, sb_t, sa_p, sb_p = sch.step_coefficients(ti)
        _lib.check(lib.s2i_sampler_step(sampler, latents.data_ptr(), noise.data_ptr(), ctx.data_ptr(), target.data_ptr(), 1, L,
                                        float(ti), 7.5, sa_t, sb_t, sa_p, sb_p, 0, guided, sch.sigma(ti), 1.6, 1,
                                        loss.data_ptr(), C.c_void_p(stream.cuda_stream)))

    for guided in (1, 0):
        with torch.cuda.stream(side):
            for _ in range(2):
                step(guided, side)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = _lib.launch_count()
            e0.record()
            for _ in range(5):
                step(guided, side)
            e1.record()
            torch.cuda.synchronize()
            launches = (_lib.launch_count() - n0) // 5
            t_stream = e0.elapsed_time(e1) / 5
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step(guided, side)
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            t_graph = e0.elapsed_time(e1) / 5
        print(f"{'guided' if guided else 'unguided'} step: {launches} launches, streamed {t_stream:.3f} ms, graph replay {t_graph:.3f} ms",
              flush=True)


if __name__ == "__main__":
    main()
