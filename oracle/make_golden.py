"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt.

Runs the reference's OWN files -- /root/reference/modules/pipeline.py (AntiGradientPipeline) and
/root/reference/modules/latent_predictor.py (LatentEdgePredictor, hook_unet) -- imported unmodified on
top of oracle/diffusers_shim, on CPU, with the seeded synthetic weights/inputs of oracle/port.py, and
stores the per-step latents it reports through ``callback`` (pipeline.py:112-115).  Only runnable in the
authoring container (needs /root/reference); the fixtures it writes travel with the repo.

    python oracle/make_golden.py tiny 4 tiny 50 sd15 4 sd15 50 tiny@dpmpp 4 tiny@dpmpp 20
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import port  # noqa: E402

REF = "/root/reference"


def run_reference(name, steps, guidance_scale=7.5, seed=port.SAMPLE_SEED, guided=True, latent_scale=1.0, kind="ddim"):
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from modules.latent_predictor import LatentEdgePredictor
    from modules.pipeline import AntiGradientPipeline

    unet = port.make_unet(name)
    lgp_o = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet, seed)
    lgp = LatentEdgePredictor(port.lgp_input_dim(unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(lgp_o.float().state_dict())
    lgp.half()
    if kind == "dpmpp":
        # the demo's scheduler, constructed with the reference's own keyword arguments (app.py:14-25)
        from diffusers import DPMSolverMultistepScheduler
        sched = DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                            num_train_timesteps=1000, trained_betas=None, predict_epsilon=True,
                                            thresholding=False, algorithm_type="dpmsolver++", solver_type="midpoint",
                                            lower_order_final=True)
    else:
        sched = port.make_scheduler()
    pipe = AntiGradientPipeline(unet=unet, scheduler=sched)
    pipe.set_prompt_embeds(emb)
    pipe.setup_lgp(lgp)
    per_step = {}
    t0 = time.perf_counter()
    pipe("synthetic", num_inference_steps=steps, guidance_scale=guidance_scale, latents=lat.clone() * latent_scale,
         sketch_image=tgt if guided else None, output_type="np",
         callback=lambda i, t, l: per_step.__setitem__(int(i), l.detach().clone().float()))
    dt = time.perf_counter() - t0
    return per_step, dt


def run_reference_sat(name="tiny21", steps=4, guidance_scale=7.5, scale=0.7, seed=port.SAMPLE_SEED):
    """Config 4: the reference's OWN SatMixin (modules/sketch_guided_attn.py, unmodified) injected into the shim UNet,
    sampled with plain CFG + v-prediction DDIM (the reference's pipeline with sketch_image=None: pipeline.py:142-143)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from modules.pipeline import AntiGradientPipeline
    from modules.sketch_guided_attn import SatMixin

    unet = port.make_unet(name)
    sat = port.make_sat(unet, SatMixin)
    lat, emb, _ = port.make_inputs(unet, seed)
    res = port.make_res_samples(unet, 2)
    sat.set_res_samples(res)
    sat.set_scale(scale)
    with torch.no_grad():
        eps = unet(torch.cat([lat] * 2), torch.tensor(501), encoder_hidden_states=emb).sample
    pipe = AntiGradientPipeline(unet=unet, scheduler=port.make_scheduler("v_prediction"))
    pipe.set_prompt_embeds(emb)
    from modules.latent_predictor import LatentEdgePredictor
    pipe.setup_lgp(LatentEdgePredictor(port.lgp_input_dim(unet), 4, port.NUM_POS_LAYERS).half())   # required by :38, unused
    per_step = {}
    pipe("synthetic", num_inference_steps=steps, guidance_scale=guidance_scale, latents=lat.clone(), sketch_image=None,
         output_type="np", callback=lambda i, t, l: per_step.__setitem__(int(i), l.detach().clone().float()))
    return {"config": name, "steps": steps, "guidance_scale": guidance_scale, "sat_scale": scale, "eps_t501": eps,
            "latents": per_step, "weight_seed": port.WEIGHT_SEED, "sat_seed": port.WEIGHT_SEED + 2, "sample_seed": seed,
            "source": "reference modules/sketch_guided_attn.py (SatMixin) + modules/pipeline.py over oracle/diffusers_shim"}


def main(argv):
    torch.set_num_threads(os.cpu_count())
    if argv and argv[0] == "sat":
        blob = run_reference_sat()
        path = os.path.join(ROOT, "tests", "golden", "tiny21_sat_4step.pt")
        torch.save(blob, path)
        print("sat fixture ->", path, "final norm", blob["latents"][blob["steps"] - 1].norm().item())
        return
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    pairs = list(zip(argv[0::2], argv[1::2]))
    for name, steps in pairs:
        steps = int(steps)
        name, _, kind = name.partition("@")          # "tiny@dpmpp": the demo's DPM-Solver++(2M) instead of DDIM
        kind = kind or "ddim"
        per_step, dt = run_reference(name, steps, kind=kind)
        keep = sorted(set(list(range(0, steps, max(1, steps // 10))) + [steps - 1]))
        # The guided loop is chaotic (DESIGN.md "Conditioning"): the SAME reference files, started from latents scaled
        # by (1 + 1e-6), drift away from the run above.  That drift is the noise floor any other implementation is
        # measured against.  The unguided run (sketch_image=None: plain CFG + DDIM) is smooth and pins the UNet +
        # scheduler over the full schedule.
        pert, _ = run_reference(name, steps, latent_scale=1.0 + 1e-6, kind=kind)
        rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
        unguided, dt_u = run_reference(name, steps, guided=False, kind=kind)
        blob = {
            "self_sensitivity": torch.tensor([rel(pert[i], per_step[i]) for i in range(steps)]),
            "unguided_latents": {i: unguided[i] for i in keep}, "unguided_cpu_seconds": dt_u,
            "config": name, "steps": steps, "guidance_scale": 7.5, "beta": 1.6, "scheduler": kind,
            "weight_seed": port.WEIGHT_SEED, "sample_seed": port.SAMPLE_SEED,
            "latents": {i: per_step[i] for i in keep},
            "norms": torch.tensor([per_step[i].norm().item() for i in range(steps)]),
            "cpu_seconds": dt, "cpu_threads": torch.get_num_threads(),
            "source": "reference modules/pipeline.py + latent_predictor.py over oracle/diffusers_shim",
        }
        path = os.path.join(out_dir, f"{name}_{steps}step.pt" if kind == "ddim" else f"{name}_{kind}_{steps}step.pt")
        torch.save(blob, path)
        print(f"{name} {steps} steps: {dt:.1f}s  final norm {per_step[steps - 1].norm():.4f} -> {path}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or ["tiny", "4", "tiny", "50"])
