"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt.

Runs the reference's OWN files -- /root/reference/modules/pipeline.py (AntiGradientPipeline) and
/root/reference/modules/latent_predictor.py (LatentEdgePredictor, hook_unet) -- imported unmodified on
top of oracle/diffusers_shim, on CPU, with the seeded synthetic weights/inputs of oracle/port.py, and
stores the per-step latents it reports through ``callback`` (pipeline.py:112-115).  Only runnable in the
authoring container (needs /root/reference); the fixtures it writes travel with the repo.

    python oracle/make_golden.py tiny 4 tiny 50 sd15 4 sd15 50 tiny@dpmpp 4 tiny@dpmpp 20
    python oracle/make_golden.py teacher tiny 50 teacher sd15 50      # every guided step, for the teacher-forced tests
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


class _Stop(Exception):
    pass


def run_reference_teacher(name, steps, guidance_scale=7.5, seed=port.SAMPLE_SEED):
    """Every GUIDED step of the reference run, for teacher-forced tests (each CUDA step restarts from the reference's own
    previous latent, so the chaotic loop cannot compound errors).

    Part 1 -- the reference's files, unmodified: AntiGradientPipeline.__call__ with ``callback`` collecting the latent after
    every step; ``apply_anti_gradient`` is wrapped (an instance attribute in front of the method, which then runs unchanged)
    to note its ``latents`` argument (the scheduler output before guidance, pipeline.py:104) and ``torch.autograd.grad`` is
    wrapped for the duration of that call to note the loss and the gradient it returns (:157-159).  The run is cut after the
    last guided step by an exception raised from the callback.
    Part 2 -- yardsticks from oracle/port.py's ``guided_step`` (asserted bit-identical to part 1 on the same input):
      * ``fp16w``: the same step with the UNet weights rounded to fp16 -- the precision the reference itself runs at
        (app.py:32-38 loads the pipeline with torch_dtype=float16); the distance of that step from the fp32 one is what ANY
        fp16-weight implementation of this step is expected to show;
      * ``pert``: the same step from latents * (1 + 1e-6) (one-step self-sensitivity)."""
    import copy
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from modules.latent_predictor import LatentEdgePredictor
    from modules.pipeline import AntiGradientPipeline

    unet = port.make_unet(name)
    lgp_o = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet, seed)
    lgp = LatentEdgePredictor(port.lgp_input_dim(unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(copy.deepcopy(lgp_o).float().state_dict())
    lgp.half()
    pipe = AntiGradientPipeline(unet=unet, scheduler=port.make_scheduler())
    pipe.set_prompt_embeds(emb)
    pipe.setup_lgp(lgp)
    n_guided = sum(1 for i in range(steps) if i <= 0.5 * steps)                 # pipeline.py:90-92
    rec, cur = {}, {}
    orig_apply, orig_grad = pipe.apply_anti_gradient, torch.autograd.grad

    def apply(latents_prev, latents, noise, timestep, target, beta):
        cur.clear()
        cur["x_ddim"] = latents.detach().clone().float()

        def grad(loss, x, *a, **k):
            out = orig_grad(loss, x, *a, **k)
            cur["loss"] = float(loss)
            cur["gnorm"] = float(torch.linalg.norm(out[0].chunk(2)[1]))
            return out

        torch.autograd.grad = grad
        try:
            return orig_apply(latents_prev, latents, noise, timestep, target, beta)
        finally:
            torch.autograd.grad = orig_grad

    pipe.apply_anti_gradient = apply

    def cb(i, t, l):
        rec[int(i)] = dict(x=l.detach().clone().float(), t=int(t), **cur)
        if int(i) == n_guided - 1:
            raise _Stop

    t0 = time.perf_counter()
    try:
        pipe("synthetic", num_inference_steps=steps, guidance_scale=guidance_scale, latents=lat.clone(), sketch_image=tgt,
             output_type="np", callback=cb)
    except _Stop:
        pass
    dt = time.perf_counter() - t0
    for h in list(getattr(pipe, "feature_blocks", [])):          # drop the reference's hooks: the port registers its own
        h._forward_hooks.clear()

    # ---- part 2: port.guided_step from the reference's own latents
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    cosv = lambda a, b: torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()
    sch = port.make_scheduler()
    sch.set_timesteps(steps)
    unet16 = copy.deepcopy(unet)
    with torch.no_grad():
        for p_ in unet16.parameters():
            p_.copy_(p_.half().float())
    taps, _ = port.register_taps(unet)
    taps16, _ = port.register_taps(unet16)
    noise = lat.clone() * sch.init_noise_sigma
    out = dict(e16=[], d16=[], cos16=[], loss16=[], e_pert=[], cos_pert=[])
    for i in range(n_guided):
        x_prev = noise if i == 0 else rec[i - 1]["x"]
        t = sch.timesteps[i]
        assert int(t) == rec[i]["t"]
        r32 = {}
        x32 = port.guided_step(unet, lgp_o, sch, emb, x_prev, noise, t, tgt, True, taps, guidance_scale, 1.6, r32)
        assert torch.equal(x32, rec[i]["x"]) and torch.equal(r32["x_ddim"], rec[i]["x_ddim"]), \
            f"port.guided_step is not bit-identical to the reference's step {i}"
        r16 = {}
        x16 = port.guided_step(unet16, lgp_o, sch, emb, x_prev, noise, t, tgt, True, taps16, guidance_scale, 1.6, r16)
        rp = {}
        xp = port.guided_step(unet, lgp_o, sch, emb, x_prev * (1.0 + 1e-6), noise, t, tgt, True, taps, guidance_scale, 1.6, rp)
        upd = x32 - r32["x_ddim"]
        out["e16"].append(rel(x16, x32))
        out["d16"].append(rel(r16["x_ddim"], r32["x_ddim"]))
        out["cos16"].append(cosv(x16 - r16["x_ddim"], upd))
        out["loss16"].append(r16["loss"])
        out["e_pert"].append(rel(xp, x32))
        out["cos_pert"].append(cosv(xp - rp["x_ddim"], upd))
        print(f"{name} teacher step {i} t={int(t)}: loss {r32['loss']:.5f} |g| {rec[i]['gnorm']:.3e} | fp16-weight step: x {out['e16'][-1]:.2e} "
              f"x_ddim {out['d16'][-1]:.2e} cos {out['cos16'][-1]:.5f} | 1e-6 restart: x {out['e_pert'][-1]:.2e} cos {out['cos_pert'][-1]:.6f}",
              flush=True)
    return {
        "config": name, "steps": steps, "guided_steps": n_guided, "guidance_scale": guidance_scale, "beta": 1.6,
        "weight_seed": port.WEIGHT_SEED, "sample_seed": seed, "t": [rec[i]["t"] for i in range(n_guided)],
        "x": {i: rec[i]["x"] for i in range(n_guided)}, "x_ddim": {i: rec[i]["x_ddim"] for i in range(n_guided)},
        "loss": torch.tensor([rec[i]["loss"] for i in range(n_guided)], dtype=torch.float64),
        "gnorm": torch.tensor([rec[i]["gnorm"] for i in range(n_guided)], dtype=torch.float64),
        "fp16w": {k: torch.tensor(out[k], dtype=torch.float64) for k in ("e16", "d16", "cos16", "loss16")},
        "pert": {k: torch.tensor(out[k], dtype=torch.float64) for k in ("e_pert", "cos_pert")},
        "cpu_seconds": dt, "cpu_threads": torch.get_num_threads(),
        "source": "reference modules/pipeline.py + latent_predictor.py over oracle/diffusers_shim (every guided step); "
                  "yardsticks from oracle/port.py guided_step (bit-identical to the reference step on the same input)",
    }


def run_reference_sketch_encoder(name="tiny21"):
    """The reference's OWN SketchEncoder class (modules/sketch_encoder.py, unmodified) over the shim, attention-free down blocks,
    on a seeded sketch latent at two timesteps."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from modules.sketch_encoder import SketchEncoder
    enc = port.make_sketch_encoder(name, cls=SketchEncoder)
    gen = torch.Generator().manual_seed(port.SAMPLE_SEED + 11)
    L = enc.config.sample_size
    x = torch.randn(2, 4, L, L, generator=gen)
    outs = {}
    with torch.no_grad():
        for t in (0, 500):
            outs[t] = [tuple(m.clone() for m in tup) for tup in enc(x, t).sample]
    return {"config": name, "x": x, "timesteps": [0, 500], "res_samples": outs, "weight_seed": port.WEIGHT_SEED + 3,
            "source": "reference modules/sketch_encoder.py (SketchEncoder, down_block_types all DownBlock2D) over oracle/diffusers_shim"}


def run_reference_vae(name="tiny"):
    """Either side of the loop with the reference's own code where it has any: the sketch target as app.py:107-109 forms it
    (``vae.encode(img).latent_dist`` -- the posterior's moments are stored, its ``sample()`` draws from the global RNG) and the
    reference's ``AntiGradientPipeline.decode_latents_L`` (modules/pipeline.py:163-174, unmodified) on a seeded latent."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from modules.pipeline import AntiGradientPipeline
    vae = port.make_vae(name)
    gen = torch.Generator().manual_seed(port.SAMPLE_SEED + 21)
    S = vae.config.sample_size
    img = torch.rand(1, 1, S, S, generator=gen).round()                  # a binary "sketch" like app.py's sketchpad image
    img = (torch.tile(img, (1, 3, 1, 1)) - 0.5) / 0.5                    # app.py:27-30, :108: ToTensor / Normalize(0.5, 0.5), 3 channels
    lat = torch.randn(1, 4, S // 8, S // 8, generator=gen) * 0.18215 * 4
    with torch.no_grad():
        dist = vae.encode(img).latent_dist
        pipe = AntiGradientPipeline(unet=port.make_unet("tiny"), scheduler=port.make_scheduler())
        pipe.vae = vae
        image_L = pipe.decode_latents_L(lat)
        decoded = vae.decode(lat / 0.18215).sample
    return {"config": name, "image": img, "moments": dist.parameters.clone(), "latents": lat, "decoded": decoded,
            "image_L": torch.from_numpy(image_L.copy()), "weight_seed": port.WEIGHT_SEED + 4,
            "source": "oracle/diffusers_shim AutoencoderKL + reference modules/pipeline.py decode_latents_L"}


def main(argv):
    torch.set_num_threads(os.cpu_count())
    if argv and argv[0] == "vae":
        blob = run_reference_vae()
        path = os.path.join(ROOT, "tests", "golden", "tiny_vae.pt")
        torch.save(blob, path)
        print("vae fixture ->", path, tuple(blob["moments"].shape), tuple(blob["decoded"].shape), blob["image_L"].shape, blob["image_L"].float().mean())
        return
    if argv and argv[0] == "sketch_encoder":
        blob = run_reference_sketch_encoder()
        path = os.path.join(ROOT, "tests", "golden", "tiny21_sketch_encoder.pt")
        torch.save(blob, path)
        print("sketch encoder fixture ->", path, [tuple(m.shape) for tup in blob["res_samples"][0] for m in tup])
        return
    if argv and argv[0] == "teacher":
        for name, steps in zip(argv[1::3], argv[2::3]):
            blob = run_reference_teacher(name, int(steps))
            path = os.path.join(ROOT, "tests", "golden", f"{name}_{int(steps)}step_teacher.pt")
            torch.save(blob, path)
            print("teacher fixture ->", path, flush=True)
        return
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
