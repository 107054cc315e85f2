"""Teacher-forced parity of EVERY guided step (reference loop body modules/pipeline.py:83-115, guidance :141-161).

The guided loop is chaotic (DESIGN.md "Conditioning"): the reference's own files, restarted from latents * (1 + 1e-6),
move 1e-3 .. 9e-3 away within ONE step, so a free-running trajectory cannot be held to a tolerance by any implementation.
Here nothing compounds: for each guided step i the CUDA sampler (one C-ABI call, s2i_sampler_step) restarts from the
REFERENCE's latent x_{i-1} (fixtures tests/golden/*_teacher.pt, written by oracle/make_golden.py from the unmodified
reference files) and its x_i is compared with the reference's x_i:

  * the scheduler output before guidance (an unguided call from the same state): 1e-3 relative -- north_star's tolerance;
  * the edge loss (modules/pipeline.py:157): 1e-3 relative (measured: at most 6.5e-4 over the 52 steps of both fixtures);
  * the length of the guidance update ||x_i - x_ddim||: 2 % (norm-ratio rule, :160);
  * its direction and the latent itself against the yardstick the fixture carries for that very step: the reference's own
    step recomputed with its UNet weights rounded to fp16 -- the precision the reference ships with (app.py:32-38,
    torch_dtype=float16).  An fp16-operand implementation is expected to sit at that distance; the assertions allow
    kMeanFactor x the yardstick on average over the guided steps and kErrFactor x at any single step.
  * run-to-run: the same call twice gives the same bits (fixed-order reductions everywhere).
"""
import copy
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# (CUDA distance) / (fp16-weight reference distance).  Both are single draws of a heavy-tailed one-step response (the tiny model's
# yardstick itself ranges 4e-3 .. 4e-2 over the 26 steps), so the MEAN over the guided steps is the sharp statement -- measured 1.29
# (tiny), 1.12 (SD1.5) -- and the per-step bound catches gross errors: measured max 2.50 (tiny, step 18), 1.24 (SD1.5).
kErrFactor = {"tiny": 4.0, "sd15": 2.0}
kMeanFactor = 1.5
kAngleFloor = 0.02      # radians: below this the yardstick itself is rounding noise


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    return F.cosine_similarity(a.double().cpu().flatten(), b.double().cpu().flatten(), dim=0).item()


def _build(name):
    from oracle import port
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    from sketch2img_b200.unet import UNet2DConditionModel
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    o_unet = port.make_unet(name)
    o_lgp = port.make_lgp(o_unet)
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    lgp = LatentEdgePredictor(port.lgp_input_dim(o_unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(copy.deepcopy(o_lgp).float().state_dict())
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    inputs = port.make_inputs(o_unet)
    del o_unet
    return pipe, inputs


@pytest.fixture(scope="module")
def tiny_pipe(cuda):
    return _build("tiny")


@pytest.fixture(scope="module")
def sd15_pipe(cuda):
    return _build("sd15")


class Stepper:
    """One denoising step through the C ABI from explicit state: what AntiGradientPipeline.__call__ does per iteration."""

    def __init__(self, pipe, emb, steps):
        from sketch2img_b200 import _lib
        self.L, self.pipe = _lib, pipe
        self.sampler = pipe._get_sampler()
        _lib.check(_lib.lib().s2i_sampler_context_changed(self.sampler))
        pipe.scheduler.set_timesteps(steps)
        self.timesteps = [int(t) for t in pipe.scheduler.timesteps]
        e = emb.cuda().float()
        self.ctx = torch.stack([e[:1], e[1:]], dim=1).reshape(2, e.shape[1], e.shape[2]).contiguous()   # (uncond, cond)
        self.loss = torch.zeros(1, device="cuda")

    def __call__(self, x_prev, noise, target, i, guided):
        L, sch = self.L, self.pipe.scheduler
        t = self.timesteps[i]
        sa_t, sb_t, sa_p, sb_p = sch.step_coefficients(t)
        lat = x_prev.cuda().float().contiguous().clone()
        L.check(L.lib().s2i_sampler_step(self.sampler, lat.data_ptr(), noise.data_ptr(), self.ctx.data_ptr(),
                                         target.data_ptr() if guided else None, 1, lat.shape[2], float(t), 7.5, sa_t, sb_t,
                                         sa_p, sb_p, sch.prediction, int(guided), sch.sigma(t), 1.6, 1, self.loss.data_ptr(),
                                         L.stream_ptr()))
        torch.cuda.synchronize()
        return lat, (self.loss.item() if guided else None)


def _teacher_forced(pipe, inputs, fix, label):
    lat, emb, tgt = inputs
    per_step = kErrFactor[fix["config"]]
    step = Stepper(pipe, emb, fix["steps"])
    assert step.timesteps[:fix["guided_steps"]] == list(fix["t"])
    noise = (lat * pipe.scheduler.init_noise_sigma).cuda().float().contiguous()
    target = tgt.cuda().float().contiguous()
    rows, ratios_e, ratios_a = [], [], []
    for i in range(fix["guided_steps"]):
        x_prev = noise if i == 0 else fix["x"][i - 1]
        x_ref, xd_ref = fix["x"][i], fix["x_ddim"][i]
        xd, _ = step(x_prev, noise, target, i, guided=False)
        x, loss = step(x_prev, noise, target, i, guided=True)
        x2, loss2 = step(x_prev, noise, target, i, guided=True)
        assert torch.equal(x, x2) and loss == loss2, f"{label} step {i}: the same call gave different bits"
        upd, upd_ref = x.cpu() - xd.cpu(), x_ref - xd_ref
        d_ddim = rel(xd, xd_ref)
        e = rel(x, x_ref)
        cs = cosine(upd, upd_ref)
        ang = math.acos(max(-1.0, min(1.0, cs)))
        e16, cos16 = float(fix["fp16w"]["e16"][i]), float(fix["fp16w"]["cos16"][i])
        ang16 = math.acos(max(-1.0, min(1.0, cos16)))
        len_ratio = upd.norm().item() / upd_ref.norm().item()
        loss_err = abs(loss - float(fix["loss"][i])) / float(fix["loss"][i])
        rows.append((i, fix["t"][i], d_ddim, loss_err, len_ratio, cs, cos16, e, e16))
        ratios_e.append(e / e16)
        ratios_a.append(ang / max(ang16, kAngleFloor))
    print(f"\n{label}: teacher-forced guided steps (each restarted from the reference's previous latent)")
    print("  i    t   x_ddim err  loss err  |upd| ratio  cos(upd)   cos(fp16w ref)   x err      fp16w ref err")
    for r in rows:
        print("  %2d  %4d  %.2e    %.2e  %.4f       %.5f    %.5f         %.2e   %.2e" % r)
    print("  mean x-err ratio to the fp16-weight yardstick %.2f (max %.2f); mean angle ratio %.2f (max %.2f)" % (
        sum(ratios_e) / len(ratios_e), max(ratios_e), sum(ratios_a) / len(ratios_a), max(ratios_a)))
    for (i, t, d_ddim, loss_err, len_ratio, cs, cos16, e, e16), re_, ra in zip(rows, ratios_e, ratios_a):
        assert d_ddim < 1e-3, f"{label} step {i}: scheduler output before guidance off by {d_ddim:.2e}"
        assert loss_err < 1e-3, f"{label} step {i}: edge loss off by {loss_err:.2e}"
        assert abs(len_ratio - 1.0) < 2e-2, f"{label} step {i}: guidance step length ratio {len_ratio:.4f}"
        assert re_ < per_step, f"{label} step {i}: latent error {e:.2e} vs fp16-weight yardstick {e16:.2e}"
        assert ra < per_step, f"{label} step {i}: update direction cosine {cs:.5f} vs yardstick {cos16:.5f}"
    assert sum(ratios_e) / len(ratios_e) < kMeanFactor
    assert sum(ratios_a) / len(ratios_a) < kMeanFactor


def test_tiny_every_guided_step_teacher_forced(tiny_pipe):
    fix = torch.load(os.path.join(GOLD, "tiny_50step_teacher.pt"))
    _teacher_forced(tiny_pipe[0], tiny_pipe[1], fix, "tiny 50-step")


def test_sd15_every_guided_step_teacher_forced(sd15_pipe):
    """BASELINE.json configs[1]: all 26 guided steps of the 50-step SD1.5 run."""
    fix = torch.load(os.path.join(GOLD, "sd15_50step_teacher.pt"))
    _teacher_forced(sd15_pipe[0], sd15_pipe[1], fix, "sd15 50-step")


@pytest.mark.parametrize("guided", [False, True])
def test_four_step_run_is_bitwise_reproducible(tiny_pipe, sd15_pipe, guided):
    """Same seed, same bits: the step has no order-dependent reduction (cluster split-K in rank order, fixed-order
    BatchNorm / loss / norm sums), eager first call, graph capture and graph replay included."""
    for label, (pipe, (lat, emb, tgt)) in (("tiny", tiny_pipe), ("sd15", sd15_pipe)):
        outs = []
        for rep in range(3):
            traj = []
            pipe("synthetic", num_inference_steps=4, guidance_scale=7.5, latents=lat.cuda(),
                 sketch_image=tgt.cuda() if guided else None, prompt_embeds=emb.cuda(), output_type="latent",
                 callback=lambda i, t, l: traj.append(l.detach().clone()))
            outs.append(torch.stack(traj))
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), f"{label} guided={guided}: runs differ"
