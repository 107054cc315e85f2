"""Parity of the CUDA sampling path (through the C ABI of include/s2i.h) against the CPU oracle (oracle/port.py)
and the golden fixtures made by the reference's own files (tests/golden/, oracle/make_golden.py).

Tolerances (relative L2, fp16 tensor-core operands with fp32 accumulation vs the fp32 CPU oracle):
  UNet eps / taps 3e-3, UNet input gradient 6e-3, LGP output 3e-3 on identical inputs; the scheduler and
  guidance-update kernels are compared bit-for-bit / to 1e-6.
Conditioning (DESIGN.md "Conditioning of the guided loop"): the reference's guidance gradient is ill-conditioned --
the oracle's OWN fp32 LGP gradient moves by 10-16 % when its taps move by 1.5e-3, its fp16 and fp32 LGP gradients
differ by 5 %, and the same reference files restarted from latents * (1 + 1e-6) drift 3-50 % apart within 4 steps
(committed as `self_sensitivity` in tests/golden/*.pt).  So: LGP gradients are checked on identical inputs against
BOTH the fp16 oracle (rounding emulated) and an fp32 restatement (exact mode) with the oracle's own fp16-vs-fp32
distance as the yardstick; UNGUIDED trajectories (smooth) are checked over the full schedule at 5e-3; GUIDED
trajectories are checked against the golden run at the oracle's own noise floor.
"""
import copy
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def tiny(cuda):
    from oracle import port
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    from sketch2img_b200.unet import UNet2DConditionModel
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    o_unet = port.make_unet("tiny")
    o_lgp = port.make_lgp(o_unet)
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    lgp = LatentEdgePredictor(port.lgp_input_dim(o_unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(copy.deepcopy(o_lgp).float().state_dict())     # nn.Module.float() converts in place
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    return dict(port=port, o_unet=o_unet, o_lgp=o_lgp, unet=unet, lgp=lgp, pipe=pipe, inputs=port.make_inputs(o_unet))


# ------------------------------------------------------------------------------------------------ UNet
def _check_unet_against_oracle(fx, t=981, label="tiny"):
    port, o_unet, eng = fx["port"], fx["o_unet"], fx["unet"].engine
    lat, emb, _ = fx["inputs"]
    x = torch.cat([lat] * 2)
    taps, handles = port.register_taps(o_unet)
    xg = x.clone().requires_grad_(True)
    with torch.enable_grad():
        eps_ref = o_unet(xg, torch.tensor(t), encoder_hidden_states=emb).sample
        tap_ref = [m.output for m in taps]
        gen = torch.Generator().manual_seed(7)
        G = [torch.randn(tr.shape, generator=gen) for tr in tap_ref]
        dx_ref = torch.autograd.grad(sum((g * tr).sum() for g, tr in zip(G, tap_ref)), xg)[0]
    for h in handles:
        h.remove()
    eps = eng.forward(x.cuda(), t, emb.cuda(), save_for_backward=True)
    e_eps = rel(eps, eps_ref)
    assert e_eps < 3e-3
    e_taps = []
    for k in range(9):
        got = eng.tap(k).permute(0, 3, 1, 2)
        assert got.shape == tap_ref[k].shape
        e_taps.append(rel(got, tap_ref[k].detach()))
        assert e_taps[-1] < 3e-3, f"tap {k}"
    Gd = [g.permute(0, 2, 3, 1).contiguous().cuda() for g in G]
    dx = eng.backward(Gd)
    e_dx = rel(dx, dx_ref)
    assert e_dx < 6e-3
    # the sampler's form: the cond sample alone (s2i_unet_backward_samples), against the oracle's gradient of that sample
    eng.forward(x.cuda(), t, emb.cuda(), save_for_backward=True)
    part = eng.backward([g[1:2].contiguous() for g in Gd], samples=(1, 1))
    e_cond = rel(part, dx_ref[1:2])
    assert e_cond < 6e-3
    print("%s UNet vs oracle: eps %.2e, taps %s, dx %.2e, cond-only dx %.2e" % (
        label, e_eps, " ".join("%.1e" % e for e in e_taps), e_dx, e_cond))
    return x, emb, Gd, dx


def test_unet_forward_taps_backward_match_oracle(tiny):
    eng = tiny["unet"].engine
    x, emb, Gd, dx = _check_unet_against_oracle(tiny)
    # linearity of the tap->input adjoint (size-independent property): J^T(2g) == 2 J^T(g)
    eng.forward(x.cuda(), 981, emb.cuda(), save_for_backward=True)
    dx2 = eng.backward([2 * g for g in Gd])
    assert rel(dx2, 2 * dx) < 3e-3
    # a second backward without a new forward is a state error, not silent garbage
    from sketch2img_b200._lib import S2IError
    with pytest.raises(S2IError):
        eng.backward(Gd)


def test_unet_backward_over_a_sample_range_equals_the_slice_of_the_whole_walk(tiny):
    """pipeline.py:159 keeps only the cond half of the latent gradient, and samples are independent computations: the
    sampler therefore walks the cond sample alone (s2i_unet_backward_samples).  It must equal that sample's slice of the
    whole-batch walk, and the oracle's autograd gradient of that sample."""
    port, o_unet, eng = tiny["port"], tiny["o_unet"], tiny["unet"].engine
    lat, emb, _ = tiny["inputs"]
    g = torch.Generator().manual_seed(13)
    x = torch.cat([lat] * 2) + 0.05 * torch.randn(2, *lat.shape[1:], generator=g)
    taps, handles = port.register_taps(o_unet)
    xg = x.clone().requires_grad_(True)
    with torch.enable_grad():
        o_unet(xg, torch.tensor(501), encoder_hidden_states=emb)
        tap_ref = [m.output for m in taps]
        G = [torch.randn(t.shape, generator=g) for t in tap_ref]
        dx_ref = torch.autograd.grad(sum((gg * t).sum() for gg, t in zip(G, tap_ref)), xg)[0]
    for h in handles:
        h.remove()
    Gd = [gg.permute(0, 2, 3, 1).contiguous().cuda() for gg in G]
    eng.forward(x.cuda(), 501, emb.cuda(), save_for_backward=True)
    whole = eng.backward(Gd)
    for b0 in (1, 0):
        eng.forward(x.cuda(), 501, emb.cuda(), save_for_backward=True)
        part = eng.backward([gg[b0:b0 + 1].contiguous() for gg in Gd], samples=(b0, 1))
        assert tuple(part.shape) == (1,) + tuple(x.shape[1:])
        assert rel(part, whole[b0:b0 + 1]) < 3e-3, f"sample {b0} vs whole-batch walk"
        assert rel(part, dx_ref[b0:b0 + 1]) < 6e-3, f"sample {b0} vs oracle"
    from sketch2img_b200._lib import S2IError
    eng.forward(x.cuda(), 501, emb.cuda(), save_for_backward=True)
    with pytest.raises(S2IError):
        eng.backward([gg[:1].contiguous() for gg in Gd], samples=(2, 1))       # outside the forward's batch


def test_unet_forward_is_reproducible_and_batch_independent(tiny):
    eng = tiny["unet"].engine
    lat, emb, _ = tiny["inputs"]
    g = torch.Generator().manual_seed(11)
    x4 = torch.randn(4, 4, lat.shape[2], lat.shape[3], generator=g).cuda()
    e4 = torch.randn(4, 77, emb.shape[2], generator=g).cuda()
    a = eng.forward(x4, 501, e4).clone()
    b = eng.forward(x4, 501, e4).clone()
    # every reduction of the forward has a fixed order (GroupNorm partials through distributed shared memory in rank order,
    # split-K partial tiles likewise): the same input gives the same bits
    assert torch.equal(a, b)
    # samples are closed computations: a batch equals its samples run alone
    for i in (0, 3):
        one = eng.forward(x4[i:i + 1], 501, e4[i:i + 1])
        assert rel(one, a[i:i + 1]) < 3e-3


# ------------------------------------------------------------------------------------------------ LGP
def _oracle_lgp(tiny, t=981):
    """Oracle LGP (fp16 = the reference; fp32 = exact-math restatement) on the oracle's own taps as detached leaves:
    the LGP's partial gradients, without the UNet paths between taps."""
    port, o_unet, o_lgp = tiny["port"], tiny["o_unet"], tiny["o_lgp"]
    lat, emb, tgt = tiny["inputs"]
    sch = port.make_scheduler()
    sch.set_timesteps(50)
    taps, handles = port.register_taps(o_unet)
    L = lat.shape[2]
    with torch.no_grad():
        o_unet(torch.cat([lat] * 2), torch.tensor(t), encoder_hidden_states=emb)
    tap_vals = [m.output.detach().clone() for m in taps]
    for h in handles:
        h.remove()
    lvl = port.noise_level(sch, lat, torch.tensor(t))
    res = {}
    for name, model in (("fp16", o_lgp), ("fp32", port.LatentEdgePredictorOracle32(o_lgp))):
        leaves = [tp.clone().requires_grad_(True) for tp in tap_vals]
        with torch.enable_grad():
            feats = torch.cat([F.interpolate(tp, size=L, mode="bilinear") for tp in leaves], dim=1)   # pipeline.py:146-151
            out = model(feats, torch.cat([lvl] * 2))
            o4 = out.reshape(2, L, L, -1).permute(0, 3, 2, 1)
            loss = F.mse_loss(tgt.float(), o4.chunk(2)[1].float())
            grads = torch.autograd.grad(loss, leaves)
        res[name] = dict(out=out.detach().float(), loss=loss.item(), grads=grads, feats=feats.detach())
    sigma = float((1 - sch.alphas_cumprod[t]) ** 0.5)
    return tap_vals, lvl, sigma, res


def test_lgp_forward_loss_backward_match_oracle(tiny):
    _check_lgp_against_oracle(tiny)


def _check_lgp_against_oracle(tiny):
    lat, _, tgt = tiny["inputs"]
    L = lat.shape[2]
    tap_vals, lvl, sigma, ref = _oracle_lgp(tiny)
    eng = tiny["lgp"].engine()
    taps_nhwc = [t.permute(0, 2, 3, 1).contiguous().cuda() for t in tap_vals]
    # the reference's own rounding noise: its fp16 LGP against the same MLP in fp32
    yard = max(rel(a, b) for a, b in zip(ref["fp16"]["grads"], ref["fp32"]["grads"]))
    print("oracle fp16-vs-fp32 LGP tap-gradient distance %.3e" % yard)
    assert yard > 5e-3          # the reference's gradient really is this noisy
    for emulate, key in ((True, "fp16"), (False, "fp32")):
        eng.set_grad_rounding(emulate)
        eng.forward_taps(taps_nhwc, 2, L, lat.cuda().contiguous(), sigma, True)
        mine = eng.output(2, L, "cuda")                     # rows in the reference's (b w h) order
        assert rel(mine, ref[key]["out"]) < 3e-3
        l, grads, scale = eng.loss_backward(tgt.cuda().contiguous(), taps_nhwc)
        assert abs(l.item() - ref[key]["loss"]) < 2e-3 * abs(ref[key]["loss"])
        errs = [rel(grads[k].permute(0, 3, 1, 2) / scale, ref[key]["grads"][k]) for k in range(9)]
        print("LGP tap gradients vs %s oracle: %s" % (key, " ".join("%.2e" % e for e in errs)))
        # on identical inputs the CUDA gradient is as close to either oracle as the two oracles are to each other
        assert max(errs) < 1.5 * yard + 1e-2, f"{key}: {errs}"
        # cond-only form (what the sampler calls): the cond sample's gradients, same values
        eng.forward_taps(taps_nhwc, 2, L, lat.cuda().contiguous(), sigma, True)
        l2, gc, scale2 = eng.loss_backward(tgt.cuda().contiguous(), taps_nhwc, cond_only=True)
        assert scale2 == scale and l2.item() == l.item()        # the loss is a fixed-order sum: bit-identical
        for k in range(9):
            assert tuple(gc[k].shape) == (1,) + tuple(grads[k].shape[1:])
            # same arithmetic; the 4096-row layer-1 dgrad may pick another tile / split-K than the 8192-row one
            assert rel(gc[k], grads[k][1:2]) < 2e-3, f"{key}: cond-only tap gradient {k} differs from the cond slice"
    eng.set_grad_rounding(True)
    # LatentEdgePredictor.forward surface (already resized + concatenated features), train-mode BN over all rows
    out2 = tiny["lgp"](ref["fp16"]["feats"].cuda(), torch.cat([lvl] * 2).cuda())
    assert out2.dtype == torch.float16 and tuple(out2.shape) == tuple(ref["fp16"]["out"].shape)
    assert rel(out2.float(), ref["fp16"]["out"]) < 3e-3


def test_lgp_eval_mode_uses_running_statistics(tiny):
    port, o_lgp = tiny["port"], tiny["o_lgp"]
    lat = tiny["inputs"][0]
    L = lat.shape[2]
    _, lvl, _, ref_ = _oracle_lgp(tiny)
    feats = ref_["fp16"]["feats"]
    ref = copy.deepcopy(o_lgp).eval()
    for m in ref.layers:
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1, generator=torch.Generator().manual_seed(5))
            m.running_var.uniform_(0.5, 1.5, generator=torch.Generator().manual_seed(6))
    want = ref(feats, torch.cat([lvl] * 2)).float()
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    mine = LatentEdgePredictor(port.lgp_input_dim(tiny["o_unet"]), 4, port.NUM_POS_LAYERS)
    mine.load_state_dict(ref.float().state_dict())
    mine.eval()
    got = mine(feats.cuda(), torch.cat([lvl] * 2).cuda()).float()
    assert rel(got, want) < 3e-3


# ------------------------------------------------------------------------------------------------ scheduler kernels
@pytest.mark.parametrize("prediction", [0, 1])
def test_cfg_ddim_step_bit_exact(cuda, prediction):
    import ctypes as C
    from sketch2img_b200 import _lib
    from sketch2img_b200.scheduler import DDIMScheduler
    lib = _lib.lib()
    sch = DDIMScheduler(prediction_type="epsilon" if prediction == 0 else "v_prediction")
    sch.set_timesteps(50)
    g = torch.Generator().manual_seed(21)
    S, n = 3, 4 * 64 * 64
    x = torch.randn(S, n, generator=g)
    eps = torch.randn(2 * S, n, generator=g)
    xd, epsd = x.cuda(), eps.cuda()
    for t in (981, 501, 1):
        sa_t, sb_t, sa_p, sb_p = sch.step_coefficients(t)
        out = torch.empty(S, n, device=cuda)
        _lib.check(lib.s2i_cfg_ddim_step(xd.data_ptr(), epsd.data_ptr(), S, n, 7.5, sb_t, sa_t, sa_p, sb_p,
                                         prediction, out.data_ptr(), _lib.stream_ptr()))
        eu, ec = eps[0::2], eps[1::2]
        e = eu + 7.5 * (ec - eu)
        f = lambda v: torch.tensor(v, dtype=torch.float32)
        if prediction == 0:
            x0 = (x - f(sb_t) * e) / f(sa_t)
        else:
            x0 = f(sa_t) * x - f(sb_t) * e
            e = f(sa_t) * e + f(sb_t) * x
        want = f(sa_p) * x0 + f(sb_p) * e
        assert torch.equal(out.cpu(), want), f"t={t}"


def test_guidance_update_matches_reference_formula(cuda):
    from sketch2img_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(22)
    S, n = 2, 4 * 64 * 64
    x_old = torch.randn(S, n, generator=g)
    x_new = x_old + 0.1 * torch.randn(S, n, generator=g)
    dx = 1e-4 * torch.randn(2 * S, n, generator=g)
    out = x_new.clone().cuda()
    scratch = torch.zeros(2 * S, dtype=torch.float64, device=cuda)
    xod, dxd = x_old.cuda(), dx.cuda()
    _lib.check(lib.s2i_guidance_update(xod.data_ptr(), out.data_ptr(), dxd.data_ptr(), S, n, 1.6,
                                       scratch.data_ptr(), _lib.stream_ptr()))
    for s in range(S):
        # modules/pipeline.py:159-161 for one sample: x_in = [x_old, x_old], g = -dx[cond]
        x_in = torch.stack([x_old[s], x_old[s]])
        gq = -dx[2 * s + 1]
        alpha = torch.linalg.norm(x_in - x_new[s]) / torch.linalg.norm(gq) * 1.6
        want = x_new[s] + alpha * gq
        assert rel(out[s], want) < 1e-6


# ------------------------------------------------------------------------------------------------ whole path
def _run_pipe(tiny, steps, **kw):
    lat, emb, tgt = tiny["inputs"]
    got = {}
    out = tiny["pipe"]("synthetic", num_inference_steps=steps, guidance_scale=7.5, latents=lat.cuda(),
                       sketch_image=kw.pop("target", tgt.cuda()), prompt_embeds=emb.cuda(), output_type="latent",
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()), **kw)
    return out, got


def _check_guided_against_golden(got, out, gold, label):
    """Guided runs are chaotic (see the module docstring): the yardstick is the drift of the reference's own files
    after a 1e-6 perturbation (`self_sensitivity`, made by oracle/make_golden.py).  The CUDA path starts 1e-3 away
    (fp16 tensor-core operands), i.e. further along the same divergence curve, hence the factor and the floor."""
    sens = gold["self_sensitivity"]
    steps = gold["steps"]
    errs = {i: rel(got[i], ref) for i, ref in gold["latents"].items()}
    print(label, "per-step rel err vs golden", {i: "%.2e" % e for i, e in errs.items()})
    print(label, "oracle self-sensitivity  ", {i: "%.2e" % float(sens[i]) for i in errs})
    for i, e in errs.items():
        floor = float(sens[min(steps - 1, i + 1)])      # one step further along the divergence
        assert e < max(4.0 * floor, 0.12), f"{label} step {i}: {e:.3e} vs noise floor {floor:.3e}"
    assert torch.isfinite(out).all()
    norms = torch.tensor([got[i].norm().item() for i in range(steps)])
    assert torch.allclose(norms, gold["norms"].float(), rtol=0.1), "latent norms leave the reference's envelope"


def _check_unguided_against_golden(pipe, inputs, gold, label, tol=5e-3):
    """sketch_image=None: apply_anti_gradient returns the DDIM latent unchanged (pipeline.py:142-143), so the run is
    plain CFG + DDIM over the full schedule -- smooth dynamics, pinned to the reference files' own trajectory."""
    lat, emb, _ = inputs
    got = {}
    out = pipe("synthetic", num_inference_steps=gold["steps"], guidance_scale=7.5, latents=lat.cuda(), sketch_image=None,
               prompt_embeds=emb.cuda(), output_type="latent",
               callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    errs = {i: rel(got[i], ref) for i, ref in gold["unguided_latents"].items()}
    print(label, "unguided per-step rel err", {i: "%.2e" % e for i, e in errs.items()})
    assert max(errs.values()) < tol
    return rel(out, gold["unguided_latents"][gold["steps"] - 1])


@pytest.mark.parametrize("steps", [4, 50])
def test_pipeline_guided_matches_reference_golden_at_noise_floor(tiny, steps):
    gold = torch.load(os.path.join(GOLD, f"tiny_{steps}step.pt"))
    out, got = _run_pipe(tiny, steps)
    _check_guided_against_golden(got, out, gold, f"tiny {steps}-step")


@pytest.mark.parametrize("steps", [4, 50])
def test_pipeline_unguided_matches_reference_golden(tiny, steps):
    gold = torch.load(os.path.join(GOLD, f"tiny_{steps}step.pt"))
    final = _check_unguided_against_golden(tiny["pipe"], tiny["inputs"], gold, f"tiny {steps}-step")
    print("tiny %d-step unguided FINAL rel err %.3e" % (steps, final))


# ---- the demo's scheduler: DPM-Solver++(2M), app.py:14-25 ---------------------------------------------------------
def _dpmpp_pipe(tiny):
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DPMSolverMultistepScheduler
    sch = DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                                      trained_betas=None, predict_epsilon=True, thresholding=False, algorithm_type="dpmsolver++",
                                      solver_type="midpoint", lower_order_final=True)
    pipe = AntiGradientPipeline(unet=tiny["unet"], scheduler=sch)
    pipe.setup_lgp(tiny["lgp"])
    return pipe


@pytest.mark.parametrize("prediction", [0, 1])
def test_cfg_dpmpp_step_bit_exact(cuda, prediction):
    """s2i_cfg_dpmpp_step against the oracle scheduler's ``step`` on the same CFG-combined model output, over a whole
    4-step (first / second / second / first order) and the head of a 50-step schedule: bit for bit."""
    from oracle import port
    from sketch2img_b200 import _lib
    from sketch2img_b200.scheduler import DPMSolverMultistepScheduler
    lib = _lib.lib()
    ptype = "epsilon" if prediction == 0 else "v_prediction"
    for n, upto in ((4, 4), (50, 5)):
        o = port.make_scheduler(prediction_type=ptype, kind="dpmpp")
        p = DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", prediction_type=ptype)
        o.set_timesteps(n)
        p.set_timesteps(n)
        g = torch.Generator().manual_seed(31 + n)
        S, nel = 3, 4 * 32 * 32
        x = torch.randn(S, nel, generator=g)
        hist = torch.full((S, nel), float("nan"), device=cuda)
        xd = x.cuda()
        for i in range(upto):
            t = o.timesteps[i]
            eps = torch.randn(2 * S, nel, generator=g)
            e = eps[0::2] + 7.5 * (eps[1::2] - eps[0::2])
            want = o.step(e, t, x).prev_sample
            pl = p.step_plan(i)
            out = torch.empty(S, nel, device=cuda)
            _lib.check(lib.s2i_cfg_dpmpp_step(xd.data_ptr(), eps.cuda().data_ptr(), hist.data_ptr(), S, nel, 7.5, pl["alpha_t"],
                                              pl["sigma_t"], pl["c_x"], pl["c_m0"], pl["c_d1"], pl["inv_r0"], pl["order"],
                                              prediction, out.data_ptr(), _lib.stream_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(out.cpu(), want), f"{n}-step schedule, step {i} (order {pl['order']})"
            assert torch.equal(hist.cpu(), o.model_outputs[-1]), "x0-prediction history"
            x, xd = want, out


@pytest.mark.parametrize("steps", [4, 20])
def test_pipeline_dpmpp_unguided_matches_reference_golden(tiny, steps):
    gold = torch.load(os.path.join(GOLD, f"tiny_dpmpp_{steps}step.pt"))
    final = _check_unguided_against_golden(_dpmpp_pipe(tiny), tiny["inputs"], gold, f"tiny dpm++ {steps}-step")
    print("tiny dpm++ %d-step unguided FINAL rel err %.3e" % (steps, final))


@pytest.mark.parametrize("steps", [4, 20])
def test_pipeline_dpmpp_guided_matches_reference_golden_at_noise_floor(tiny, steps):
    gold = torch.load(os.path.join(GOLD, f"tiny_dpmpp_{steps}step.pt"))
    lat, emb, tgt = tiny["inputs"]
    got = {}
    out = _dpmpp_pipe(tiny)("synthetic", num_inference_steps=steps, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
                            prompt_embeds=emb.cuda(), output_type="latent",
                            callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    _check_guided_against_golden(got, out, gold, f"tiny dpm++ {steps}-step")


def test_guided_step_stage_by_stage(tiny):
    """ONE guided step from identical state, stage by stage: CFG + DDIM latent, UNet adjoint applied to the ORACLE's
    tap gradients, and the update direction of the full CUDA chain."""
    port, o_unet, o_lgp = tiny["port"], tiny["o_unet"], tiny["o_lgp"]
    lat, emb, tgt = tiny["inputs"]
    eng, leng = tiny["unet"].engine, tiny["lgp"].engine()
    L = lat.shape[2]
    sch = port.make_scheduler()
    sch.set_timesteps(50)
    t = torch.tensor(981)
    taps, handles = port.register_taps(o_unet)
    x_in = torch.cat([lat] * 2).requires_grad_(True)
    with torch.enable_grad():
        eps_ref = o_unet(x_in, t, encoder_hidden_states=emb).sample
        tap_ref = [m.output for m in taps]
    for h in handles:
        h.remove()
    eu, ec = eps_ref.detach().chunk(2)
    x_ddim = sch.step(eu + 7.5 * (ec - eu), t, lat, eta=0.0).prev_sample
    leaves = [tp.detach().clone().requires_grad_(True) for tp in tap_ref]
    with torch.enable_grad():
        feats = torch.cat([F.interpolate(lf, size=L, mode="bilinear") for lf in leaves], dim=1)
        lvl = port.noise_level(sch, lat, t)
        out = o_lgp(feats, torch.cat([lvl] * 2))
        loss = F.mse_loss(tgt.float(), out.reshape(2, L, L, -1).permute(0, 3, 2, 1).chunk(2)[1].float())
        g_tap = torch.autograd.grad(loss, leaves)
        dx_ref = torch.autograd.grad(tap_ref, x_in, grad_outputs=[g.to(tp.dtype) for g, tp in zip(g_tap, tap_ref)])[0]
    # UNet adjoint alone (identical tap gradients in): tight
    eng.forward(x_in.detach().cuda(), 981, emb.cuda(), save_for_backward=True)
    dx_a = eng.backward([g.permute(0, 2, 3, 1).contiguous().cuda() for g in g_tap])
    assert rel(dx_a, dx_ref) < 6e-3
    # whole CUDA chain: one C-ABI sampler step from the same state
    first = {}
    tiny["pipe"]("synthetic", num_inference_steps=50, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
                 prompt_embeds=emb.cuda(), output_type="latent",
                 callback=lambda i, t_, l: first.setdefault("x", l.detach().cpu().clone()) if int(i) == 0 else None)
    x = first["x"]
    gq = (-dx_ref).chunk(2)[1]
    alpha = torch.linalg.norm(x_in.detach() - x_ddim) / torch.linalg.norm(gq) * 1.6
    x_new_ref = x_ddim + alpha * gq
    upd = x - x_ddim                           # what guidance added on top of the (accurate) DDIM latent
    cos = F.cosine_similarity(upd.flatten().double(), (alpha * gq).flatten().double(), dim=0).item()
    print("guided step: update-direction cosine %.5f, |update| mine %.4f oracle %.4f, x_new rel err %.3e" % (
        cos, upd.norm().item(), (alpha * gq).norm().item(), rel(x, x_new_ref)))
    assert abs(upd.norm().item() / (alpha * gq).norm().item() - 1) < 2e-2     # the norm-ratio step size (pipeline.py:160)
    assert cos > 0.97            # direction: limited by the LGP's x100 input sensitivity (module docstring)


def test_pipeline_without_sketch_skips_guidance(tiny):
    """target None => apply_anti_gradient returns latents unchanged (pipeline.py:142-143): plain CFG + DDIM."""
    port, o_unet = tiny["port"], tiny["o_unet"]
    lat, emb, _ = tiny["inputs"]
    out, _ = _run_pipe(tiny, 4, target=None)
    sch = port.make_scheduler()
    sch.set_timesteps(4)
    x = lat.clone()
    with torch.no_grad():
        for t in sch.timesteps:
            eps = o_unet(torch.cat([x] * 2), t, encoder_hidden_states=emb).sample
            eu, ec = eps.chunk(2)
            x = sch.step(eu + 7.5 * (ec - eu), t, x, eta=0.0).prev_sample
    assert rel(out, x) < 3e-3


def test_batched_samples_equal_independent_calls(tiny):
    """a batch of B is B independent batch-1 reference calls (per-sample BN statistics and step size)."""
    pipe = tiny["pipe"]
    lat, emb, tgt = tiny["inputs"]
    g = torch.Generator().manual_seed(33)
    lat2 = torch.cat([lat, torch.randn(lat.shape, generator=g)])
    emb_b = torch.randn(emb.shape, generator=g)
    tgt2 = torch.cat([tgt, torch.randn(tgt.shape, generator=g)])
    embs = torch.cat([emb[:1], emb_b[:1], emb[1:], emb_b[1:]])          # [uncond..., cond...]
    both = pipe(["a", "b"], num_inference_steps=4, latents=lat2.cuda(), sketch_image=tgt2.cuda(),
                prompt_embeds=embs.cuda(), output_type="latent").clone()
    one0 = pipe("a", num_inference_steps=4, latents=lat2[:1].cuda(), sketch_image=tgt2[:1].cuda(),
                prompt_embeds=emb.cuda(), output_type="latent").clone()
    one1 = pipe("b", num_inference_steps=4, latents=lat2[1:].cuda(), sketch_image=tgt2[1:].cuda(),
                prompt_embeds=emb_b.cuda(), output_type="latent").clone()
    # guided: equal up to the chaotic loop's noise floor after 4 steps; unguided (smooth): equal to fp16 noise
    gold = torch.load(os.path.join(GOLD, "tiny_4step.pt"))
    floor = max(4.0 * float(gold["self_sensitivity"][3]), 0.12)
    assert rel(both[:1], one0) < floor and rel(both[1:], one1) < floor
    ub = pipe(["a", "b"], num_inference_steps=4, latents=lat2.cuda(), prompt_embeds=embs.cuda(), output_type="latent").clone()
    u0 = pipe("a", num_inference_steps=4, latents=lat2[:1].cuda(), prompt_embeds=emb.cuda(), output_type="latent").clone()
    u1 = pipe("b", num_inference_steps=4, latents=lat2[1:].cuda(), prompt_embeds=emb_b.cuda(), output_type="latent").clone()
    assert rel(ub[:1], u0) < 3e-3 and rel(ub[1:], u1) < 3e-3


def test_pipeline_error_behaviour(tiny):
    pipe = tiny["pipe"]
    lat, emb, tgt = tiny["inputs"]
    with pytest.raises(NotImplementedError):
        pipe("a", num_inference_steps=2, guidance_scale=1.0, latents=lat.cuda(), prompt_embeds=emb.cuda())
    with pytest.raises(ValueError):
        pipe("a", height=100, width=64, num_inference_steps=2, latents=lat.cuda(), prompt_embeds=emb.cuda())
    with pytest.raises(ValueError):
        pipe("a", num_inference_steps=2, latents=torch.zeros(1, 4, 3, 3), prompt_embeds=emb.cuda())
    with pytest.raises(ValueError):
        pipe("a", num_inference_steps=2, latents=lat.cuda(), prompt_embeds=emb[:1].cuda())


def test_hook_unet_taps_expose_reference_shapes(tiny):
    """hook_unet contract (latent_predictor.py:47-81): 9 blocks, `.output` = fp32 NCHW feature of the last forward."""
    pipe, o_unet = tiny["pipe"], tiny["o_unet"]
    lat, emb, _ = tiny["inputs"]
    pipe.unet(torch.cat([lat] * 2).cuda(), 981, emb.cuda())
    boc = o_unet.config.block_out_channels
    L = lat.shape[2]
    want = [(boc[0], L // 2), (boc[1], L // 4), (boc[2], L // 8), (boc[3], L // 8), (boc[3], L // 8), (boc[3], L // 8),
            (boc[3], L // 4), (boc[2], L // 2), (boc[1], L)]
    assert len(pipe.feature_blocks) == 9
    for blk, (c, s) in zip(pipe.feature_blocks, want):
        assert tuple(blk.output.shape) == (2, c, s, s) and blk.output.dtype == torch.float32


# ------------------------------------------------------------------------------------------------ SD1.5 size
@pytest.fixture(scope="module")
def sd15(cuda):
    from oracle import port
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    from sketch2img_b200.unet import UNet2DConditionModel
    o_unet = port.make_unet("sd15")
    o_lgp = port.make_lgp(o_unet)
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    lgp = LatentEdgePredictor(port.lgp_input_dim(o_unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(copy.deepcopy(o_lgp).float().state_dict())     # nn.Module.float() converts in place
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    inputs = port.make_inputs(o_unet)
    return dict(port=port, o_unet=o_unet, o_lgp=o_lgp, unet=unet, lgp=lgp, pipe=pipe, inputs=inputs)


@pytest.mark.parametrize("steps", [4, 50])
def test_sd15_guided_matches_reference_golden_at_noise_floor(sd15, steps):
    """BASELINE.json configs[0] (4-step) and configs[1] (50-step) against the fixtures made by the reference's
    pipeline.py + latent_predictor.py on CPU."""
    gold = torch.load(os.path.join(GOLD, f"sd15_{steps}step.pt"))
    lat, emb, tgt = sd15["inputs"]
    got = {}
    out = sd15["pipe"]("synthetic", num_inference_steps=steps, guidance_scale=7.5, latents=lat.cuda(),
                       sketch_image=tgt.cuda(), prompt_embeds=emb.cuda(), output_type="latent",
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    _check_guided_against_golden(got, out, gold, f"sd15 {steps}-step")


@pytest.mark.parametrize("steps", [4, 50])
def test_sd15_unguided_matches_reference_golden(sd15, steps):
    """north_star's tolerance -- final-latent relative L2 within 1e-3 of the reference -- asserted on the configuration it is
    quoted for (50 steps).  The 4-step schedule (configs[0]) strides 250 timesteps per step: each step's new latent is
    ~0.6 x0-prediction, so one forward's fp16-operand error (eps: 1.2e-3 relative at SD1.5 size) reaches the latent almost
    undamped and the floor is the single-forward floor; it is asserted at 2e-3 and printed."""
    gold = torch.load(os.path.join(GOLD, f"sd15_{steps}step.pt"))
    final = _check_unguided_against_golden(sd15["pipe"], sd15["inputs"], gold, f"sd15 {steps}-step", tol=2e-3)
    print("sd15 %d-step unguided FINAL rel err %.3e (north_star target 1e-3)" % (steps, final))
    assert final < (1e-3 if steps == 50 else 2e-3)


# ---- SD1.5-size kernels against the oracle on identical inputs (the bisecting tools of round 1, now collected by pytest)
def test_sd15_unet_forward_taps_backward_match_oracle(sd15):
    """s2i_unet_forward / s2i_unet_backward(_samples) at the size BASELINE.json's metric is quoted on: eps and the 9 taps
    3e-3, the tap -> input adjoint 6e-3 (whole batch and the sampler's cond-only walk)."""
    _check_unet_against_oracle(sd15, label="sd15")


def test_sd15_lgp_forward_loss_backward_match_oracle(sd15):
    """The 8192 x 9320 feature matrix, the 9320 -> 512 layer and its cond-only input gradient, train-mode BatchNorm over
    8192 rows, the separable resize adjoint: against the fp16 oracle and its fp32 restatement (same yardstick as tiny)."""
    _check_lgp_against_oracle(sd15)
