"""Parity of the CUDA sampling path (through the C ABI of include/s2i.h) against the CPU oracle (oracle/port.py)
and the golden fixtures made by the reference's own files (tests/golden/, oracle/make_golden.py).

Tolerances (relative L2, fp16 tensor-core operands with fp32 accumulation vs the fp32 CPU oracle):
  UNet eps / taps 3e-3, UNet input gradient 6e-3, LGP output 3e-3, LGP tap gradients 2e-2 (fp16 autograd rounding
  of O(1e-5) values), final latent of a guided run: see each test (north_star target 1e-3 on the SD1.5 job).
The scheduler / guidance-update kernels are compared bit-for-bit or to 1e-6.
"""
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
    lgp.load_state_dict(o_lgp.float().state_dict())
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    return dict(port=port, o_unet=o_unet, o_lgp=o_lgp, unet=unet, lgp=lgp, pipe=pipe, inputs=port.make_inputs(o_unet))


# ------------------------------------------------------------------------------------------------ UNet
def test_unet_forward_taps_backward_match_oracle(tiny):
    port, o_unet, eng = tiny["port"], tiny["o_unet"], tiny["unet"].engine
    lat, emb, _ = tiny["inputs"]
    x = torch.cat([lat] * 2)
    taps, handles = port.register_taps(o_unet)
    xg = x.clone().requires_grad_(True)
    with torch.enable_grad():
        eps_ref = o_unet(xg, torch.tensor(981), encoder_hidden_states=emb).sample
        tap_ref = [m.output for m in taps]
        gen = torch.Generator().manual_seed(7)
        G = [torch.randn(t.shape, generator=gen) for t in tap_ref]
        dx_ref = torch.autograd.grad(sum((g * t).sum() for g, t in zip(G, tap_ref)), xg)[0]
    for h in handles:
        h.remove()
    eps = eng.forward(x.cuda(), 981, emb.cuda(), save_for_backward=True)
    assert rel(eps, eps_ref) < 3e-3
    for k in range(9):
        got = eng.tap(k).permute(0, 3, 1, 2)
        assert got.shape == tap_ref[k].shape
        assert rel(got, tap_ref[k].detach()) < 3e-3, f"tap {k}"
    Gd = [g.permute(0, 2, 3, 1).contiguous().cuda() for g in G]
    dx = eng.backward(Gd)
    assert rel(dx, dx_ref) < 6e-3
    # linearity of the tap->input adjoint (size-independent property): J^T(2g) == 2 J^T(g)
    eng.forward(x.cuda(), 981, emb.cuda(), save_for_backward=True)
    dx2 = eng.backward([2 * g for g in Gd])
    assert rel(dx2, 2 * dx) < 1e-3
    # a second backward without a new forward is a state error, not silent garbage
    from sketch2img_b200._lib import S2IError
    with pytest.raises(S2IError):
        eng.backward(Gd)


def test_unet_forward_is_reproducible_and_batch_independent(tiny):
    eng = tiny["unet"].engine
    lat, emb, _ = tiny["inputs"]
    g = torch.Generator().manual_seed(11)
    x4 = torch.randn(4, 4, lat.shape[2], lat.shape[3], generator=g).cuda()
    e4 = torch.randn(4, 77, emb.shape[2], generator=g).cuda()
    a = eng.forward(x4, 501, e4).clone()
    b = eng.forward(x4, 501, e4).clone()
    assert rel(a, b) < 1e-5
    # samples are closed computations: a batch equals its samples run alone
    for i in (0, 3):
        one = eng.forward(x4[i:i + 1], 501, e4[i:i + 1])
        assert rel(one, a[i:i + 1]) < 1e-4


# ------------------------------------------------------------------------------------------------ LGP
def _oracle_lgp(tiny, t=981):
    port, o_unet, o_lgp = tiny["port"], tiny["o_unet"], tiny["o_lgp"]
    lat, emb, tgt = tiny["inputs"]
    sch = port.make_scheduler()
    sch.set_timesteps(50)
    taps, handles = port.register_taps(o_unet)
    x = torch.cat([lat] * 2).requires_grad_(True)
    L = lat.shape[2]
    with torch.enable_grad():
        o_unet(x, torch.tensor(t), encoder_hidden_states=emb)
        tap_out = [m.output for m in taps]
        feats = port.lgp_features(taps, L)
        lvl = port.noise_level(sch, lat, torch.tensor(t))
        out = o_lgp(feats, torch.cat([lvl] * 2))
        o4 = out.reshape(2, L, L, -1).permute(0, 3, 2, 1)
        loss = F.mse_loss(tgt.float(), o4.chunk(2)[1].float())
        g_ref = torch.autograd.grad(loss, tap_out)
    for h in handles:
        h.remove()
    sigma = float((1 - sch.alphas_cumprod[t]) ** 0.5)
    return tap_out, feats.detach(), lvl, out.detach(), loss.item(), g_ref, sigma


def test_lgp_forward_loss_backward_match_oracle(tiny):
    lat, _, tgt = tiny["inputs"]
    L = lat.shape[2]
    tap_out, feats, lvl, out, loss, g_ref, sigma = _oracle_lgp(tiny)
    eng = tiny["lgp"].engine()
    taps_nhwc = [t.detach().permute(0, 2, 3, 1).contiguous().cuda() for t in tap_out]
    eng.forward_taps(taps_nhwc, 2, L, lat.cuda().contiguous(), sigma, True)
    mine = eng.output(2, L, "cuda")                     # rows in the reference's (b w h) order
    assert rel(mine, out.float()) < 3e-3
    l, grads, scale = eng.loss_backward(tgt.cuda().contiguous(), taps_nhwc)
    assert abs(l.item() - loss) < 2e-3 * abs(loss)
    for k in range(9):
        gm = grads[k].permute(0, 3, 1, 2) / scale
        assert rel(gm, g_ref[k]) < 2e-2, f"tap grad {k}"
    # LatentEdgePredictor.forward surface (already resized + concatenated features), train-mode BN over all rows
    out2 = tiny["lgp"](feats.cuda(), torch.cat([lvl] * 2).cuda())
    assert out2.dtype == torch.float16 and out2.shape == out.shape
    assert rel(out2.float(), out.float()) < 3e-3


def test_lgp_eval_mode_uses_running_statistics(tiny):
    port, o_lgp = tiny["port"], tiny["o_lgp"]
    lat = tiny["inputs"][0]
    L = lat.shape[2]
    _, feats, lvl, _, _, _, _ = _oracle_lgp(tiny)
    import copy
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
    for t in (981, 501, 1):
        sa_t, sb_t, sa_p, sb_p = sch.step_coefficients(t)
        out = torch.empty(S, n, device=cuda)
        _lib.check(lib.s2i_cfg_ddim_step(x.cuda().data_ptr(), eps.cuda().data_ptr(), S, n, 7.5, sb_t, sa_t, sa_p, sb_p,
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
    _lib.check(lib.s2i_guidance_update(x_old.cuda().data_ptr(), out.data_ptr(), dx.cuda().data_ptr(), S, n, 1.6,
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


def test_pipeline_4_steps_matches_reference_golden(tiny):
    gold = torch.load(os.path.join(GOLD, "tiny_4step.pt"))
    out, got = _run_pipe(tiny, 4)
    for i, ref in gold["latents"].items():
        assert rel(got[i], ref) < 3e-3, f"step {i}"
    assert rel(out, gold["latents"][3]) < 3e-3


def test_pipeline_50_steps_matches_reference_golden(tiny):
    gold = torch.load(os.path.join(GOLD, "tiny_50step.pt"))
    out, got = _run_pipe(tiny, 50)
    errs = {i: rel(got[i], ref) for i, ref in gold["latents"].items()}
    print("tiny 50-step per-step rel err", {i: "%.2e" % e for i, e in errs.items()})
    assert rel(out, gold["latents"][49]) < 1e-2
    norms = torch.tensor([got[i].norm().item() for i in range(50)])
    assert torch.allclose(norms, gold["norms"].float(), rtol=5e-3)


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
    assert rel(both[:1], one0) < 2e-4 and rel(both[1:], one1) < 2e-4
    gold = torch.load(os.path.join(GOLD, "tiny_4step.pt"))
    assert rel(both[:1], gold["latents"][3]) < 3e-3


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
    lgp.load_state_dict(o_lgp.float().state_dict())
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler())
    pipe.setup_lgp(lgp)
    inputs = port.make_inputs(o_unet)
    del o_unet
    return dict(pipe=pipe, inputs=inputs)


def test_sd15_4_steps_matches_reference_golden(sd15):
    """BASELINE.json configs[0] (SD1.5 64x64-latent 4-step DDIM + LGP) against the fixture made by the reference's
    pipeline.py + latent_predictor.py on CPU."""
    gold = torch.load(os.path.join(GOLD, "sd15_4step.pt"))
    lat, emb, tgt = sd15["inputs"]
    got = {}
    out = sd15["pipe"]("synthetic", num_inference_steps=4, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
                       prompt_embeds=emb.cuda(), output_type="latent",
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    errs = {i: rel(got[i], ref) for i, ref in gold["latents"].items()}
    print("sd15 4-step per-step rel err", {i: "%.2e" % e for i, e in errs.items()})
    assert rel(out, gold["latents"][3]) < 3e-3


def test_sd15_50_steps_matches_reference_golden(sd15):
    """BASELINE.json configs[1]: the full 50-step job; north_star target is 1e-3 relative L2 on the final latent."""
    gold = torch.load(os.path.join(GOLD, "sd15_50step.pt"))
    lat, emb, tgt = sd15["inputs"]
    got = {}
    out = sd15["pipe"]("synthetic", num_inference_steps=50, guidance_scale=7.5, latents=lat.cuda(), sketch_image=tgt.cuda(),
                       prompt_embeds=emb.cuda(), output_type="latent",
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    errs = {i: rel(got[i], ref) for i, ref in gold["latents"].items()}
    print("sd15 50-step per-step rel err", {i: "%.2e" % e for i, e in errs.items()})
    final = rel(out, gold["latents"][49])
    print("sd15 50-step FINAL rel err %.3e" % final)
    assert final < 1e-2
