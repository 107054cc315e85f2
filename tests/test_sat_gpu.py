"""Parity of the injected sketch attention path (SD2.1-shaped UNet with linear projections,
d_head 64, v-prediction DDIM, SatMixin active, no LGP) against the CPU oracle (oracle/port.py: SatMixinOracle, pinned
bit-exactly to the reference's own modules/sketch_guided_attn.py) and the fixture made by the reference files.
Tolerance: 3e-3 relative L2 on eps (fp16 tensor-core operands, fp32 accumulation), 5e-3 on the 4-step latent."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _build(name, cuda):
    from oracle import port
    from sketch2img_b200.sketch_guided_attn import SatMixin
    from sketch2img_b200.unet import UNet2DConditionModel
    o_unet = port.make_unet(name)
    sd = o_unet.state_dict()
    unet = UNet2DConditionModel(vars(o_unet.config), sd)
    o_sat = port.make_sat(o_unet)
    sat = SatMixin(unet)
    sat.load_state_dict(o_sat.state_dict())
    return port, o_unet, o_sat, unet, sat


def test_sat_forward_matches_oracle_and_fixture(cuda):
    port, o_unet, o_sat, unet, sat = _build("tiny21", cuda)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    lat, emb, _ = port.make_inputs(o_unet)
    res = port.make_res_samples(o_unet, 2)
    x = torch.cat([lat] * 2)
    gold = torch.load(os.path.join(GOLD, "tiny21_sat_4step.pt"))
    # without features the blocks run unmodified (sketch_guided_attn.py:120)
    with torch.no_grad():
        plain = o_unet(x, torch.tensor(501), encoder_hidden_states=emb).sample
    assert rel(unet(x.cuda(), 501, emb.cuda()).sample, plain) < 3e-3
    for scale in (0.7, 1.0):
        o_sat.set_res_samples(res)
        o_sat.set_scale(scale)
        sat.set_res_samples([tuple(t.cuda() for t in tup) for tup in res])
        sat.set_scale(scale)
        with torch.no_grad():
            want = o_unet(x, torch.tensor(501), encoder_hidden_states=emb).sample
        got = unet(x.cuda(), 501, emb.cuda()).sample
        assert rel(got, want) < 3e-3, f"scale {scale}"
        assert rel(want, plain) > 0.05                      # the injected attention really changes the prediction
        if scale == 0.7:
            assert rel(got, gold["eps_t501"]) < 3e-3        # fixture from the reference's own SatMixin
    # a feature whose batch / token count does not match the forward is an error, not silent garbage
    from sketch2img_b200._lib import S2IError
    with pytest.raises(S2IError):
        unet(x[:1].cuda(), 501, emb[:1].cuda())
    # removing the features restores the plain block
    for blk in sat.blocks:
        blk.set_res_sample(None)
    assert rel(unet(x.cuda(), 501, emb.cuda()).sample, plain) < 3e-3


def test_sat_sampling_matches_reference_fixture(cuda):
    """Config-4 loop: CFG 7.5 + v-prediction DDIM with SatMixin active, through the drop-in pipeline."""
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    port, o_unet, o_sat, unet, sat = _build("tiny21", cuda)
    lat, emb, _ = port.make_inputs(o_unet)
    gold = torch.load(os.path.join(GOLD, "tiny21_sat_4step.pt"))
    sat.set_res_samples([tuple(t.cuda() for t in tup) for tup in port.make_res_samples(o_unet, 2)])
    sat.set_scale(gold["sat_scale"])
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler(prediction_type="v_prediction"))
    got = {}
    pipe("synthetic", num_inference_steps=gold["steps"], guidance_scale=gold["guidance_scale"], latents=lat.cuda(),
         sketch_image=None, prompt_embeds=emb.cuda(), output_type="latent",
         callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    errs = {i: rel(got[i], ref) for i, ref in gold["latents"].items()}
    print("tiny21 + SatMixin 4-step per-step rel err vs reference fixture", {i: "%.2e" % e for i, e in errs.items()})
    assert max(errs.values()) < 5e-3


def test_sat_sampling_two_images_per_call_matches_oracle(cuda):
    """Two images per call (config 4 runs batches): the reference concatenates [uncond_0, uncond_1, cond_0, cond_1]
    (pipeline.py:85) and its sketch features come in that order; the engine's batch is sample-major too, so the features
    pass through unpermuted.  Oracle: oracle/port.py's loop with SatMixinOracle on the same batch, on CPU."""
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    port, o_unet, o_sat, unet, sat = _build("tiny21", cuda)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    g = torch.Generator().manual_seed(77)
    L, D = o_unet.config.sample_size, o_unet.config.cross_attention_dim
    lat = torch.randn(2, 4, L, L, generator=g)
    emb = torch.randn(4, 77, D, generator=g)                      # [uncond_0, uncond_1, cond_0, cond_1]
    res = port.make_res_samples(o_unet, 4)
    o_sat.set_res_samples(res)
    o_sat.set_scale(0.7)
    sat.set_res_samples([tuple(t.cuda() for t in tup) for tup in res])
    sat.set_scale(0.7)
    want = {}
    port.guided_sample(o_unet, None, port.make_scheduler("v_prediction"), emb, lat.clone(), None, num_steps=4,
                       callback=lambda i, t, l: want.__setitem__(int(i), l.detach().clone()))
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler(prediction_type="v_prediction"))
    got = {}
    pipe(["a", "b"], num_inference_steps=4, guidance_scale=7.5, latents=lat.cuda(), sketch_image=None, prompt_embeds=emb.cuda(),
         output_type="latent", callback=lambda i, t, l: got.__setitem__(int(i), l.detach().float().cpu().clone()))
    errs = {i: rel(got[i], want[i]) for i in want}
    print("tiny21 + SatMixin, 2 images per call, per-step rel err vs oracle", {i: "%.2e" % e for i, e in errs.items()})
    assert max(errs.values()) < 5e-3
    # and each image equals its own single-image call (samples are closed computations)
    for s in range(2):
        sat.set_res_samples([tuple(t[[s, 2 + s]].cuda() for t in tup) for tup in res])
        one = pipe("a", num_inference_steps=4, guidance_scale=7.5, latents=lat[s:s + 1].cuda(), sketch_image=None,
                   prompt_embeds=emb[[s, 2 + s]].cuda(), output_type="latent")
        assert rel(one, got[3][s:s + 1]) < 3e-3


def test_sat_sd21_shape_forward_matches_oracle(cuda):
    """The real SD2.1 topology (320/640/1280 channels, 5/10/20/20 heads of 64, context 1024, linear projections) at a
    48 x 48 latent: every block through the fused attention kernel (N = 2304 .. 36 tokens)."""
    port, o_unet, o_sat, unet, sat = _build("sd21", cuda)
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(5)
    L = 48
    x = torch.randn(2, 4, L, L, generator=gen)
    emb = torch.randn(2, 77, 1024, generator=gen)
    res = port.make_res_samples(o_unet, 2, size=L)
    o_sat.set_res_samples(res)
    sat.set_res_samples([tuple(t.cuda() for t in tup) for tup in res])
    with torch.no_grad():
        want = o_unet(x, torch.tensor(301), encoder_hidden_states=emb).sample
    got = unet(x.cuda(), 301, emb.cuda()).sample
    assert rel(got, want) < 3e-3
