"""the VAE either side of the sampling loop (app.py:107-109 sketch -> target latent; modules/pipeline.py:118,
:163-174 latent -> image) on the engine, against the CPU oracle (oracle/diffusers_shim AutoencoderKL) and the fixture whose
image was written by the reference's own ``decode_latents_L``.  Tolerance: 3e-3 relative L2 (fp16 operands, fp32 accumulation)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _build(name):
    from oracle import port
    from sketch2img_b200.vae import AutoencoderKL
    o_vae = port.make_vae(name)
    return port, o_vae, AutoencoderKL(vars(o_vae.config), o_vae.state_dict())


def test_tiny_vae_matches_fixture_and_reference_decode_latents_L(cuda):
    from sketch2img_b200.pipeline import AntiGradientPipeline
    port, o_vae, vae = _build("tiny")
    gold = torch.load(os.path.join(GOLD, "tiny_vae.pt"))
    dist = vae.encode(gold["image"].cuda()).latent_dist
    e_m = rel(dist.parameters, gold["moments"])
    dec = vae.decode(gold["latents"].cuda() / 0.18215).sample
    e_d = rel(dec, gold["decoded"])
    print("tiny VAE: moments rel err %.2e, decoded image rel err %.2e" % (e_m, e_d))
    assert e_m < 3e-3 and e_d < 3e-3
    assert tuple(dist.mean.shape) == (1, 4, 8, 8) and torch.equal(dist.mode(), dist.mean)
    g = torch.Generator(device="cuda").manual_seed(3)
    s = dist.sample(generator=g)
    assert tuple(s.shape) == (1, 4, 8, 8) and rel(s - dist.mean, dist.std * torch.randn(s.shape, generator=torch.Generator(device="cuda").manual_seed(3), device="cuda")) < 1e-6
    # the reference's decode_latents_L (pipeline.py:163-174): uint8 image with everything below 0.5 zeroed
    pipe = AntiGradientPipeline(unet=None, scheduler=None, vae=vae)
    img = torch.from_numpy(pipe.decode_latents_L(gold["latents"].cuda()))
    want = gold["image_L"]
    assert img.shape == want.shape and img.dtype == torch.uint8
    diff = (img.int() - want.int()).abs()
    # a pixel within 1e-3 of the 0.5 threshold may fall on the other side (value jumps by ~128); everything else is within one step
    flips = (diff > 1).float().mean().item()
    print("decode_latents_L: %.4f %% of the pixels differ by more than one grey level (threshold flips)" % (100 * flips))
    assert flips < 2e-3
    # same bits twice
    assert torch.equal(vae.decode(gold["latents"].cuda() / 0.18215).sample, dec)


@pytest.mark.parametrize("side", [256, 512])
def test_sd_vae_matches_oracle(cuda, side):
    """The SD VAE topology (128 / 256 / 512 / 512 channels, one 512-wide attention head over (side/8)^2 tokens) at 256 x 256
    and at the 512 x 512 of BASELINE.json's metric, against the oracle on the host cores."""
    port, o_vae, vae = _build("sd")
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(side)
    img = (torch.rand(1, 3, side, side, generator=g) - 0.5) / 0.5
    lat = torch.randn(1, 4, side // 8, side // 8, generator=g)
    with torch.no_grad():
        want_m = o_vae.encode(img).latent_dist.parameters
        want_d = o_vae.decode(lat).sample
    e_m = rel(vae.encode(img.cuda()).latent_dist.parameters, want_m)
    e_d = rel(vae.decode(lat.cuda()).sample, want_d)
    print("SD VAE %d x %d: moments rel err %.2e, decoded image rel err %.2e, arena %.2f GB" % (side, side, e_m, e_d, vae.arena_bytes() / 1e9))
    assert e_m < 3e-3 and e_d < 3e-3


def test_sketch_to_image_call_surface(cuda):
    """app.py:107-122 end to end on the engine: sketch image -> vae.encode(...).latent_dist -> sketch_image= of the pipeline ->
    guided sampling -> vae.decode -> numpy image (output_type "np")."""
    import copy
    from oracle import port
    from sketch2img_b200.latent_predictor import LatentEdgePredictor
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    from sketch2img_b200.unet import UNet2DConditionModel
    from sketch2img_b200.vae import AutoencoderKL
    o_unet = port.make_unet("tiny")
    o_lgp = port.make_lgp(o_unet)
    o_vae = port.make_vae("tiny")
    vae = AutoencoderKL(vars(o_vae.config), o_vae.state_dict())
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    lgp = LatentEdgePredictor(port.lgp_input_dim(o_unet), 4, port.NUM_POS_LAYERS)
    lgp.load_state_dict(copy.deepcopy(o_lgp).float().state_dict())
    pipe = AntiGradientPipeline(unet=unet, scheduler=DDIMScheduler(), vae=vae)
    pipe.setup_lgp(lgp)
    lat, emb, _ = port.make_inputs(o_unet)
    side = 8 * lat.shape[2]
    g = torch.Generator().manual_seed(5)
    sketch = (torch.tile(torch.rand(1, 1, side, side, generator=g).round(), (1, 3, 1, 1)) - 0.5) / 0.5
    target = vae.encode(sketch.cuda()).latent_dist.sample() * 0.18215           # app.py:109
    assert tuple(target.shape) == tuple(lat.shape)
    image = pipe("synthetic", num_inference_steps=4, guidance_scale=7.5, latents=lat.cuda(), sketch_image=target,
                 prompt_embeds=emb.cuda(), output_type="np")
    assert image.shape == (1, side, side, 3) and (image >= 0).all() and (image <= 1).all()
    # a sketch latent at another resolution is refused like the reference's mse_loss would
    with pytest.raises(ValueError):
        pipe("synthetic", num_inference_steps=2, latents=lat.cuda(), sketch_image=target[:, :, :4, :4], prompt_embeds=emb.cuda(),
             output_type="latent")
