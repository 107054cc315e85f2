"""the sketch feature encoder (modules/sketch_encoder.py:11-98) on the CUDA engine, against the fixture
written by the reference's own SketchEncoder class and against the CPU oracle, alone and as the producer of SatMixin's
features (sketch_guided_attn.py:29-40).  Tolerance 3e-3 relative L2 per feature map (fp16 operands, fp32 accumulation)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_sketch_encoder_matches_reference_fixture(cuda):
    from oracle import port
    from sketch2img_b200.sketch_encoder import SketchEncoder
    gold = torch.load(os.path.join(GOLD, "tiny21_sketch_encoder.pt"))
    o_enc = port.make_sketch_encoder(gold["config"])
    enc = SketchEncoder(vars(o_enc.config), o_enc.state_dict())
    for t in gold["timesteps"]:
        out = enc(gold["x"].cuda(), t).sample
        want = gold["res_samples"][t]
        assert [len(tup) for tup in out] == [3, 3, 3, 2]
        errs = [rel(m, n) for a, b in zip(out, want) for m, n in zip(a, b)]
        print("sketch encoder t=%d per-map rel err %s" % (t, " ".join("%.1e" % e for e in errs)))
        assert all(tuple(m.shape) == tuple(n.shape) for a, b in zip(out, want) for m, n in zip(a, b))
        assert max(errs) < 3e-3
    # same bits twice; a single sample equals its slice of the batch
    a = enc(gold["x"].cuda(), 500).sample
    b = enc(gold["x"].cuda(), 500).sample
    assert all(torch.equal(m, n) for u, v in zip(a, b) for m, n in zip(u, v))
    one = enc(gold["x"][1:].cuda(), 500).sample
    assert max(rel(m, n[1:]) for u, v in zip(one, a) for m, n in zip(u, v)) < 3e-3
    # the reference's forward cannot run cross-attention down blocks (no encoder_hidden_states reaches them): refuse them too
    with pytest.raises(ValueError):
        SketchEncoder(dict(vars(o_enc.config), down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",)), o_enc.state_dict())


def test_sketch_encoder_feeds_satmixin(cuda):
    """SketchEncoder -> SatMixin.set_res_samples -> UNet forward, all on the engine, against the same chain on the oracle."""
    from oracle import port
    from sketch2img_b200.sketch_encoder import SketchEncoder
    from sketch2img_b200.sketch_guided_attn import SatMixin
    from sketch2img_b200.unet import UNet2DConditionModel
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    o_unet = port.make_unet("tiny21")
    unet = UNet2DConditionModel(vars(o_unet.config), o_unet.state_dict())
    o_sat = port.make_sat(o_unet)
    sat = SatMixin(unet)
    sat.load_state_dict(o_sat.state_dict())
    o_enc = port.make_sketch_encoder("tiny21")
    enc = SketchEncoder(vars(o_enc.config), o_enc.state_dict())
    lat, emb, _ = port.make_inputs(o_unet)
    g = torch.Generator().manual_seed(21)
    sketch = torch.randn(2, 4, lat.shape[2], lat.shape[3], generator=g)       # VAE latent of the sketch, per CFG half
    x = torch.cat([lat] * 2)
    o_sat.set_res_samples(port.sketch_encoder_forward(o_enc, sketch, 0))
    o_sat.set_scale(0.8)
    with torch.no_grad():
        want = o_unet(x, torch.tensor(401), encoder_hidden_states=emb).sample
    sat.set_res_samples(enc(sketch.cuda(), 0).sample)
    sat.set_scale(0.8)
    got = unet(x.cuda(), 401, emb.cuda()).sample
    assert rel(got, want) < 3e-3
