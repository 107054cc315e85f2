"""CPU tests of the oracle (test infrastructure): the port restates the reference loop bit-exactly, matches the
committed golden fixtures, and the reference's own files still agree when /root/reference is present."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _run_port(name, steps):
    from oracle import port
    unet = port.make_unet(name)
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    got = {}
    out = port.guided_sample(unet, lgp, port.make_scheduler(), emb, lat.clone(), tgt, num_steps=steps,
                             callback=lambda i, t, l: got.__setitem__(int(i), l.detach().clone()))
    return out, got


@pytest.mark.parametrize("steps", [4, 50])
def test_port_matches_golden_tiny(steps):
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, f"tiny_{steps}step.pt"))
    out, got = _run_port("tiny", steps)
    for i, ref in gold["latents"].items():
        assert torch.equal(got[i], ref), f"step {i} differs from the fixture made by the reference files"
    assert abs(out.norm().item() - gold["norms"][-1].item()) < 1e-3 * gold["norms"][-1].item()


@pytest.mark.parametrize("steps", [4, 20])
def test_port_matches_golden_dpmpp(steps):
    """The demo's scheduler (DPM-Solver++(2M), app.py:14-25): the port's loop over the shim scheduler reproduces the fixture
    the unmodified reference pipeline wrote with that scheduler, bit for bit."""
    from oracle import port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, f"tiny_dpmpp_{steps}step.pt"))
    assert gold["scheduler"] == "dpmpp"
    unet = port.make_unet("tiny")
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    got = {}
    port.guided_sample(unet, lgp, port.make_scheduler(kind="dpmpp"), emb, lat.clone(), tgt, num_steps=steps,
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().clone()))
    for i, ref in gold["latents"].items():
        assert torch.equal(got[i], ref), f"step {i} differs from the fixture made by the reference files"


def test_golden_fixture_metadata():
    for name in ["tiny_4step", "tiny_50step", "sd15_4step", "sd15_50step", "tiny_dpmpp_4step", "tiny_dpmpp_20step"]:
        g = torch.load(os.path.join(GOLD, name + ".pt"))
        assert g["source"].startswith("reference modules/pipeline.py")
        assert g["steps"] - 1 in g["latents"]
        assert torch.isfinite(g["norms"]).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/modules"), reason="reference sources only exist in the authoring container")
def test_reference_files_over_shim_equal_port():
    """The reference's pipeline.py + latent_predictor.py, imported UNMODIFIED over oracle/diffusers_shim, give
    bit-identical per-step latents to oracle/port.py."""
    from oracle import port
    from oracle.make_golden import run_reference
    ref_steps, _ = run_reference("tiny", 4)
    _, got = _run_port("tiny", 4)
    for i in range(4):
        assert torch.equal(ref_steps[i], got[i])


def test_reference_quirks_restated():
    """train-mode BatchNorm, 3-of-4 guided steps, noise level from the initial noise."""
    from oracle import port
    unet = port.make_unet("tiny")
    lgp = port.make_lgp(unet)
    assert lgp.training and all(p.dtype == torch.float16 for p in lgp.parameters())
    assert port.lgp_input_dim(unet) == 64 + 128 + 256 * 3 + 256 + 256 + 256 + 128 + 40
    sch = port.make_scheduler()
    sch.set_timesteps(4)
    assert sch.timesteps.tolist() == [751, 501, 251, 1]
    sch.set_timesteps(50)
    assert sch.timesteps[:2].tolist() == [981, 961] and sch.timesteps[-1].item() == 1
    assert abs(sch.alphas_cumprod[981].item() - 0.0057755) < 1e-6
    assert sum(1 for i in range(50) if i <= 0.5 * 50) == 26


# ------------------------------------------------------------------------------------------------ injected sketch attention
def _port_sat_run(steps=4):
    from oracle import port
    unet = port.make_unet("tiny21")
    sat = port.make_sat(unet)
    lat, emb, _ = port.make_inputs(unet)
    sat.set_res_samples(port.make_res_samples(unet, 2))
    sat.set_scale(0.7)
    with torch.no_grad():
        eps = unet(torch.cat([lat] * 2), torch.tensor(501), encoder_hidden_states=emb).sample
    got = {}
    port.guided_sample(unet, None, port.make_scheduler("v_prediction"), emb, lat.clone(), None, num_steps=steps,
                       callback=lambda i, t, l: got.__setitem__(int(i), l.detach().clone()))
    return eps, got, sat


def test_port_sat_matches_golden():
    """oracle/port.py's SatMixin restatement + plain CFG / v-prediction DDIM reproduce the fixture made by the reference's
    own sketch_guided_attn.py + pipeline.py bit for bit."""
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, "tiny21_sat_4step.pt"))
    eps, got, sat = _port_sat_run(gold["steps"])
    assert torch.equal(eps, gold["eps_t501"])
    for i, ref in gold["latents"].items():
        assert torch.equal(got[i], ref), f"step {i}"
    assert len(sat.blocks) == 16 and sat.blocks[6].name == "sketch_attn_up_blocks_1_attentions_0_transformer_blocks_0"


@pytest.mark.skipif(not os.path.isdir("/root/reference/modules"), reason="reference sources only exist in the authoring container")
def test_reference_satmixin_over_shim_equals_port():
    from oracle import port
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from modules.sketch_guided_attn import SatMixin
    unet = port.make_unet("tiny21")
    sat = port.make_sat(unet, SatMixin)
    lat, emb, _ = port.make_inputs(unet)
    sat.set_res_samples(port.make_res_samples(unet, 2))
    sat.set_scale(0.7)
    with torch.no_grad():
        eps_ref = unet(torch.cat([lat] * 2), torch.tensor(501), encoder_hidden_states=emb).sample
    eps, _, sat_port = _port_sat_run(1)
    assert torch.equal(eps, eps_ref)
    assert list(sat.state_dict().keys()) == list(sat_port.state_dict().keys())


# ------------------------------------------------------------------------------------------------ teacher-forced fixtures
def test_port_guided_step_reproduces_teacher_fixture():
    """tests/golden/tiny_50step_teacher.pt holds every guided step of the reference run (latent before / after guidance,
    loss, gradient norm).  oracle/port.py's guided_step restarted from the fixture's x_{i-1} must give the fixture's x_i bit
    for bit: the fixture is the reference files' own trajectory, and the port's single step is the reference's step."""
    from oracle import port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    fix = torch.load(os.path.join(GOLD, "tiny_50step_teacher.pt"))
    assert fix["guided_steps"] == 26 and fix["source"].startswith("reference modules/pipeline.py")
    unet = port.make_unet("tiny")
    lgp = port.make_lgp(unet)
    lat, emb, tgt = port.make_inputs(unet)
    sch = port.make_scheduler()
    sch.set_timesteps(fix["steps"])
    taps, _ = port.register_taps(unet)
    noise = lat * sch.init_noise_sigma
    for i in (0, 1, 13, 25):
        x_prev = noise if i == 0 else fix["x"][i - 1]
        rec = {}
        x = port.guided_step(unet, lgp, sch, emb, x_prev, noise, sch.timesteps[i], tgt, True, taps, 7.5, 1.6, rec)
        assert int(sch.timesteps[i]) == fix["t"][i]
        assert torch.equal(x, fix["x"][i]) and torch.equal(rec["x_ddim"], fix["x_ddim"][i]), f"step {i}"
        assert abs(rec["loss"] - float(fix["loss"][i])) <= 1e-6 * float(fix["loss"][i])
    # the yardsticks the GPU tests lean on are what the docstring of make_golden.run_reference_teacher says they are
    assert (fix["fp16w"]["e16"] > 1e-3).all() and (fix["fp16w"]["d16"] < 1e-4).all()
    assert (fix["pert"]["e_pert"] > 5e-4).all()        # a 1e-6 restart moves the next latent by ~1e-3: chaotic


def test_teacher_fixture_sd15_metadata():
    fix = torch.load(os.path.join(GOLD, "sd15_50step_teacher.pt"))
    assert fix["config"] == "sd15" and fix["steps"] == 50 and fix["guided_steps"] == 26
    assert fix["t"][0] == 981 and fix["t"][25] == 481
    assert all(tuple(fix["x"][i].shape) == (1, 4, 64, 64) for i in range(26))
    assert torch.isfinite(fix["loss"]).all() and torch.isfinite(fix["gnorm"]).all()


def test_fp16_operand_floor_of_one_forward():
    """Why a 4-step unguided run cannot reach 1e-3 (it is asserted at 2e-3; the 50-step run is asserted at 1e-3): the ORACLE
    itself, evaluated with its weights and every Conv2d / Linear input rounded to fp16 -- exact fp32 accumulation, exact
    norms / softmax -- is already ~1.2e-3 away from its fp32 self in ONE forward.  That is the floor of the fp16 tensor-core
    operand format (10 mantissa bits, the same as TF32), not of the kernels; the CUDA forward measures 1.2-1.3e-3."""
    import copy
    from oracle import port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    unet = port.make_unet("tiny")
    lat, emb, _ = port.make_inputs(unet)
    x = torch.cat([lat] * 2)
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    with torch.no_grad():
        ref = unet(x, torch.tensor(751), encoder_hidden_states=emb).sample
        u16 = copy.deepcopy(unet)
        for p_ in u16.parameters():
            p_.copy_(p_.half().float())
        e_w = rel(u16(x, torch.tensor(751), encoder_hidden_states=emb).sample, ref)
        pre = lambda m, inp: tuple(i.half().float() if torch.is_tensor(i) and i.is_floating_point() else i for i in inp)
        for m in u16.modules():
            if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
                m.register_forward_pre_hook(pre)
        e_wa = rel(u16(x, torch.tensor(751), encoder_hidden_states=emb).sample, ref)
    print("oracle one-forward eps error: fp16 weights %.3e, fp16 weights + fp16 GEMM inputs %.3e" % (e_w, e_wa))
    assert 5e-4 < e_w < 1.5e-3 and 8e-4 < e_wa < 2.5e-3


# ------------------------------------------------------------------------------------------------ sketch feature encoder
def test_port_sketch_encoder_matches_golden():
    """oracle/port.py's restatement of SketchEncoder.forward (sketch_encoder.py:50-98) reproduces the fixture written by the
    reference's own class bit for bit."""
    from oracle import port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, "tiny21_sketch_encoder.pt"))
    enc = port.make_sketch_encoder(gold["config"])
    for t in gold["timesteps"]:
        got = port.sketch_encoder_forward(enc, gold["x"], t)
        want = gold["res_samples"][t]
        assert [len(tup) for tup in got] == [3, 3, 3, 2] == [len(tup) for tup in want]
        for a, b in zip(got, want):
            for m, n in zip(a, b):
                assert torch.equal(m, n)


@pytest.mark.skipif(not os.path.isdir("/root/reference/modules"), reason="reference sources only exist in the authoring container")
def test_reference_sketch_encoder_over_shim_equals_port():
    from oracle import port
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    from modules.sketch_encoder import SketchEncoder
    ref = port.make_sketch_encoder("tiny21", cls=SketchEncoder)
    enc = port.make_sketch_encoder("tiny21")
    x = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = ref(x, 321).sample
    got = port.sketch_encoder_forward(enc, x, 321)
    for a, b in zip(got, want):
        for m, n in zip(a, b):
            assert torch.equal(m, n)
    # with the standard cross-attention down blocks the reference's forward cannot run (no encoder_hidden_states reaches them)
    bad = SketchEncoder(**port.CONFIGS["tiny21"])
    with pytest.raises(Exception):
        bad(x, 321)


# ------------------------------------------------------------------------------------------------ VAE either side of the loop
def test_vae_fixture_is_what_the_oracle_computes():
    """tests/golden/tiny_vae.pt: posterior moments / decoded image of the shim AutoencoderKL and the uint8 image the
    reference's decode_latents_L (pipeline.py:163-174) makes of the same latent; the port's restatement reproduces it."""
    from oracle import port
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, "tiny_vae.pt"))
    vae = port.make_vae(gold["config"])
    with torch.no_grad():
        assert torch.equal(vae.encode(gold["image"]).latent_dist.parameters, gold["moments"])
        assert torch.equal(vae.decode(gold["latents"] / 0.18215).sample, gold["decoded"])
        img = torch.from_numpy(port.decode_latents_L(vae, gold["latents"]))
    assert torch.equal(img, gold["image_L"]) and img.dtype == torch.uint8 and ((img == 0) | (img >= 127)).all()
    keys = list(vae.state_dict().keys())
    assert "encoder.down_blocks.0.resnets.0.norm1.weight" in keys and "decoder.mid_block.attentions.0.proj_attn.bias" in keys
    assert "quant_conv.weight" in keys and "encoder.down_blocks.2.downsamplers.0.conv.weight" in keys
