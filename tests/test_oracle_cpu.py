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
