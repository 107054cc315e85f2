"""CPU tests of the host-side logic and the C-ABI boundary (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from sketch2img_b200 import build
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "s2i.h")).read()
    declared = sorted(set(re.findall(r"\b(s2i_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/s2i.h but not exported by libs2i.so"


def test_ctypes_struct_matches_header_field_order():
    from sketch2img_b200 import _lib
    header = open(os.path.join(ROOT, "include", "s2i.h")).read()
    body = header[header.index("typedef struct s2i_gemm_desc {"):header.index("} s2i_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split("{", 1)[1].split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        for part in stmt.split(","):
            names.append(re.findall(r"([A-Za-z_0-9]+)\s*$", part.strip())[0])
    assert names == [f[0] for f in _lib.GemmDesc._fields_]


def test_scheduler_matches_oracle_scheduler():
    from oracle import port
    from sketch2img_b200.scheduler import DDIMScheduler
    mine, ref = DDIMScheduler(), port.make_scheduler()
    assert torch.equal(mine.alphas_cumprod, ref.alphas_cumprod)
    for n in (4, 25, 50):
        mine.set_timesteps(n)
        ref.set_timesteps(n)
        assert mine.timesteps.tolist() == ref.timesteps.tolist()
    # the fused step's fp32 formula, evaluated with torch ops, reproduces the oracle step bit for bit
    g = torch.Generator().manual_seed(3)
    x, eu, ec = (torch.randn(1, 4, 8, 8, generator=g) for _ in range(3))
    for t in mine.timesteps.tolist():
        sa_t, sb_t, sa_p, sb_p = (torch.tensor(v, dtype=torch.float32) for v in mine.step_coefficients(t))
        eps = eu + 7.5 * (ec - eu)
        want = ref.step(eps, torch.tensor(t), x, eta=0.0).prev_sample
        x0 = (x - sb_t * eps) / sa_t
        got = sa_p * x0 + sb_p * eps
        assert torch.equal(got, want)
        assert mine.sigma(t) == float((1 - ref.alphas_cumprod[t]) ** 0.5)


def test_pipeline_input_validation_without_gpu():
    from sketch2img_b200.pipeline import AntiGradientPipeline
    from sketch2img_b200.scheduler import DDIMScheduler
    pipe = AntiGradientPipeline(unet=None, scheduler=DDIMScheduler())
    with pytest.raises(ValueError):
        pipe.check_inputs(3, 512, 512, 1)
    with pytest.raises(ValueError):
        pipe.check_inputs("a", 500, 512, 1)
    with pytest.raises(ValueError):
        pipe.check_inputs("a", 512, 512, 0)
    pipe.check_inputs(["a", "b"], 512, 512, 1)
    with pytest.raises(ValueError):
        pipe.prepare_latents(1, 4, 512, 512, torch.float32, "cpu", None, torch.zeros(1, 4, 32, 32))
    lat = pipe.prepare_latents(2, 4, 64, 64, torch.float32, "cpu", torch.Generator().manual_seed(0))
    assert lat.shape == (2, 4, 8, 8)
    nl = pipe.get_noise_level(torch.ones(1, 4, 2, 2), torch.tensor(981))
    assert abs(nl[0, 0, 0, 0].item() - (1 - 0.0057755) ** 0.5) < 1e-4


def test_lgp_module_state_dict_contract():
    """edge_predictor.pt compatibility: keys of latent_predictor.py:15-29."""
    from sketch2img_b200.latent_predictor import LatentEdgePredictor, TAP_NAMES
    m = LatentEdgePredictor(9320, 4, 9)
    keys = set(m.state_dict().keys())
    for i in (0, 3, 6, 9, 12):
        assert {f"layers.{i}.weight", f"layers.{i}.bias"} <= keys
    for i in (2, 5, 8, 11):
        assert {f"layers.{i}.weight", f"layers.{i}.bias", f"layers.{i}.running_mean", f"layers.{i}.running_var",
                f"layers.{i}.num_batches_tracked"} <= keys
    assert m.layers[0].in_features == 9320 and m.layers[12].out_features == 4
    assert m.training and all(float(b.abs().sum()) == 0 for n, b in m.named_parameters() if n.endswith("bias") and "layers.1" != n[:8] and m.get_submodule(n.rsplit(".", 1)[0]).__class__.__name__ == "Linear")
    assert len(TAP_NAMES) == 9 and TAP_NAMES[3] == "mid_block.attentions.0"
    with pytest.raises(Exception):
        m(torch.zeros(2, 9280, 8, 8), torch.zeros(2, 4, 8, 8))     # CPU tensors: no CPU fallback


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sketch2img_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), f"{fn} references the oracle"


def test_satmixin_state_dict_contract():
    """Same sub-module names, order (down, up, mid) and parameter names/shapes as the reference's SatMixin
    (sketch_guided_attn.py:14-27, :62-72), so its checkpoints load into the drop-in unchanged."""
    from types import SimpleNamespace
    from oracle import port
    from sketch2img_b200.sketch_guided_attn import AttnModule, SatMixin
    from sketch2img_b200.unet import transformer_block_handles, transformer_block_paths
    o_unet = port.make_unet("tiny21")
    o_sat = port.make_sat(o_unet)
    cfg = {k: v for k, v in vars(o_unet.config).items() if not k.startswith("_")}
    fake = SimpleNamespace(config=o_unet.config, engine=None)
    fake.named_modules = lambda: iter([("", fake)] + transformer_block_handles(fake, cfg))
    sat = SatMixin(fake)
    assert len(sat.blocks) == 16 and [b.name for b in sat.blocks] == [b.name for b in o_sat.blocks]
    want = {k: tuple(v.shape) for k, v in o_sat.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in sat.state_dict().items()}
    assert got == want
    sat.load_state_dict(o_sat.state_dict())
    assert transformer_block_paths(cfg)[6] == "up_blocks.1.attentions.0" and transformer_block_paths(cfg)[-1] == "mid_block.attentions.0"
    # the reference's constructor signature: AttnModule(sat_name, base_layer) (sketch_guided_attn.py:48), width / heads read
    # off base_layer.attn1 like :51-60
    name, blk = transformer_block_handles(fake, cfg)[7]
    m = AttnModule("x", blk)
    assert m.sketch_norm.normalized_shape == (blk.attn1.to_q.in_features,) and m.sketch_attn.heads == blk.attn1.heads
    if os.path.isdir("/root/reference/modules"):
        # the reference's OWN SatMixin accepts the drop-in UNet's module tree and builds the same parameter set
        import sys
        port.add_shim_to_path()
        if "/root/reference" not in sys.path:
            sys.path.insert(0, "/root/reference")
        from modules.sketch_guided_attn import SatMixin as RefSatMixin
        ref = RefSatMixin(fake)
        assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == want


def test_dpmpp_host_scalars_reproduce_the_oracle_scheduler():
    """sketch2img_b200.scheduler.DPMSolverMultistepScheduler (constructed with the reference's own keyword arguments,
    app.py:14-25) hands the CUDA step the scalars of diffusers' first / second-order DPM-Solver++ updates: applied with the
    kernel's rounding order on CPU they reproduce the oracle scheduler's ``step`` bit for bit, over whole schedules
    (4 steps: first, second, second, lower-order-final first; 20 / 50 steps: first, then second order)."""
    import torch
    from oracle import port
    from sketch2img_b200.scheduler import DPMSolverMultistepScheduler
    o = port.make_scheduler(kind="dpmpp")
    p = DPMSolverMultistepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", num_train_timesteps=1000,
                                    trained_betas=None, predict_epsilon=True, thresholding=False, algorithm_type="dpmsolver++",
                                    solver_type="midpoint", lower_order_final=True)
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    for n, orders in ((4, [1, 2, 2, 1]), (20, [1] + [2] * 19), (50, [1] + [2] * 49)):
        o.set_timesteps(n)
        p.set_timesteps(n)
        assert torch.equal(o.timesteps, p.timesteps)
        assert [p.step_plan(i)["order"] for i in range(n)] == orders
        g = torch.Generator().manual_seed(n)
        x, m1 = torch.randn(1, 4, 8, 8, generator=g), None
        for i, t in enumerate(o.timesteps):
            e = torch.randn(1, 4, 8, 8, generator=g)
            want = o.step(e, t, x).prev_sample
            pl = p.step_plan(i)
            m0 = (x - f(pl["sigma_t"]) * e) / f(pl["alpha_t"])
            r = f(pl["c_x"]) * x - f(pl["c_m0"]) * m0
            if pl["order"] == 2:
                r = r - f(pl["c_d1"]) * (f(pl["inv_r0"]) * (m0 - m1))
            assert torch.equal(r, want), f"{n}-step schedule, step {i}"
            assert p.sigma(int(t)) == float((1 - o.alphas_cumprod[t]) ** 0.5)       # the LGP noise level of pipeline.py:133
            x, m1 = want, m0
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(solver_order=3)
    with pytest.raises(NotImplementedError):
        DPMSolverMultistepScheduler(algorithm_type="dpmsolver")
