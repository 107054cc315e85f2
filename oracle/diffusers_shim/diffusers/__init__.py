"""ORACLE / TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the `diffusers` package.

The reference (/root/reference/modules/{pipeline,latent_predictor,sketch_guided_attn}.py) imports
diffusers, which is absent from this image and cannot be installed (no network).  Putting this
directory's parent on sys.path lets the reference's own files run UNMODIFIED on CPU; see
oracle/make_golden.py and SURVEY.md section 8(c).  Nothing under sketch2img_b200/ imports this.
"""
from .models.unet_2d_condition import UNet2DConditionModel  # noqa: F401
from .models.vae import AutoencoderKL  # noqa: F401
from .schedulers import DDIMScheduler, DPMSolverMultistepScheduler  # noqa: F401
from .pipeline_sd import StableDiffusionPipeline  # noqa: F401

__version__ = "0.13.0+oracle-shim"
