"""ORACLE / TEST INFRASTRUCTURE ONLY."""
from .unet_2d_condition import UNet2DConditionModel, UNet2DConditionOutput  # noqa: F401
from .attention import CrossAttention, BasicTransformerBlock  # noqa: F401
from .vae import AutoencoderKL  # noqa: F401
