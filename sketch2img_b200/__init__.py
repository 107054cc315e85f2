"""sketch2img_b200 -- Blackwell (sm_100a) native sketch-guided Stable-Diffusion sampling path.

Drop-in for the hot path of Mikubill/sketch2img (modules/pipeline.py, modules/latent_predictor.py,
modules/sketch_guided_attn.py): Python/PyTorch host code over hand-written CUDA behind a C ABI
(include/s2i.h, sketch2img_b200/libs2i.so).
"""
__version__ = "0.1.0"
