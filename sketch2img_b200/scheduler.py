"""Host half of the DDIM scheduler used on the sketch-guided path.

The reference drives a diffusers scheduler object (``set_timesteps`` / ``scale_model_input`` / ``step`` /
``alphas_cumprod`` at /root/reference/modules/pipeline.py:60,86,104,133).  Here the timetable and the
alpha-bar table live on the host; the per-element arithmetic of ``step`` is fused with the CFG combine in
``s2i_cfg_ddim_step`` / ``s2i_sampler_step``.  Stable-Diffusion configuration: scaled_linear betas
0.00085..0.012, clip_sample=False, set_alpha_to_one=False, steps_offset=1, eta=0.
"""
from types import SimpleNamespace

import numpy as np
import torch


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 steps_offset=1, set_alpha_to_one=False, prediction_type="epsilon", clip_sample=False):
        if clip_sample:
            raise NotImplementedError("clip_sample=True is not part of the SD configuration")
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      prediction_type=prediction_type)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts) + self.config.steps_offset      # stays on the host

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step_coefficients(self, t):
        """fp32 scalars of DDIM step t -> t_prev, rounded like the reference's 0-dim fp32 tensor ops:
        (sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev))."""
        t = int(t)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        return (float(a_t ** 0.5), float((1 - a_t) ** 0.5), float(a_p ** 0.5), float((1 - a_p) ** 0.5))

    def sigma(self, t):
        """sqrt(1 - alpha_bar_t) in fp32 (pipeline.py:133)."""
        return float((1 - self.alphas_cumprod[int(t)]) ** 0.5)

    @property
    def prediction(self):
        return {"epsilon": 0, "v_prediction": 1}[self.config.prediction_type]

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps].to(original_samples.dtype)
        sa, sb = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
