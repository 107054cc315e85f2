"""Host halves of the schedulers used on the sketch-guided path: DDIM (BASELINE.json's configuration) and the
DPM-Solver++(2M) scheduler the reference's demo constructs (/root/reference/app.py:14-25).

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


class DPMSolverMultistepScheduler:
    """Host half of diffusers' ``DPMSolverMultistepScheduler`` with the constructor arguments the reference's demo passes
    (/root/reference/app.py:14-25, evaluation.py:21-32): ``algorithm_type="dpmsolver++"``, ``solver_type="midpoint"``,
    ``solver_order`` 1 or 2, ``lower_order_final``, no thresholding.  The timetable, the alpha / sigma / lambda tables and the
    per-step scalars live here (fp32 0-dim torch arithmetic in diffusers' order, so the scalars round like the
    reference's); the per-element update -- and the x0-prediction history it needs -- runs in ``s2i_cfg_dpmpp_step`` /
    ``s2i_sampler_step_dpmpp``."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, sample_max_value=1.0, algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, predict_epsilon=None):
        if predict_epsilon is not None:            # deprecated spelling used at app.py:20
            prediction_type = "epsilon" if predict_epsilon else "sample"
        if prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError(f"prediction_type {prediction_type!r} is not supported by the fused step")
        if algorithm_type != "dpmsolver++" or solver_type != "midpoint" or solver_order not in (1, 2) or thresholding:
            raise NotImplementedError("the fused step implements dpmsolver++ / midpoint / solver_order <= 2 without "
                                      "thresholding (the configuration of app.py:14-25)")
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                      prediction_type=prediction_type, algorithm_type=algorithm_type, solver_type=solver_type,
                                      lower_order_final=lower_order_final, thresholding=thresholding)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ts = (np.linspace(0, self.config.num_train_timesteps - 1, self.num_inference_steps + 1).round()[::-1][:-1].copy()
              .astype(np.int64))
        self.timesteps = torch.from_numpy(ts)                                  # stays on the host

    def scale_model_input(self, sample, timestep=None):
        return sample

    def sigma(self, t):
        """sqrt(1 - alpha_bar_t) in fp32 (pipeline.py:133)."""
        return float((1 - self.alphas_cumprod[int(t)]) ** 0.5)

    @property
    def prediction(self):
        return {"epsilon": 0, "v_prediction": 1}[self.config.prediction_type]

    def step_plan(self, i):
        """Scalars of step i of the current timetable (diffusers ``step`` -> ``dpm_solver_first_order_update`` /
        ``multistep_dpm_solver_second_order_update``): dict(order, alpha_t, sigma_t, c_x, c_m0, c_d1, inv_r0)."""
        ts = self.timesteps
        n = len(ts)
        t = int(ts[i])
        prev = 0 if i == n - 1 else int(ts[i + 1])
        lower_final = i == n - 1 and self.config.lower_order_final and n < 15
        first = self.config.solver_order == 1 or i == 0 or lower_final
        lam_p, lam_t = self.lambda_t[prev], self.lambda_t[t]
        h = lam_p - lam_t
        c_x = self.sigma_t[prev] / self.sigma_t[t]
        c_m0 = self.alpha_t[prev] * (torch.exp(-h) - 1.0)
        plan = {"order": 1, "alpha_t": float(self.alpha_t[t]), "sigma_t": float(self.sigma_t[t]), "c_x": float(c_x),
                "c_m0": float(c_m0), "c_d1": 0.0, "inv_r0": 0.0}
        if not first:
            h_0 = lam_t - self.lambda_t[int(ts[i - 1])]
            r0 = h_0 / h
            plan.update(order=2, c_d1=float(0.5 * c_m0), inv_r0=float(1.0 / r0))
        return plan

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps].to(original_samples.dtype)
        sa, sb = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
