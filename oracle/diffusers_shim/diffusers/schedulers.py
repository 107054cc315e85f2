"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped CUDA path.

CPU restatement of diffusers' ``DDIMScheduler`` in the Stable-Diffusion configuration (scaled_linear
betas 0.00085..0.012, clip_sample=False, set_alpha_to_one=False, steps_offset=1), the scheduler named by
BASELINE.json and called by the reference at /root/reference/modules/pipeline.py:60,86,104,133.
Algorithm: SURVEY.md Appendix A.6 (diffusers is un-vendored; ~v0.12-0.13).
"""
from types import SimpleNamespace

import numpy as np
import torch


class DDIMSchedulerOutput(dict):
    def __init__(self, prev_sample, pred_original_sample=None):
        super().__init__(prev_sample=prev_sample, pred_original_sample=pred_original_sample)
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                 steps_offset=1, prediction_type="epsilon"):
        if beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                                      clip_sample=clip_sample, prediction_type=prediction_type)
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)
        self.timesteps += self.config.steps_offset

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False,
             generator=None, variance_noise=None, return_dict=True):
        prev_t = timestep - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        elif self.config.prediction_type == "v_prediction":
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            model_output = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            raise ValueError(self.config.prediction_type)
        if self.config.clip_sample:
            x0 = torch.clamp(x0, -1, 1)
        b_prev = 1 - a_prev
        variance = (b_prev / b_t) * (1 - a_t / a_prev)
        std = eta * variance ** 0.5
        direction = (1 - a_prev - std ** 2) ** 0.5 * model_output
        prev = a_prev ** 0.5 * x0 + direction
        if eta > 0:
            noise = variance_noise if variance_noise is not None else torch.randn(
                model_output.shape, generator=generator, dtype=model_output.dtype).to(model_output.device)
            prev = prev + std * noise
        return DDIMSchedulerOutput(prev, x0)

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps].to(original_samples.dtype)
        sa, sb = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise


class SchedulerOutput(dict):
    def __init__(self, prev_sample):
        super().__init__(prev_sample=prev_sample)
        self.prev_sample = prev_sample


class DPMSolverMultistepScheduler:
    """CPU restatement of diffusers' ``DPMSolverMultistepScheduler`` as the reference's demo configures it
    (/root/reference/app.py:14-25, evaluation.py:21-32: scaled_linear betas 0.00085..0.012, ``algorithm_type=
    "dpmsolver++"``, ``solver_order=2`` (default), ``solver_type="midpoint"``, ``lower_order_final=True``,
    ``thresholding=False``, epsilon prediction) -- DPM-Solver++(2M), Lu et al. 2022, Algorithm 2, in diffusers' (~v0.12)
    arithmetic order: data prediction x0 = (x - sigma_t eps) / alpha_t kept as the multistep history,
        first order :  x' = (sigma_p / sigma_t) x - alpha_p (exp(-h) - 1) m0
        second order:  x' = (sigma_p / sigma_t) x - alpha_p (exp(-h) - 1) m0 - 0.5 alpha_p (exp(-h) - 1) (m0 - m1) / r0
    with lambda = log alpha - log sigma, h = lambda_p - lambda_t, r0 = (lambda_t - lambda_prev_t) / h; the first step
    (and, for fewer than 15 steps, the last) is first order.  PARITY UNPINNED against genuine diffusers (un-vendored)."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, solver_order=2, prediction_type="epsilon", thresholding=False,
                 dynamic_thresholding_ratio=0.995, sample_max_value=1.0, algorithm_type="dpmsolver++",
                 solver_type="midpoint", lower_order_final=True, predict_epsilon=None):
        if predict_epsilon is not None:            # deprecated spelling the reference still uses (app.py:20)
            prediction_type = "epsilon" if predict_epsilon else "sample"
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if algorithm_type != "dpmsolver++" or solver_type != "midpoint" or solver_order not in (1, 2) or thresholding:
            raise NotImplementedError("oracle shim: only dpmsolver++ / midpoint / order <= 2 / no thresholding (app.py:14-25)")
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.init_noise_sigma = 1.0
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                      prediction_type=prediction_type, algorithm_type=algorithm_type, solver_type=solver_type,
                                      lower_order_final=lower_order_final, thresholding=thresholding)
        self.num_inference_steps = None
        ts = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy()
        self.timesteps = torch.from_numpy(ts)
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ts = (np.linspace(0, self.config.num_train_timesteps - 1, num_inference_steps + 1).round()[::-1][:-1].copy()
              .astype(np.int64))
        self.timesteps = torch.from_numpy(ts).to(device)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0

    def convert_model_output(self, model_output, timestep, sample):
        alpha_t, sigma_t = self.alpha_t[timestep], self.sigma_t[timestep]
        if self.config.prediction_type == "epsilon":
            return (sample - sigma_t * model_output) / alpha_t
        if self.config.prediction_type == "sample":
            return model_output
        if self.config.prediction_type == "v_prediction":
            return alpha_t * sample - sigma_t * model_output
        raise ValueError(self.config.prediction_type)

    def dpm_solver_first_order_update(self, model_output, timestep, prev_timestep, sample):
        lambda_t, lambda_s = self.lambda_t[prev_timestep], self.lambda_t[timestep]
        alpha_t = self.alpha_t[prev_timestep]
        sigma_t, sigma_s = self.sigma_t[prev_timestep], self.sigma_t[timestep]
        h = lambda_t - lambda_s
        return (sigma_t / sigma_s) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * model_output

    def multistep_dpm_solver_second_order_update(self, model_output_list, timestep_list, prev_timestep, sample):
        t, s0, s1 = prev_timestep, timestep_list[-1], timestep_list[-2]
        m0, m1 = model_output_list[-1], model_output_list[-2]
        lambda_t, lambda_s0, lambda_s1 = self.lambda_t[t], self.lambda_t[s0], self.lambda_t[s1]
        alpha_t = self.alpha_t[t]
        sigma_t, sigma_s0 = self.sigma_t[t], self.sigma_t[s0]
        h, h_0 = lambda_t - lambda_s0, lambda_s0 - lambda_s1
        r0 = h_0 / h
        D0, D1 = m0, (1.0 / r0) * (m0 - m1)
        return ((sigma_t / sigma_s0) * sample - (alpha_t * (torch.exp(-h) - 1.0)) * D0
                - 0.5 * (alpha_t * (torch.exp(-h) - 1.0)) * D1)

    def step(self, model_output, timestep, sample, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        step_index = (self.timesteps == timestep).nonzero()
        step_index = len(self.timesteps) - 1 if len(step_index) == 0 else step_index.item()
        prev_timestep = 0 if step_index == len(self.timesteps) - 1 else self.timesteps[step_index + 1]
        lower_order_final = ((step_index == len(self.timesteps) - 1) and self.config.lower_order_final
                             and len(self.timesteps) < 15)
        model_output = self.convert_model_output(model_output, timestep, sample)
        for i in range(self.config.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = model_output
        if self.config.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev_sample = self.dpm_solver_first_order_update(model_output, timestep, prev_timestep, sample)
        else:
            timestep_list = [self.timesteps[step_index - 1], timestep]
            prev_sample = self.multistep_dpm_solver_second_order_update(self.model_outputs, timestep_list, prev_timestep,
                                                                       sample)
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        return SchedulerOutput(prev_sample)

    def add_noise(self, original_samples, noise, timesteps):
        a = self.alphas_cumprod.to(original_samples.device)[timesteps].to(original_samples.dtype)
        sa, sb = a ** 0.5, (1 - a) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise
