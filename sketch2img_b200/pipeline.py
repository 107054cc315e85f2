"""Drop-in for /root/reference/modules/pipeline.py (``AntiGradientPipeline``) on the B200 engine.

Same call surface (pipeline.py:15-37, :132-174): ``setup_lgp``, ``__call__(prompt, height, width,
num_inference_steps, guidance_scale, negative_prompt, num_images_per_prompt, eta, generator, latents,
output_type, return_dict, callback, callback_steps, sketch_image)``, ``get_noise_level``,
``apply_anti_gradient``, ``decode_latents_L``.  The loop body (:83-115) runs as ONE C-ABI call per step
(``s2i_sampler_step``): CFG-doubled UNet forward, CFG combine + DDIM step, and on guided steps the LGP edge-loss
gradient through the UNet and the norm-ratio update.

Batch semantics follow SURVEY Q1: a batch of B prompts is B independent batch-1 reference calls (per-sample
BatchNorm statistics and per-sample step size).  Classifier-free guidance must be on (the reference's guided
path only executes that way).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .latent_predictor import LatentEdgePredictor, hook_unet


class _Progress:
    def __init__(self, total):
        self.total, self.n = total, 0

    def update(self, n=1):
        self.n += n

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class AntiGradientPipeline:
    def __init__(self, unet, scheduler, vae=None, text_encoder=None, tokenizer=None, safety_checker=None,
                 feature_extractor=None, requires_safety_checker=False):
        self.unet, self.scheduler, self.vae = unet, scheduler, vae
        self.text_encoder, self.tokenizer = text_encoder, tokenizer
        self.safety_checker, self.feature_extractor = safety_checker, feature_extractor
        self.vae_scale_factor = 8
        self.lgp_model = None
        self.feature_blocks = None
        self._sampler = None
        self._sampler_key = None
        self.max_samples_per_launch = 4      # bounds the activation arena; larger batches run in chunks per step
        self.last_losses = None

    # ------------------------------------------------------------------ reference surface
    def setup_lgp(self, lgp):
        """pipeline.py:15-17."""
        self.lgp_model: LatentEdgePredictor = lgp
        self.feature_blocks = hook_unet(self.unet)

    def to(self, device):
        return self

    @property
    def _execution_device(self):
        return self.unet.device

    def progress_bar(self, iterable=None, total=None):
        return _Progress(total)

    def check_inputs(self, prompt, height, width, callback_steps):
        if not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps} of type"
                             f" {type(callback_steps)}.")

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance, negative_prompt=None):
        """[uncond..., cond...] CLIP embeddings like diffusers; needs a tokenizer + text encoder."""
        if self.text_encoder is None or self.tokenizer is None:
            raise RuntimeError("no text encoder attached: pass prompt_embeds=[2B,77,D] ([uncond..., cond...])")
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        neg = negative_prompt if negative_prompt is not None else [""] * len(prompts)
        neg = [neg] * len(prompts) if isinstance(neg, str) else list(neg)

        def enc(texts):
            ids = self.tokenizer(texts, padding="max_length", max_length=self.tokenizer.model_max_length,
                                 truncation=True, return_tensors="pt").input_ids
            e = self.text_encoder(ids.to(self.text_encoder.device))[0]
            return e.repeat_interleave(num_images_per_prompt, dim=0)

        with torch.no_grad():
            cond = enc(prompts)
            if not do_classifier_free_guidance:
                return cond.to(device)
            return torch.cat([enc(neg), cond]).to(device)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
        else:
            if tuple(latents.shape) != shape:
                raise ValueError(f"Unexpected latents shape, got {tuple(latents.shape)}, expected {shape}")
            latents = latents.to(device=device, dtype=dtype)
        return latents * self.scheduler.init_noise_sigma

    def get_noise_level(self, noise, timesteps):
        """pipeline.py:132-139."""
        s = ((1 - self.scheduler.alphas_cumprod[timesteps]) ** 0.5).flatten()
        while len(s.shape) < len(noise.shape):
            s = s.unsqueeze(-1)
        return s.to(noise.device) * noise

    def _get_sampler(self):
        lgp_eng = self.lgp_model.engine() if self.lgp_model is not None else None
        key = (id(self.unet.engine), id(lgp_eng))
        if self._sampler is None or key != self._sampler_key:
            if self._sampler is not None:
                _lib.lib().s2i_sampler_destroy(self._sampler)
            h = C.c_void_p()
            _lib.check(_lib.lib().s2i_sampler_create(self.unet.engine._h, lgp_eng.handle if lgp_eng else None, C.byref(h)))
            self._sampler, self._sampler_key, self._lgp_eng = h, key, lgp_eng
        return self._sampler

    def __del__(self):
        try:
            if self._sampler is not None:
                _lib.lib().s2i_sampler_destroy(self._sampler)
                self._sampler = None
        except Exception:
            pass

    # ------------------------------------------------------------------ the sampling loop
    @torch.no_grad()
    def __call__(self, prompt, height=None, width=None, num_inference_steps=50, guidance_scale=7.5,
                 negative_prompt=None, num_images_per_prompt=1, eta=0.0, generator=None, latents=None,
                 output_type="pil", return_dict=True, callback=None, callback_steps=1, sketch_image=None,
                 prompt_embeds=None):
        lib = _lib.lib()
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        self.check_inputs(prompt, height, width, callback_steps)
        batch_size = 1 if isinstance(prompt, str) else len(prompt)
        device = self._execution_device
        if not guidance_scale > 1.0:
            raise NotImplementedError("the sketch-guided engine runs with classifier-free guidance on (guidance_scale > 1)")
        multistep = hasattr(self.scheduler, "step_plan")        # DPM-Solver++(2M): the demo's scheduler (app.py:14-25)
        if eta != 0.0 and not multistep:
            raise NotImplementedError("DDIM eta must be 0 on the fused step")
        if height != width and sketch_image is not None:
            raise RuntimeError("sketch guidance needs square latents (reference: pipeline.py:147 resizes to shape[2] only)")

        S = batch_size * num_images_per_prompt
        if prompt_embeds is None:
            prompt_embeds = self._encode_prompt(prompt, device, num_images_per_prompt, True, negative_prompt)
        emb = prompt_embeds.to(device, torch.float32)
        if emb.shape[0] != 2 * S:
            raise ValueError(f"prompt_embeds must be [2*{S},77,D] ordered [uncond..., cond...], got {tuple(emb.shape)}")
        # s2i_sampler_step takes the context as (uncond_s, cond_s) pairs (it permutes to its sample-major batch internally)
        ctx = torch.stack([emb[:S], emb[S:]], dim=1).reshape(2 * S, emb.shape[1], emb.shape[2]).contiguous()

        self.scheduler.set_timesteps(num_inference_steps, device=device)
        timesteps = self.scheduler.timesteps
        latents = self.prepare_latents(S, self.unet.in_channels, height, width, torch.float32, device, generator, latents)
        latents = latents.contiguous().clone()
        noise = latents.detach().clone()                                            # pipeline.py:75
        L = latents.shape[2]

        target = None
        if sketch_image is not None:
            if self.lgp_model is None:
                raise RuntimeError("call setup_lgp(lgp) before sampling with a sketch_image")
            target = sketch_image.to(device, torch.float32)
            if target.shape[0] == 1 and S > 1:
                target = target.expand(S, -1, -1, -1)
            if tuple(target.shape) != (S, self.unet.in_channels, L, L):
                # the reference fails here too (F.mse_loss broadcasting error, pipeline.py:157): a sketch latent encoded at
                # another resolution, or a batch that is neither 1 nor the number of samples
                raise ValueError(f"sketch_image must be [{S} or 1, {self.unet.in_channels}, {L}, {L}] (the VAE latent of the sketch at "
                                 f"the sampling resolution), got {tuple(sketch_image.shape)}")
            target = target.contiguous()
        sampler = self._get_sampler()
        _lib.check(lib.s2i_sampler_context_changed(sampler))       # new prompt embeddings: re-project the context K/V
        train = int(self.lgp_model.training) if self.lgp_model is not None else 1
        loss = torch.zeros(S, device=device, dtype=torch.float32)
        stream = _lib.stream_ptr()
        chunk = max(1, int(self.max_samples_per_launch))
        step_stop = 0.5 * len(timesteps)                                            # pipeline.py:90
        self.last_losses = []
        x0_hist = torch.zeros_like(latents) if multistep else None                  # the solver's x0-prediction history
        with self.progress_bar(total=num_inference_steps) as progress_bar:
            for i, t in enumerate(timesteps):
                ti = int(t)
                guided = int(i <= step_stop and target is not None)                 # pipeline.py:89-92, :108
                if multistep:
                    p = self.scheduler.step_plan(i)
                    for s0 in range(0, S, chunk):
                        s1 = min(S, s0 + chunk)
                        _lib.check(lib.s2i_sampler_step_dpmpp(
                            sampler, latents[s0:s1].data_ptr(), noise[s0:s1].data_ptr(), ctx[2 * s0:2 * s1].data_ptr(),
                            target[s0:s1].data_ptr() if target is not None else None, x0_hist[s0:s1].data_ptr(), s1 - s0, L,
                            float(ti), float(guidance_scale), p["alpha_t"], p["sigma_t"], p["c_x"], p["c_m0"], p["c_d1"],
                            p["inv_r0"], p["order"], self.scheduler.prediction, guided, 1.6, train, loss[s0:s1].data_ptr(),
                            stream))
                    progress_bar.update()
                    if callback is not None and i % callback_steps == 0:
                        callback(i, t, latents)
                    continue
                sa_t, sb_t, sa_p, sb_p = self.scheduler.step_coefficients(ti)
                sigma = self.scheduler.sigma(ti)
                for s0 in range(0, S, chunk):
                    s1 = min(S, s0 + chunk)
                    _lib.check(lib.s2i_sampler_step(
                        sampler, latents[s0:s1].data_ptr(), noise[s0:s1].data_ptr(), ctx[2 * s0:2 * s1].data_ptr(),
                        target[s0:s1].data_ptr() if target is not None else None, s1 - s0, L, float(ti),
                        float(guidance_scale), sa_t, sb_t, sa_p, sb_p, self.scheduler.prediction, guided, sigma, 1.6,
                        train, loss[s0:s1].data_ptr(), stream))
                progress_bar.update()
                if callback is not None and i % callback_steps == 0:
                    callback(i, t, latents)
        self.final_latents = latents
        if output_type == "latent":
            return latents if return_dict else (latents, None)
        image = self.decode_latents(latents)
        has_nsfw_concept = None
        if output_type == "pil":
            image = self.numpy_to_pil(image)
        if not return_dict:
            return (image, has_nsfw_concept)
        return image                                                                # pipeline.py:130 (bare list)

    # ------------------------------------------------------------------ guidance as a standalone call
    def apply_anti_gradient(self, latents_prev, latents, noise, timestep, target, beta):
        """pipeline.py:141-161 for the taps of the LAST ``unet(..., save_for_backward=True)`` forward on
        ``latents_prev`` ([2,4,h,w] = the CFG-doubled input)."""
        if target is None:
            return latents
        eng = self.unet.engine
        lgp = self.lgp_model.engine()
        taps = eng.taps()
        B, _, L, _ = latents_prev.shape
        if B != 2:
            # the reference only runs for one CFG pair: `latents_prev - latents` is [2B, ...] - [B, ...] (pipeline.py:160;
            # SURVEY Q1).  Batches go through __call__, which treats every sample as its own batch-1 reference call.
            raise RuntimeError(f"apply_anti_gradient handles one (uncond, cond) pair, got a batch of {B} (reference: "
                               f"pipeline.py:160 does not broadcast for more than one sample)")
        lgp.forward_taps(taps, B, L, noise.float().contiguous(), self.scheduler.sigma(int(timestep)), self.lgp_model.training)
        _, grads, _ = lgp.loss_backward(target.float().contiguous(), taps)
        dx = eng.backward(grads)
        out = latents.float().contiguous().clone()
        S = B // 2
        n = out[0].numel()
        scratch = torch.zeros(2 * S, device=out.device, dtype=torch.float64)
        x_old = latents_prev.float().reshape(S, 2, -1)[:, 0].contiguous()
        _lib.check(_lib.lib().s2i_guidance_update(x_old.data_ptr(), out.data_ptr(), dx.data_ptr(), S, n, float(beta),
                                                  scratch.data_ptr(), _lib.stream_ptr()))
        return out

    # ------------------------------------------------------------------ either side of the path
    def decode_latents(self, latents):
        if self.vae is None:
            return latents.detach().cpu().permute(0, 2, 3, 1).float().numpy()
        image = self.vae.decode(latents / 0.18215).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        return image.cpu().permute(0, 2, 3, 1).float().numpy()

    def decode_latents_L(self, latents):
        """pipeline.py:163-174."""
        image = self.vae.decode(1 / 0.18215 * latents).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        image = image.detach().cpu().permute(0, 2, 3, 1).float().numpy()
        image[image < 0.5] = 0
        image = image.squeeze(0) * 255
        return image.astype(np.uint8)

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None]
        images = (np.clip(images, 0, 1) * 255).round().astype("uint8")
        return [Image.fromarray(im[..., :3]) for im in images]
