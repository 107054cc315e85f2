"""ORACLE / TEST INFRASTRUCTURE ONLY -- not part of the shipped CUDA path.

The handful of ``StableDiffusionPipeline`` helpers the reference's subclass calls
(/root/reference/modules/pipeline.py:13,44,55,65,78,82,118,121,125).
No tokenizer/CLIP/VAE weights exist offline: ``_encode_prompt`` returns caller-installed synthetic
embeddings (``set_prompt_embeds``) and ``decode_latents`` degrades to a latent->numpy view when no
VAE is attached.
"""
import contextlib
import inspect

import numpy as np
import torch


class _Bar:
    def update(self, n=1):
        pass


class StableDiffusionPipeline:
    def __init__(self, vae=None, text_encoder=None, tokenizer=None, unet=None, scheduler=None,
                 safety_checker=None, feature_extractor=None, requires_safety_checker=False):
        self.vae, self.text_encoder, self.tokenizer = vae, text_encoder, tokenizer
        self.unet, self.scheduler = unet, scheduler
        self.safety_checker, self.feature_extractor = safety_checker, feature_extractor
        self.vae_scale_factor = 8
        self._prompt_embeds = None

    def to(self, device):
        self.unet.to(device)
        return self

    @property
    def _execution_device(self):
        return self.unet.device

    def set_prompt_embeds(self, embeds):
        """[2B,77,D] ordered [uncond, cond] (what _encode_prompt returns under CFG)."""
        self._prompt_embeds = embeds

    def check_inputs(self, prompt, height, width, callback_steps):
        if not isinstance(prompt, (str, list)):
            raise ValueError(f"`prompt` has to be of type `str` or `list` but is {type(prompt)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if callback_steps is None or not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError(f"`callback_steps` has to be a positive integer but is {callback_steps}")

    def _encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance,
                       negative_prompt=None):
        if self._prompt_embeds is None:
            raise RuntimeError("oracle shim has no text encoder: call set_prompt_embeds() first")
        e = self._prompt_embeds.to(device)
        return e if do_classifier_free_guidance else e.chunk(2)[1]

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_channels_latents, height // self.vae_scale_factor, width // self.vae_scale_factor)
        if latents is None:
            latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)
        else:
            if latents.shape != shape:
                raise ValueError(f"Unexpected latents shape, got {latents.shape}, expected {shape}")
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def prepare_extra_step_kwargs(self, generator, eta):
        params = set(inspect.signature(self.scheduler.step).parameters.keys())
        kw = {}
        if "eta" in params:
            kw["eta"] = eta
        if "generator" in params:
            kw["generator"] = generator
        return kw

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    def decode_latents(self, latents):
        if self.vae is None:
            return latents.detach().cpu().permute(0, 2, 3, 1).float().numpy()
        image = self.vae.decode(latents / 0.18215).sample
        image = (image / 2 + 0.5).clamp(0, 1)
        return image.cpu().permute(0, 2, 3, 1).float().numpy()

    def run_safety_checker(self, image, device, dtype):
        return image, None

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None]
        images = (np.clip(images, 0, 1) * 255).round().astype("uint8")
        return [Image.fromarray(im[..., :3]) for im in images]
