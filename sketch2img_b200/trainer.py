"""The LGP training step of /root/reference/trainer.py:208-252 on the CUDA engine.

The reference's loop body: noised latents -> ``unet(noisy_latents, timesteps, encoder_hidden_states)`` (frozen, no CFG) -> the 9
hooked features resized to the latent size and concatenated -> ``edge_predictor(features, noise_level)`` ->
``mse_loss(result, sketchs)`` -> ``backward`` -> ``optimizer.step()``.  As shipped that body raises ``NameError`` at :240
(``intermidiate_result`` / ``intermediate_result`` typo); this module implements what it evidently
means -- the concatenated features of :237-244 go into the predictor.

Here: one engine forward per latent (each has its own timestep, trainer.py:228), the taps gathered into batch tensors, then
``s2i_lgp_forward_taps_batch`` (BatchNorm statistics over all ``bsz * L * L`` rows, like ``nn.BatchNorm1d`` in train mode on the
trainer's batch) and ``s2i_lgp_train_step`` (loss, tcgen05 weight-gradient GEMMs, AdamW on the fp32 masters).  The reference's
optimizer is bitsandbytes ``AdamW8bit``; the update rule here is plain AdamW (fp32 moments) -- the 8-bit state quantisation is
not reproduced.  No CPU fallback.
"""
import torch


class AdamWState:
    """Hyper-parameters + step counter of the engine-side AdamW (the moments live on the device inside the LGP engine)."""

    def __init__(self, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0


def get_noise_level(noise, alphas_cumprod, timesteps):
    """trainer.py:197-204."""
    s = ((1 - alphas_cumprod[timesteps]) ** 0.5).flatten()
    while len(s.shape) < len(noise.shape):
        s = s.unsqueeze(-1)
    return s.to(noise.device) * noise


@torch.no_grad()
def training_step(unet, edge_predictor, noisy_latents, timesteps, encoder_hidden_states, noise_level, sketchs, optimizer):
    """trainer.py:233-251 for one batch.  noisy_latents / noise_level / sketchs: [bsz, 4, L, L]; timesteps: [bsz] ints;
    encoder_hidden_states: [bsz, 77, D].  Returns the loss (python float); ``edge_predictor``'s engine holds the updated weights
    (``edge_predictor.pull_from_engine()`` copies them back into the ``nn.Module``)."""
    dev = unet.device
    bsz, _, L, L2 = noisy_latents.shape
    if L != L2:
        raise RuntimeError("the LGP resizes every feature to latents.shape[2] (trainer.py:238): square latents only")
    x = noisy_latents.to(dev, torch.float32)
    ctx = encoder_hidden_states.to(dev, torch.float32)
    taps = None
    for b in range(bsz):                                    # per-latent timestep: one frozen-UNet forward each (trainer.py:235)
        unet.engine.forward(x[b:b + 1], float(timesteps[b]), ctx[b:b + 1])
        tb = [t.contiguous() for t in unet.engine.taps()]
        if taps is None:
            taps = [torch.empty((bsz,) + tuple(t.shape[1:]), device=dev, dtype=torch.float32) for t in tb]
        for k in range(9):
            taps[k][b] = tb[k][0]
    eng = edge_predictor.engine()
    eng.forward_taps_batch(taps, bsz, L, noise_level.to(dev, torch.float32).contiguous())
    optimizer.step_count += 1
    return eng.train_step(sketchs.to(dev, torch.float32).contiguous(), optimizer.lr, optimizer.betas, optimizer.eps,
                          optimizer.weight_decay, optimizer.step_count)
