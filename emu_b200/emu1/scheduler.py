"""Host-side PNDM (PLMS) scheduler tables for the Emu1 generation pipeline.

Same configuration surface as the scheduler the reference loads from `<checkpoint>/scheduler` (diffusers ``PNDMScheduler``
in its Stable-Diffusion-1.5 configuration: scaled-linear betas, ``skip_prk_steps=True``, ``set_alpha_to_one=False``,
``steps_offset=1``, epsilon prediction) and the attribute names the reference pipeline touches
(Emu1/models/pipeline.py:94-127): ``set_timesteps``, ``timesteps``, ``init_noise_sigma``, ``scale_model_input`` (identity).
``step`` itself is fused into the CUDA denoise iteration (emu_denoise_step_multistep); ``step_coefficients(i)`` gives the eight
scalars of iteration ``i`` — the 4th-order Adams-Bashforth weights over the stored noise predictions and the two factors of
``_get_prev_sample``.  diffusers is not vendored under /root/reference: restated from the published algorithm ("parity
unpinned", like the Euler tables of Emu2-Gen).
"""
import json
import os

import torch


class PNDMScheduler:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 skip_prk_steps=True, set_alpha_to_one=False, prediction_type="epsilon", steps_offset=1,
                 timestep_spacing="leading", **_unused):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or not skip_prk_steps \
                or timestep_spacing != "leading":
            raise NotImplementedError("only the Stable-Diffusion-1.5 PNDM configuration of the Emu1 pipeline is implemented")
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else float(self.alphas_cumprod[0])
        self.timesteps = None
        self.num_inference_steps = None

    @classmethod
    def from_config(cls, path):
        if os.path.isdir(path):
            path = os.path.join(path, "scheduler_config.json")
        cfg = json.load(open(path))
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    from_pretrained = from_config

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        base = (torch.arange(0, num_inference_steps) * ratio).round().long() + self.steps_offset
        # PLMS without the Runge-Kutta warm-up repeats the second-to-last timestep: N + 1 UNet evaluations for N steps
        plms = torch.cat([base[:-1], base[-2:-1], base[-1:]]).flip(0)
        self.timesteps = plms
        return self

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _prev_sample_factors(self, t, t_prev):
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[t_prev]) if t_prev >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff, -(a_p - a_t) / denom

    def step_coefficients(self, i):
        """(a, b, wc, w0, w1, w2, 0, flags) of iteration i over self.timesteps (see emu_denoise_step_multistep)."""
        ratio = self.num_train_timesteps // self.num_inference_steps
        t = int(self.timesteps[i])
        counter = i
        n_hist = min(i - 1, 3) if i >= 2 else 0   # stored predictions BEFORE this iteration pushes (iteration 1 does not push)
        if counter == 1:        # the repeated timestep: average with the first prediction, restart from the saved sample
            t_prev, t_eff = t, t + ratio
            wc, w = 0.5, (0.5, 0.0, 0.0)
            flags = 2           # x = saved sample, no push
        else:
            t_prev, t_eff = t - ratio, t
            stored = {0: 0, 2: 1, 3: 2}.get(i, 3)  # history length before the push at iterations 0, 2, 3, >= 4
            if stored == 0:
                wc, w = 1.0, (0.0, 0.0, 0.0)
            elif stored == 1:
                wc, w = 1.5, (-0.5, 0.0, 0.0)
            elif stored == 2:
                wc, w = 23.0 / 12.0, (-16.0 / 12.0, 5.0 / 12.0, 0.0)
            else:
                wc, w = 55.0 / 24.0, (-59.0 / 24.0, 37.0 / 24.0, -9.0 / 24.0)
            flags = 1 | (4 if counter == 0 else 0)   # push; the first iteration also saves its input sample
        del n_hist
        a, b = self._prev_sample_factors(t_eff, t_prev)
        return (a, b, wc, w[0], w[1], w[2], 0.0, float(flags))
