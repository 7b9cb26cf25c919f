"""Host-side Euler discrete scheduler tables for Emu2-Gen.

Same configuration surface as the scheduler the reference instantiates from
Emu2/emu/conf/diffusion_config/scheduler/scheduler_config.json (diffusers EulerDiscreteScheduler: scaled-linear
betas, "leading" timestep spacing, steps_offset 1, epsilon prediction, linear sigma interpolation) and the same
attribute names the reference pipeline touches (Emu2/emu/diffusion.py:116-149): set_timesteps, timesteps, sigmas,
init_noise_sigma.  scale_model_input / step are fused into the CUDA denoise step (emu_denoise_step).
"""
import json
import os

import torch


class EulerDiscreteScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon", interpolation_type="linear", timestep_spacing="leading", steps_offset=1,
                 use_karras_sigmas=False, **_unused):
        if beta_schedule != "scaled_linear" or prediction_type != "epsilon" or use_karras_sigmas \
                or interpolation_type != "linear":
            raise NotImplementedError("only the Emu2-Gen scheduler configuration is implemented")
        self.num_train_timesteps = num_train_timesteps
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self._sigmas_all = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.timesteps = None
        self.sigmas = torch.cat([self._sigmas_all.flip(0), torch.zeros(1)])

    @classmethod
    def from_config(cls, path):
        if os.path.isdir(path):
            path = os.path.join(path, "scheduler_config.json")
        cfg = json.load(open(path))
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def set_timesteps(self, num_inference_steps, device=None):
        n = self.num_train_timesteps
        if self.timestep_spacing == "leading":
            ratio = n // num_inference_steps
            ts = (torch.arange(0, num_inference_steps) * ratio).round().flip(0).float() + self.steps_offset
        elif self.timestep_spacing == "trailing":
            ratio = n / num_inference_steps
            ts = (torch.arange(n, 0, -ratio)).round().float() - 1
        else:  # linspace
            ts = torch.linspace(0, n - 1, num_inference_steps).flip(0).float()
        lo = ts.floor().long().clamp(max=n - 1)
        hi = (lo + 1).clamp(max=n - 1)
        frac = ts - lo.float()
        sig = self._sigmas_all[lo] * (1 - frac) + self._sigmas_all[hi] * frac
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.timesteps = ts
        return self

    @property
    def init_noise_sigma(self):
        if self.timestep_spacing in ("linspace", "trailing"):
            return float(self.sigmas.max())
        return float((self.sigmas.max() ** 2 + 1) ** 0.5)
