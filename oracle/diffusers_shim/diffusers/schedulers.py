"""DDPMScheduler, diffusers 0.15 semantics (import site: train.py:8,32-36,96-98).

Defaults: beta_start 1e-4, beta_end 0.02, variance_type "fixed_small", clip_sample True.
"""
import torch


class _Cfg(dict):
    __getattr__ = dict.__getitem__


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, prediction_type="epsilon"):
        assert beta_schedule == "linear" and prediction_type == "epsilon" and variance_type == "fixed_small"
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                           beta_schedule=beta_schedule, variance_type=variance_type, clip_sample=clip_sample,
                           prediction_type=prediction_type)
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).flip(0)
        self.timesteps = ts.to(device) if device is not None else ts

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa = sa.unsqueeze(-1)
            sb = sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def step(self, model_output, timestep, sample, generator=None, noise=None):
        t = int(timestep)
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.config.clip_sample:
            x0 = x0.clamp(-1, 1)
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        ct = cur_alpha ** 0.5 * b_prev / b_t
        mean = c0 * x0 + ct * sample
        if t > 0:
            var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype,
                                    device=model_output.device)
            mean = mean + var ** 0.5 * noise
        return mean
