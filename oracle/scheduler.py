"""oracle/scheduler.py -- scalar restatement of the DDPM sampling arithmetic the reference's pipeline runs.
TEST INFRASTRUCTURE ONLY.  **Parity unpinned** (diffusers==0.20.0 ``DDPMScheduler`` is a third-party object absent
here; the reference builds it from the SD-2.1-base scheduler config, decoders/sd.py:48-50, and the patched pipeline
drives it through set_timesteps / scale_model_input / step, sd_pipeline_monkey_patch.py:153-154, 190, 216-218).

Written per step in float64 from the published DDPM equations (Ho et al. 2020, eq. 6-7, with the ``leading`` timestep
grid + ``steps_offset`` and the ``fixed_small`` variance of the scheduler config) -- deliberately not sharing a line
with the product's scheduler.py."""
from __future__ import annotations

import math

import torch


def sd21_alphas_cumprod(num_train=1000, beta_start=0.00085, beta_end=0.012):
    """``scaled_linear``: betas = linspace(sqrt(b0), sqrt(b1), T)^2 in fp32, cumprod of (1 - beta) in fp32."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def leading_timesteps(num_inference_steps, num_train=1000, steps_offset=1):
    ratio = num_train // num_inference_steps
    return [i * ratio + steps_offset for i in reversed(range(num_inference_steps))]


def ddpm_step_ref(eps, t, x_t, noise, num_inference_steps, acp=None, num_train=1000, prediction_type="epsilon"):
    """One ancestral step x_t -> x_{t-1}; ``noise`` is the standard-normal draw (ignored at t == 0)."""
    acp = sd21_alphas_cumprod(num_train) if acp is None else acp
    prev_t = t - num_train // num_inference_steps
    a_t = float(acp[t])
    a_prev = float(acp[prev_t]) if prev_t >= 0 else 1.0
    alpha_t, beta_t = a_t / a_prev, 1.0 - a_t / a_prev
    x_t, eps = x_t.double(), eps.double()
    if prediction_type == "epsilon":
        x0 = (x_t - math.sqrt(1.0 - a_t) * eps) / math.sqrt(a_t)
    elif prediction_type == "v_prediction":
        x0 = math.sqrt(a_t) * x_t - math.sqrt(1.0 - a_t) * eps
    else:
        raise NotImplementedError(prediction_type)
    mean = (math.sqrt(a_prev) * beta_t / (1.0 - a_t)) * x0 + (math.sqrt(alpha_t) * (1.0 - a_prev) / (1.0 - a_t)) * x_t
    if t > 0:
        var = max((1.0 - a_prev) / (1.0 - a_t) * beta_t, 1e-20)
        mean = mean + math.sqrt(var) * noise.double()
    return mean
