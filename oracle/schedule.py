"""Schedule tables of the reference samplers (TEST INFRASTRUCTURE).

Restates, with the same dtypes and operation order (so the results are
bit-identical):
  * make_beta_schedule            dynamic/diffusionmodules/util.py:23-43
  * Schedule_DDPM.register_schedule   diffusion/sampler/ddpm_sampler.py:25-103
  * make_ddim_timesteps / make_ddim_sampling_parameters   util.py:46-74
  * DDIMSampler.make_schedule     diffusion/sampler/ddim_plms_sampler.py:38-81
"""
import numpy as np
import torch


def beta_schedule(schedule, n, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 numpy betas (util.py:23-43)."""
    if schedule == "linear":
        b = torch.linspace(linear_start**0.5, linear_end**0.5, n, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        ts = torch.arange(n + 1, dtype=torch.float64) / n + cosine_s
        a = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        a = a / a[0]
        b = 1 - a[1:] / a[:-1]
        b = np.clip(b, a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        b = torch.linspace(linear_start, linear_end, n, dtype=torch.float64)
    elif schedule == "sqrt":
        b = torch.linspace(linear_start, linear_end, n, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return b.numpy()


def ddpm_tables(num_timesteps, beta_schedule_name="linear", linear_start=1e-4, linear_end=2e-2,
                cosine_s=8e-3, v_posterior=0.0):
    """fp32 buffers registered by Schedule_DDPM (ddpm_sampler.py:25-103)."""
    betas = beta_schedule(beta_schedule_name, num_timesteps, linear_start, linear_end, cosine_s)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(ac),
        "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / ac)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / ac - 1)),
        "posterior_variance": f32(post_var),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(post_var, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(ac_prev) / (1.0 - ac)),
        "posterior_mean_coef2": f32((1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac)),
    }


def ddim_timesteps(num_ddim, num_ddpm):
    """'uniform' discretisation, +1 shift (util.py:46-60)."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def ddim_tables(alphas_cumprod, num_ddim, num_ddpm, eta):
    """DDIMSampler.make_schedule (ddim_plms_sampler.py:38-81).

    `alphas_cumprod` is the fp32 torch buffer of the DDPM schedule.  Dtypes follow
    the reference: alphas is an fp32 torch tensor, alphas_prev a float64 numpy
    array, sigmas a float64 torch tensor, sqrt(1-alphas) fp32 — each is rounded to
    fp32 only when broadcast with torch.full_like(x, v) (:360-366).
    """
    ts = ddim_timesteps(num_ddim, num_ddpm)
    ac = alphas_cumprod.cpu()
    alphas = ac[ts]
    alphas_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return {
        "timesteps": ts,
        "alphas": alphas,
        "alphas_prev": alphas_prev,
        "sigmas": sigmas,
        "sqrt_one_minus_alphas": np.sqrt(1.0 - alphas),
    }
