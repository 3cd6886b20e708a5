"""DDPM ("native"), DDIM and PLMS reverse loops (TEST INFRASTRUCTURE).

Restates diffusion/sampler/ddpm_sampler.py:154-238, ddim_plms_sampler.py:302-525,
diffusion_utils/util.py:70-82,99-100 and diffusion/ddpm.py:108-122, with the random
draws replaced by a host-supplied NOISE TAPE so that the reference, this oracle and
the CUDA path consume identical noise (SURVEY.md §8c caveat iii):

    tape = {'x_T': [B,C,H,W] fp32, 'noise': [n_steps,B,C,H,W] fp32}

noise[k] is the k-th per-step draw in loop order (the reference draws one
randn(shape) per step even when it is multiplied by zero).
"""
import numpy as np
import torch

from . import schedule as S


def clip_x0(pred_x0, clip_denoised, dtp):
    """clip_x0_minus_one_to_one (diffusion_utils/util.py:70-82)."""
    if dtp < 1.0:
        s = torch.quantile(pred_x0.flatten(1).abs(), dtp, dim=-1)
        s = s.clamp(min=1.0).view(-1, *([1] * (pred_x0.dim() - 1)))
        return pred_x0.clamp(-s, s) / s
    if clip_denoised:
        return pred_x0.clamp(-1.0, 1.0)
    return pred_x0


def to_uint8(img):
    """clip_unnormalize_to_zero_to_255 (diffusion_utils/util.py:99-100)."""
    return ((img + 1) * 127.5).clamp(0, 255).to(torch.uint8)


def _log_indices(total, log_num_per_prog):
    return torch.linspace(0, total, log_num_per_prog, dtype=torch.int).cpu().numpy().tolist()


def _ext(a, t, x):
    """extract_into_tensor (util.py:96-99)."""
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))


def ddpm_sample(eps_fn, tape, tables, sampling_kwargs):
    """Schedule_DDPM.sample / p_sample / p_mean_variance (ddpm_sampler.py:154-238)."""
    T = sampling_kwargs["num_timesteps"]
    assert tables["alphas_cumprod"].shape[0] == T  # ddpm_sampler.py:37-38,49
    temperature = sampling_kwargs["temperature"]
    if type(temperature) == float:
        temperature = [temperature] * T
    logs = _log_indices(T, sampling_kwargs["log_num_per_prog"])
    x = tape["x_T"].clone()
    b = x.shape[0]
    out = dict(pred_x0=[], x_inter=[])
    k = 0
    for i in reversed(range(T)):
        t = torch.full((b,), i, dtype=torch.long)
        eps = eps_fn(x, t)
        x0 = _ext(tables["sqrt_recip_alphas_cumprod"], t, x) * x - _ext(tables["sqrt_recipm1_alphas_cumprod"], t, x) * eps
        x0 = clip_x0(x0, sampling_kwargs["clip_denoised"], sampling_kwargs["dtp"])
        mean = _ext(tables["posterior_mean_coef1"], t, x) * x0 + _ext(tables["posterior_mean_coef2"], t, x) * x
        logvar = _ext(tables["posterior_log_variance_clipped"], t, x)
        noise = tape["noise"][k] * temperature[i]
        if sampling_kwargs["noise_dropout"] > 0:  # F.dropout(noise, p) (ddpm_sampler.py:184-185), factor from the tape
            noise = noise * tape["dropout_mul"][k]
        k += 1
        nonzero = (1 - (t == 0).float()).reshape(b, *((1,) * (x.dim() - 1)))
        x = mean + nonzero * (0.5 * logvar).exp() * noise
        if i in logs:
            out["pred_x0"].append(x0.unsqueeze(0))
            out["x_inter"].append(x.unsqueeze(0))
    out = {k_: torch.cat(v, 0) for k_, v in out.items()}
    return x, out


def _ddim_update(x, e_t, dt, index, noise, sampling_kwargs, dropout_mul=None):
    """p_sample_ddim / p_sample_plms arithmetic (ddim_plms_sampler.py:360-391,493-525)."""
    a_t = torch.full_like(x, dt["alphas"][index])
    a_prev = torch.full_like(x, dt["alphas_prev"][index])
    sigma_t = torch.full_like(x, dt["sigmas"][index])
    s1m = torch.full_like(x, dt["sqrt_one_minus_alphas"][index])
    x0 = (x - s1m * e_t) / a_t.sqrt()
    x0 = clip_x0(x0, sampling_kwargs["clip_denoised"], sampling_kwargs["dtp"])
    dir_xt = (1.0 - a_prev - sigma_t**2).sqrt() * e_t
    nz = sigma_t * noise * sampling_kwargs["temperature"]
    if sampling_kwargs["noise_dropout"] > 0:  # F.dropout(noise, p) (ddim_plms_sampler.py:388-389), factor from the tape
        nz = nz * dropout_mul
    return a_prev.sqrt() * x0 + dir_xt + nz, x0


def ddim_sample(eps_fn, tape, alphas_cumprod, num_ddpm, sampling_kwargs):
    """DDIMSampler.ddim_sampling core loop (ddim_plms_sampler.py:302-343)."""
    dt = S.ddim_tables(alphas_cumprod, sampling_kwargs["num_timesteps"], num_ddpm, sampling_kwargs["ddim_eta"])
    ts = dt["timesteps"]
    total = ts.shape[0]
    logs = _log_indices(total, sampling_kwargs["log_num_per_prog"])
    x = tape["x_T"].clone()
    b = x.shape[0]
    out = dict(pred_x0=[], x_inter=[])
    for i, step in enumerate(np.flip(ts)):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        e_t = eps_fn(x, t)
        x, x0 = _ddim_update(x, e_t, dt, index, tape["noise"][i], sampling_kwargs,
                             tape["dropout_mul"][i] if "dropout_mul" in tape else None)
        if index in logs:
            out["x_inter"].append(x.unsqueeze(0))
            out["pred_x0"].append(x0.unsqueeze(0))
    out = {k_: torch.cat(v, 0) for k_, v in out.items()}
    return x, out


def plms_sample(eps_fn, tape, alphas_cumprod, num_ddpm, sampling_kwargs):
    """DDIMSampler.plms_sampling (ddim_plms_sampler.py:393-480); eta forced to 0 (:41-46).
    The first step draws TWO noises (two p_sample_plms calls); every later step one."""
    kw = dict(sampling_kwargs)
    kw["ddim_eta"] = 0
    dt = S.ddim_tables(alphas_cumprod, kw["num_timesteps"], num_ddpm, 0)
    ts = dt["timesteps"]
    total = ts.shape[0]
    time_range = np.flip(ts)
    logs = _log_indices(total, kw["log_num_per_prog"])
    x = tape["x_T"].clone()
    b = x.shape[0]
    old = []
    k = 0
    out = dict(pred_x0=[], x_inter=[])
    for i, step in enumerate(time_range):
        index = total - i - 1
        t = torch.full((b,), int(step), dtype=torch.long)
        t_next = torch.full((b,), int(time_range[min(i + 1, len(time_range) - 1)]), dtype=torch.long)
        e_t = eps_fn(x, t)
        if len(old) == 0:
            x_prev, _ = _ddim_update(x, e_t, dt, index, tape["noise"][k], kw,
                                     tape["dropout_mul"][k] if "dropout_mul" in tape else None)
            k += 1
            e_next = eps_fn(x_prev, t_next)
            e_p = (e_t + e_next) / 2
        elif len(old) == 1:
            e_p = (3 * e_t - old[-1]) / 2
        elif len(old) == 2:
            e_p = (23 * e_t - 16 * old[-1] + 5 * old[-2]) / 12
        else:
            e_p = (55 * e_t - 59 * old[-1] + 37 * old[-2] - 9 * old[-3]) / 24
        x, x0 = _ddim_update(x, e_p, dt, index, tape["noise"][k], kw,
                             tape["dropout_mul"][k] if "dropout_mul" in tape else None)
        k += 1
        old.append(e_t)
        if len(old) >= 4:
            old.pop(0)
        if index in logs:
            out["pred_x0"].append(x0.unsqueeze(0))
            out["x_inter"].append(x.unsqueeze(0))
    out = {k_: torch.cat(v, 0) for k_, v in out.items()}
    return x, out


def pndm_sample(eps_fn, tape, num_ddpm, sampling_kwargs, beta_start=1e-4, beta_end=2e-2):
    """PNDM_Sampler.pndm_sampling + PNDMScheduler (diffusion/sampler/pndm_sampler.py:13-141,187-211): float32
    `np.linspace` betas (NOT the DDPM's sqrt-linear schedule), float32 cumprod with a trailing 0.0, four Runge-Kutta
    warm-up stages x 3, then the 4-term linear multistep; only x_T is drawn."""
    n = sampling_kwargs["num_timesteps"]
    betas = np.linspace(beta_start, beta_end, num_ddpm, dtype=np.float32)
    ac = torch.from_numpy(np.array(list(np.cumprod(1.0 - betas, axis=0)) + [0.0], dtype=np.float32))
    step = num_ddpm // n
    times = list(range(0, num_ddpm, step))
    w = np.array(times[-4:]).repeat(2) + np.tile(np.array([0, step // 2]), 4)
    warm = list(reversed(w[:-1].repeat(2)[1:-1]))
    steps = list(reversed(times[:-3]))

    def transfer(x, t, t_next, et):
        at, at_next = ac[t + 1].view(-1, 1, 1, 1), ac[t_next + 1].view(-1, 1, 1, 1)
        x_delta = (at_next - at) * ((1 / (at.sqrt() * (at.sqrt() + at_next.sqrt()))) * x
                                    - 1 / (at.sqrt() * (((1 - at_next) * at).sqrt() + ((1 - at) * at_next).sqrt())) * et)
        return x + x_delta

    image = tape["x_T"].clone()
    b = image.shape[0]
    cur_residual, cur_image, ets = 0, None, []
    for t in range(len(warm)):
        residual = eps_fn(image, torch.full((b,), int(warm[t]), dtype=torch.long))
        t_prev, t_next = warm[t // 4 * 4], warm[min(t + 1, len(warm) - 1)]
        if t % 4 == 0:
            cur_residual = cur_residual + 1 / 6 * residual
            ets.append(residual)
            cur_image = image
        elif (t - 1) % 4 == 0 or (t - 2) % 4 == 0:
            cur_residual = cur_residual + 1 / 3 * residual
        else:
            residual = cur_residual + 1 / 6 * residual
            cur_residual = 0
        image = transfer(cur_image, t_prev, t_next, residual)
    for t in range(len(steps)):
        t_prev, t_next = steps[t], steps[min(t + 1, len(steps) - 1)]
        ets.append(eps_fn(image, torch.full((b,), int(steps[t]), dtype=torch.long)))
        residual = (1 / 24) * (55 * ets[-1] - 59 * ets[-2] + 37 * ets[-3] - 9 * ets[-4])
        image = transfer(image, t_prev, t_next, residual)
    return image, dict(pred_x0=image.unsqueeze(0), x_inter=image.unsqueeze(0))


def p_sample_loop(method, eps_fn, tape, diffusion_cfg, sampling_kwargs):
    """LatentDiffusion.p_sample_loop (diffusion/ddpm.py:108-122): dispatch + uint8."""
    T = diffusion_cfg["num_timesteps"]
    tables = S.ddpm_tables(
        T,
        diffusion_cfg.get("beta_schedule", "linear"),
        diffusion_cfg.get("linear_start", 1e-4),
        diffusion_cfg.get("linear_end", 2e-2),
        diffusion_cfg.get("cosine_s", 8e-3),
        diffusion_cfg.get("v_posterior", 0.0),
    )
    if method == "native":
        x, inter = ddpm_sample(eps_fn, tape, tables, sampling_kwargs)
    elif method == "ddim":
        x, inter = ddim_sample(eps_fn, tape, tables["alphas_cumprod"], T, sampling_kwargs)
    elif method == "plms":
        x, inter = plms_sample(eps_fn, tape, tables["alphas_cumprod"], T, sampling_kwargs)
    elif method == "pndm":
        x, inter = pndm_sample(eps_fn, tape, T, sampling_kwargs, diffusion_cfg.get("linear_start", 1e-4),
                               diffusion_cfg.get("linear_end", 2e-2))
    else:
        raise KeyError(method)
    inter = dict(inter)
    inter["pred_x0"] = to_uint8(inter["pred_x0"])
    return to_uint8(x), inter, x
