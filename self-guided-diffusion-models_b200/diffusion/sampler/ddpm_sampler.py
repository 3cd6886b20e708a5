"""Drop-in for diffusion/sampler/ddpm_sampler.py `Schedule_DDPM` (the 'native' sampler).

Schedule buffers are built on the host with the reference's float64 numpy formulas
(ddpm_sampler.py:25-103) and are bit-identical; each reverse step is ONE fused CUDA kernel
(guidance mix + predict_start_from_noise + clamp + q_posterior + noise add,
ddpm_sampler.py:121-192) fed by the batched cond||uncond UNet pass.
"""
from functools import partial

import numpy as np
import torch
from torch import nn

from ... import _lib
from ...diffusion_utils import dict2obj
from ._common import GuidedEps, NoiseSource, StepExtras, check_supported, coef6, log_indices, make_beta_schedule


class Schedule_DDPM(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.hparams = dict2obj(kwargs)
        self.register_schedule(
            given_betas=self.hparams.given_betas, beta_schedule=self.hparams.beta_schedule,
            timesteps=self.hparams.num_timesteps, linear_start=self.hparams.linear_start,
            linear_end=self.hparams.linear_end, cosine_s=self.hparams.cosine_s)

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000,
                          linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
        if given_betas is not None:
            betas = given_betas
        else:
            betas = make_beta_schedule(beta_schedule, self.hparams.num_timesteps, linear_start=linear_start,
                                       linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1.0 - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
        if timesteps < self.hparams.num_timesteps:
            raise NotImplementedError  # ddpm_sampler.py:37-38
        self.linear_start, self.linear_end = linear_start, linear_end
        assert alphas_cumprod.shape[0] == timesteps, "alphas have to be defined for each timestep"
        dev = self.hparams.device
        reg = lambda name, arr: self.register_buffer(name, torch.tensor(arr, dtype=torch.float32).to(dev))
        v_post = self.hparams.v_posterior
        reg("betas", betas)
        reg("alphas_cumprod", alphas_cumprod)
        reg("alphas_cumprod_prev", alphas_cumprod_prev)
        reg("sqrt_alphas_cumprod", np.sqrt(alphas_cumprod))
        reg("sqrt_one_minus_alphas_cumprod", np.sqrt(1.0 - alphas_cumprod))
        reg("log_one_minus_alphas_cumprod", np.log(1.0 - alphas_cumprod))
        reg("sqrt_recip_alphas_cumprod", np.sqrt(1.0 / alphas_cumprod))
        reg("sqrt_recipm1_alphas_cumprod", np.sqrt(1.0 / alphas_cumprod - 1))
        posterior_variance = (1 - v_post) * betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod) + v_post * betas
        reg("posterior_variance", posterior_variance)
        reg("posterior_log_variance_clipped", np.log(np.maximum(posterior_variance, 1e-20)))
        reg("posterior_mean_coef1", betas * np.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod))
        reg("posterior_mean_coef2", (1.0 - alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - alphas_cumprod))
        # Training-side buffers (ddpm_sampler.py:85-103).  The sampling path never reads them, but `snr_derivative`
        # and `SNR` are persistent, i.e. part of every reference checkpoint's `diffusion.sampler.*` keys: a strict
        # load_state_dict needs them here.  `lvlb_weights` is non-persistent like the reference's.
        if self.hparams.parameterization == "eps":
            lvlb_weights = self.betas ** 2 / (2 * self.posterior_variance *
                                              torch.tensor(alphas, dtype=torch.float32).to(dev) * (1 - self.alphas_cumprod))
        elif self.hparams.parameterization == "x0":
            lvlb_weights = 0.5 * np.sqrt(torch.Tensor(alphas_cumprod)) / (2.0 * 1 - torch.Tensor(alphas_cumprod))
        else:
            raise NotImplementedError("mu not supported")
        lvlb_weights[0] = lvlb_weights[1]
        self.register_buffer("lvlb_weights", lvlb_weights.to(dev), persistent=False)
        assert not torch.isnan(self.lvlb_weights).all()
        self.register_buffer("snr_derivative", torch.zeros(1000, dtype=torch.float32).to(dev))
        self.register_buffer("SNR", torch.zeros(1000, dtype=torch.float32).to(dev))

    def _x0_coefs(self, tab, i):
        """(c_recip, c_recipm1) of `x_recon = c_recip * x - c_recipm1 * model_out` (ddpm_sampler.py:158-163):
        predict_start_from_noise for parameterization 'eps'; for 'x0' the model output IS x_recon, which the same
        fused kernel reproduces exactly with (0, -1): 0 * x - (-1 * out) == out bit for bit."""
        if self.hparams.parameterization == "x0":
            return 0.0, -1.0
        return tab["sqrt_recip_alphas_cumprod"][i], tab["sqrt_recipm1_alphas_cumprod"][i]

    @torch.no_grad()
    def p_sample(self, x, t, clip_denoised=False, repeat_noise=False, temperature=1.0, noise_dropout=0.0,
                 sampling_kwargs=None, denoise_sample_fn=None, denoise_sample_fn_kwargs=None, noise=None,
                 index=None, **kwargs):
        """One reverse step (ddpm_sampler.py:175-192): guided eps + fused posterior update.
        `t` must be batch-uniform (it always is while sampling); pass `index` = that timestep to
        avoid reading it back from the device.  `noise` is an optional host-supplied draw."""
        check_supported(sampling_kwargs)
        if repeat_noise:
            raise NotImplementedError("repeat_noise")
        i = int(t[0]) if index is None else int(index)
        device = x.device
        if not hasattr(self, "_step_tab") or self._step_tab[0] is not self.posterior_log_variance_clipped:
            tab = {k: getattr(self, k).detach().cpu() for k in
                   ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                    "posterior_mean_coef2", "posterior_log_variance_clipped")}
            tab["sigma"] = (0.5 * tab["posterior_log_variance_clipped"]).exp()
            self._step_tab = (self.posterior_log_variance_clipped, tab)
        tab = self._step_tab[1]
        eps_src = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device, fresh_weights=False)
        x = x.detach().float().contiguous()
        pc, pu, w, w_ptr, st = eps_src(x, t.to(device=device, dtype=torch.long).contiguous())
        nz = torch.randn(x.shape, device=device) if noise is None else noise.to(device, torch.float32).contiguous()
        out, x0 = torch.empty_like(x), torch.empty_like(x)
        c = coef6(*self._x0_coefs(tab, i), tab["posterior_mean_coef1"][i], tab["posterior_mean_coef2"][i],
                  tab["sigma"][i] if i != 0 else 0.0, temperature)
        stream = _lib.current_stream(device)
        extras = StepExtras(sampling_kwargs, x, noise_dropout=noise_dropout)
        dyn, mul = extras.pointers(stream, 0, (pc, pu, w, w_ptr, st), c, x, x.shape[0], x.shape[1:].numel())
        _lib.check(_lib.lib().sgdm_ddpm_step_ex(stream, pc, pu, w, w_ptr, st, c,
                                                1 if sampling_kwargs["clip_denoised"] else 0, x.data_ptr(), nz.data_ptr(),
                                                out.data_ptr(), x0.data_ptr(), x.shape[0], x.shape[1:].numel(), dyn, mul))
        return out, x0, None

    @torch.no_grad()
    def sample(self, shape, sampling_kwargs=None, **kwargs):
        """Schedule_DDPM.sample (ddpm_sampler.py:194-238) -> (x, {'pred_x0','x_inter'})."""
        device = torch.device(self.hparams.device)
        if device.type != "cuda":
            raise _lib.SgdmError(f"sampler device is {device}: sgdm_b200 has no CPU path")
        with torch.cuda.device(device):  # kernels launch on the CURRENT device: make it the sampler's
            return self._sample(shape, sampling_kwargs=sampling_kwargs, **kwargs)

    def _sample(self, shape, sampling_kwargs=None, denoise_sample_fn=None, denoise_sample_fn_kwargs=None,
                condition_kwargs=None, noise_tape=None, **kwargs):
        check_supported(sampling_kwargs)
        temperature = sampling_kwargs["temperature"]
        timesteps = sampling_kwargs["num_timesteps"]
        self.register_schedule(timesteps=timesteps, given_betas=self.hparams.given_betas,
                               beta_schedule=self.hparams.beta_schedule, linear_start=self.hparams.linear_start,
                               linear_end=self.hparams.linear_end, cosine_s=self.hparams.cosine_s)
        device = torch.device(self.hparams.device)
        B = shape[0]
        lib, stream = _lib.lib(), _lib.current_stream(device)
        noise = NoiseSource(shape, device, noise_tape)
        img = noise.x_T().contiguous()
        nxt = torch.empty_like(img)
        if type(temperature) == float:
            temperature = [temperature] * timesteps
        logs = log_indices(timesteps, sampling_kwargs["log_num_per_prog"])
        # per-step scalars, from the fp32 buffers exactly as extract_into_tensor would read them
        tab = {k: getattr(self, k).detach().cpu() for k in
               ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1",
                "posterior_mean_coef2", "posterior_log_variance_clipped")}
        sigma = (0.5 * tab["posterior_log_variance_clipped"]).exp()
        eps_src = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device)
        clip = 1 if sampling_kwargs["clip_denoised"] else 0
        per_sample = img.shape[1:].numel()
        extras = StepExtras(sampling_kwargs, img, noise)
        out = dict(pred_x0=[], x_inter=[])
        for i in reversed(range(0, timesteps)):
            ts = torch.full((B,), i, device=device, dtype=torch.long)
            pc, pu, w, w_ptr, st = eps_src(img, ts)
            nz = noise.next()
            x0 = torch.empty_like(img) if i in logs else None
            c = coef6(*self._x0_coefs(tab, i), tab["posterior_mean_coef1"][i], tab["posterior_mean_coef2"][i],
                      sigma[i] if i != 0 else 0.0, temperature[i])
            dyn, mul = extras.pointers(stream, 0, (pc, pu, w, w_ptr, st), c, img, B, per_sample)
            _lib.check(lib.sgdm_ddpm_step_ex(stream, pc, pu, w, w_ptr, st, c, clip, img.data_ptr(), nz.data_ptr(),
                                             nxt.data_ptr(), _lib.ptr(x0), B, per_sample, dyn, mul))
            img, nxt = nxt, img
            if i in logs:
                out["pred_x0"].append(x0.unsqueeze(0))
                out["x_inter"].append(img.clone().unsqueeze(0))
        out["pred_x0"] = torch.cat(out["pred_x0"], 0)
        out["x_inter"] = torch.cat(out["x_inter"], 0)
        return img, out
