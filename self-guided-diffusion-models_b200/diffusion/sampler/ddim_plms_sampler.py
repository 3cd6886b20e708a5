"""Drop-in for diffusion/sampler/ddim_plms_sampler.py `DDIMSampler` (ddim and plms).

make_schedule reproduces the reference's mixed fp32 / float64 dtype chain
(ddim_plms_sampler.py:38-81, diffusionmodules/util.py:46-74) so every per-step scalar is
bit-identical after its torch.full_like(x, v) rounding (:360-366).  Each step is one fused
CUDA kernel (guidance mix + Eq.12 update, :358-391).  Intermediates are returned on the
CPU like the reference (:331-335) but copied once at the end instead of 9 blocking D2H
copies inside the loop.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ._common import (GuidedEps, NoiseSource, StepExtras, check_supported, coef6, log_indices, make_ddim_sampling_parameters,
                      make_ddim_timesteps)


class DDIMSampler(object):
    def __init__(self, ddpm_num_timesteps, device, sampler_type):
        super().__init__()
        self.ddpm_num_timesteps = ddpm_num_timesteps
        self.device = device
        self.sampler_type = sampler_type

    def make_schedule(self, sampling_kwargs, ddim_discretize="uniform", **kwargs):
        ddim_num_steps = sampling_kwargs["num_timesteps"]
        ddim_eta = sampling_kwargs["ddim_eta"]
        if ddim_eta != 0 and self.sampler_type in ["plms"]:
            ddim_eta = 0  # ddim_plms_sampler.py:41-46
        alphas_cumprod = sampling_kwargs["alphas_cumprod"]
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps)
        assert alphas_cumprod.shape[0] == self.ddpm_num_timesteps, "alphas have to be defined for each timestep"
        sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(
            alphacums=alphas_cumprod.cpu(), ddim_timesteps=self.ddim_timesteps, eta=ddim_eta)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sigmas, alphas, alphas_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)
        # fp32 per-step scalars exactly as torch.full_like(x, v) + the fp32 tensor ops produce them
        n = len(self.ddim_timesteps)
        a_t = torch.stack([torch.full_like(torch.zeros(()), alphas[i]) for i in range(n)])
        a_prev = torch.stack([torch.full_like(torch.zeros(()), alphas_prev[i]) for i in range(n)])
        sig = torch.stack([torch.full_like(torch.zeros(()), sigmas[i]) for i in range(n)])
        s1m = torch.stack([torch.full_like(torch.zeros(()), self.ddim_sqrt_one_minus_alphas[i]) for i in range(n)])
        self._coefs = dict(s1m=s1m, sqrt_at=a_t.sqrt(), sqrt_a_prev=a_prev.sqrt(),
                           dir=(1.0 - a_prev - sig**2).sqrt(), sigma=sig)

    @torch.no_grad()
    def sample(self, shape, sampling_kwargs=None, **kwargs):
        self.make_schedule(sampling_kwargs=sampling_kwargs)
        device = torch.device(self.device)
        if device.type != "cuda":
            raise _lib.SgdmError(f"sampler device is {device}: sgdm_b200 has no CPU path")
        with torch.cuda.device(device):  # kernels launch on the CURRENT device: make it the sampler's
            if self.sampler_type == "ddim":
                return self.ddim_sampling(shape, sampling_kwargs=sampling_kwargs, **kwargs)
            elif self.sampler_type == "plms":
                return self.plms_sampling(shape, sampling_kwargs=sampling_kwargs, **kwargs)
        raise NotImplementedError

    def _setup(self, shape, sampling_kwargs, noise_tape):
        check_supported(sampling_kwargs)
        device = torch.device(self.device)
        if device.type != "cuda":
            raise _lib.SgdmError(f"sampler device is {device}: sgdm_b200 has no CPU path")
        noise = NoiseSource(shape, device, noise_tape)
        img = noise.x_T().contiguous()
        self._extras = StepExtras(sampling_kwargs, img, noise)
        return device, noise, img

    def _step(self, stream, eps, index, clip, temperature, img, nz, nxt, x0, B, per_sample, eps_out=None):
        pc, pu, w, w_ptr, st = eps
        k = self._coefs
        c = coef6(k["s1m"][index], k["sqrt_at"][index], k["sqrt_a_prev"][index], k["dir"][index], k["sigma"][index],
                  temperature)
        dyn, mul = self._extras.pointers(stream, 1, eps, c, img, B, per_sample)
        _lib.check(_lib.lib().sgdm_ddim_step_ex(stream, pc, pu, w, w_ptr, st, c, clip, img.data_ptr(), nz.data_ptr(),
                                                nxt.data_ptr(), _lib.ptr(x0), _lib.ptr(eps_out), B, per_sample, dyn, mul))

    @torch.no_grad()
    def p_sample_ddim(self, x, t, index, condition_kwargs=None, sampling_kwargs=None, denoise_sample_fn=None,
                      denoise_sample_fn_kwargs=None, repeat_noise=False, noise=None):
        """One DDIM step, the reference's per-step entry point (ddim_plms_sampler.py:345-391) -> (x_prev, pred_x0,
        None): guided eps + the fused Eq.12 update.  make_schedule() must have run; `t` is batch-uniform and
        `index` its position in ddim_timesteps.  `noise` is an optional host-supplied draw (default: torch.randn on
        the device, drawn even when sigma = 0 like the reference).  The reference's third return value
        (pred_x0_unclipped) is not materialised."""
        check_supported(sampling_kwargs)
        if repeat_noise:
            raise NotImplementedError("repeat_noise")
        device = x.device
        with torch.cuda.device(device):
            x = x.detach().float().contiguous()
            eps = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device, fresh_weights=False)(
                x, t.to(device=device, dtype=torch.long).contiguous())
            nz = torch.randn(x.shape, device=device) if noise is None else noise.to(device, torch.float32).contiguous()
            self._extras = StepExtras(sampling_kwargs, x)
            out, x0 = torch.empty_like(x), torch.empty_like(x)
            self._step(_lib.current_stream(device), eps, int(index), 1 if sampling_kwargs["clip_denoised"] else 0,
                       sampling_kwargs["temperature"], x, nz, out, x0, x.shape[0], x.shape[1:].numel())
        return out, x0, None

    @torch.no_grad()
    def ddim_sampling(self, shape, sampling_kwargs, denoise_sample_fn=None, denoise_sample_fn_kwargs=None,
                      condition_kwargs=None, noise_tape=None, device_intermediates=False, **kwargs):
        """ddim_sampling (ddim_plms_sampler.py:302-343).  The reference moves the logged intermediates to the CPU
        inside the loop (:331-335); here they stay on the device until the end and are then copied once
        (`device_intermediates=True`: not at all — LatentDiffusion.p_sample_loop converts pred_x0 to uint8 on the
        device first and moves both through pinned memory)."""
        device, noise, img = self._setup(shape, sampling_kwargs, noise_tape)
        B = shape[0]
        stream = _lib.current_stream(device)
        timesteps = self.ddim_timesteps
        total = timesteps.shape[0]
        logs = log_indices(total, sampling_kwargs["log_num_per_prog"])
        eps_src = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device)
        clip = 1 if sampling_kwargs["clip_denoised"] else 0
        per_sample = img.shape[1:].numel()
        nxt = torch.empty_like(img)
        out = dict(pred_x0=[], x_inter=[])
        for i, step in enumerate(np.flip(timesteps)):
            index = total - i - 1
            ts = torch.full((B,), int(step), device=device, dtype=torch.long)
            eps = eps_src(img, ts)
            nz = noise.next()
            x0 = torch.empty_like(img) if index in logs else None
            self._step(stream, eps, index, clip, sampling_kwargs["temperature"], img, nz, nxt, x0, B, per_sample)
            img, nxt = nxt, img
            if index in logs:
                out["x_inter"].append(img.clone().unsqueeze(0))
                out["pred_x0"].append(x0.unsqueeze(0))
        out["x_inter"] = torch.cat(out["x_inter"], 0)
        out["pred_x0"] = torch.cat(out["pred_x0"], 0)
        if not device_intermediates:
            out = {k: v.cpu() for k, v in out.items()}
        return img, out

    @torch.no_grad()
    def plms_sampling(self, shape, sampling_kwargs, denoise_sample_fn=None, denoise_sample_fn_kwargs=None,
                      condition_kwargs=None, noise_tape=None, **kwargs):
        """plms_sampling (ddim_plms_sampler.py:393-480): Adams-Bashforth on the GUIDED eps."""
        device, noise, img = self._setup(shape, sampling_kwargs, noise_tape)
        B = shape[0]
        lib, stream = _lib.lib(), _lib.current_stream(device)
        timesteps = self.ddim_timesteps
        total = timesteps.shape[0]
        time_range = np.flip(timesteps)
        logs = log_indices(total, sampling_kwargs["log_num_per_prog"])
        eps_src = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device)
        clip = 1 if sampling_kwargs["clip_denoised"] else 0
        per_sample, n = img.shape[1:].numel(), img.numel()
        temperature = sampling_kwargs["temperature"]
        old = []
        out = dict(pred_x0=[], x_inter=[])

        def guided(x, ts):
            """materialised guided eps (the history needs it): mix kernel only"""
            pc, pu, w, w_ptr, st = eps_src(x, ts)
            e = torch.empty_like(x)
            _lib.check(lib.sgdm_mix(stream, pc, pu, w, w_ptr, st, e.data_ptr(), B, per_sample))
            return e

        def lincomb(terms, coefs, div):
            o = torch.empty_like(img)
            ptrs = (C.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
            cf = (C.c_float * len(terms))(*coefs)
            _lib.check(lib.sgdm_lincomb(stream, len(terms), ptrs, cf, float(div), o.data_ptr(), n))
            return o

        for i, step in enumerate(time_range):
            index = total - i - 1
            ts = torch.full((B,), int(step), device=device, dtype=torch.long)
            ts_next = torch.full((B,), int(time_range[min(i + 1, len(time_range) - 1)]), device=device, dtype=torch.long)
            e_t = guided(img, ts)
            if len(old) == 0:
                x_prev = torch.empty_like(img)
                self._step(stream, (e_t.data_ptr(), None, 0.0, None, 0), index, clip, temperature, img, noise.next(),
                           x_prev, None, B, per_sample)
                e_t_next = guided(x_prev, ts_next)
                e_p = lincomb([e_t, e_t_next], [1.0, 1.0], 2.0)
            elif len(old) == 1:
                e_p = lincomb([e_t, old[-1]], [3.0, -1.0], 2.0)
            elif len(old) == 2:
                e_p = lincomb([e_t, old[-1], old[-2]], [23.0, -16.0, 5.0], 12.0)
            else:
                e_p = lincomb([e_t, old[-1], old[-2], old[-3]], [55.0, -59.0, 37.0, -9.0], 24.0)
            nxt = torch.empty_like(img)
            x0 = torch.empty_like(img) if index in logs else None
            self._step(stream, (e_p.data_ptr(), None, 0.0, None, 0), index, clip, temperature, img, noise.next(), nxt,
                       x0, B, per_sample)
            img = nxt
            old.append(e_t)
            if len(old) >= 4:
                old.pop(0)
            if index in logs:
                out["pred_x0"].append(x0.unsqueeze(0))
                out["x_inter"].append(img.clone().unsqueeze(0))
        out["pred_x0"] = torch.cat(out["pred_x0"], 0)
        out["x_inter"] = torch.cat(out["x_inter"], 0)
        return img, out
