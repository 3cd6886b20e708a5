"""Shared pieces of the sampler mirrors: schedule tables and the fused step driver."""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ...dynamic.diffusionmodules._unet_base import EngineUNet, _is_number


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 numpy betas, same expressions as dynamic/diffusionmodules/util.py:23-43."""
    if schedule == "linear":
        betas = torch.linspace(linear_start**0.5, linear_end**0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        timesteps = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        alphas = timesteps / (1 + cosine_s) * np.pi / 2
        alphas = torch.cos(alphas).pow(2)
        alphas = alphas / alphas[0]
        betas = 1 - alphas[1:] / alphas[:-1]
        betas = np.clip(betas, a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=False):
    """util.py:46-60."""
    if ddim_discr_method == "uniform":
        c = num_ddpm_timesteps // num_ddim_timesteps
        ddim_timesteps = np.asarray(list(range(0, num_ddpm_timesteps, c)))
    elif ddim_discr_method == "quad":
        ddim_timesteps = ((np.linspace(0, np.sqrt(num_ddpm_timesteps * 0.8), num_ddim_timesteps)) ** 2).astype(int)
    else:
        raise NotImplementedError(f'There is no ddim discretization method called "{ddim_discr_method}"')
    return ddim_timesteps + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=False):
    """util.py:63-74 — alphas: fp32 torch; alphas_prev: float64 numpy; sigmas: their mix."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def log_indices(total, log_num_per_prog):
    return torch.linspace(0, total, log_num_per_prog, dtype=torch.int).cpu().numpy().tolist()


class GuidedEps:
    """Binds (model, cond, layout, cond_scale) for one trajectory and produces, per step, the
    eps pointers the fused update kernels consume.

    If `denoise_sample_fn` is the bound `forward_with_cond_scale` of an sgdm_b200 UNet, the
    conditional and unconditional scores stay in engine memory and the guidance mix is fused
    into the sampler update kernel.  Any other callable is invoked as the reference does
    (`denoise_sample_fn(x, t, **kwargs)`) and its result is fed to the same kernel.
    """

    def __init__(self, denoise_sample_fn, kwargs, device, fresh_weights=True):
        self.fn = denoise_sample_fn
        self.kwargs = dict(kwargs or {})
        self.device = torch.device(device)
        model = getattr(denoise_sample_fn, "__sgdm_model__", None)
        if model is None:
            owner = getattr(denoise_sample_fn, "__self__", None)
            if isinstance(owner, EngineUNet) and getattr(denoise_sample_fn, "__name__", "") == "forward_with_cond_scale":
                model = owner
        extra = set(self.kwargs) - {"cond", "layout", "cond_scale"}
        self.model = model if (model is not None and not extra) else None
        self.w, self.w_ptr, self.mode = 0.0, None, "generic"
        self._w_tensor, self._w_keep = None, None
        self._keep = []
        if self.model is not None:
            m = self.model
            cs = self.kwargs.get("cond_scale")
            if _is_number(cs, m._FLOAT_SHORTCUT) and cs == 1:
                self.mode = "cond"
            elif _is_number(cs, m._FLOAT_SHORTCUT) and cs == 0:
                self.mode = "uncond"
            else:
                self.mode = "guided"
                if torch.is_tensor(cs):
                    # bound to the batch lazily in __call__ (B is known there): one entry per sample
                    self._w_tensor = cs.detach().to(device=self.device, dtype=torch.float32).reshape(-1)
                else:
                    self.w = float(cs)
            self.scale_type = m._scale_type()
            if fresh_weights:  # once per trajectory: the device-side fingerprint also catches `.data` writes (ema_scope)
                m.refresh_weights()
            else:
                m.sync_weights()
            self._prepared = None

    def _prepare(self, x, t):
        m = self.model
        if self._prepared is None:  # cond / layout are constant over the trajectory: convert once
            _, _, cond, layout = m._prep_inputs(x, t, self.kwargs.get("cond"), self.kwargs.get("layout"))
            self._prepared = (cond, layout)
        return self._prepared

    def __call__(self, x, t):
        """-> (eps_c_ptr, eps_u_ptr or None, w, w_ptr, scale_type)"""
        if self.model is None:
            eps = self.fn(x, t, **self.kwargs)
            eps = eps.detach().float().contiguous()
            self._keep = [eps]
            return eps.data_ptr(), None, 0.0, None, 0
        m = self.model
        cond, layout = self._prepare(x, t)
        if self.mode == "guided":
            if self._w_tensor is not None and self.w_ptr is None:
                # same contract as forward_with_cond_scale: a 1-element tensor broadcasts, anything else must hold
                # exactly one weight per sample (the kernel indexes w_per_sample[b] for b < B)
                wt = self._w_tensor
                if wt.numel() == 1:
                    wt = wt.expand(x.shape[0])
                if wt.numel() != x.shape[0]:
                    raise AssertionError("tensor cond_scale must have one entry per sample "
                                         f"(got {wt.numel()} for batch {x.shape[0]})")
                wt = wt.contiguous()
                self._w_keep = wt
                self.w_ptr = wt.data_ptr()
            pc, pu = m.guided_pair_ptrs(x, t, cond, layout)
            return pc, pu, self.w, self.w_ptr, self.scale_type
        B = x.shape[0]
        drop = torch.full((B,), 1 if self.mode == "uncond" else 0, dtype=torch.uint8, device=x.device)
        eps = torch.empty_like(x)
        _lib.check(_lib.lib().sgdm_forward(m._h, _lib.current_stream(x.device), x.data_ptr(), t.data_ptr(),
                                           _lib.ptr(cond), _lib.ptr(layout), drop.data_ptr(), B, eps.data_ptr()))
        self._keep = [eps, drop]
        return eps.data_ptr(), None, 0.0, None, 0


def coef6(*vals):
    return (C.c_float * 6)(*[float(v) for v in vals])


class StepExtras:
    """The optional parts of an update (sampling_kwargs `dtp` < 1, `noise_dropout` > 0) as device pointers
    for sgdm_ddpm_step_ex / sgdm_ddim_step_ex.

    dtp < 1 (clip_x0_minus_one_to_one, diffusion_utils/util.py:70-82): sgdm_dyn_threshold runs the first half of
    the same update (unclipped pred_x0 into a scratch tensor) and a per-sample quantile kernel; the update then
    clamps to [-s, s] and divides by s.  noise_dropout (ddpm_sampler.py:184-185, ddim_plms_sampler.py:388-389):
    the F.dropout factor {0, 1/(1-p)} is drawn by the host exactly where the reference draws it (after the
    step's noise) and multiplied onto the scaled noise inside the kernel."""

    def __init__(self, sampling_kwargs, like, noise_source=None, noise_dropout=None):
        self.dtp = float(sampling_kwargs.get("dtp", 1))
        p = sampling_kwargs.get("noise_dropout", 0) if noise_dropout is None else noise_dropout
        self.p = float(p)
        self.noise = noise_source
        self.scratch = torch.empty_like(like) if self.dtp < 1.0 else None
        self.s = torch.empty((like.shape[0],), device=like.device, dtype=torch.float32) if self.dtp < 1.0 else None
        self._mul = None

    def pointers(self, stream, kind, eps, coefs, x, B, per_sample):
        """-> (dyn_s pointer or None, noise_mul pointer or None) for the update that follows"""
        dyn = None
        if self.dtp < 1.0:
            pc, pu, w, w_ptr, st = eps
            _lib.check(_lib.lib().sgdm_dyn_threshold(stream, kind, pc, pu, w, w_ptr, st, coefs, x.data_ptr(), self.dtp,
                                                     self.scratch.data_ptr(), self.s.data_ptr(), B, per_sample))
            dyn = self.s.data_ptr()
        mul = None
        if self.p > 0.0:
            if self.noise is not None:
                self._mul = self.noise.dropout_mul(self.p)
            else:
                self._mul = torch.nn.functional.dropout(torch.ones_like(x), p=self.p)
            mul = self._mul.data_ptr()
        return dyn, mul


def check_supported(sampling_kwargs):
    vis = sampling_kwargs.get("vis", None)
    for flag in ("condscale", "interp", "chainvis", "scoremix_vis"):
        if vis is not None and hasattr(vis, flag) and getattr(vis, flag):
            raise NotImplementedError(f"paper-visualisation branch vis.{flag} is out of scope")


class NoiseSource:
    """Per-trajectory noise: a host-supplied tape {'x_T', 'noise'[k]} (parity runs) or the
    same torch.randn draws the reference makes on the device."""

    def __init__(self, shape, device, tape=None):
        self.shape, self.device, self.k = tuple(shape), torch.device(device), 0
        self.tape = tape
        if tape is not None:
            self._noise = tape["noise"].to(self.device, torch.float32, non_blocking=True)

    def x_T(self):
        if self.tape is not None:
            return self.tape["x_T"].to(self.device, torch.float32).contiguous().clone()
        return torch.randn(self.shape, device=self.device)

    def next(self):
        if self.tape is not None:
            n = self._noise[self.k]
            self.k += 1
            return n
        return torch.randn(self.shape, device=self.device)

    def dropout_mul(self, p):
        """The F.dropout factor {0, 1/(1-p)} belonging to the noise returned by the last next(): from the tape
        (`dropout_mul`, parity runs) or drawn here — F.dropout's draw depends on the shape only."""
        if self.tape is not None and "dropout_mul" in self.tape:
            return self.tape["dropout_mul"][self.k - 1].to(self.device, torch.float32).contiguous()
        return torch.nn.functional.dropout(torch.ones(self.shape, device=self.device), p=p)
