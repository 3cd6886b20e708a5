"""Drop-in for diffusion/sampler/pndm_sampler.py `PNDM_Sampler` (F-PNDM: 4 Runge-Kutta warm-up stages x 3 + linear
multistep, pndm_sampler.py:97-141,166-211).

The schedule tables reproduce the reference's float32 numpy chain (`np.linspace` betas — NOT the sqrt-linear DDPM
schedule — float32 cumprod, a trailing 0.0, pndm_sampler.py:30-47); every combination of residuals and the transfer
x_next = x + (a_next - a_t) (A x - B e_t) run as fused CUDA kernels with the reference's fp32 operation order
(sgdm_lincomb / sgdm_lincomb_scaled / sgdm_pndm_transfer), fed by the guided eps of the batched cond||uncond pass.
"""
import ctypes as C

import numpy as np
import torch

from ... import _lib
from ._common import GuidedEps, NoiseSource, check_supported


class PNDMScheduler:
    """Host-side tables and step bookkeeping of the reference's PNDMScheduler (tensor_format='pt')."""

    def __init__(self, timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear"):
        if beta_schedule != "linear":
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.timesteps = timesteps
        self.betas = np.linspace(beta_start, beta_end, timesteps, dtype=np.float32)
        self.alphas = 1.0 - self.betas
        cum = np.cumprod(self.alphas, axis=0)
        self.alphas_cumprod = torch.from_numpy(np.array(list(cum) + [0.0], dtype=np.float32))  # :44-45
        self.pndm_order = 4

    def get_warmup_time_steps(self, n):
        step = self.timesteps // n
        inference_step_times = list(range(0, self.timesteps, step))
        w = np.array(inference_step_times[-self.pndm_order:]).repeat(2) + np.tile(np.array([0, step // 2]), self.pndm_order)
        return list(reversed(w[:-1].repeat(2)[1:-1]))

    def get_time_steps(self, n):
        inference_step_times = list(range(0, self.timesteps, self.timesteps // n))
        return list(reversed(inference_step_times[:-3]))

    def transfer_scalars(self, t, t_next):
        """(d, A, B) of x_next = x + d * (A x - B e_t), each an fp32 value computed with the reference's fp32 tensor
        expression order (pndm_sampler.py:131-138)."""
        ac = self.alphas_cumprod
        at, at_next = ac[t + 1], ac[t_next + 1]
        d = at_next - at
        A = 1 / (at.sqrt() * (at.sqrt() + at_next.sqrt()))
        B = 1 / (at.sqrt() * (((1 - at_next) * at).sqrt() + ((1 - at) * at_next).sqrt()))
        return float(d), float(A), float(B)


class PNDM_Sampler(object):
    def __init__(self, ddpm_num_timesteps, beta_start, beta_end, beta_schedule="linear", tensor_format="pt", device="cuda"):
        super().__init__()
        self.ddpm_num_timesteps = ddpm_num_timesteps
        self.device = device
        self.beta_start = beta_start
        self.beta_end = beta_end
        self.beta_schedule = beta_schedule
        self.tensor_format = tensor_format

    @torch.no_grad()
    def sample(self, shape, sampling_kwargs, log_num_per_prog=100, **kwargs):
        check_supported(sampling_kwargs)
        self.num_inference_steps = sampling_kwargs["num_timesteps"]
        # (the reference builds the scheduler with its default 'linear' schedule whatever the DDPM's is, :177-178)
        self.noise_scheduler = PNDMScheduler(timesteps=self.ddpm_num_timesteps, beta_start=self.beta_start,
                                             beta_end=self.beta_end)
        device = torch.device(self.device)
        if device.type != "cuda":
            raise _lib.SgdmError(f"sampler device is {device}: sgdm_b200 has no CPU path")
        with torch.cuda.device(device):
            return self.pndm_sampling(shape, sampling_kwargs=sampling_kwargs, **kwargs)

    @torch.no_grad()
    def pndm_sampling(self, shape, denoise_sample_fn=None, denoise_sample_fn_kwargs=None, condition_kwargs=None,
                      sampling_kwargs=None, noise_tape=None, **kwargs):
        device = torch.device(self.device)
        lib, stream = _lib.lib(), _lib.current_stream(device)
        sch, S = self.noise_scheduler, self.num_inference_steps
        B = shape[0]
        image = NoiseSource(shape, device, noise_tape).x_T().contiguous()
        n, per_sample = image.numel(), image.shape[1:].numel()
        eps_src = GuidedEps(denoise_sample_fn, denoise_sample_fn_kwargs, device)

        def guided(x, t_value):
            """materialised guided eps ("residual"): the history and the Runge-Kutta sums need it"""
            ts = torch.full((B,), int(t_value), device=device, dtype=torch.long)
            pc, pu, w, w_ptr, st = eps_src(x, ts)
            e = torch.empty_like(x)
            _lib.check(lib.sgdm_mix(stream, pc, pu, w, w_ptr, st, e.data_ptr(), B, per_sample))
            return e

        def lincomb(terms, coefs, scale=None):
            o = torch.empty_like(image)
            ptrs = (C.c_void_p * len(terms))(*[t.data_ptr() for t in terms])
            cf = (C.c_float * len(terms))(*coefs)
            if scale is None:
                _lib.check(lib.sgdm_lincomb(stream, len(terms), ptrs, cf, 1.0, o.data_ptr(), n))
            else:
                _lib.check(lib.sgdm_lincomb_scaled(stream, len(terms), ptrs, cf, float(scale), o.data_ptr(), n))
            return o

        def transfer(x, t, t_next, et):
            d, A, Bc = sch.transfer_scalars(int(t), int(t_next))
            o = torch.empty_like(x)
            _lib.check(lib.sgdm_pndm_transfer(stream, x.data_ptr(), et.data_ptr(), d, A, Bc, o.data_ptr(), n))
            return o

        c6, c3 = 1 / 6, 1 / 3  # python floats: torch rounds them to fp32 when multiplying an fp32 tensor
        # ---- Runge-Kutta warm-up (step_prk, :97-116)
        warm = sch.get_warmup_time_steps(S)
        cur_residual, cur_image, ets = None, None, []
        for t in range(len(warm)):
            residual = guided(image, warm[t])
            t_prev = warm[t // 4 * 4]
            t_next = warm[min(t + 1, len(warm) - 1)]
            if t % 4 == 0:
                cur_residual = lincomb([residual], [c6])  # 0 + 1/6 r
                ets.append(residual)
                cur_image = image
            elif (t - 1) % 4 == 0 or (t - 2) % 4 == 0:
                cur_residual = lincomb([cur_residual, residual], [1.0, c3])
            else:
                residual = lincomb([cur_residual, residual], [1.0, c6])
                cur_residual = None
            image = transfer(cur_image, t_prev, t_next, residual)
        # ---- linear multistep (step_plms, :118-127)
        steps = sch.get_time_steps(S)
        for t in range(len(steps)):
            t_prev = steps[t]
            t_next = steps[min(t + 1, len(steps) - 1)]
            ets.append(guided(image, steps[t]))
            residual = lincomb([ets[-1], ets[-2], ets[-3], ets[-4]], [55.0, -59.0, 37.0, -9.0], scale=1 / 24)
            image = transfer(image, t_prev, t_next, residual)
            ets = ets[-4:]
        return image, dict(pred_x0=image)
