#!/usr/bin/env python
"""ms per guided step (CFG UNet eps + fused DDPM update) for every BASELINE.json config, one GPU, CUDA events.
Not the bench line (bench.py measures configs[1]); this table documents the other configs in DESIGN.md.
   python tools/bench_configs.py [--steps 10]          (needs a B200)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import bench  # noqa: E402
from sgdm_b200 import _lib, synthetic  # noqa: E402
from sgdm_b200.diffusion.ddpm import LatentDiffusion  # noqa: E402
from sgdm_b200.diffusion.sampler._common import GuidedEps, coef6  # noqa: E402
from test_host_mirror import build_model  # noqa: E402

BASE = dict(in_channels=3, out_channels=3, num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4],
            num_heads=8, scale_type="imagen")
# (name, cfg, batch, GFLOP per guided sample-step: SURVEY.md 8d)
CONFIGS = [
    ("cfg1 CIFAR-10 32x32 unet_fast mc=64 label", dict(BASE, kind="unet_fast", image_size=32, model_channels=64,
     resblock_updown=True, cond_dim=10, condition_method="label", layout_dim=0, context_dim=None, cond_token_num=0), 16, 9.874),
    ("cfg2 ImageNet-64 unet_fast mc=128 label 1000", dict(BASE, kind="unet_fast", image_size=64, model_channels=128,
     resblock_updown=True, cond_dim=1000, condition_method="label", layout_dim=0, context_dim=None, cond_token_num=0), 256, 158.534),
    ("cfg3 ImageNet-64 unet_fast mc=128 cluster 5000", dict(BASE, kind="unet_fast", image_size=64, model_channels=128,
     resblock_updown=True, cond_dim=5000, condition_method="cluster", layout_dim=0, context_dim=None, cond_token_num=0), 256, 158.538),
    ("cfg4 VOC-64 unetca_fast clusterlayout ctx=32", dict(BASE, kind="unetca_fast", image_size=64, model_channels=128,
     resblock_updown=False, cond_dim=100, condition_method="clusterlayout", layout_dim=1, context_dim=32, cond_token_num=1), 256, 135.290),
    ("cfg5 COCO-Stuff-64 unetca_fast stegoclusterlayout", dict(BASE, kind="unetca_fast", image_size=64, model_channels=128,
     resblock_updown=False, cond_dim=27, condition_method="stegoclusterlayout", layout_dim=27, context_dim=32, cond_token_num=1), 128, 135.780),
]


def run(name, cfg, B, gflop, steps, families=False):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = build_model(cfg)
    bench.build_reference_init_state(m)
    m = m.to(dev).eval()
    T, H = 250, cfg["image_size"]
    ld = LatentDiffusion(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                         v_posterior=0.0, parameterization="eps", device=str(dev), num_timesteps=T, loss_type="l2")
    ld.set_denoise_fn(m.forward, m.forward_with_cond_scale)
    data = synthetic.synthetic_batch(cfg["condition_method"], B, cfg["cond_dim"], H, cfg["layout_dim"], seed=4321)
    if cfg["condition_method"] == "clusterlayout":
        kw = dict(cond=data["cluster"].float().to(dev), layout=data["lostbboxmask"].float().to(dev))
    elif cfg["condition_method"] == "stegoclusterlayout":
        kw = dict(cond=data["stego_attr"].float().to(dev), layout=data["stegomask"].float().to(dev))
    else:
        kw = dict(cond=data[cfg["condition_method"]].to(dev))
    kw["cond_scale"] = 2.0
    tape = synthetic.noise_tape((B, 3, H, H), 1, seed=1234)
    x, noise = tape["x_T"].to(dev), tape["noise"][0].to(dev)
    nxt = torch.empty_like(x)
    sampler = ld.sampler
    tabs = {k: getattr(sampler, k).detach().cpu() for k in
            ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
             "posterior_log_variance_clipped")}
    sigma = (0.5 * tabs["posterior_log_variance_clipped"]).exp()
    eps_src = GuidedEps(ld.denoise_sample_fn, kw, dev)
    lib, stream = _lib.lib(), torch.cuda.current_stream()

    def step(xc, xn, i):
        ts = torch.full((B,), i, device=dev, dtype=torch.long)
        pc, pu, w, w_ptr, st = eps_src(xc, ts)
        c = coef6(tabs["sqrt_recip_alphas_cumprod"][i], tabs["sqrt_recipm1_alphas_cumprod"][i],
                  tabs["posterior_mean_coef1"][i], tabs["posterior_mean_coef2"][i], sigma[i] if i else 0.0, 1.0)
        _lib.check(lib.sgdm_ddpm_step(stream.cuda_stream, pc, pu, w, w_ptr, st, c, 1, xc.data_ptr(), noise.data_ptr(),
                                      xn.data_ptr(), None, B, x[0].numel()))

    for _ in range(3):
        step(x, nxt, T - 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(x, nxt, T - 2 - k)
        x, nxt = nxt, x
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    row = dict(config=name, batch=B, ms_per_step=round(ms, 3), samples_per_s_250_steps=round(B / (250 * ms / 1e3), 2),
               tflops=round(B * gflop / ms, 1))
    if families:
        # one profiled replay (CUDA events around every launch): per-family split and the slowest launches
        import ctypes as C
        _lib.check(lib.sgdm_set_profiling(m._h, 1))
        step(x, nxt, T - 2 - steps)
        torch.cuda.synchronize()
        _lib.check(lib.sgdm_set_profiling(m._h, 0))
        kind, pms, fl, by = C.c_char_p(), C.c_double(), C.c_double(), C.c_double()
        fam, ops = {}, []
        for j in range(lib.sgdm_profile_count(m._h)):
            _lib.check(lib.sgdm_profile_get(m._h, j, C.byref(kind), C.byref(pms), C.byref(fl), C.byref(by)))
            f = fam.setdefault(kind.value.decode(), [0.0, 0, 0.0])
            f[0] += pms.value; f[1] += 1; f[2] += fl.value
            ops.append((round(pms.value, 3), j, kind.value.decode(), round(fl.value / 1e9), round(by.value / 1e6)))
        row["families_ms"] = {k: [round(v[0], 3), v[1], round(v[2] / (v[0] * 1e-3) / 1e12) if v[2] else None] for k, v in fam.items()}
        row["slowest"] = sorted(ops, reverse=True)[:12]
    print(json.dumps(row), flush=True)
    del m, ld, eps_src
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--families", action="store_true", help="add the per-kernel-family split of one profiled step")
    a = ap.parse_args()
    for name, cfg, B, gflop in CONFIGS:
        if a.only and a.only not in name:
            continue
        run(name, cfg, B, gflop, a.steps, a.families)
