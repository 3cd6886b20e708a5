"""CPU forecast of the trajectory-level error budget: runs the oracle with operand-precision emulation
(oracle/unet.py `emu`) on a named-config trajectory golden and prints PSNR / x_inter rel-L2 against the
reference's fp32 trajectory.  Which rounding point costs how much, before any kernel is changed.

    python tools/precision_forecast.py traj_cfg1 ddim10_eta0 [variant ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_npz, load_unet_case, psnr_u8, rel_l2  # noqa: E402
from oracle import sampler as osamp  # noqa: E402
from oracle import unet as ounet  # noqa: E402
from sgdm_b200 import synthetic  # noqa: E402

VARIANTS = {
    "fp32": None,
    "shipped": dict(a="f16", w="f16", h1="f16", q="f16", p="f16", o="f16"),        # what the kernels do today
    "w32": dict(a="f16", h1="f16", q="f16", p="f16", o="f16"),                     # exact weights
    "a32": dict(w="f16"),                                                 # exact activations (weights rounded only)
    "h1_32": dict(a="f16", w="f16", q="f16", p="f16", o="f16"),                    # h1 kept in fp32
    "wx2": dict(a="f16", w="f16x2", h1="f16", q="f16", p="f16", o="f16"),          # weights as hi + lo
    "ax2": dict(a="f16x2", w="f16", q="f16", p="f16", o="f16x2"),                    # conv/GEMM activations as hi + lo, h1 fp32
    "awx2": dict(a="f16x2", w="f16x2", q="f16", p="f16", o="f16x2"),
    "x3": dict(a="f16x2", w="f16x2", q="f16", p="f16", o="f16"),         # the engine's precision=1 mode
                   # 3-term split in every conv/GEMM; attention fp16
    "attn_only": dict(q="f16", p="f16", o="f16"),                                  # only the attention kernel's operands rounded
    "bf16": dict(a="bf16", w="bf16", h1="bf16", q="bf16", p="bf16", o="bf16"),
}


def main():
    tname, run = sys.argv[1], sys.argv[2]
    names = sys.argv[3:] or list(VARIANTS)
    meta, g = load_npz(f"{tname}.npz")
    umeta, _ = load_unet_case(meta["unet_case"])
    cfg = umeta["cfg"]
    sd = synthetic.synthetic_state_dict([(n, tuple(s)) for n, s in umeta["named_shapes"]], umeta["weight_seed"])
    method, T, over = meta["runs"][run]
    B, H = meta["batch"], cfg["image_size"]
    kw = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("kw_")}
    skw = dict(ddim_eta=0.0, log_num_per_prog=10, clip_denoised=True, dtp=1, temperature=1.0, noise_dropout=0)
    skw.update(over)
    S = skw["num_timesteps"]
    tape = synthetic.noise_tape((B, 3, H, H), S + 1 if method == "plms" else S, seed=meta["tape_seed"])
    ref_u8 = torch.from_numpy(g[f"{run}_samples"])
    ref_xi = torch.from_numpy(g[f"{run}_x_inter"])
    torch.set_num_threads(int(os.environ.get("THREADS", "8")))
    for name in names:
        emu = VARIANTS[name]
        eps_fn = lambda x, t: ounet.forward_with_cond_scale(sd, cfg, x, t, meta["cond_scale"], emu=emu, **kw)
        with torch.no_grad():
            u8, inter, x = osamp.p_sample_loop(method, eps_fn, tape, dict(num_timesteps=T), skw)
        per = [rel_l2(inter["x_inter"][k], ref_xi[k]) for k in range(ref_xi.shape[0])]
        print(f"{tname}/{run} {name:8s} PSNR {psnr_u8(u8, ref_u8):6.2f} dB   x_inter rel-L2 "
              + " ".join(f"{e:.1e}" for e in per), flush=True)


if __name__ == "__main__":
    main()
