"""How well-conditioned is a golden trajectory?  Re-runs it with the fp32 ORACLE (same arithmetic as the
reference) after perturbing x_T by ONE fp32 ulp per element (torch.nextafter), and prints the PSNR / x_inter
rel-L2 against the reference's own fp32 golden.  A trajectory whose 1-ulp-perturbed fp32 re-run already misses
40 dB cannot be tracked to 40 dB by any implementation whose arithmetic is not bit-identical to the reference's
(a different cuDNN algorithm or summation order is a perturbation of that size at every layer of every step).

    python tools/trajectory_sensitivity.py traj_cfg2 ddim250_eta0 [ulps]
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


def main():
    tname, run = sys.argv[1], sys.argv[2]
    ulps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
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
    x = tape["x_T"]
    for _ in range(ulps):
        x = torch.nextafter(x, torch.full_like(x, float("inf")))
    tape["x_T"] = x
    torch.set_num_threads(int(os.environ.get("THREADS", "8")))
    eps_fn = lambda xx, t: ounet.forward_with_cond_scale(sd, cfg, xx, t, meta["cond_scale"], **kw)
    with torch.no_grad():
        u8, inter, _ = osamp.p_sample_loop(method, eps_fn, tape, dict(num_timesteps=T), skw)
    ref_xi = torch.from_numpy(g[f"{run}_x_inter"])
    per = [rel_l2(inter["x_inter"][k], ref_xi[k]) for k in range(ref_xi.shape[0])]
    print(f"{tname}/{run}: fp32 oracle with x_T moved by {ulps} ulp -> PSNR {psnr_u8(u8, torch.from_numpy(g[f'{run}_samples'])):.2f} dB, "
          "x_inter rel-L2 " + " ".join(f"{e:.1e}" for e in per), flush=True)


if __name__ == "__main__":
    main()
