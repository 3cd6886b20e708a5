"""Pins the CPU oracle against outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import sampler as osamp
from oracle import schedule as osched
from oracle import unet as ounet
from oracle import condition as ocond
from sgdm_b200 import synthetic

from common import UNET_CASES, kwargs_from_arrays, load_npz, load_unet_case, rel_l2

TOL = 2e-5  # fp32 vs fp32, different op order (functional vs nn.Module): rounding only


@pytest.mark.parametrize("name", UNET_CASES)
def test_unet_eps_matches_reference(name):
    if name.startswith("cfg") and name != "cfg1_cifar_label":
        torch.set_num_threads(8)
    meta, a = load_unet_case(name)
    cfg = meta["cfg"]
    sd = synthetic.synthetic_state_dict([(n, tuple(s)) for n, s in meta["named_shapes"]], meta["weight_seed"])
    kw = kwargs_from_arrays(a)
    x, t = a["x"], a["t"]
    with torch.no_grad():
        g = ounet.forward_with_cond_scale(sd, cfg, x, t, meta["cond_scale"], **kw)
        assert rel_l2(g, a["eps_guided"]) < TOL
        if name.endswith("tiny") or name == "cfg1_cifar_label":
            c = ounet.forward_with_cond_scale(sd, cfg, x, t, 1, **kw)
            u = ounet.forward_with_cond_scale(sd, cfg, x, t, 0, **kw)
            assert rel_l2(c, a["eps_cond"]) < TOL
            assert rel_l2(u, a["eps_uncond"]) < TOL
            mask = torch.ones(x.shape[0], dtype=torch.bool)
            mask[0] = False
            m = ounet.unet_forward(sd, cfg, x, t, kw.get("cond"), kw.get("layout"), mask)
            assert rel_l2(m, a["eps_masked"]) < TOL
            w = a["w_tensor"]
            gw = ounet.forward_with_cond_scale(sd, cfg, x, t, w, **kw)
            assert rel_l2(gw, a["eps_guided_tensor_w"]) < TOL


def test_unetca_float_one_takes_doubled_path():
    # openaimodel_ca.py:882 tests isinstance(cond_scale, int): float 1.0 is NOT short-circuited
    meta, a = load_unet_case("unetca_clusterlayout_tiny")
    cfg = meta["cfg"]
    sd = synthetic.synthetic_state_dict([(n, tuple(s)) for n, s in meta["named_shapes"]], meta["weight_seed"])
    kw = kwargs_from_arrays(a)
    with torch.no_grad():
        c_int = ounet.forward_with_cond_scale(sd, cfg, a["x"], a["t"], 1, **kw)
        c_flt = ounet.forward_with_cond_scale(sd, cfg, a["x"], a["t"], 1.0, **kw)
    # (1-1.0)*eps_u + 1.0*eps_c == eps_c up to the batched-vs-single conv rounding
    assert rel_l2(c_flt, c_int) < TOL


def test_condition_lookup_bit_exact():
    for name in UNET_CASES[:4]:
        meta, a = load_unet_case(name)
        cfg = meta["cfg"]
        batch = {k[5:]: v for k, v in a.items() if k.startswith("data_")}
        kw = ocond.denoise_kwargs_for_sampling(cfg["condition_method"], batch, 2.0)
        ref = kwargs_from_arrays(a)
        assert kw.pop("cond_scale") == 2.0
        assert set(kw) == set(ref)
        for k in ref:
            assert kw[k].dtype == ref[k].dtype and torch.equal(kw[k], ref[k]), k


def test_stego_attr_nhot():
    batch = synthetic.synthetic_batch("stegoclusterlayout", 3, 27, 16, 27, seed=5)
    cls = batch["stegomask"].argmax(1)
    for b in range(3):
        want = torch.zeros(27, dtype=torch.long)
        want[torch.unique(cls[b])] = 1  # stegomask_to_attr_nhot (complex_ds_common_util.py:126-133)
        assert torch.equal(batch["stego_attr"][b], want)
        assert torch.equal(ocond.stego_attr_nhot(batch["stegomask"])[b], want)


def test_schedules_bit_exact():
    _, g = load_npz("schedules.npz")
    for T in (10, 250, 1000):
        tab = osched.ddpm_tables(T)
        for k, v in tab.items():
            ref = g[f"ddpm{T}_{k}"]
            assert v.dtype == torch.float32
            assert np.array_equal(v.numpy().view(np.uint32), ref.view(np.uint32)), (T, k)
    ac = osched.ddpm_tables(1000)["alphas_cumprod"]
    x = torch.zeros(1)
    for S, eta in ((10, 0.0), (50, 0.0), (250, 0.0), (250, 1.0), (10, 1.0)):
        d = osched.ddim_tables(ac, S, 1000, eta)
        tag = f"ddim1000_{S}_{eta}"
        assert np.array_equal(d["timesteps"], g[tag + "_timesteps"])
        for k_mine, k_ref in (("alphas", "ddim_alphas"), ("alphas_prev", "ddim_alphas_prev"),
                              ("sigmas", "ddim_sigmas"), ("sqrt_one_minus_alphas", "ddim_sqrt_one_minus_alphas")):
            mine = np.asarray([torch.full_like(x, d[k_mine][i]).item() for i in range(S)], dtype=np.float32)
            assert np.array_equal(mine.view(np.uint32), g[tag + "_" + k_ref].view(np.uint32)), (tag, k_mine)


def test_ddim_timesteps_quirks():
    # S must divide T: S=300 on T=1000 yields 334 steps ending at 1000 (out of range) [SURVEY §8a S3]
    ts = osched.ddim_timesteps(300, 1000)
    assert len(ts) == 334 and ts[-1] == 1000
    assert list(osched.ddim_timesteps(10, 1000)) == [1, 101, 201, 301, 401, 501, 601, 701, 801, 901]


@pytest.mark.parametrize("run", ["ddim10_eta0", "ddim10_eta1", "native10", "plms10", "ddim10_dtp", "native10_dtp_dropout",
                                 "ddim10_eta1_dropout"])
def test_sampling_matches_reference(run):
    meta, g = load_npz("sampling_tiny.npz")
    umeta, _ = load_unet_case(meta["unet_case"])
    cfg = umeta["cfg"]
    sd = synthetic.synthetic_state_dict([(n, tuple(s)) for n, s in umeta["named_shapes"]], umeta["weight_seed"])
    method, T, over = meta["runs"][run]
    B = meta["batch"]
    cond = torch.from_numpy(g["data_label"])
    skw = dict(ddim_eta=0.0, log_num_per_prog=10, clip_denoised=True, dtp=1, temperature=1.0, noise_dropout=0)
    skw.update(over)
    tape = synthetic.noise_tape((B, 3, cfg["image_size"], cfg["image_size"]), 11 if method == "plms" else 10,
                                seed=meta["tape_seed"], noise_dropout=skw["noise_dropout"])
    eps_fn = lambda x, t: ounet.forward_with_cond_scale(sd, cfg, x, t, meta["cond_scale"], cond=cond)
    with torch.no_grad():
        u8, inter, x = osamp.p_sample_loop(method, eps_fn, tape, dict(num_timesteps=T), skw)
    ref_u8 = torch.from_numpy(g[f"{run}_samples"])
    # identical noise, fp32 both sides: allow off-by-one on the truncating uint8 cast
    diff = (u8.int() - ref_u8.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 0.01
    assert inter["pred_x0"].shape == g[f"{run}_pred_x0"].shape
    assert rel_l2(inter["x_inter"], torch.from_numpy(g[f"{run}_x_inter"])) < 1e-4


@pytest.mark.parametrize("tname,run", [("traj_cfg1", "ddim10_eta0"), ("traj_cfg1", "native10"), ("traj_cfg1", "plms10"),
                                       ("traj_cfg1_pndm", "pndm10"),
                                       ("traj_cfg4", "ddim10_eta0"), ("traj_cfg5", "ddim10_eta0")])
def test_named_config_trajectories_match_reference(tname, run):
    """The oracle on the trajectory goldens of the NAMED BASELINE configs (config 1 exactly as stated: 32x32, B=16,
    DDIM-10; unetca_fast clusterlayout / stegoclusterlayout at true shapes).  The 250-step config-2 goldens are
    checked on the GPU only (minutes of CPU time per run)."""
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
    tape = synthetic.noise_tape((B, 3, H, H), S + 1 if method == "plms" else 0 if method == "pndm" else S, seed=meta["tape_seed"])
    eps_fn = lambda x, t: ounet.forward_with_cond_scale(sd, cfg, x, t, meta["cond_scale"], **kw)
    torch.set_num_threads(8)
    with torch.no_grad():
        u8, inter, x = osamp.p_sample_loop(method, eps_fn, tape, dict(num_timesteps=T), skw)
    ref_u8 = torch.from_numpy(g[f"{run}_samples"])
    diff = (u8.int() - ref_u8.int()).abs()
    assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 0.01
    if f"{run}_x_inter" in g:  # (PNDM returns no x_inter)
        assert rel_l2(inter["x_inter"], torch.from_numpy(g[f"{run}_x_inter"])) < 1e-4


@pytest.mark.parametrize("name", UNET_CASES)
def test_oracle_param_inventory_matches_reference(name):
    """oracle.unet.param_shapes (what bench.py's CPU arm builds its weights from) == the reference module's state_dict."""
    meta, _ = load_unet_case(name)
    ref = {n: tuple(s) for n, s in meta["named_shapes"]}
    got = dict(ounet.param_shapes(meta["cfg"]))
    assert got == ref, (sorted(set(ref) ^ set(got))[:8], [k for k in ref if k in got and got[k] != ref[k]][:8])


def test_subpixel_upconv_identity():
    """The algebra behind ConvDesc::up2 (csrc/conv.cuh), in float64 on the CPU: conv3x3(nearest_2x(x)) equals, for output
    parity (dy, dx), a 2x2 conv of x whose tap (a, b) is the sum of the 3x3 taps r in V(dy, a), s in V(dx, b) with
    V(0,0) = {0}, V(0,1) = {1,2}, V(1,0) = {0,1}, V(1,1) = {2} — the nearest-2x upsample + conv of every up ResBlock
    (openaimodel.py:214-216,304-307) and of unetca_fast's Upsample (openaimodel_ca.py:101-133) at 4/9 of the MACs."""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)
    w = torch.randn(4, 5, 3, 3, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    V = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
    H, W = x.shape[2:]
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.empty_like(ref)
    for dy in (0, 1):
        for dx in (0, 1):
            k = torch.zeros(4, 5, 2, 2, dtype=torch.float64)
            for a in (0, 1):
                for c in (0, 1):
                    k[:, :, a, c] = sum(w[:, :, r, s] for r in V[(dy, a)] for s in V[(dx, c)])
            out[:, :, dy::2, dx::2] = F.conv2d(xp[:, :, dy:dy + H + 1, dx:dx + W + 1], k, b)
    assert torch.allclose(out, ref, rtol=0, atol=1e-12)
