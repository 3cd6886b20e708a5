"""CPU tests of the host-side mirror and of the C-ABI surface (no compute calls)."""
import ctypes
import os
import re
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from common import UNET_CASES, kwargs_from_arrays, load_npz, load_unet_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "self-guided-diffusion-models_b200")

CONDITION = NS(scale_type="imagen", clusterlayout=NS(layout_dim=1, how="lost"),
               stegoclusterlayout=NS(layout_dim=27), layout=NS(layout_dim=21))


def build_model(cfg, precision=None):
    from sgdm_b200.dynamic.diffusionmodules import openaimodel, openaimodel_ca

    common = dict(image_size=cfg["image_size"], in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
                  model_channels=cfg["model_channels"], attention_resolutions=cfg["attention_resolutions"],
                  num_res_blocks=cfg["num_res_blocks"], channel_mult=cfg["channel_mult"], num_heads=cfg["num_heads"],
                  use_scale_shift_norm=True, use_checkpoint=False, use_fp16=False, cond_dim=cfg["cond_dim"],
                  condition_method=cfg["condition_method"], condition=CONDITION, precision=precision)
    if cfg["kind"] == "unet_fast":  # kwargs of config/dynamic/unet_fast.yaml
        return openaimodel.UNetModel(dropout=0.1, resblock_updown=True, **common)
    return openaimodel_ca.UNetModel(dropout=0.0, use_ca_block=True, transformer_depth=1, legacy=False,
                                    cond_token_num=cfg["cond_token_num"], context_dim=cfg["context_dim"],
                                    use_cls_token_as_pooled=cfg.get("use_cls_token_as_pooled", True), **common)


def test_library_exports_every_declared_symbol():
    from sgdm_b200 import _lib

    header = open(os.path.join(ROOT, "include", "sgdm_b200.h")).read()
    declared = set(re.findall(r"\b(sgdm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in include/sgdm_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert _lib.lib().sgdm_operand_dtype() in (b"f16", b"bf16")


@pytest.mark.parametrize("name", UNET_CASES)
def test_state_dict_matches_reference_inventory(name):
    meta, _ = load_unet_case(name)
    m = build_model(meta["cfg"])
    ref = {n: tuple(s) for n, s in meta["named_shapes"]}
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == ref
    # a reference checkpoint loads key for key
    from sgdm_b200 import synthetic

    sd = synthetic.synthetic_state_dict(meta["named_shapes"], seed=3)
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    # LayerNorm.beta of Attention_LR is a buffer, null embeddings are frozen parameters
    params = dict(m.named_parameters())
    assert all(not k.endswith(".beta") for k in params)
    for k, p in params.items():
        assert p.requires_grad == (("null_cond_emb" not in k) and ("null_layout_emb" not in k)), k


def test_fresh_module_has_reference_zero_init():
    meta, _ = load_unet_case("unet_fast_label_tiny")
    m = build_model(meta["cfg"])
    sd = m.state_dict()
    for k, v in sd.items():
        zero = ".out_layers.3." in k or ".proj_out." in k or k.startswith("out.2.") or "null_" in k
        if zero:
            assert v.abs().max() == 0, k


def test_compute_without_cuda_fails_loudly():
    from sgdm_b200 import _lib

    meta, a = load_unet_case("unet_fast_label_tiny")
    m = build_model(meta["cfg"])
    if torch.cuda.is_available():
        pytest.skip("needs a CUDA-less host")
    with pytest.raises(_lib.SgdmError):
        m.forward_with_cond_scale(a["x"], a["t"], 2.0, cond=a["kw_cond"])
    from sgdm_b200.diffusion.ddpm import LatentDiffusion

    ld = LatentDiffusion(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                         v_posterior=0.0, parameterization="eps", device="cpu", num_timesteps=10, loss_type="l2")
    ld.set_denoise_fn(m.forward, m.forward_with_cond_scale)
    skw = dict(sampling_method="native", num_timesteps=10, ddim_eta=0.0, log_num_per_prog=10, clip_denoised=True,
               dtp=1, temperature=1.0, noise_dropout=0)
    with pytest.raises(_lib.SgdmError):
        ld.p_sample_loop("native", (2, 3, 16, 16), skw, denoise_sample_fn_kwargs=dict(cond=a["kw_cond"], cond_scale=2.0))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "/root/reference" not in src, os.path.join(dirpath, f)


def test_schedule_mirror_bit_exact():
    from sgdm_b200.diffusion.ddpm import LatentDiffusion

    _, g = load_npz("schedules.npz")
    for T in (10, 250, 1000):
        ld = LatentDiffusion(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                             cosine_s=8e-3, v_posterior=0.0, parameterization="eps", device="cpu", num_timesteps=T,
                             loss_type="l2")
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                  "posterior_mean_coef2"):
            mine = getattr(ld.sampler, k).numpy()
            assert np.array_equal(mine.view(np.uint32), g[f"ddpm{T}_{k}"].view(np.uint32)), (T, k)
        if T == 1000:
            s = ld.sampler
            for S, eta in ((10, 0.0), (50, 0.0), (250, 0.0), (250, 1.0), (10, 1.0)):
                d = ld.sampler_list["ddim"]
                d.make_schedule(dict(num_timesteps=S, ddim_eta=eta, alphas_cumprod=s.alphas_cumprod))
                tag = f"ddim1000_{S}_{eta}"
                assert np.array_equal(d.ddim_timesteps, g[tag + "_timesteps"])
                c = d._coefs
                ref_at = torch.from_numpy(g[tag + "_ddim_alphas"])
                ref_ap = torch.from_numpy(g[tag + "_ddim_alphas_prev"])
                ref_sg = torch.from_numpy(g[tag + "_ddim_sigmas"])
                ref_s1 = torch.from_numpy(g[tag + "_ddim_sqrt_one_minus_alphas"])
                eq = lambda a, b: np.array_equal(a.numpy().view(np.uint32), b.numpy().view(np.uint32))
                assert eq(c["s1m"], ref_s1) and eq(c["sigma"], ref_sg)
                assert eq(c["sqrt_at"], ref_at.sqrt()) and eq(c["sqrt_a_prev"], ref_ap.sqrt())
                assert eq(c["dir"], (1.0 - ref_ap - ref_sg**2).sqrt())


def test_native_sampler_requires_matching_T():
    from sgdm_b200.diffusion.sampler.ddpm_sampler import Schedule_DDPM

    s = Schedule_DDPM(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                      v_posterior=0.0, parameterization="eps", device="cpu", num_timesteps=1000, loss_type="l2")
    with pytest.raises(NotImplementedError):  # ddpm_sampler.py:37-38
        s.register_schedule(timesteps=250)


def test_condition_mirror_bit_exact():
    from sgdm_b200.dynamic_input.condition import prepare_denoise_fn_kwargs_4sampling

    for name in UNET_CASES[:4]:
        meta, a = load_unet_case(name)
        cfg = meta["cfg"]
        pl = NS(hparams=NS(cond_dim=cfg["cond_dim"], condition_method=cfg["condition_method"], cond_drop_prob=0.1,
                           condition=CONDITION), training=False, device=torch.device("cpu"))
        batch = {k[5:]: v for k, v in a.items() if k.startswith("data_")}
        kw = prepare_denoise_fn_kwargs_4sampling(pl, batch, dict(random_sample_condition=False), cond_scale=2.0)
        ref = kwargs_from_arrays(a)
        assert kw.pop("cond_scale") == 2.0 and set(kw) == set(ref)
        for k in ref:
            assert kw[k].dtype == ref[k].dtype and torch.equal(kw[k], ref[k])


def test_bench_reference_arm_never_maps_the_product_library():
    """`bench.py --impl reference` is the CPU arm: the oracle port only, weights from the oracle's own inventory.
    The product's CUDA library must not even be mapped into that process (the driver records loaded .so files)."""
    import json
    import subprocess
    import sys

    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', '1', '--steps', '2', '--warmup', '0'];\n"
            "import bench; bench.main();\n"
            "print('MAPPED', 'libsgdm_b200' in open('/proc/self/maps').read())")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "MAPPED False", lines[-1]
    line = json.loads(lines[-2])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["config_id"] == 1


def test_cached_tensor_list_follows_parameter_replacement():
    """EngineUNet caches the (name -> tensor) list the weight sync walks; replacing a Parameter OBJECT (not just its
    data) or converting the module must drop that cache."""
    import torch

    meta, _ = load_unet_case("unet_fast_label_tiny")
    m = build_model(meta["cfg"])
    t0 = m._tensors()
    assert m._tensors() is t0  # cached
    name = "input_blocks.1.0.in_layers.2.weight"
    node = m.get_submodule("input_blocks.1.0.in_layers.2")
    new = torch.nn.Parameter(torch.zeros_like(node.weight))
    node.weight = new
    assert m._tensors()[name] is new
    m.null_cond_emb = torch.nn.Parameter(torch.ones_like(m.null_cond_emb), requires_grad=False)
    assert m._tensors()["null_cond_emb"] is m.null_cond_emb
    t1 = m._tensors()
    m.double()
    assert m._tensors() is not t1 and m._tensors()[name].dtype == torch.float64
