"""Generate golden vectors by running the UNMODIFIED reference (/root/reference).

Run in the authoring container only:   python tests/golden/make_golden.py
Writes tests/golden/*.npz (small, committed).  Weights are NOT stored: they are
regenerated from (name, shape, seed) by sgdm_b200.synthetic.synthetic_state_dict,
and the (name, shape) inventory of the reference module is stored so that the
drop-in's state_dict layout can be checked key for key.

Cases
  unet_*  : eps for guided / cond-only / uncond-only / masked forward, per config
  sched   : every schedule table for T in {10, 250, 1000} and DDIM(S, eta) variants (bit-exact targets)
  sample_*: full trajectories (DDIM eta=0, DDIM eta=1, native DDPM, PLMS) on a tiny model
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refshim  # noqa: E402

refshim.install()

from sgdm_b200 import synthetic  # noqa: E402

import diffusion.ddpm as ref_ddpm  # noqa: E402
from dynamic.diffusionmodules import openaimodel as ref_unet  # noqa: E402
from dynamic.diffusionmodules import openaimodel_ca as ref_unetca  # noqa: E402
from dynamic_input.condition import prepare_denoise_fn_kwargs_4sampling  # noqa: E402
from diffusion_utils.util import dict2obj  # noqa: E402

torch.set_num_threads(8)

CASES = {
    # name: (cfg, batch)
    "unet_fast_label_tiny": (
        dict(kind="unet_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=True, cond_dim=10, condition_method="label", layout_dim=0,
             context_dim=None, cond_token_num=0, scale_type="imagen"), 2),
    "unet_fast_clusterlayout_tiny": (
        dict(kind="unet_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=True, cond_dim=100, condition_method="clusterlayout", layout_dim=1,
             context_dim=None, cond_token_num=0, scale_type="imagen"), 2),
    "unetca_clusterlayout_tiny": (
        dict(kind="unetca_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=100, condition_method="clusterlayout", layout_dim=1,
             context_dim=32, cond_token_num=1, scale_type="imagen"), 2),
    "unetca_stego_tiny": (
        dict(kind="unetca_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=27, condition_method="stegoclusterlayout", layout_dim=27,
             context_dim=32, cond_token_num=1, scale_type="imagen"), 2),
    # layout-only guidance (cond_token_num = 0, cond_dim = 0: the reference's test_unittest.py `condition_method=layout` runs)
    "unetca_layout_tiny": (
        dict(kind="unetca_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=0, condition_method="layout", layout_dim=21,
             context_dim=32, cond_token_num=0, scale_type="imagen"), 2),
    # cond_token_num > 1: a [B, N, cond_dim] token condition through to_cond_tokens_2d (openaimodel_ca.py:988-1012), pooled by
    # the CLS token (the config's use_cls_token_as_pooled=True) or by the mean over the tokens
    "unetca_tokens4_cls_tiny": (
        dict(kind="unetca_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=24, condition_method="patchfeat", layout_dim=0,
             context_dim=32, cond_token_num=4, use_cls_token_as_pooled=True, scale_type="imagen"), 2),
    "unetca_tokens12_mean_tiny": (
        dict(kind="unetca_fast", image_size=16, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=24, condition_method="patchfeat", layout_dim=0,
             context_dim=32, cond_token_num=12, use_cls_token_as_pooled=False, scale_type="imagen"), 2),
    # the attention layout of config/dynamic/unet.yaml (attention at ds 2 and 4, 32 heads: head dims 8 and 16) through the
    # constructor the reference actually has (unet.yaml as written passes `num_classes` / `cond_mlp_divide`, which
    # UNetModel.__init__ rejects with a TypeError)
    "unet_heads32_ds24_tiny": (
        dict(kind="unet_fast", image_size=16, in_channels=3, out_channels=3, model_channels=128,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[2, 4], num_heads=32,
             resblock_updown=True, cond_dim=10, condition_method="label", layout_dim=0,
             context_dim=None, cond_token_num=0, scale_type="imagen"), 2),
    # BASELINE.json configs at their true shapes (batch kept small: CPU reference)
    "cfg1_cifar_label": (
        dict(kind="unet_fast", image_size=32, in_channels=3, out_channels=3, model_channels=64,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=True, cond_dim=10, condition_method="label", layout_dim=0,
             context_dim=None, cond_token_num=0, scale_type="imagen"), 4),
    "cfg2_in64_label": (
        dict(kind="unet_fast", image_size=64, in_channels=3, out_channels=3, model_channels=128,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=True, cond_dim=1000, condition_method="label", layout_dim=0,
             context_dim=None, cond_token_num=0, scale_type="imagen"), 2),
    "cfg4_voc_clusterlayout": (
        dict(kind="unetca_fast", image_size=64, in_channels=3, out_channels=3, model_channels=128,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=100, condition_method="clusterlayout", layout_dim=1,
             context_dim=32, cond_token_num=1, scale_type="imagen"), 2),
    "cfg5_coco_stego": (
        dict(kind="unetca_fast", image_size=64, in_channels=3, out_channels=3, model_channels=128,
             num_res_blocks=2, channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8,
             resblock_updown=False, cond_dim=27, condition_method="stegoclusterlayout", layout_dim=27,
             context_dim=32, cond_token_num=1, scale_type="imagen"), 2),
}


def build_reference_unet(cfg):
    common = dict(
        image_size=cfg["image_size"], in_channels=cfg["in_channels"], out_channels=cfg["out_channels"],
        model_channels=cfg["model_channels"], attention_resolutions=cfg["attention_resolutions"],
        num_res_blocks=cfg["num_res_blocks"], channel_mult=cfg["channel_mult"], num_heads=cfg["num_heads"],
        use_scale_shift_norm=True, use_checkpoint=False, use_fp16=False, cond_dim=cfg["cond_dim"],
        condition_method=cfg["condition_method"], condition=refshim.condition_obj(),
    )
    if cfg["kind"] == "unet_fast":  # config/dynamic/unet_fast.yaml
        m = ref_unet.UNetModel(dropout=0.1, resblock_updown=True, **common)
    else:  # config/dynamic/unetca_fast.yaml + README overrides
        m = ref_unetca.UNetModel(dropout=0.0, use_ca_block=True, transformer_depth=1, legacy=False,
                                 cond_token_num=cfg["cond_token_num"], context_dim=cfg["context_dim"],
                                 use_cls_token_as_pooled=cfg.get("use_cls_token_as_pooled", True), **common)
    return m.eval()


def load_synthetic(module, seed):
    named_shapes = [(k, tuple(v.shape)) for k, v in module.state_dict().items()]
    sd = synthetic.synthetic_state_dict(named_shapes, seed)
    missing, unexpected = module.load_state_dict(sd, strict=True)
    return named_shapes


class FakeModule:
    """The slice of TaoDiffusion that prepare_denoise_fn_kwargs_4sampling reads."""

    def __init__(self, cfg):
        self.hparams = dict2obj(dict(
            cond_dim=cfg["cond_dim"], condition_method=cfg["condition_method"], cond_drop_prob=0.1,
            condition=dict(clusterlayout=dict(how="lost"), layout=dict(how="stego"))))
        self.training = False
        self.device = torch.device("cpu")


def make_inputs(cfg, batch, seed):
    g = torch.Generator().manual_seed(seed)
    H = cfg["image_size"]
    x = torch.randn(batch, cfg["in_channels"], H, H, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    data = synthetic.synthetic_batch(cfg["condition_method"], batch, cfg["cond_dim"], H,
                                     cfg["layout_dim"], seed=seed + 1, cond_token_num=cfg.get("cond_token_num") or 1)
    return x, t, data


@torch.no_grad()
def gen_unet_case(name, cfg, batch):
    torch.manual_seed(0)
    model = build_reference_unet(cfg)
    named_shapes = load_synthetic(model, seed=7)
    x, t, data = make_inputs(cfg, batch, seed=11)
    kw = prepare_denoise_fn_kwargs_4sampling(
        FakeModule(cfg), dict(data), dict(random_sample_condition=False), cond_scale=2.0)
    cond_scale = kw.pop("cond_scale")
    out = {}
    out["eps_guided"] = model.forward_with_cond_scale(x, t, cond_scale=cond_scale, **kw)
    out["eps_cond"] = model.forward_with_cond_scale(x, t, cond_scale=1, **kw)
    out["eps_uncond"] = model.forward_with_cond_scale(x, t, cond_scale=0, **kw)
    # forward() with an explicit per-sample drop probability in {0,1}: sample 0 keeps, rest drop
    p = torch.ones(batch)
    p[0] = 0.0
    out["eps_masked"] = model.forward(x=x, timesteps=t, cond_drop_prob=p, **kw)[0]
    # per-sample tensor cond_scale [B,1,1,1] (ddim_plms_sampler.py:117-119)
    w = torch.linspace(0.5, 3.0, batch).view(batch, 1, 1, 1)
    out["eps_guided_tensor_w"] = model.forward_with_cond_scale(x, t, cond_scale=w, **kw)
    arrays = {k: v.numpy() for k, v in out.items()}
    arrays.update(x=x.numpy(), t=t.numpy(), w_tensor=w.numpy())
    for k, v in data.items():
        arrays["data_" + k] = v.numpy()
    for k, v in kw.items():
        if torch.is_tensor(v):
            arrays["kw_" + k] = v.numpy()
    arrays["meta"] = np.frombuffer(json.dumps(dict(
        cfg=cfg, batch=batch, weight_seed=7, input_seed=11, cond_scale=cond_scale,
        named_shapes=[[n, list(s)] for n, s in named_shapes])).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"unet_{name}.npz"), **arrays)
    print(name, {k: (tuple(v.shape), float(np.abs(v).max())) for k, v in arrays.items() if k.startswith("eps")})
    return model


def diffusion_kwargs(T):
    return dict(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2,
                cosine_s=8e-3, v_posterior=0.0, parameterization="eps", device="cpu",
                num_timesteps=T, loss_type="l2")


def gen_schedules():
    arrays = {}
    for T in (10, 250, 1000):
        ld = ref_ddpm.LatentDiffusion(**diffusion_kwargs(T))
        s = ld.sampler
        for k in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod",
                  "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            arrays[f"ddpm{T}_{k}"] = getattr(s, k).numpy()
        if T == 1000:
            for S, eta in ((10, 0.0), (50, 0.0), (250, 0.0), (250, 1.0), (10, 1.0)):
                d = ld.sampler_list["ddim"]
                d.make_schedule(dict(num_timesteps=S, ddim_eta=eta, alphas_cumprod=s.alphas_cumprod,
                                     betas=s.betas, alphas_cumprod_prev=s.alphas_cumprod_prev))
                tag = f"ddim1000_{S}_{eta}"
                arrays[tag + "_timesteps"] = np.asarray(d.ddim_timesteps)
                x = torch.zeros(1)
                # the values the sampler actually uses: fp32 after torch.full_like(x, v)
                for k in ("ddim_alphas", "ddim_alphas_prev", "ddim_sigmas", "ddim_sqrt_one_minus_alphas"):
                    tab = getattr(d, k)
                    arrays[tag + "_" + k] = np.asarray(
                        [torch.full_like(x, tab[i]).item() for i in range(len(d.ddim_timesteps))], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), **arrays)
    print("schedules", len(arrays))


class Tape:
    """Replaces torch.randn inside the reference samplers with a recorded tape."""

    def __init__(self, tape):
        self.items = [tape["x_T"]] + [n for n in tape["noise"]]
        self.k = 0

    def __call__(self, *shape, **kw):
        if len(shape) == 1 and not isinstance(shape[0], int):
            shape = tuple(shape[0])
        out = self.items[self.k]
        assert tuple(out.shape) == tuple(shape), (out.shape, shape)
        self.k += 1
        return out.clone()


@torch.no_grad()
def gen_sampling(model, cfg, batch):
    H = cfg["image_size"]
    shape = (batch, 3, H, H)
    data = synthetic.synthetic_batch(cfg["condition_method"], batch, cfg["cond_dim"], H, cfg["layout_dim"], seed=21)
    arrays = {"data_label": data["label"].numpy()}
    runs = {
        "ddim10_eta0": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
        "ddim10_eta1": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=1.0)),
        "native10": ("native", 10, dict(num_timesteps=10, ddim_eta=0.0)),
        "plms10": ("plms", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
        # dynamic thresholding (the reference's own torch.quantile) and noise dropout; F.dropout's
        # random factor is replaced by the tape's `dropout_mul` exactly like torch.randn by the noise
        "ddim10_dtp": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=1.0, dtp=0.9)),
        "native10_dtp_dropout": ("native", 10, dict(num_timesteps=10, dtp=0.95, noise_dropout=0.25)),
        "ddim10_eta1_dropout": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=1.0, noise_dropout=0.25)),
    }
    for rname, (method, T, over) in runs.items():
        ld = ref_ddpm.LatentDiffusion(**diffusion_kwargs(T))
        ld.set_denoise_fn(model.forward, model.forward_with_cond_scale)
        skw = dict(sampling_method=method, vis=None, ddim_eta=0.0, log_num_per_prog=10, clip_denoised=True,
                   dtp=1, temperature=1.0, noise_dropout=0, random_sample_condition=False,
                   return_inter_dict=False, disable_tqdm=True)
        skw.update(over)
        kw = prepare_denoise_fn_kwargs_4sampling(FakeModule(cfg), dict(data), skw, cond_scale=2.0)
        n_draws = 11 if method == "plms" else 10
        tape = synthetic.noise_tape(shape, n_draws, seed=1234, noise_dropout=skw["noise_dropout"])
        real_randn, real_dropout = torch.randn, torch.nn.functional.dropout
        torch.randn = Tape(tape)
        if skw["noise_dropout"] > 0:  # the factor belonging to the noise drawn last
            def taped_dropout(x, p=0.5, training=True, inplace=False):
                if not training:  # the UNet's nn.Dropout modules in eval mode
                    return x
                assert p == skw["noise_dropout"]
                return x * tape["dropout_mul"][torch.randn.k - 2]

            torch.nn.functional.dropout = taped_dropout
        try:
            samples, inter = ld.p_sample_loop(method, shape, skw, denoise_sample_fn_kwargs=kw,
                                              condition_kwargs=dict(cond_scale=2.0, condition_method="label"))
            used = torch.randn.k
        finally:
            torch.randn, torch.nn.functional.dropout = real_randn, real_dropout
        assert used == 1 + n_draws, (rname, used)
        arrays[f"{rname}_samples"] = samples.numpy()
        arrays[f"{rname}_pred_x0"] = inter["pred_x0"].numpy()
        arrays[f"{rname}_x_inter"] = inter["x_inter"].numpy()
        print(rname, samples.shape, inter["pred_x0"].shape, samples.float().mean().item())
    arrays["meta"] = np.frombuffer(json.dumps(dict(
        runs={k: [v[0], v[1], v[2]] for k, v in runs.items()}, batch=batch, tape_seed=1234, data_seed=21,
        cond_scale=2.0, unet_case="unet_fast_label_tiny")).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "sampling_tiny.npz"), **arrays)




# ---------------------------------------------------------------------------------------------------
# Trajectory goldens at the NAMED BASELINE configs (round 2):  python tests/golden/make_golden.py traj [name...]
#   traj_cfg1   : config 1 exactly as BASELINE.json states it - 32x32, mc=64, B=16, DDIM-10 eta=0 (+ native-10, PLMS-10)
#   traj_cfg2   : config 2 at B=2 - 250-step native DDPM (T=250) and DDIM-250 eta=0 (T=1000)
#   traj_cfg4/5 : one DDIM-10 trajectory each of unetca_fast clusterlayout / stegoclusterlayout at true shapes
# Every run is the unmodified reference's LatentDiffusion.p_sample_loop (diffusion/ddpm.py:108-122) with
# torch.randn replaced by the seeded tape.
TRAJ = {
    "traj_cfg1": ("cfg1_cifar_label", 16, {
        "ddim10_eta0": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
        "native10": ("native", 10, dict(num_timesteps=10)),
        "plms10": ("plms", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
    }),
    "traj_cfg2": ("cfg2_in64_label", 2, {
        "native250": ("native", 250, dict(num_timesteps=250)),
        "ddim250_eta0": ("ddim", 1000, dict(num_timesteps=250, ddim_eta=0.0)),
    }),
    # PNDM (pndm_sampler.py): 12 Runge-Kutta warm-up evaluations + 7 multistep ones for num_timesteps = 10
    "traj_cfg1_pndm": ("cfg1_cifar_label", 16, {
        "pndm10": ("pndm", 1000, dict(num_timesteps=10)),
    }),
    "traj_cfg4": ("cfg4_voc_clusterlayout", 2, {
        "ddim10_eta0": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
    }),
    "traj_cfg5": ("cfg5_coco_stego", 2, {
        "ddim10_eta0": ("ddim", 1000, dict(num_timesteps=10, ddim_eta=0.0)),
    }),
}


@torch.no_grad()
def gen_trajectories(tname):
    import time

    case, batch, runs = TRAJ[tname]
    cfg, _ = CASES[case]
    torch.manual_seed(0)
    model = build_reference_unet(cfg)
    named_shapes = load_synthetic(model, seed=7)
    H = cfg["image_size"]
    shape = (batch, 3, H, H)
    data = synthetic.synthetic_batch(cfg["condition_method"], batch, cfg["cond_dim"], H, cfg["layout_dim"], seed=21)
    arrays = {}
    for rname, (method, T, over) in runs.items():
        ld = ref_ddpm.LatentDiffusion(**diffusion_kwargs(T))
        ld.set_denoise_fn(model.forward, model.forward_with_cond_scale)
        skw = dict(sampling_method=method, vis=None, ddim_eta=0.0, log_num_per_prog=10, clip_denoised=True,
                   dtp=1, temperature=1.0, noise_dropout=0, random_sample_condition=False,
                   return_inter_dict=False, disable_tqdm=True)
        skw.update(over)
        kw = prepare_denoise_fn_kwargs_4sampling(FakeModule(cfg), dict(data), skw, cond_scale=2.0)
        S = skw["num_timesteps"]
        n_draws = S + 1 if method == "plms" else 0 if method == "pndm" else S
        tape = synthetic.noise_tape(shape, n_draws, seed=1234)
        real_randn = torch.randn
        torch.randn = Tape(tape)
        t0 = time.time()
        try:
            samples, inter = ld.p_sample_loop(method, shape, skw, denoise_sample_fn_kwargs=kw,
                                              condition_kwargs=dict(cond_scale=2.0,
                                                                    condition_method=cfg["condition_method"]))
            used = torch.randn.k
        finally:
            torch.randn = real_randn
        assert used == 1 + n_draws, (rname, used)
        arrays[f"{rname}_samples"] = samples.numpy()
        arrays[f"{rname}_pred_x0"] = inter["pred_x0"].numpy()
        if "x_inter" in inter:  # (PNDM returns dict(pred_x0=image) only)
            arrays[f"{rname}_x_inter"] = inter["x_inter"].numpy()
        for k, v in kw.items():
            if torch.is_tensor(v):
                arrays["kw_" + k] = v.numpy()
        print(tname, rname, tuple(samples.shape), tuple(inter["pred_x0"].shape), f"{time.time() - t0:.1f}s",
              "mean", samples.float().mean().item(), flush=True)
    arrays["meta"] = np.frombuffer(json.dumps(dict(
        runs={k: [v[0], v[1], v[2]] for k, v in runs.items()}, batch=batch, tape_seed=1234, data_seed=21,
        cond_scale=2.0, weight_seed=7, unet_case=case)).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"{tname}.npz"), **arrays)


if __name__ == "__main__":
    if sys.argv[1:2] == ["traj"]:
        for tname in (sys.argv[2:] or list(TRAJ)):
            gen_trajectories(tname)
        sys.exit(0)
    only = sys.argv[1:]
    gen_schedules()
    for name, (cfg, batch) in CASES.items():
        if only and name not in only:
            continue
        m = gen_unet_case(name, cfg, batch)
        if name == "unet_fast_label_tiny":
            gen_sampling(m, cfg, batch=2)
