#!/usr/bin/env python
"""bench.py — guided 64x64 samples/sec (CFG UNet step) on B200, ms per UNet step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

Workload (BASELINE.json configs[1]): ImageNet-64 label guidance, unet_fast model_channels=128,
cond_dim=1000, cond_scale=2, 250-step DDPM ("native"), batch 256 per GPU, random-init weights
(zero-initialised tensors re-randomised), synthetic one-hot labels, host-seeded noise.

A "step" is one pass of the hot path over one batch: the batched cond||uncond UNet eps
prediction + the fused guidance-mix / posterior update.  `value` = trajectory samples per
second = (batch over all ranks) / (250 * seconds per step); `ms_per_step` is the per-step time.

  value      inputs resident in HBM, fused sampler path (what p_sample_loop runs)
  e2e        the same step through the reference-facing API with HOST (pinned) buffers:
             H2D of x_t, t, noise (and cond), forward_with_cond_scale + p_sample, D2H of x_{t-1}
  roofline   the dominant kernel (tcgen05 implicit-GEMM conv): algorithmic FLOPs of its
             launches in one step / their CUDA-event time, vs MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the reference algorithm on the host CPU (oracle port, torch fp32, all cores),
             on a bounded sample (small batch), same metric definition
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

STEPS_PER_SAMPLE = 250
GFLOP_PER_GUIDED_SAMPLE_STEP = 158.534  # SURVEY.md §8d / BASELINE.md §2, config 2
METRIC = "guided 64x64 samples/sec (CFG UNet step)"
UNIT = "samples/s (250-step DDPM trajectories; one step = CFG UNet eps + posterior update)"

CFG = dict(kind="unet_fast", image_size=64, in_channels=3, out_channels=3, model_channels=128, num_res_blocks=2,
           channel_mult=[1, 2, 4], attention_resolutions=[4], num_heads=8, resblock_updown=True, cond_dim=1000,
           condition_method="label", layout_dim=0, context_dim=None, cond_token_num=0, scale_type="imagen")


def ncu_conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per conv_gemm launch, from the committed `ncu --set full`
    capture of one step (profiles/*ncu_conv_gemm*.csv, written by tools/ncu_summary.py); None if absent."""
    import csv
    import glob

    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_conv_gemm*.csv")))
    if not files:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        hdr = rows[0]
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot, n = 0.0, 0
        for r in rows[1:]:
            b = 0.0
            for name in ("dram_read", "dram_write"):
                j = next(i for i, h in enumerate(hdr) if h.startswith(name))
                u = hdr[j].split("[")[-1].rstrip("]")
                b += float(r[j]) * unit.get(u, 1.0)
            tot += b
            n += 1
        return (tot / n if n else None), os.path.basename(files[-1])
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=p.get("bf16_tflops_sustained", p["bf16_tflops"]), hbm=p["hbm_gbs"], src="measured (MEASURED_PEAKS.json, sustained)")
    except Exception:
        return dict(tflops=1400.0, hbm=6650.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [a.strip() for a in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # "under load": upper half of the samples (the sampler also sees the idle edges)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return dict(sm_mhz=med, sm_max_mhz=mx, samples=len(sm), reasons=sorted(reasons))


def build_reference_init_state(model, seed=0):
    """reference-style init, then every zero-initialised tensor re-randomised N(0, 0.02)."""
    import torch

    g = torch.Generator().manual_seed(seed + 1)
    torch.manual_seed(seed)
    with torch.no_grad():
        for name, p in list(model.named_parameters()):
            if p.requires_grad and p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def make_model(device):
    import torch
    from test_host_mirror import build_model

    torch.manual_seed(0)
    m = build_model(CFG)
    build_reference_init_state(m)
    return m.to(device).eval()


def cpu_reference_arm(steps, warmup, batch, threads=None):
    """The reference algorithm on the host CPU: oracle port, fp32, all host threads."""
    import torch
    from oracle import sampler as osamp  # noqa: F401  (the checker doubles as the timed CPU baseline)
    from oracle import schedule as osched
    from oracle import unet as ounet
    from sgdm_b200 import synthetic
    from test_host_mirror import build_model

    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: set the count explicitly)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    torch.set_num_threads(threads or avail)
    cores = torch.get_num_threads()
    torch.manual_seed(0)
    m = build_model(CFG)
    build_reference_init_state(m)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    tab = osched.ddpm_tables(STEPS_PER_SAMPLE)
    tape = synthetic.noise_tape((batch, 3, 64, 64), 1, seed=1234)
    cond = synthetic.synthetic_batch("label", batch, 1000, 64, seed=4321)["label"]
    x = tape["x_T"]
    kw = dict(clip_denoised=True, dtp=1)

    def step(x, i):
        t = torch.full((batch,), i, dtype=torch.long)
        eps = ounet.forward_with_cond_scale(sd, CFG, x, t, 2.0, cond=cond)
        x0 = osamp._ext(tab["sqrt_recip_alphas_cumprod"], t, x) * x - osamp._ext(tab["sqrt_recipm1_alphas_cumprod"], t, x) * eps
        x0 = osamp.clip_x0(x0, True, 1)
        mean = osamp._ext(tab["posterior_mean_coef1"], t, x) * x0 + osamp._ext(tab["posterior_mean_coef2"], t, x) * x
        return mean + (0.5 * osamp._ext(tab["posterior_log_variance_clipped"], t, x)).exp() * tape["noise"][0]

    with torch.no_grad():
        for w in range(warmup):
            x = step(x, STEPS_PER_SAMPLE - 1 - w)
        t0 = time.perf_counter()
        for k in range(steps):
            x = step(x, STEPS_PER_SAMPLE - 1 - warmup - k)
        dt = time.perf_counter() - t0
    ms = dt / steps * 1e3
    value = batch / (STEPS_PER_SAMPLE * ms / 1e3)
    return dict(value=value, ms_per_step=ms, cores=cores, batch=batch,
                sample=f"{steps} guided steps at batch {batch} (of the batch-256 workload), after {warmup} warm-up")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sgdm_b200", choices=["sgdm_b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (weak scaling)")
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=0, help="cpu_baseline steps (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ops", default="", help="write the per-launch profile of one step to this JSON file")
    ap.add_argument("--ncu", action="store_true", help="minimal run for profiling under ncu: warm-up + timed steps only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    config = dict(workload="ImageNet-64 label guidance, unet_fast mc=128, cond_dim=1000, cond_scale=2, "
                           "250-step DDPM (native), batch 256 per GPU [BASELINE.json configs[1]]",
                  per_gpu_batch=args.batch, global_batch=args.batch * world, steps_per_sample=STEPS_PER_SAMPLE,
                  image="3x64x64", parallelism=f"batch-sharded x{world}, no data-path collective, final all-gather of uint8 samples",
                  l2="per-step working set (activations, several GB at batch 256) is far larger than the 126 MB L2; no extra flush")

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_arm(args.steps, args.warmup, args.cpu_batch)
        line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32", data="synthetic", impl="reference", config=config,
                    cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                    e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    note="reference = the reference algorithm (oracle port of the pure-PyTorch path) on the host CPU; "
                         "ms_per_step is for the bounded sample batch, value is normalised per sample")
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist

    from sgdm_b200 import _lib, synthetic
    from sgdm_b200.diffusion.ddpm import LatentDiffusion

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (sgdm_b200 has no CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.lib()
    B = args.batch
    model = make_model(device)
    ld = LatentDiffusion(given_betas=None, beta_schedule="linear", linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3,
                         v_posterior=0.0, parameterization="eps", device=str(device), num_timesteps=STEPS_PER_SAMPLE,
                         loss_type="l2")
    ld.set_denoise_fn(model.forward, model.forward_with_cond_scale)
    sampler = ld.sampler
    skw = dict(sampling_method="native", num_timesteps=STEPS_PER_SAMPLE, ddim_eta=0.0, log_num_per_prog=10,
               clip_denoised=True, dtp=1, temperature=1.0, noise_dropout=0)
    # synthetic inputs: this rank's shard of the global batch
    tape = synthetic.noise_tape((B, 3, 64, 64), 1, seed=1234 + rank)
    cond_host = synthetic.synthetic_batch("label", B, 1000, 64, seed=4321 + rank)["label"]
    x = tape["x_T"].to(device)
    noise = tape["noise"][0].to(device)
    cond = cond_host.to(device)
    kw = dict(cond=cond, cond_scale=2.0)
    stream = torch.cuda.current_stream()

    from sgdm_b200.diffusion.sampler._common import GuidedEps, coef6

    eps_src = GuidedEps(ld.denoise_sample_fn, kw, device)
    tabs = {k: getattr(sampler, k).detach().cpu() for k in
            ("sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2",
             "posterior_log_variance_clipped")}
    sigma = (0.5 * tabs["posterior_log_variance_clipped"]).exp()
    nxt = torch.empty_like(x)
    per_sample = x[0].numel()

    def fused_step(xc, xn, i):
        """exactly what Schedule_DDPM.sample does per step"""
        ts = torch.full((B,), i, device=device, dtype=torch.long)
        pc, pu, w, w_ptr, st = eps_src(xc, ts)
        c = coef6(tabs["sqrt_recip_alphas_cumprod"][i], tabs["sqrt_recipm1_alphas_cumprod"][i],
                  tabs["posterior_mean_coef1"][i], tabs["posterior_mean_coef2"][i], sigma[i] if i else 0.0, 1.0)
        _lib.check(lib.sgdm_ddpm_step(stream.cuda_stream, pc, pu, w, w_ptr, st, c, 1, xc.data_ptr(), noise.data_ptr(),
                                      xn.data_ptr(), None, B, per_sample))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (builds the plan, packs weights)
    i = STEPS_PER_SAMPLE - 1
    for _ in range(args.warmup):
        fused_step(x, nxt, i)
        x, nxt = nxt, x
        i -= 1
    barrier()
    # ---- timed region: K steps + the trajectory-end uint8 conversion and all-gather
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    barrier()
    launches0 = lib.sgdm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        fused_step(x, nxt, max(i, 0))
        x, nxt = nxt, x
        i -= 1
    u8 = torch.empty(x.shape, dtype=torch.uint8, device=device)
    _lib.check(lib.sgdm_to_uint8(stream.cuda_stream, x.data_ptr(), u8.data_ptr(), x.numel()))
    if world > 1:
        gathered = torch.empty((world,) + tuple(u8.shape), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(gathered, u8)
    e1.record(stream)
    barrier()
    launches = lib.sgdm_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = B * world / (STEPS_PER_SAMPLE * ms_step / 1e3)

    if args.ncu:
        if rank == 0:
            print(json.dumps(dict(ncu_mode=True, ms_per_step_under_profiler=ms_step)), flush=True)
        return

    # ---- e2e: reference-facing API with host buffers (H2D + forward_with_cond_scale + p_sample + D2H per step)
    hx = tape["x_T"].clone().pin_memory()
    hn = tape["noise"][0].clone().pin_memory()
    hcond = cond_host.clone().pin_memory()
    hout = torch.empty_like(hx).pin_memory()
    ht = torch.empty((B,), dtype=torch.long).pin_memory()

    def e2e_step(i):
        nonlocal hx, hout
        ht.fill_(i)
        dx = hx.to(device, non_blocking=True)
        dn = hn.to(device, non_blocking=True)
        dt_ = ht.to(device, non_blocking=True)
        dc = hcond.to(device, non_blocking=True)
        out, _, _ = sampler.p_sample(dx, dt_, temperature=1.0, sampling_kwargs=skw, denoise_sample_fn=ld.denoise_sample_fn,
                                     denoise_sample_fn_kwargs=dict(cond=dc, cond_scale=2.0), noise=dn, index=i)
        hout.copy_(out, non_blocking=True)
        stream.synchronize()
        hx, hout = hout, hx  # the step's result is the next step's input: swap the pinned buffers, no host copy

    for _ in range(2):
        e2e_step(STEPS_PER_SAMPLE - 1)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    n_e2e = max(3, min(args.steps, 10))
    for k in range(n_e2e):
        e2e_step(STEPS_PER_SAMPLE - 2 - k)
    e1.record(stream)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e  # host clock: the D2H + sync are part of the step
    t = torch.tensor([e2e_ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = B * world / (STEPS_PER_SAMPLE * e2e_ms / 1e3)
    h2d = hx.numel() * 4 + hn.numel() * 4 + ht.numel() * 8 + hcond.numel() * 8
    d2h = hout.numel() * 4

    # ---- roofline of the dominant kernel: one profiled step (events around every launch)
    roof = None
    fam = {}
    if rank == 0:
        _lib.check(lib.sgdm_set_profiling(model._h, 1))
        fused_step(x, nxt, 5)
        torch.cuda.synchronize()
        _lib.check(lib.sgdm_set_profiling(model._h, 0))
        kind, ms, fl, by = C.c_char_p(), C.c_double(), C.c_double(), C.c_double()
        ops = []
        for j in range(lib.sgdm_profile_count(model._h)):
            _lib.check(lib.sgdm_profile_get(model._h, j, C.byref(kind), C.byref(ms), C.byref(fl), C.byref(by)))
            ops.append(dict(i=j, kind=kind.value.decode(), ms=round(ms.value, 4), gflop=round(fl.value / 1e9, 2),
                            mbytes=round(by.value / 1e6, 1),
                            tflops=round(fl.value / (ms.value * 1e-3) / 1e12, 1) if ms.value > 0 and fl.value else None,
                            gbs=round(by.value / (ms.value * 1e-3) / 1e9, 1) if ms.value > 0 and by.value else None))
            f = fam.setdefault(kind.value.decode(), dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
            f["ms"] += ms.value; f["flops"] += fl.value; f["bytes"] += by.value; f["launches"] += 1
        if args.dump_ops:
            with open(args.dump_ops, "w") as fh:
                json.dump(ops, fh, indent=0)
        pk = peaks()
        conv = dict(ms=0.0, flops=0.0, launches=0)
        for k_ in ("conv3x3", "gemm1x1"):
            if k_ in fam:
                for q in conv:
                    conv[q] += fam[k_][q]
        total_ms = sum(f["ms"] for f in fam.values())
        ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        traffic, traffic_src = ncu_conv_traffic()
        roof = dict(bound="tensor", kernel="conv_gemm_kernel (tcgen05 implicit GEMM: conv3x3 + 1x1/linear GEMMs)",
                    achieved=ach, peak=pk["tflops"], unit="TFLOP/s", frac=ach / pk["tflops"], traffic=traffic,
                    traffic_unit="bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one step)",
                    traffic_source=traffic_src, algorithmic_bytes_per_launch=(fam.get("conv3x3", {}).get("bytes", 0.0) + fam.get("gemm1x1", {}).get("bytes", 0.0)) / max(conv["launches"], 1),
                    peak_source=pk["src"], launches_per_step=conv["launches"],
                    avg_launch_ms=conv["ms"] / max(conv["launches"], 1),
                    flops_per_step=conv["flops"], share_of_step=conv["ms"] / total_ms if total_ms else None,
                    whole_step_tflops=GFLOP_PER_GUIDED_SAMPLE_STEP * 1e9 * B / (ms_step * 1e-3) / 1e12,
                    families={k_: dict(ms=round(v["ms"], 3), launches=v["launches"],
                                       tflops=round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 and v["flops"] else None,
                                       gbs=round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 and v["bytes"] else None)
                              for k_, v in fam.items()})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cs = args.cpu_steps or 30  # ~10-30 s of host work at batch 4
        r = cpu_reference_arm(cs, 1, args.cpu_batch)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                   ms_per_step_at_sample_batch=r["ms_per_step"])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=f"{lib.sgdm_operand_dtype().decode()} operands, f32 accumulate / residual stream / sampler state",
                    data="synthetic", config=config, clocks=clk,
                    e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=e2e_ms, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
