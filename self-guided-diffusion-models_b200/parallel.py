"""Batch-sharded multi-GPU sampling.

Every sample's reverse trajectory is independent (GroupNorm / LayerNorm / attention are
per-sample, the CFG pair stays on one rank), so the path shards by batch with NO data-path
collective: rank r of R takes a contiguous slice of the batch (its slice of x_T, of the noise
tape, of `cond` / `layout`), runs the whole trajectory locally, and the only communication is
ONE all-gather of the final uint8 samples (NCCL over NVLink on GPUs; the same code runs on
gloo for the CPU tests).  The reference has no gather at all: under DDP each rank writes its
own PNG directory (eval/test_exps/common_stuff.py:127-130).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous split of n items; the first n % world ranks hold one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_tree(obj, lo, hi, n):
    """Slice every tensor whose leading dimension is the batch (n); leave scalars alone.
    Per-sample tensor cond_scale [n,1,1,1] is sliced too."""
    if torch.is_tensor(obj):
        return obj[lo:hi] if obj.dim() > 0 and obj.shape[0] == n else obj
    if isinstance(obj, dict):
        return {k: shard_tree(v, lo, hi, n) for k, v in obj.items()}
    return obj


def shard_tape(tape, lo, hi):
    if tape is None:
        return None
    out = {"x_T": tape["x_T"][lo:hi], "noise": tape["noise"][:, lo:hi]}
    if "dropout_mul" in tape:  # the F.dropout factors of noise_dropout > 0 runs travel with the noise
        out["dropout_mul"] = tape["dropout_mul"][:, lo:hi]
    return out


def all_gather_samples(local, n_total, group=None):
    """local [b_r, ...] uint8 (ragged b_r allowed) -> [n_total, ...] on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bmax = max(shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world))
    pad = torch.zeros((bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * bmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * bmax: r * bmax + (hi - lo)])
    return torch.cat(parts, 0)


@torch.no_grad()
def sample_sharded(diffusion, sampling_method, shape, sampling_kwargs, denoise_sample_fn_kwargs=None,
                   condition_kwargs=None, noise_tape=None, group=None, gather_intermediates=False, presharded=False):
    """`LatentDiffusion.p_sample_loop` over a batch sharded across the ranks of `group`.

    shape is the GLOBAL shape [B, C, H, W]; kwargs / tape are global and sliced here — or, with
    `presharded=True`, already this rank's shard (each rank loads only its own conditions: at batch 1024 the global
    stegoclusterlayout one-hot layout alone is 450 MB).
    Returns (samples uint8 [B, C, H, W] on every rank, intermediates of the LOCAL shard, or
    gathered pred_x0 when gather_intermediates)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = shape[0]
    lo, hi = shard_bounds(n, rank, world)
    local_shape = (hi - lo,) + tuple(shape[1:])
    if presharded:
        kw, tape = dict(denoise_sample_fn_kwargs or {}), noise_tape
    else:
        kw, tape = shard_tree(denoise_sample_fn_kwargs or {}, lo, hi, n), shard_tape(noise_tape, lo, hi)
    samples, inter = diffusion.p_sample_loop(sampling_method, local_shape, sampling_kwargs,
                                             denoise_sample_fn_kwargs=kw, condition_kwargs=condition_kwargs,
                                             **({"noise_tape": tape} if tape is not None else {}))
    if world == 1:
        return samples, inter
    full = all_gather_samples(samples, n, group)
    if gather_intermediates:
        p0 = inter["pred_x0"]  # [n_log, b_r, C, H, W] uint8
        g = all_gather_samples(p0.transpose(0, 1).contiguous().to(samples.device), n, group)
        inter = dict(inter)
        inter["pred_x0"] = g.transpose(0, 1).contiguous()
    return full, inter
