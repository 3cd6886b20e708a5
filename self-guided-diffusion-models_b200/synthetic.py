"""Deterministic synthetic weights, conditions and noise tapes.

There is no network for checkpoints or datasets, so every parity test, smoke() and
bench.py run on seeded synthetic inputs of the named shapes.  Each tensor is drawn
from its own generator seeded by (seed, crc32(name)), so the values do not depend
on iteration order and can be regenerated anywhere from a (name, shape) list.
"""
import zlib

import torch
import torch.nn.functional as F

# Tensors the reference zero-initialises (zero_module: openaimodel.py:273-276,357,833-834).
# A freshly built reference UNet therefore outputs exactly 0; parity
# tests must re-randomise them or they are vacuous.
# They are matched by name below: '.out_layers.3.', '.proj_out.', top-level 'out.2.'.


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2**63 - 1))
    return g


def synthetic_tensor(name, shape, seed=0):
    g = _gen(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("null_kv"):
        return torch.randn(shape, generator=g)
    if "null_cond_emb" in name or "null_layout_emb" in name:
        # zeros in the reference (requires_grad=False); small values here so that the
        # null-substitution path is actually exercised by parity tests
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        if ".out_layers.3." in name or ".proj_out." in name or name.startswith("out.2."):
            return 0.02 * torch.randn(shape, generator=g)
        bound = fan_in**-0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    if leaf in ("gamma",) or (leaf == "weight" and len(shape) == 1):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf == "beta":
        return 0.05 * torch.randn(shape, generator=g)
    return 0.05 * torch.randn(shape, generator=g)  # biases


def synthetic_state_dict(named_shapes, seed=0):
    """named_shapes: iterable of (name, shape).  Returns {name: fp32 tensor}."""
    return {n: synthetic_tensor(n, s, seed) for n, s in named_shapes}


def noise_tape(shape, n_draws, seed=1234, noise_dropout=0.0):
    """Host-supplied randomness of one trajectory: x_T, the per-step noise and (noise_dropout > 0) the
    F.dropout factor {0, 1/(1-p)} of every step, drawn after the noise so shorter tapes are prefixes."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x_T = torch.randn(shape, generator=g)
    noise = torch.randn((n_draws, *shape), generator=g)
    tape = {"x_T": x_T, "noise": noise}
    if noise_dropout > 0.0:
        keep = torch.rand((n_draws, *shape), generator=g) >= noise_dropout
        tape["dropout_mul"] = keep.float() * torch.ones(()).div(1.0 - noise_dropout)
    return tape


def synthetic_batch(condition_method, batch, cond_dim, image_size, layout_dim=0, seed=4321, cond_token_num=1):
    """A dataset-batch dict in the reference's formats:
    label/cluster: int64 one-hot [B,cond_dim]; clusterlayout: cluster one-hot +
    'lostbboxmask' {0,1} [B,1,H,W]; stegoclusterlayout: 'stegomask' one-hot
    [B,27,H,W] + 'stego_attr' n-hot [B,27]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    H = image_size
    out = {}
    if condition_method in ("label", "cluster"):
        idx = torch.randint(0, cond_dim, (batch,), generator=g)
        out[condition_method] = F.one_hot(idx, cond_dim)
    elif condition_method == "clusterlayout":
        idx = torch.randint(0, cond_dim, (batch,), generator=g)
        out["cluster"] = F.one_hot(idx, cond_dim)
        c = torch.randint(0, H, (batch, 4), generator=g)
        x0, x1 = torch.minimum(c[:, 0], c[:, 1]), torch.maximum(c[:, 0], c[:, 1])
        y0, y1 = torch.minimum(c[:, 2], c[:, 3]), torch.maximum(c[:, 2], c[:, 3])
        ar = torch.arange(H)
        mx = (ar[None, :] >= x0[:, None]) & (ar[None, :] <= x1[:, None])
        my = (ar[None, :] >= y0[:, None]) & (ar[None, :] <= y1[:, None])
        out["lostbboxmask"] = (my[:, :, None] & mx[:, None, :]).long()[:, None]
    elif condition_method in ("stegoclusterlayout", "layout"):
        k = layout_dim
        blk = max(H // 4, 1)
        cls = torch.randint(0, k, (batch, H // blk, H // blk), generator=g)
        cls = cls.repeat_interleave(blk, 1).repeat_interleave(blk, 2)
        onehot = F.one_hot(cls, k).permute(0, 3, 1, 2).contiguous()
        out["stegomask"] = onehot
        out["stego_attr"] = (onehot.flatten(2).sum(-1) > 0).long()
    elif condition_method in ("attr", "feat", "patchfeat"):
        # float features: one vector per sample, or `cond_token_num` > 1 tokens per sample ([B, N, cond_dim], the
        # input of unetca_fast's to_cond_tokens_2d branch)
        shape = (batch, cond_dim) if cond_token_num <= 1 else (batch, cond_token_num, cond_dim)
        out[condition_method] = torch.randn(shape, generator=g)
    elif condition_method is None:
        pass
    else:
        raise ValueError(condition_method)
    return out
