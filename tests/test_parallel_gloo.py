"""World-size-2 gloo test of the batch-sharding host logic (no GPU): shard bounds, slicing of
cond / layout / per-sample cond_scale / noise tape, and the final all-gather with a ragged split."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeDiffusion:
    """Stands in for LatentDiffusion: the 'sample' encodes which inputs each row saw."""

    def p_sample_loop(self, method, shape, skw, denoise_sample_fn_kwargs=None, condition_kwargs=None, noise_tape=None):
        b = shape[0]
        cond = denoise_sample_fn_kwargs["cond"]
        w = denoise_sample_fn_kwargs["cond_scale"]
        assert cond.shape[0] == b and noise_tape["x_T"].shape[0] == b and noise_tape["noise"].shape[1] == b
        assert denoise_sample_fn_kwargs["layout"].shape[0] == b and w.shape[0] == b
        label = cond.argmax(1).to(torch.uint8)
        tag = (noise_tape["x_T"][:, 0, 0, 0] * 10).round().to(torch.uint8)
        out = torch.zeros(shape, dtype=torch.uint8)
        out[:, 0] = label.view(b, 1, 1)
        out[:, 1] = tag.view(b, 1, 1)
        out[:, 2] = (w.view(b) * 10).round().to(torch.uint8).view(b, 1, 1)
        return out, {"pred_x0": out.unsqueeze(0).repeat(3, 1, 1, 1, 1)}


def _worker(rank, world, port, q, B=7):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sgdm_b200 import parallel

    # B = 7: ragged, 4 + 3;  B = 1: the second rank's shard is EMPTY (fewer samples than ranks)
    cond = torch.nn.functional.one_hot(torch.arange(B) % 5, 5)
    layout = torch.arange(B).float().view(B, 1, 1, 1).expand(B, 1, 4, 4).contiguous()
    w = torch.linspace(0.5, 2.0, B).view(B, 1, 1, 1)
    tape = {"x_T": torch.arange(B).float().view(B, 1, 1, 1).expand(B, 3, 4, 4) / 10, "noise": torch.zeros(2, B, 3, 4, 4)}
    full, inter = parallel.sample_sharded(FakeDiffusion(), "ddim", (B, 3, 4, 4), {}, dict(cond=cond, layout=layout, cond_scale=w),
                                          noise_tape=tape, gather_intermediates=True)
    ok = tuple(full.shape) == (B, 3, 4, 4)
    ok &= torch.equal(full[:, 0, 0, 0], (torch.arange(B) % 5).to(torch.uint8))
    ok &= torch.equal(full[:, 1, 0, 0], torch.arange(B).to(torch.uint8))
    ok &= torch.equal(full[:, 2, 0, 0], (w.view(B) * 10).round().to(torch.uint8))
    ok &= tuple(inter["pred_x0"].shape) == (3, B, 3, 4, 4) and torch.equal(inter["pred_x0"][1], full)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    sys.path.insert(0, ROOT)
    from sgdm_b200.parallel import shard_bounds

    for n in (1, 7, 8, 256, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_bounds(1024, 3, 8) == (384, 512)  # config 5: 128 per GPU


import pytest  # noqa: E402


@pytest.mark.parametrize("B", [7, 1])
def test_sample_sharded_world2_gloo(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (50 if B == 1 else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, B)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
