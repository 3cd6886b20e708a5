#!/usr/bin/env python
"""Micro-benchmark of the conv/GEMM kernel through the C ABI: epilogue variants of a few UNet layer shapes.
   python tools/bench_conv.py            (needs a B200; timings with CUDA events, L2-sized rotation of buffers)"""
import math, os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from sgdm_b200 import _lib

L = _lib.lib()
OP = torch.float16 if L.sgdm_operand_dtype() == b"f16" else torch.bfloat16
S = lambda: torch.cuda.current_stream().cuda_stream
P = lambda t: None if t is None else t.data_ptr()

def pack(w, bn):
    Co, Ci, ks, _ = w.shape
    ktot = ks * ks * Ci
    npad = (Co + bn - 1) // bn * bn
    dst = torch.zeros(npad, ktot, dtype=OP, device="cuda")
    assert L.sgdm_k_pack_weight(S(), P(w.contiguous()), P(dst), Co, Ci, ks, Ci, ktot, 0) == 0
    return dst

def run(name, B, H, Cin, Cout, ks, out, res, stats, pair, iters=10, halo=-1):
    bn = next(b for b in (256, 128, 64, 32) if Cout % b == 0)
    x = torch.randn(B, H, H, Cin, device="cuda").to(OP)
    w = torch.randn(Cout, Cin, ks, ks, device="cuda") / math.sqrt(Cin * ks * ks)
    wp = pack(w, bn)
    bias = torch.randn(Cout, device="cuda")
    r = torch.randn(B, H, H, Cout, device="cuda") if res else None
    o32 = torch.empty(B, H, H, Cout, device="cuda") if out == "f32" else None
    oop = torch.empty(B, H, H, Cout, dtype=OP, device="cuda") if out == "op" else None
    st = torch.empty((B * H * H + 31) // 32, Cout // 4, 2, device="cuda") if stats else None
    L.sgdm_debug_set_conv_pair(pair)
    L.sgdm_debug_set_conv_halo(halo)
    def go():
        rc = L.sgdm_k_conv_stats(S(), P(x), B, H, H, Cin, None, 0, P(wp), ks, 1, H, H, Cout, P(bias), P(r), 1 if res else 0,
                                 P(o32), P(oop), None, 0, 0, P(st), 4, None, None, 0)
        assert rc == 0, L.sgdm_last_error().decode()
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # one more launch with the cycle counters on
    tm = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.sgdm_debug_set_conv_timing(tm.data_ptr()); go(); torch.cuda.synchronize(); L.sgdm_debug_set_conv_timing(None)
    t = tm.tolist()
    nch = max(t[7], 1)
    names = ["acc", "buf", "res", "tmem", "stage", "stats", "store"]
    epi = " ".join(f"{n}={t[k]/nch:6.0f}" for k, n in enumerate(names))
    fl = 2.0 * B * H * H * Cout * Cin * ks * ks
    by = x.numel() * 2 + B * H * H * Cout * ((4 if out == "f32" else 2) + (4 if res else 0))
    L.sgdm_debug_set_conv_pair(-1)
    L.sgdm_debug_set_conv_halo(-1)
    print(f"halo={halo:2d} {name:34s} out={out:3s} res={int(res)} stats={int(stats)} pair={pair:2d}  {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s  {by/ms/1e6:6.0f} GB/s | clk/chunk: {epi} | mma_wait_smem={t[8]/1e6:.1f}M mma_wait_acc={t[9]/1e6:.1f}M prod_wait={t[10]/1e6:.1f}M", flush=True)

def run_up2(name, B, H, Cin, Cout, out, stats, pair, iters=10):
    """upsample + 3x3 conv in the sub-pixel mode; H = low resolution.  TFLOP/s on the EXECUTED MACs (4/9 of the definition's)."""
    x = torch.randn(B, H, H, Cin, device="cuda").to(OP)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / math.sqrt(Cin * 9)
    bias = torch.randn(Cout, device="cuda")
    scratch = torch.zeros(4 * Cout * 9 * Cin, dtype=OP, device="cuda")
    o32 = torch.empty(B, 2 * H, 2 * H, Cout, device="cuda") if out == "f32" else None
    oop = torch.empty(B, 2 * H, 2 * H, Cout, dtype=OP, device="cuda") if out == "op" else None
    st = torch.empty(B * 4 * H * H // 32, Cout // 4, 2, device="cuda") if stats else None
    L.sgdm_debug_set_conv_pair(pair)
    def go(mode=2):
        rc = L.sgdm_k_conv_up2(S(), P(x), B, H, H, Cin, P(w), P(scratch), P(bias), P(o32), P(oop), Cout, P(st), 4, mode)
        assert rc == 0, L.sgdm_last_error().decode()
    go(0)
    for _ in range(3): go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tm = torch.zeros(16, dtype=torch.int64, device="cuda")
    L.sgdm_debug_set_conv_timing(tm.data_ptr()); go(); torch.cuda.synchronize(); L.sgdm_debug_set_conv_timing(None)
    t = tm.tolist()
    nch = max(t[7], 1)
    names = ["acc", "buf", "res", "tmem", "stage", "stats", "store"]
    epi = " ".join(f"{n}={t[k]/nch:6.0f}" for k, n in enumerate(names))
    fl = 2.0 * B * 4 * H * H * Cout * Cin * 4
    L.sgdm_debug_set_conv_pair(-1)
    print(f"up2 {name:30s} out={out:3s} stats={int(stats)} pair={pair:2d}  {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s executed | clk/chunk: {epi} | mma_wait_smem={t[8]/1e6:.1f}M mma_wait_acc={t[9]/1e6:.1f}M prod_wait={t[10]/1e6:.1f}M", flush=True)


if os.environ.get("UP2"):
    for name, (B, H, Cin, Cout) in {"256->256 @32->64 B512": (512, 32, 256, 256), "384->384 @16->32 B512": (512, 16, 384, 384),
                                    "512->512 @16->32 B512": (512, 16, 512, 512), "512->512 @8->16 B512": (512, 8, 512, 512)}.items():
        for out, stats, pair in (("op", True, -1), ("op", False, -1), ("f32", False, -1), ("op", True, 0)):
            try:
                run_up2(name, B, H, Cin, Cout, out, stats, pair)
            except AssertionError as e:
                print("up2", name, out, stats, pair, "->", e)
    sys.exit(0)

SHAPES = {
    "proj 512->512 1x1 @16 B512": (512, 16, 512, 512, 1),
    "qkv 512->1536 1x1 @16 B512": (512, 16, 512, 1536, 1),
    "conv 128->128 3x3 @64 B512": (512, 64, 128, 128, 3),
    "conv 256->256 3x3 @32 B512": (512, 32, 256, 256, 3),
    "conv 512->512 3x3 @16 B512": (512, 16, 512, 512, 3),
    "first 64->128 3x3 @64 B512": (512, 64, 64, 128, 3),
    "first conv as the K=64 im2col GEMM @64 B512": (512, 64, 64, 128, 1),
}
if os.environ.get("RESIDUAL"):
    # residual-epilogue experiments: fp32 out + residual + statistics
    for name in sys.argv[1:] or ["proj 512->512 1x1 @16 B512", "conv 128->128 3x3 @64 B512", "conv 512->512 3x3 @16 B512"]:
        B, H, Cin, Cout, ks = SHAPES[name]
        run(name, B, H, Cin, Cout, ks, "f32", True, True, -1)
    sys.exit(0)
if os.environ.get("HALO"):
    SHAPES["last 128->3 3x3 @64 B512"] = (512, 64, 128, 3, 3)
    for name in ("conv 512->512 3x3 @16 B512", "conv 256->256 3x3 @32 B512", "conv 128->128 3x3 @64 B512", "first 64->128 3x3 @64 B512"):
        B, H, Cin, Cout, ks = SHAPES[name]
        for out, res, stats in (("op", False, True), ("f32", True, True)):
            for halo in (0, 1):
                try:
                    run(name, B, H, Cin, Cout, ks, out, res, stats, -1, halo=halo)
                except AssertionError as e:
                    print(f"halo={halo} {name} out={out} res={int(res)}: {e}")
    sys.exit(0)
sel = sys.argv[1:] or list(SHAPES)
for name in sel:
    B, H, Cin, Cout, ks = SHAPES[name]
    for out, res, stats in (("op", False, False), ("op", False, True), ("f32", False, False), ("f32", False, True), ("f32", True, False), ("f32", True, True)):
        for pair in ((-1,) if Cout == 128 else (0, 1)):
            run(name, B, H, Cin, Cout, ks, out, res, stats, pair)
