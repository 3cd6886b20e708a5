#!/usr/bin/env python
"""Micro-benchmark of the self-attention kernels through the C ABI (needs a B200): one attention site of config 2
(B = 512 samples x 8 heads, T = 256, d = 64), CUDA events, 20 launches."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sgdm_b200 import _lib

L = _lib.lib()
OP = torch.float16 if L.sgdm_operand_dtype() == b"f16" else torch.bfloat16
S = lambda: torch.cuda.current_stream().cuda_stream

def run(B, T, H, D, mqa=False, nx=0, iters=20):
    C = H * D
    if not mqa:
        qkv = torch.randn(B, T, 3 * C, device="cuda").to(OP)
        out = torch.zeros(B, T, C, dtype=OP, device="cuda")
        go = lambda: L.sgdm_k_attention(S(), qkv.data_ptr(), 3 * C, 3 * D, qkv.data_ptr() + 2 * D, 3 * C, 3 * D,
                                        qkv.data_ptr() + 4 * D, 3 * C, 3 * D, None, None, 0, out.data_ptr(), C, B, T, H, D,
                                        1 / math.sqrt(D))
    else:
        nq = C + 2 * D
        buf = torch.randn(B, T, nq, device="cuda").to(OP)
        kx = torch.randn(B, nx, D, device="cuda").to(OP); vx = torch.randn(B, nx, D, device="cuda").to(OP)
        out = torch.zeros(B, T, C, dtype=OP, device="cuda")
        go = lambda: L.sgdm_k_attention(S(), buf.data_ptr(), nq, D, buf.data_ptr() + 2 * C, nq, 0, buf.data_ptr() + 2 * (C + D), nq, 0,
                                        kx.data_ptr(), vx.data_ptr(), nx, out.data_ptr(), C, B, T, H, D, 1 / math.sqrt(D))
    for _ in range(3): assert go() == 0, L.sgdm_last_error().decode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * T * (T + nx) * D
    print(f"B={B} T={T} H={H} D={D} mqa={int(mqa)} nx={nx}: {ms*1e3:8.1f} us  {fl/ms/1e9:6.0f} TFLOP/s  dbg={os.environ.get('SGDM_ATTN_DBG', '0')}", flush=True)

if __name__ == "__main__":
    run(512, 256, 8, 64)
    if not os.environ.get("SGDM_ATTN_DBG"):
        run(512, 256, 8, 64, True, 17)
