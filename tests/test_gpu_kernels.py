"""Unit parity of every CUDA kernel, called through the C ABI (ctypes), against torch fp32.

The 16-bit operand tensors are rounded ONCE (to the library's operand dtype) and the torch
reference consumes the rounded values in fp32, so the comparison isolates the kernel's own
arithmetic (fp32 accumulation order): tolerances are ~1e-4 relative, far below the
operand-rounding error budget measured end to end in test_gpu_e2e.py.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from sgdm_b200 import _lib

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = _lib.lib()
    lib._op = torch.float16 if lib.sgdm_operand_dtype() == b"f16" else torch.bfloat16
    return lib


def S():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def ck(lib, rc):
    assert rc == 0, lib.sgdm_last_error().decode()


def out_tol(lib, fp16_tol):
    """Tolerance for a kernel OUTPUT stored in the operand type: its final rounding is 8x coarser in the bf16 build."""
    return fp16_tol if lib._op == torch.float16 else 8 * fp16_tol


def relerr(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def pack_weight(lib, w, extra=None, block_n=None):
    """torch conv weight [Co,Ci,k,k] (+ optional 1x1 skip weight [Co,C2,1,1]) -> packed op tensor."""
    Co, Ci, ks, _ = w.shape
    C2 = 0 if extra is None else extra.shape[1]
    ktot = ks * ks * Ci + C2
    bn = block_n or next((b for b in (256, 128, 64, 32) if Co % b == 0), 16)
    npad = (Co + bn - 1) // bn * bn
    dst = torch.zeros(npad, ktot, dtype=lib._op, device="cuda")
    ck(lib, lib.sgdm_k_pack_weight(S(), P(w.contiguous()), P(dst), Co, Ci, ks, Ci, ktot, 0))
    if extra is not None:
        ck(lib, lib.sgdm_k_pack_weight(S(), P(extra.contiguous()), P(dst), Co, C2, 1, C2, ktot, ks * ks * Ci))
    return dst, bn


def run_conv(lib, x_nhwc, w_packed, bn, ks, stride, Cout, Hout, Wout, bias=None, in2=None, res=None, res_mode=0,
             out="f32", naive=0):
    B, Hin, Win, Cin = x_nhwc.shape
    C2 = 0 if in2 is None else in2.shape[-1]
    o32 = oop = onchw = None
    if out == "f32":
        o32 = torch.full((B, Hout, Wout, Cout), float("nan"), device="cuda")
    elif out == "op":
        oop = torch.zeros((B, Hout, Wout, Cout), dtype=lib._op, device="cuda")
    else:
        onchw = torch.full((B, Cout, Hout, Wout), float("nan"), device="cuda")
    ck(lib, lib.sgdm_k_conv(S(), P(x_nhwc), B, Hin, Win, Cin, P(in2), C2, P(w_packed), ks, stride, Hout, Wout, Cout,
                            P(bias), P(res), res_mode, P(o32), P(oop), P(onchw), bn, naive))
    torch.cuda.synchronize()
    return o32 if o32 is not None else (oop if oop is not None else onchw)


CONV_CASES = [
    # (B, H, W, Cin, Cout, ks, stride, skipC, res_mode, out, note)
    (2, 16, 16, 64, 128, 3, 1, 0, 0, "f32", "3x3 basic, tile = half image"),
    (1, 64, 64, 64, 128, 3, 1, 0, 1, "f32", "64x64: tile = 2 image rows, residual"),
    (4, 8, 8, 128, 256, 3, 1, 0, 0, "f32", "8x8: tile spans 2 images, N=256"),
    (16, 4, 4, 256, 256, 3, 1, 0, 1, "f32", "4x4: tile spans 8 images"),
    (3, 32, 32, 128, 128, 3, 1, 0, 0, "op", "16-bit output, odd batch"),
    (2, 16, 16, 192, 64, 3, 1, 0, 0, "f32", "Cin=192 (3 chunks), N=64"),
    (2, 32, 32, 128, 128, 3, 2, 0, 0, "f32", "stride 2 (TMA elementStrides)"),
    (2, 16, 16, 256, 256, 3, 2, 0, 0, "f32", "stride 2 -> 8x8"),
    (2, 16, 16, 128, 256, 3, 1, 384, 0, "f32", "fused 1x1 skip source (K = 9*128 + 384)"),
    (2, 16, 16, 128, 128, 3, 1, 0, 2, "f32", "residual from nearest-2x upsampled source"),
    (2, 16, 16, 512, 1536, 1, 1, 0, 0, "op", "1x1 qkv GEMM"),
    (2, 16, 16, 512, 640, 1, 1, 0, 0, "op", "1x1, N=640 (5 tiles of 128)"),
    (2, 16, 16, 256, 320, 1, 1, 0, 0, "op", "1x1, N=320 (5 tiles of 64)"),
    (200, 1, 1, 384, 512, 1, 1, 0, 0, "f32", "linear, M=200 (ragged last tile)"),
    (90, 16, 16, 512, 1536, 1, 1, 0, 0, "op", "1x1 qkv GEMM, 180 m-tiles x 6 n-tiles (A-stationary: slot reuse across m-tiles)"),
    (80, 16, 16, 256, 768, 1, 1, 0, 1, "f32", "1x1, K=256, 3 n-tiles, residual, 160 m-tiles"),
    (5, 1, 1, 768, 1024, 1, 1, 0, 0, "f32", "linear, tiny M"),
    (2, 32, 32, 64, 3, 3, 1, 0, 0, "nchw", "final conv: N=3 padded to 16, NCHW epilogue"),
    (1, 64, 64, 128, 3, 3, 1, 0, 0, "nchw", "final conv at 64x64"),
    (33, 16, 16, 64, 64, 3, 1, 0, 0, "f32", "66 tiles, ragged batch (B=33)"),
    (5, 64, 64, 64, 128, 3, 1, 0, 1, "f32", "160 tiles > 148 CTAs: second tile on some CTAs"),
    (40, 32, 32, 128, 256, 3, 1, 128, 0, "f32", "320 tiles x N=256: persistent loop, both TMEM stages, phase flips"),
    (37, 16, 16, 512, 512, 1, 1, 0, 1, "f32", "74 m-tiles x 2 n-tiles"),
    (3, 64, 64, 384, 128, 3, 1, 0, 0, "op", "Cout=128: 384->128 at 64x64, 16-bit out (swap-AB candidate)"),
    (3, 8, 8, 128, 128, 3, 1, 0, 1, "f32", "Cout=128 at 8x8: 256-pixel tile spans 4 images, ragged"),
    (40, 32, 32, 128, 128, 3, 2, 0, 0, "f32", "Cout=128 stride 2, 320 tiles"),
    (2, 32, 32, 256, 128, 3, 1, 384, 0, "f32", "Cout=128 with fused 1x1 skip"),
    (5, 16, 16, 256, 256, 3, 1, 0, 1, "f32", "halo candidate: 256->256 at 16x16, residual, odd batch"),
    (3, 32, 32, 192, 256, 3, 1, 0, 0, "op", "halo candidate: Cin=192 at 32x32, 16-bit out"),
    (2, 64, 64, 128, 512, 3, 1, 0, 1, "f32", "halo candidate: 64x64 (2-row tiles), two n-tiles, residual"),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[-1] for c in CONV_CASES])
def test_conv_tcgen05_vs_torch(L, case):
    B, H, W, Cin, Cout, ks, stride, skipC, res_mode, out, note = case
    g = torch.Generator(device="cuda").manual_seed(hash(note) % 2**31)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    w = (torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks))
    bias = torch.randn(Cout, device="cuda", generator=g)
    Ho, Wo = H // stride, W // stride
    in2 = wskip = None
    if skipC:
        in2 = torch.randn(B, Ho, Wo, skipC, device="cuda", generator=g).to(L._op)
        wskip = torch.randn(Cout, skipC, 1, 1, device="cuda", generator=g) / math.sqrt(skipC)
    res = None
    if res_mode == 1:
        res = torch.randn(B, Ho, Wo, Cout, device="cuda", generator=g)
    elif res_mode == 2:
        res = torch.randn(B, Ho // 2, Wo // 2, Cout, device="cuda", generator=g)
    wp, bn = pack_weight(L, w, wskip)
    # reference on the rounded operands
    wr = w.to(L._op).float()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wr, bias, stride=stride, padding=1 if ks == 3 else 0)
    if skipC:
        ref = ref + F.conv2d(in2.float().permute(0, 3, 1, 2), wskip.to(L._op).float())
    if res_mode == 1:
        ref = ref + res.permute(0, 3, 1, 2)
    elif res_mode == 2:
        ref = ref + F.interpolate(res.permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # (label, block_n argument, naive, CTA-pair mode, halo mode); -1 = the library's policy
    modes = [("naive", bn, 1, 0, 0), ("tcgen05 1-CTA", bn, 0, 0, 0)]
    if bn >= 64:
        modes.append(("tcgen05 CTA pair (cta_group::2)", bn, 0, 1, 0))
    if Cout == 128 and out != "nchw" and res_mode != 2:
        modes.append(("tcgen05 swap-AB (block_n=0: engine policy)", 0, 0, -1, 0))
    if ks == 3 and stride == 1 and (H * W) % 128 == 0 and W % 8 == 0:
        modes.append(("tcgen05 halo (3 vertical taps share one staged tile), policy pair/swap", 0 if Cout == 128 and out != "nchw" and res_mode != 2 else bn, 0, -1, 1))
        modes.append(("tcgen05 halo, 1-CTA", bn, 0, 0, 1))
        # K blocks of 32 channels (64-byte rows, SWIZZLE_64B operands), forced: 1-CTA and, for Cout = 128, swap-AB
        modes.append(("tcgen05 halo, 32-channel K blocks, 1-CTA", bn, 0, 0, 1, 1))
        if Cout == 128 and out != "nchw" and res_mode != 2:
            modes.append(("tcgen05 halo, 32-channel K blocks, swap-AB", 0, 0, 0, 1, 1))
    if ks == 1 and not skipC and Cin <= 512 and out != "nchw":
        # A-stationary main loop (resident activation K blocks, weight-only ring), forced: 1-CTA and CTA pair
        modes.append(("tcgen05 A-stationary, 1-CTA", bn, 0, 0, 0, -1, 1))
        if bn >= 64:
            modes.append(("tcgen05 A-stationary, CTA pair", bn, 0, 1, 0, -1, 1))
    for mode in modes:
        label, bn_arg, naive, pair, halo = mode[:5]
        k32 = mode[5] if len(mode) > 5 else -1
        L.sgdm_debug_set_conv_astat(mode[6] if len(mode) > 6 else -1)
        L.sgdm_debug_set_conv_pair(pair)
        L.sgdm_debug_set_conv_halo(halo)
        L.sgdm_debug_set_conv_k32(k32)
        try:
            got = run_conv(L, x, wp, bn_arg, ks, stride, Cout, Ho, Wo, bias, in2, res, res_mode, out, naive)
        except AssertionError as e:
            if halo == 1 and ("halo mode needs" in str(e) or "shared memory budget" in str(e) or "k32 needs" in str(e)) or "A-stationary mode" in str(e):
                print(f"[conv {note}] {label}: not applicable ({e})")
                continue
            raise
        finally:
            L.sgdm_debug_set_conv_pair(-1)
            L.sgdm_debug_set_conv_halo(-1)
            L.sgdm_debug_set_conv_k32(-1)
            L.sgdm_debug_set_conv_astat(-1)
        got = got.float() if out == "nchw" else got.float().permute(0, 3, 1, 2)
        assert torch.isfinite(got).all(), f"non-finite output ({label})"
        e = relerr(got, ref)
        print(f"[conv {note}] {label} rel_l2={e:.3e}")
        assert e < (out_tol(L, 2e-3) if out == "op" else 2e-5), f"{label} conv mismatch {e}"


HFOLD_CASES = [
    (3, 64, 64, 128, 3, "64x64 head of config 2 (2-row tiles)"),
    (2, 32, 32, 64, 3, "32x32 head of config 1"),
    (5, 16, 16, 128, 3, "16x16, odd batch"),
    (300, 16, 16, 64, 3, "600 tiles > 148 CTAs: ring and exchange-buffer reuse"),
    (2, 32, 32, 128, 5, "Cout=5: 15 folded columns"),
    (1, 128, 128, 64, 4, "128-wide rows (one row per tile), Cout=4"),
]


@pytest.mark.parametrize("case", HFOLD_CASES, ids=[c[-1] for c in HFOLD_CASES])
def test_conv_head_folded_horizontal_taps(L, case):
    """The output head with the three horizontal taps folded into the GEMM N dimension (ConvDesc::hfold)
    against torch conv2d on the same rounded operands; also against the plain per-tap kernel path."""
    B, H, W, Cin, Cout, note = case
    g = torch.Generator(device="cuda").manual_seed(hash(note) % 2**31)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(Cin * 9)
    bias = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(L._op).float(), bias, padding=1)
    out = torch.full((B, Cout, H, W), float("nan"), device="cuda")
    scratch = torch.empty(16 * 3 * Cin, dtype=L._op, device="cuda")
    ck(L, L.sgdm_k_conv_head_hfold(S(), P(x), B, H, W, Cin, P(w), P(scratch), P(bias), P(out), Cout))
    torch.cuda.synchronize()
    e = relerr(out, ref)
    print(f"[conv head hfold {note}] rel_l2={e:.3e}")
    assert e < 2e-5
    if Cout <= 3:
        wp, bn = pack_weight(L, w)
        out2 = run_conv(L, x, wp, bn, 3, 1, Cout, H, W, bias=bias, out="nchw")
        assert relerr(out, out2) < 2e-6


BN192_CASES = [
    # (B, H, W, Cin, Cout, ks, res, out, note)
    (3, 16, 16, 384, 384, 3, True, "f32", "3x3 384 -> 384 at 16x16, residual, two 192-column tiles"),
    (2, 32, 32, 128, 192, 3, False, "op", "3x3 128 -> 192 at 32x32, 16-bit out (three 64-channel chunks)"),
    (5, 16, 16, 384, 1152, 1, False, "op", "1x1 qkv 384 -> 1152: six tiles, A-stationary"),
    (4, 8, 8, 768, 384, 3, False, "f32", "3x3 768 -> 384 at 8x8 (tile spans two images)"),
]
PARTIAL_TILE_CASES = [
    (5, 16, 16, 512, 640, 1, False, "op", "1x1 q|kv 512 -> 640 as 256-column tiles: partial last tile"),
    (3, 16, 16, 256, 896, 1, True, "f32", "1x1 256 -> 896, fp32 + residual, partial last tile"),
    (2, 16, 16, 128, 640, 3, False, "op", "3x3 128 -> 640 (halo), partial last tile"),
]


@pytest.mark.parametrize("case", BN192_CASES + PARTIAL_TILE_CASES, ids=[c[-1] for c in BN192_CASES + PARTIAL_TILE_CASES])
def test_conv_192_column_tiles(L, case):
    """block_n = 192 (the engine's choice for 192 / 384 / 576 output channels): one CTA, CTA pair, policy; with the
    epilogue's GroupNorm statistics."""
    B, H, W, Cin, Cout, ks, with_res, out, note = case
    g = torch.Generator(device="cuda").manual_seed(abs(hash(note)) % 2**31)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    w = torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks)
    bias = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g) if with_res else None
    bn_t = 256 if case in PARTIAL_TILE_CASES else 192
    wp, _ = pack_weight(L, w, None, bn_t)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(L._op).float(), bias, padding=ks // 2)
    if with_res:
        ref = ref + res.permute(0, 3, 1, 2)
    M = B * H * W
    nblk = (M + 31) // 32
    for label, pair in (("1-CTA", 0), ("CTA pair", 1), ("policy", -1)):
        o32 = torch.full((B, H, W, Cout), float("nan"), device="cuda") if out == "f32" else None
        oop = torch.zeros((B, H, W, Cout), dtype=L._op, device="cuda") if out == "op" else None
        stats = torch.full((nblk, Cout // 4, 2), float("nan"), device="cuda")
        L.sgdm_debug_set_conv_pair(pair)
        try:
            ck(L, L.sgdm_k_conv_stats(S(), P(x), B, H, W, Cin, None, 0, P(wp), ks, 1, H, W, Cout, P(bias), P(res),
                                      1 if with_res else 0, P(o32), P(oop), None, bn_t, 0, P(stats), 4, None, None, 0))
        finally:
            L.sgdm_debug_set_conv_pair(-1)
        torch.cuda.synchronize()
        got = (o32 if o32 is not None else oop).float()
        e = relerr(got.permute(0, 3, 1, 2), ref)
        print(f"[conv block_n 192: {note}] {label} rel_l2={e:.3e}")
        assert e < (out_tol(L, 2e-3) if out == "op" else 2e-5), label
        blk = got.double().reshape(nblk, 32, Cout // 4, 4)
        want = torch.stack([blk.sum((1, 3)), (blk * blk).sum((1, 3))], -1)
        assert relerr(stats, want) < 1e-5, label


UP2_CASES = [
    # (B, H, W, Cin, Cout, out, gran, note)  -- H, W are the LOW resolution
    (3, 16, 16, 128, 256, "f32", 4, "16x16 -> 32x32, N = 4 x 256, odd batch"),
    (2, 32, 32, 256, 256, "op", 4, "32x32 -> 64x64, 16-bit out"),
    (2, 32, 32, 512, 512, "op", 0, "config-2 shape 512 -> 512, two n-tiles per parity, no statistics"),
    (1, 64, 64, 64, 128, "f32", 2, "64x64 -> 128x128 (2-row tiles), block_n 128"),
    (5, 8, 16, 192, 64, "f32", 2, "8x16 images, Cin = 192, N = 64"),
    (3, 16, 16, 384, 384, "op", 4, "config-2 shape 384 -> 384 at 16x16: 192-column tiles, two per parity"),
    (40, 16, 16, 128, 256, "op", 4, "80 m-tiles x 4 parities: persistent loop, both accumulator stages"),
    (6, 8, 8, 256, 256, "f32", 4, "8x8 -> 16x16: dense geometry (plain 2x2 conv per parity), tile = 2 images"),
    (5, 8, 8, 128, 512, "op", 4, "8x8, odd batch: ragged last tile, two n-tiles per parity, 16-bit out"),
    (9, 4, 8, 64, 64, "f32", 2, "4x8 images: tile = 4 images, ragged, N = 64"),
    (64, 8, 8, 512, 512, "op", 4, "config-2 shape 512 -> 512 at 8x8, 32 m-tiles"),
]


@pytest.mark.parametrize("case", UP2_CASES, ids=[c[-1] for c in UP2_CASES])
def test_conv_subpixel_upsample_vs_torch(L, case):
    """"nearest-2x upsample, then 3x3 conv" run as four 2x2 parity convs on the low-resolution tensor (ConvDesc::up2):
    (a) against torch conv2d on the upsampled tensor (the definition; differs only by the single rounding of the summed
    weights), (b) tightly against the same parity sums formed in torch, (c) the epilogue's GroupNorm statistics."""
    B, H, W, Cin, Cout, out, gran, note = case
    g = torch.Generator(device="cuda").manual_seed(abs(hash(note)) % 2**31)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(Cin * 9)
    bias = torch.randn(Cout, device="cuda", generator=g)
    xu = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest")
    ref_def = F.conv2d(xu, w, bias, padding=1)                                   # (a) unrounded weights
    # (b) parity (dy, dx): 2x2 kernel over the low-resolution input, taps at offsets (dy-1+a, dx-1+b)
    V = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}
    ref_par = torch.empty_like(ref_def)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))
    for dy in (0, 1):
        for dx in (0, 1):
            k = torch.zeros(Cout, Cin, 2, 2, device="cuda")
            for a in (0, 1):
                for b in (0, 1):
                    k[:, :, a, b] = sum(w[:, :, r, s] for r in V[(dy, a)] for s in V[(dx, b)])
            k = k.to(L._op).float()
            o = F.conv2d(xp[:, :, dy:dy + H + 1, dx:dx + W + 1], k, bias)       # [B, Cout, H, W]
            ref_par[:, :, dy::2, dx::2] = o
    scratch = torch.empty(4 * Cout * 9 * Cin, dtype=L._op, device="cuda")
    nblk = B * 4 * H * W // 32
    for label, naive, pair in (("naive", 1, -1), ("tcgen05 1-CTA", 0, 0), ("tcgen05 CTA pair", 0, 1), ("policy", 0, -1)):
        o32 = torch.full((B, 2 * H, 2 * W, Cout), float("nan"), device="cuda") if out == "f32" else None
        oop = torch.zeros((B, 2 * H, 2 * W, Cout), dtype=L._op, device="cuda") if out == "op" else None
        stats = torch.full((nblk, Cout // gran, 2), float("nan"), device="cuda") if gran and not naive else None
        L.sgdm_debug_set_conv_pair(pair)
        try:
            ck(L, L.sgdm_k_conv_up2(S(), P(x), B, H, W, Cin, P(w.contiguous()), P(scratch), P(bias), P(o32), P(oop), Cout,
                                    P(stats), gran or 4, naive))
        finally:
            L.sgdm_debug_set_conv_pair(-1)
        torch.cuda.synchronize()
        got = (o32 if o32 is not None else oop).float()
        assert torch.isfinite(got).all(), label
        e_def = relerr(got.permute(0, 3, 1, 2), ref_def)
        e_par = relerr(got.permute(0, 3, 1, 2), ref_par)
        print(f"[conv up2 {note}] {label}: vs definition {e_def:.3e}, vs parity sums {e_par:.3e}")
        assert e_par < (out_tol(L, 2e-3) if out == "op" else 2e-5), label
        assert e_def < out_tol(L, 3e-3), label
        if stats is not None:
            # sample n's 4HW/32 row blocks: parity-major, then 32-pixel blocks of the LOW-resolution raster
            q = got.double().reshape(B, H, 2, W, 2, Cout).permute(0, 2, 4, 1, 3, 5).reshape(B, 4, H * W // 32, 32, Cout // gran, gran)
            want = torch.stack([q.sum((3, 5)), (q * q).sum((3, 5))], -1).reshape(nblk, Cout // gran, 2)
            assert torch.isfinite(stats).all(), label
            assert relerr(stats, want) < 1e-5, label
            # and per sample they add up to the sample's plain sums (what gn_finalize consumes)
            tot = stats.double().reshape(B, -1, Cout // gran, 2).sum(1)
            q2 = got.double().reshape(B, -1, Cout // gran, gran)
            assert relerr(tot[..., 0], q2.sum((1, 3))) < 1e-6 or float(q2.sum((1, 3)).abs().max()) < 1e-3


STATS_CASES = [
    # (B, H, W, Cin, Cout, ks, res_mode, out, block_n, gran, note)
    (3, 16, 16, 64, 256, 3, 1, "f32", 256, 4, "fp32 out + residual, N=256, gran 4"),
    (3, 16, 16, 64, 256, 3, 0, "op", 256, 2, "16-bit out, gran 2"),
    (2, 32, 32, 128, 128, 3, 1, "f32", 0, 4, "swap-AB fp32 out + residual"),
    (3, 32, 32, 128, 128, 3, 0, "op", 0, 4, "swap-AB 16-bit out (h1), ragged tile"),
    (2, 32, 32, 128, 128, 3, 0, "op", 0, 2, "swap-AB 16-bit out, gran 2"),
    (5, 8, 8, 128, 64, 3, 0, "f32", 64, 2, "N=64, 8x8 images (two row blocks per sample)"),
    (2, 16, 16, 512, 512, 1, 1, "f32", 256, 4, "1x1 proj + residual, two n-tiles"),
]


@pytest.mark.parametrize("case", STATS_CASES, ids=[c[-1] for c in STATS_CASES])
def test_conv_epilogue_groupnorm_stats(L, case):
    """The conv epilogue's GroupNorm partial sums equal sums over the tensor it wrote."""
    B, H, W, Cin, Cout, ks, res_mode, out, bn, gran, note = case
    g = torch.Generator(device="cuda").manual_seed(abs(hash(note)) % 2**31)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    w = torch.randn(Cout, Cin, ks, ks, device="cuda", generator=g) / math.sqrt(Cin * ks * ks)
    bias = torch.randn(Cout, device="cuda", generator=g)
    res = torch.randn(B, H, W, Cout, device="cuda", generator=g) if res_mode else None
    wp, bn_auto = pack_weight(L, w, None, bn or None)
    M = B * H * W
    nblk = (M + 31) // 32
    stats = torch.full((nblk, Cout // gran, 2), float("nan"), device="cuda")
    o32 = torch.full((B, H, W, Cout), float("nan"), device="cuda") if out == "f32" else None
    oop = torch.zeros((B, H, W, Cout), dtype=L._op, device="cuda") if out == "op" else None
    copy16 = torch.zeros((B, H, W, Cout), dtype=L._op, device="cuda") if out == "f32" else None  # second output
    L.sgdm_debug_set_conv_pair(1 if B % 2 else 0)  # odd-batch cases run as CTA pairs, the others as single CTAs
    try:
        ck(L, L.sgdm_k_conv_stats(S(), P(x), B, H, W, Cin, None, 0, P(wp), ks, 1, H, W, Cout, P(bias), P(res), res_mode,
                                  P(o32), P(oop), None, bn, 0, P(stats), gran, P(copy16), None, 0))
    finally:
        L.sgdm_debug_set_conv_pair(-1)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(L._op).float(), bias, padding=1 if ks == 3 else 0)
    if res_mode:
        ref = ref + res.permute(0, 3, 1, 2)
    ref = ref.permute(0, 2, 3, 1).reshape(M, Cout)
    got = (o32 if o32 is not None else oop).float().reshape(M, Cout)
    assert relerr(got, ref) < (2e-3 if out == "op" else 2e-5)
    if copy16 is not None:  # the 16-bit copy is exactly the rounded fp32 output
        assert torch.equal(copy16.reshape(M, Cout), o32.reshape(M, Cout).to(L._op))
    # the statistics describe the values the kernel WROTE (for a 16-bit output: the rounded values)
    pad = nblk * 32 - M
    gotp = torch.cat([got, torch.zeros(pad, Cout, device="cuda")]) if pad else got
    blk = gotp.double().reshape(nblk, 32, Cout // gran, gran)
    want = torch.stack([blk.sum((1, 3)), (blk * blk).sum((1, 3))], -1)
    assert torch.isfinite(stats).all()
    e = relerr(stats, want)
    print(f"[conv stats {note}] rel_l2={e:.3e}")
    assert e < 1e-5


def test_conv_skip_source_from_two_tensors(L):
    """Fused 1x1 skip whose source is the channel concat of two 16-bit tensors (never materialised)."""
    B, H, W, Cin, Cout, Ca, Cb = 3, 16, 16, 128, 256, 192, 64
    g = torch.Generator(device="cuda").manual_seed(21)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g).to(L._op)
    sa = torch.randn(B, H, W, Ca, device="cuda", generator=g).to(L._op)
    sb = torch.randn(B, H, W, Cb, device="cuda", generator=g).to(L._op)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / math.sqrt(Cin * 9)
    ws = torch.randn(Cout, Ca + Cb, 1, 1, device="cuda", generator=g) / math.sqrt(Ca + Cb)
    bias = torch.randn(Cout, device="cuda", generator=g)
    wp, bn = pack_weight(L, w, ws)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(L._op).float(), bias, padding=1)
    ref = ref + F.conv2d(torch.cat([sa, sb], -1).float().permute(0, 3, 1, 2), ws.to(L._op).float())
    for halo in (0, 1):
        o32 = torch.full((B, H, W, Cout), float("nan"), device="cuda")
        L.sgdm_debug_set_conv_halo(halo)
        try:
            ck(L, L.sgdm_k_conv_stats(S(), P(x), B, H, W, Cin, P(sa), Ca, P(wp), 3, 1, H, W, Cout, P(bias), None, 0,
                                      P(o32), None, None, bn, 0, None, 4, None, P(sb), Cb))
        finally:
            L.sgdm_debug_set_conv_halo(-1)
        torch.cuda.synchronize()
        e = relerr(o32.permute(0, 3, 1, 2), ref)
        print(f"[conv skip source = concat of two tensors, halo={halo}] rel_l2={e:.3e}")
        assert e < 2e-5


def torch_partial_stats(t, gran):
    """[B,H,W,C] fp32 -> the conv-epilogue statistics layout [rows/32, C/gran, 2]."""
    C = t.shape[-1]
    blk = t.double().reshape(-1, 32, C // gran, gran)
    return torch.stack([blk.sum((1, 3)), (blk * blk).sum((1, 3))], -1).float().contiguous()


FUSED_GN_CASES = [
    # (B, H, W, C0, C1, gran, half_in, note)
    (3, 16, 16, 128, 0, 4, False, "C=128 gran 4"),
    (2, 8, 8, 256, 128, 4, False, "concat 256+128: 12 ch/group straddles the sources"),
    (2, 16, 16, 128, 64, 2, False, "concat 128+64 = 192 (6 ch/group), gran 2"),
    (2, 8, 8, 64, 0, 2, True, "C=64 (2 ch/group), 16-bit source"),
    (1, 64, 64, 128, 0, 4, True, "64x64 16-bit source (deep-unroll path)"),
    (2, 16, 16, 512, 512, 4, False, "concat 1024"),
    (2, 16, 16, 256, 128, 4, True, "16-bit concat 256+128 (both sources are epilogue copies)"),
    (3, 8, 8, 128, 64, 2, True, "16-bit concat 192, gran 2"),
]


@pytest.mark.parametrize("case", FUSED_GN_CASES, ids=[c[-1] for c in FUSED_GN_CASES])
def test_groupnorm_from_epilogue_stats(L, case):
    B, H, W, C0, C1, gran, half_in, note = case
    C = C0 + C1
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(B, H, W, C0, device="cuda", generator=g) * 2 + 0.5
    b = torch.randn(B, H, W, C1, device="cuda", generator=g) * 0.5 - 1 if C1 else None
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    st0 = torch_partial_stats(a, gran)  # statistics of the unrounded values, as the producing conv emits them
    st1 = torch_partial_stats(b, gran) if C1 else None
    src = a.to(L._op) if half_in else a
    if half_in and C1:
        b = b.to(L._op)
    out = torch.zeros(B, H, W, C, dtype=L._op, device="cuda")
    ck(L, L.sgdm_k_groupnorm_fused(S(), P(src), 1 if half_in else 0, P(b), B, H, W, C0, C1, P(gamma), P(beta), None, 0,
                                   1, 0, P(st0), P(st1), gran, P(out), None, None))
    torch.cuda.synchronize()
    x = torch.cat([a, b.float()], -1) if C1 else a
    ref = F.silu(F.group_norm(x.permute(0, 3, 1, 2), 32, gamma, beta, eps=1e-5))
    e = relerr(out.float().permute(0, 3, 1, 2), ref)
    print(f"[gn fused stats {note}] rel_l2={e:.3e}")
    assert e < out_tol(L, 1e-3)


def test_conv_rejects_bad_shapes(L):
    x = torch.zeros(1, 16, 16, 48, dtype=L._op, device="cuda")
    w = torch.zeros(64, 9 * 48, dtype=L._op, device="cuda")
    o = torch.zeros(1, 16, 16, 64, device="cuda")
    rc = L.sgdm_k_conv(S(), P(x), 1, 16, 16, 48, None, 0, P(w), 3, 1, 16, 16, 64, None, None, 0, P(o), None, None, 64, 0)
    assert rc != 0 and b"multiples of 64" in L.sgdm_last_error()


GN_CASES = [
    # (B, H, W, C0, C1, film, silu, resample, raw, note)
    (2, 16, 16, 128, 0, False, 1, 0, False, "plain GN+SiLU"),
    (3, 32, 32, 64, 0, True, 1, 0, False, "C=64 (2 ch/group), FiLM"),
    (2, 16, 16, 256, 128, False, 1, 0, True, "concat 256+128 (12 ch/group straddles the sources), raw copy"),
    (2, 8, 8, 512, 256, True, 1, 0, False, "concat 512+256 = 768 (24 ch/group)"),
    (2, 8, 8, 512, 512, False, 1, 0, True, "concat 1024"),
    (2, 16, 16, 128, 64, False, 1, 0, True, "concat 192 (6 ch/group)"),
    (2, 16, 16, 256, 0, False, 0, 0, False, "no SiLU (attention norm)"),
    (2, 32, 32, 128, 0, False, 1, 1, False, "avg-pool 2x2 after SiLU + pooled raw"),
    (2, 8, 8, 256, 0, False, 1, 2, False, "nearest 2x after SiLU"),
    (1, 64, 64, 128, 0, True, 1, 0, False, "64x64, B=1 (multi-chunk stats)"),
    (5, 4, 4, 256, 0, False, 1, 0, False, "4x4 images"),
]


@pytest.mark.parametrize("case", GN_CASES, ids=[c[-1] for c in GN_CASES])
def test_groupnorm(L, case):
    B, H, W, C0, C1, film, silu, resample, raw, note = case
    C = C0 + C1
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(B, H, W, C0, device="cuda", generator=g) * 2 + 0.5
    b = torch.randn(B, H, W, C1, device="cuda", generator=g) * 0.5 - 1 if C1 else None
    gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    beta = 0.1 * torch.randn(C, device="cuda", generator=g)
    fl = torch.randn(B, 2 * C + 8, device="cuda", generator=g) * 0.3 if film else None
    Ho, Wo = (H // 2, W // 2) if resample == 1 else (H * 2, W * 2) if resample == 2 else (H, W)
    out = torch.zeros(B, Ho, Wo, C, dtype=L._op, device="cuda")
    raw_out = torch.zeros(B, H, W, C, dtype=L._op, device="cuda") if raw else None
    pool = torch.zeros(B, Ho, Wo, C, device="cuda") if resample == 1 else None
    half_in = C1 == 0 and not raw and resample != 1 and (B + H) % 2 == 0  # also cover the 16-bit-source variant
    if half_in:
        a16 = a.to(L._op)
        a = a16.float()
    ck(L, L.sgdm_k_groupnorm(S(), P(a16 if half_in else a), 1 if half_in else 0, P(b), B, H, W, C0, C1, P(gamma), P(beta),
                             P(fl), 2 * C + 8 if film else 0, silu, resample, P(out), P(raw_out), P(pool)))
    torch.cuda.synchronize()
    x = torch.cat([a, b], -1) if C1 else a
    xc = x.permute(0, 3, 1, 2)
    ref = F.group_norm(xc, 32, gamma, beta, eps=1e-5)
    if film:
        ref = ref * (1 + fl[:, :C, None, None]) + fl[:, C:2 * C, None, None]
    if silu:
        ref = F.silu(ref)
    if resample == 1:
        ref = F.avg_pool2d(ref, 2, 2)
    elif resample == 2:
        ref = F.interpolate(ref, scale_factor=2, mode="nearest")
    e = relerr(out.float().permute(0, 3, 1, 2), ref)
    print(f"[gn {note}{' (16-bit source)' if half_in else ''}] rel_l2={e:.3e}")
    assert e < out_tol(L, 1e-3)  # output is rounded to the 16-bit operand type (2^-11 relative for fp16)
    if raw:
        assert relerr(raw_out.float(), x) < out_tol(L, 1e-3)
    if pool is not None:
        assert relerr(pool.permute(0, 3, 1, 2), F.avg_pool2d(xc, 2, 2)) < 1e-6


ATTN_CASES = [
    # (B, T, heads, D, n_extra, mqa, note)
    (2, 256, 8, 64, 0, False, "legacy qkv layout, T=256 d=64 (cfg2)"),
    (40, 256, 8, 64, 0, False, "legacy, T=256 d=64, 320 (sample, head) pairs > 148 CTAs (tcgen05 ring reuse)"),
    (3, 64, 8, 32, 0, False, "legacy, T=64 d=32 (cfg1)"),
    (2, 16, 8, 32, 0, False, "T=16 < one query tile"),
    (2, 256, 8, 64, 17, True, "multi-query + 17 context/null keys (Attention_LR)"),
    (40, 256, 8, 64, 17, True, "Attention_LR, 320 (sample, head) pairs > 148 CTAs (tcgen05 barrier ring reuse)"),
    (3, 256, 8, 64, 32, True, "MQA, 32 extra keys (full extra tile)"),
    (3, 256, 4, 64, 1, True, "MQA, a single extra key"),
    (2, 16, 8, 32, 17, True, "MQA, T=16 d=32"),
    (1, 100, 4, 64, 5, True, "ragged T=100"),
    (2, 256, 8, 128, 0, False, "legacy, T=256 d=128 (unet_fast_s64: mc=256)"),
    (2, 64, 8, 128, 0, False, "legacy, T=64 d=128"),
    (2, 256, 32, 8, 0, False, "32 heads on 256 channels: d=8 (zero-extended to 16)"),
    (2, 64, 32, 16, 0, False, "32 heads on 512 channels: d=16"),
    (1, 1024, 32, 8, 0, False, "T=1024 d=8 (64x64 images, attention at ds 2)"),
    (1, 1024, 32, 16, 0, False, "T=1024 d=16"),
    (2, 100, 4, 16, 5, True, "MQA d=16, ragged T, 5 extra keys"),
    (2, 256, 8, 64, 49, True, "MQA, 49 extra keys (cond_token_num 40: beyond the tcgen05 kernel's 32)"),
    (1, 1024, 8, 64, 0, False, "T=1024 d=64 (128x128 images, attention at ds 4): keys staged in blocks of 256"),
    (1, 1024, 4, 64, 17, True, "MQA T=1024 + 17 extra keys: ragged last key block"),
    (2, 576, 2, 128, 0, False, "T=576 d=128: three key blocks, the last one partial"),
]


@pytest.mark.parametrize("case", ATTN_CASES, ids=[c[-1] for c in ATTN_CASES])
def test_attention(L, case):
    B, T, H, D, nx, mqa, note = case
    g = torch.Generator(device="cuda").manual_seed(11)
    C = H * D
    if not mqa:
        qkv = torch.randn(B, T, 3 * C, device="cuda", generator=g).to(L._op)
        out = torch.zeros(B, T, C, dtype=L._op, device="cuda")
        outs = {}
        for mode in (0, -1):  # mma.sync kernel, then the tcgen05 kernel where it applies (T = 256, D = 64)
            L.sgdm_debug_set_attn_tc(mode)
            try:
                out.zero_()
                ck(L, L.sgdm_k_attention(S(), qkv.data_ptr(), 3 * C, 3 * D, qkv.data_ptr() + 2 * D, 3 * C, 3 * D,
                                         qkv.data_ptr() + 4 * D, 3 * C, 3 * D, None, None, 0, P(out), C, B, T, H, D,
                                         1 / math.sqrt(D)))
                torch.cuda.synchronize()
                outs[mode] = out.clone()
            finally:
                L.sgdm_debug_set_attn_tc(-1)
        # QKVAttentionLegacy on [N, H*3*D, T]
        x = qkv.float().permute(0, 2, 1)
        q, k, v = x.reshape(B * H, 3 * D, T).split(D, dim=1)
        s = 1 / math.sqrt(math.sqrt(D))
        w = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), -1)
        ref = torch.einsum("bts,bcs->bct", w, v).reshape(B, C, T).permute(0, 2, 1)
    else:
        nq = C + 2 * D
        buf = torch.randn(B, T, nq, device="cuda", generator=g).to(L._op)
        kx = torch.randn(B, nx, D, device="cuda", generator=g).to(L._op)
        vx = torch.randn(B, nx, D, device="cuda", generator=g).to(L._op)
        out = torch.zeros(B, T, C, dtype=L._op, device="cuda")
        outs = {}
        for mode in (0, -1):  # mma.sync kernel, then the tcgen05 Attention_LR kernel where it applies (T = 256, D = 64)
            L.sgdm_debug_set_attn_tc(mode)
            try:
                out.zero_()
                ck(L, L.sgdm_k_attention(S(), buf.data_ptr(), nq, D, buf.data_ptr() + 2 * C, nq, 0,
                                         buf.data_ptr() + 2 * (C + D), nq, 0, P(kx), P(vx), nx, P(out), C, B, T, H, D,
                                         D ** -0.5))
                torch.cuda.synchronize()
                outs[mode] = out.clone()
            finally:
                L.sgdm_debug_set_attn_tc(-1)
        q = buf[..., :C].float().reshape(B, T, H, D).permute(0, 2, 1, 3) * D ** -0.5
        k = torch.cat([kx.float(), buf[..., C:C + D].float()], 1)
        v = torch.cat([vx.float(), buf[..., C + D:].float()], 1)
        attn = torch.einsum("bhid,bjd->bhij", q, k).softmax(-1)
        ref = torch.einsum("bhij,bjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(B, T, C)
    e0 = relerr(outs[0].float(), ref)
    print(f"[attn {note}] mma.sync kernel rel_l2={e0:.3e}")
    assert e0 < out_tol(L, 3e-3)
    e = relerr(out.float(), ref)
    print(f"[attn {note}] rel_l2={e:.3e}")
    assert torch.isfinite(out.float()).all()
    assert e < out_tol(L, 3e-3)  # P and the output are rounded to 16 bits


def test_layernorm(L):
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, C in ((500, 512), (33, 256), (7, 1024), (64, 128)):
        x = torch.randn(rows, C, device="cuda", generator=g) * 3 + 1
        ga = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
        be = 0.1 * torch.randn(C, device="cuda", generator=g)
        res = torch.randn(rows, C, device="cuda", generator=g)
        o1 = torch.zeros(rows, C, dtype=L._op, device="cuda")
        o2 = torch.zeros(rows, C, device="cuda")
        ck(L, L.sgdm_k_layernorm(S(), P(x), P(ga), P(be), None, P(o1), None, rows, C))
        ck(L, L.sgdm_k_layernorm(S(), P(x), P(ga), P(be), P(res), None, P(o2), rows, C))
        torch.cuda.synchronize()
        ref = F.layer_norm(x, (C,), ga, be)
        assert relerr(o1.float(), ref) < out_tol(L, 1e-3)
        assert relerr(o2, ref + res) < 1e-5


def test_layernorm_with_groupnorm_statistics(L):
    """LayerNorm + residual that also emits the conv-epilogue-format GroupNorm partial sums of its output."""
    g = torch.Generator(device="cuda").manual_seed(31)
    for rows, C, gran in ((512, 512, 4), (64, 256, 2), (96, 1024, 4), (32, 128, 2)):
        x = torch.randn(rows, C, device="cuda", generator=g) * 2 + 0.5
        ga = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
        be = 0.1 * torch.randn(C, device="cuda", generator=g)
        res = torch.randn(rows, C, device="cuda", generator=g)
        out = torch.zeros(rows, C, device="cuda")
        st = torch.zeros(rows // 32, C // gran, 2, device="cuda")
        ck(L, L.sgdm_k_layernorm_stats(S(), P(x), P(ga), P(be), P(res), P(out), P(st), gran, rows, C))
        ref2 = torch.zeros(rows, C, device="cuda")
        ck(L, L.sgdm_k_layernorm(S(), P(x), P(ga), P(be), P(res), None, P(ref2), rows, C))
        torch.cuda.synchronize()
        assert relerr(out, F.layer_norm(x, (C,), ga, be) + res) < 1e-5
        assert torch.equal(out, ref2)  # same arithmetic as the plain kernel
        blk = out.double().view(rows // 32, 32, C // gran, gran)
        assert relerr(st[..., 0].double(), blk.sum((1, 3))) < 1e-5
        assert relerr(st[..., 1].double(), (blk * blk).sum((1, 3))) < 1e-5


def test_split_precision_operand_layout(L):
    """[hi | hi | lo] rows (sgdm_config.precision = 1): hi + lo reproduces the fp32 value to ~2^-22."""
    g = torch.Generator(device="cuda").manual_seed(32)
    rows, C = 96, 256
    x = torch.randn(rows, C, device="cuda", generator=g) * 3
    ga = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
    be = 0.1 * torch.randn(C, device="cuda", generator=g)
    o = torch.zeros(rows, 3 * C, dtype=L._op, device="cuda")
    ck(L, L.sgdm_k_layernorm_split3(S(), P(x), P(ga), P(be), P(o), rows, C))
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (C,), ga, be)
    hi, hi2, lo = o[:, :C].float(), o[:, C:2 * C].float(), o[:, 2 * C:].float()
    assert torch.equal(hi, hi2)
    assert relerr(hi, ref) < out_tol(L, 1e-3) and relerr(hi + lo, ref) < (1e-5 if L._op == torch.float16 else 1e-4)


def test_linear_f32(L):
    g = torch.Generator(device="cuda").manual_seed(5)
    for M, N, K in ((32, 512, 128), (5, 256, 1000), (130, 70, 33), (512, 256, 5000)):
        x = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
        b = torch.randn(N, device="cuda", generator=g)
        out = torch.zeros(M, N + 3, device="cuda")
        ck(L, L.sgdm_k_linear_f32(S(), P(x), K, P(W), P(b), P(out), N + 3, M, N, K, 1, 0))
        ck(L, L.sgdm_k_linear_f32(S(), P(x), K, P(W), P(b), P(out), N + 3, M, N, K, 0, 1))
        torch.cuda.synchronize()
        torch.backends.cuda.matmul.allow_tf32 = False
        lin = F.linear(x.double(), W.double(), b.double())
        ref = (F.silu(lin) + lin).float()
        assert relerr(out[:, :N], ref) < 1e-5
        assert out[:, N:].abs().max() == 0
    # split-K (the engine's path for the K = 1000 / 5000 condition MLPs at small batch): same values to fp32 rounding,
    # bit-identical from run to run, exact for one-hot inputs
    for M, N, K, splits in ((64, 256, 5000, 16), (64, 256, 5000, -1), (32, 512, 1000, 3), (7, 70, 1030, 4)):
        x = torch.randn(M, K, device="cuda", generator=g)
        W = torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)
        b = torch.randn(N, device="cuda", generator=g)
        part = torch.full((16 * M * N,), float("nan"), device="cuda")
        outs = []
        for _ in range(2):
            out = torch.zeros(M, N + 3, device="cuda")
            ck(L, L.sgdm_k_linear_f32_splitk(S(), P(x), K, P(W), P(b), P(out), N + 3, M, N, K, 1, 0, P(part), splits))
            ck(L, L.sgdm_k_linear_f32_splitk(S(), P(x), K, P(W), P(b), P(out), N + 3, M, N, K, 0, 1, P(part), splits))
            torch.cuda.synchronize()
            outs.append(out)
        lin = F.linear(x.double(), W.double(), b.double())
        assert relerr(outs[0][:, :N], (F.silu(lin) + lin).float()) < 1e-5
        assert torch.equal(outs[0], outs[1]) and outs[0][:, N:].abs().max() == 0
    idx = torch.randint(0, 5000, (8,), device="cuda", generator=g)
    oh = F.one_hot(idx, 5000).float()
    W = torch.randn(256, 5000, device="cuda", generator=g)
    b = torch.randn(256, device="cuda", generator=g)
    out = torch.zeros(8, 256, device="cuda")
    part = torch.empty(16 * 8 * 256, device="cuda")
    ck(L, L.sgdm_k_linear_f32_splitk(S(), P(oh), 5000, P(W), P(b), P(out), 256, 8, 256, 5000, 0, 0, P(part), -1))
    torch.cuda.synchronize()
    assert torch.equal(out, W[:, idx].t() + b)
    # one-hot input: the first Linear is an exact column gather
    idx = torch.randint(0, 1000, (16,), device="cuda", generator=g)
    oh = F.one_hot(idx, 1000).float()
    W = torch.randn(256, 1000, device="cuda", generator=g)
    b = torch.randn(256, device="cuda", generator=g)
    out = torch.zeros(16, 256, device="cuda")
    ck(L, L.sgdm_k_linear_f32(S(), P(oh), 1000, P(W), P(b), P(out), 256, 16, 256, 1000, 0, 0))
    torch.cuda.synchronize()
    assert torch.equal(out, W[:, idx].t() + b)


def test_cast_and_upsample(L):
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(3, 8, 8, 128, device="cuda", generator=g)
    a = torch.zeros(3, 8, 8, 128, dtype=L._op, device="cuda")
    b = torch.zeros(3, 16, 16, 128, dtype=L._op, device="cuda")
    ck(L, L.sgdm_k_cast(S(), P(x), P(a), 3, 8, 8, 128, 0))
    ck(L, L.sgdm_k_cast(S(), P(x), P(b), 3, 8, 8, 128, 1))
    torch.cuda.synchronize()
    assert torch.equal(a, x.to(L._op))
    up = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(b, up.to(L._op))


def test_sampler_update_kernels_bit_exact(L):
    """The fused mix+update kernels reproduce the unfused torch fp32 arithmetic of the
    reference bit for bit (oracle/sampler.py restates ddim_plms_sampler.py:360-391 and
    ddpm_sampler.py:154-192)."""
    import ctypes as C

    from oracle import sampler as osamp
    from oracle import schedule as osched

    g = torch.Generator().manual_seed(21)
    B, shape = 4, (4, 3, 16, 16)
    x, ec, eu, nz = (torch.randn(shape, generator=g) for _ in range(4))
    xd, ecd, eud, nzd = (t.cuda() for t in (x, ec, eu, nz))
    per = x[0].numel()
    tab = osched.ddpm_tables(1000)
    for eta in (0.0, 1.0):
        dt = osched.ddim_tables(tab["alphas_cumprod"], 10, 1000, eta)
        for index in (0, 4, 9):
            for w in (2.0, 0.3):
                e = (1 - w) * eu + w * ec
                kw = dict(clip_denoised=True, dtp=1, temperature=1.0, noise_dropout=0)
                ref_x, ref_x0 = osamp._ddim_update(x, e, dt, index, nz, kw)
                z = torch.zeros(())
                a_t = torch.full_like(z, dt["alphas"][index]); a_p = torch.full_like(z, dt["alphas_prev"][index])
                sg = torch.full_like(z, dt["sigmas"][index]); s1 = torch.full_like(z, dt["sqrt_one_minus_alphas"][index])
                coef = (C.c_float * 6)(s1.item(), a_t.sqrt().item(), a_p.sqrt().item(),
                                       (1.0 - a_p - sg**2).sqrt().item(), sg.item(), 1.0)
                xo, x0o, eo = (torch.empty_like(xd) for _ in range(3))
                ck(L, L.sgdm_ddim_step(S(), P(ecd), P(eud), w, None, 0, coef, 1, P(xd), P(nzd), P(xo), P(x0o), P(eo),
                                       B, per))
                torch.cuda.synchronize()
                assert torch.equal(eo.cpu(), e), "guidance mix not bit-exact"
                assert torch.equal(x0o.cpu(), ref_x0), "pred_x0 not bit-exact"
                assert torch.equal(xo.cpu(), ref_x), "x_prev not bit-exact"
    # DDPM
    for i in (0, 1, 500, 999):
        w = 2.0
        e = (1 - w) * eu + w * ec
        t = torch.full((B,), i, dtype=torch.long)
        x0 = osamp._ext(tab["sqrt_recip_alphas_cumprod"], t, x) * x - osamp._ext(tab["sqrt_recipm1_alphas_cumprod"], t, x) * e
        x0 = x0.clamp(-1, 1)
        mean = osamp._ext(tab["posterior_mean_coef1"], t, x) * x0 + osamp._ext(tab["posterior_mean_coef2"], t, x) * x
        logvar = osamp._ext(tab["posterior_log_variance_clipped"], t, x)
        nonzero = (1 - (t == 0).float()).reshape(B, 1, 1, 1)
        ref = mean + nonzero * (0.5 * logvar).exp() * (nz * 1.0)
        sig = (0.5 * tab["posterior_log_variance_clipped"]).exp()
        coef = (C.c_float * 6)(tab["sqrt_recip_alphas_cumprod"][i].item(), tab["sqrt_recipm1_alphas_cumprod"][i].item(),
                               tab["posterior_mean_coef1"][i].item(), tab["posterior_mean_coef2"][i].item(),
                               sig[i].item() if i else 0.0, 1.0)
        xo, x0o = torch.empty_like(xd), torch.empty_like(xd)
        ck(L, L.sgdm_ddpm_step(S(), P(ecd), P(eud), w, None, 0, coef, 1, P(xd), P(nzd), P(xo), P(x0o), B, per))
        torch.cuda.synchronize()
        assert torch.equal(x0o.cpu(), x0)
        assert torch.equal(xo.cpu(), ref)
    # sampling_kwargs extras: dynamic thresholding and noise dropout inside the fused updates
    mul = (torch.rand(shape, generator=g) >= 0.25).float() * torch.ones(()).div(0.75)
    muld = mul.cuda()
    s_dev = torch.empty(B, device="cuda")
    scratch = torch.empty_like(xd)
    big = x * torch.tensor([0.5, 1.0, 2.0, 4.0]).view(B, 1, 1, 1)  # samples below and above the s > 1 regime
    bigd = big.cuda()
    for dtp in (0.9, 0.995, 0.37):
        w = 2.0
        e = (1 - w) * eu + w * ec
        kw = dict(clip_denoised=True, dtp=dtp, temperature=0.7, noise_dropout=0.25)
        dt = osched.ddim_tables(tab["alphas_cumprod"], 10, 1000, 1.0)
        index = 4
        ref_x, ref_x0 = osamp._ddim_update(big, e, dt, index, nz, kw, mul)
        z = torch.zeros(())
        a_t = torch.full_like(z, dt["alphas"][index]); a_p = torch.full_like(z, dt["alphas_prev"][index])
        sg = torch.full_like(z, dt["sigmas"][index]); s1 = torch.full_like(z, dt["sqrt_one_minus_alphas"][index])
        coef = (C.c_float * 6)(s1.item(), a_t.sqrt().item(), a_p.sqrt().item(), (1.0 - a_p - sg**2).sqrt().item(),
                               sg.item(), 0.7)
        ck(L, L.sgdm_dyn_threshold(S(), 1, P(ecd), P(eud), w, None, 0, coef, P(bigd), dtp, P(scratch), P(s_dev), B, per))
        raw = (big - s1 * e) / a_t.sqrt()
        s_ref = torch.quantile(raw.flatten(1).abs(), dtp, dim=-1).clamp(min=1.0)
        assert torch.equal(scratch.cpu(), raw), "unclipped pred_x0 not bit-exact"
        assert torch.equal(s_dev.cpu(), s_ref), "dynamic threshold differs from torch.quantile"
        assert torch.equal(s_dev, torch.quantile(scratch.flatten(1).abs(), dtp, dim=-1).clamp(min=1.0)), "vs CUDA torch.quantile"
        xo, x0o = torch.empty_like(xd), torch.empty_like(xd)
        ck(L, L.sgdm_ddim_step_ex(S(), P(ecd), P(eud), w, None, 0, coef, 1, P(bigd), P(nzd), P(xo), P(x0o), None, B, per,
                                  P(s_dev), P(muld)))
        assert torch.equal(x0o.cpu(), ref_x0), "dynamically thresholded pred_x0 not bit-exact"
        assert torch.equal(xo.cpu(), ref_x), "x_prev with dtp + noise dropout not bit-exact"
        # DDPM flavour of the same
        i = 500
        t = torch.full((B,), i, dtype=torch.long)
        x0 = osamp._ext(tab["sqrt_recip_alphas_cumprod"], t, x) * big - osamp._ext(tab["sqrt_recipm1_alphas_cumprod"], t, x) * e
        x0 = osamp.clip_x0(x0, True, dtp)
        mean = osamp._ext(tab["posterior_mean_coef1"], t, x) * x0 + osamp._ext(tab["posterior_mean_coef2"], t, x) * big
        logvar = osamp._ext(tab["posterior_log_variance_clipped"], t, x)
        ref = mean + (0.5 * logvar).exp() * ((nz * 0.7) * mul)
        sig = (0.5 * tab["posterior_log_variance_clipped"]).exp()
        coef = (C.c_float * 6)(tab["sqrt_recip_alphas_cumprod"][i].item(), tab["sqrt_recipm1_alphas_cumprod"][i].item(),
                               tab["posterior_mean_coef1"][i].item(), tab["posterior_mean_coef2"][i].item(), sig[i].item(), 0.7)
        ck(L, L.sgdm_dyn_threshold(S(), 0, P(ecd), P(eud), w, None, 0, coef, P(bigd), dtp, P(scratch), P(s_dev), B, per))
        ck(L, L.sgdm_ddpm_step_ex(S(), P(ecd), P(eud), w, None, 0, coef, 1, P(bigd), P(nzd), P(xo), P(x0o), B, per,
                                  P(s_dev), P(muld)))
        assert torch.equal(x0o.cpu(), x0) and torch.equal(xo.cpu(), ref), "DDPM update with dtp + noise dropout"
    # quantile kernel alone: ties, constants, exact-rank q, odd lengths
    for n, q in ((1, 0.5), (2, 0.5), (777, 0.25), (12288, 0.9), (12288, 1.0 - 2**-20), (1001, 0.5)):
        v = (torch.randn(3, n, generator=g) * 2).round() / 2 if n > 2 else torch.randn(3, n, generator=g) * 3
        vd = v.cuda()
        so = torch.empty(3, device="cuda")
        ck(L, L.sgdm_k_quantile_abs(S(), P(vd), 3, n, q, P(so)))
        assert torch.equal(so.cpu(), torch.quantile(v.abs(), q, dim=-1).clamp(min=1.0)), (n, q)
    # per-sample tensor cond_scale and the 'cfg' scale type
    wt = torch.linspace(0.5, 3.0, B)
    out = torch.empty_like(xd)
    ck(L, L.sgdm_mix(S(), P(ecd), P(eud), 0.0, P(wt.cuda()), 0, P(out), B, per))
    assert torch.equal(out.cpu(), (1 - wt.view(B, 1, 1, 1)) * eu + wt.view(B, 1, 1, 1) * ec)
    ck(L, L.sgdm_mix(S(), P(ecd), P(eud), 0.1, None, 1, P(out), B, per))
    assert torch.equal(out.cpu(), (1 + 0.1) * ec - 0.1 * eu)
    # uint8 conversion
    v = torch.randn(10000, generator=g) * 1.5
    o8 = torch.empty(10000, dtype=torch.uint8, device="cuda")
    ck(L, L.sgdm_to_uint8(S(), P(v.cuda()), P(o8), 10000))
    assert torch.equal(o8.cpu(), osamp.to_uint8(v))
    # PLMS combination
    ptrs = (C.c_void_p * 4)(P(xd), P(ecd), P(eud), P(nzd))
    cf = (C.c_float * 4)(55.0, -59.0, 37.0, -9.0)
    ck(L, L.sgdm_lincomb(S(), 4, ptrs, cf, 24.0, P(out), xd.numel()))
    assert torch.equal(out.cpu(), (55 * x - 59 * ec + 37 * eu - 9 * nz) / 24)
