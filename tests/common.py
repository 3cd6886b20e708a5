"""Shared helpers for the parity tests."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

UNET_CASES = [
    "unet_fast_label_tiny", "unet_fast_clusterlayout_tiny", "unetca_clusterlayout_tiny",
    "unetca_stego_tiny", "unetca_layout_tiny", "unet_heads32_ds24_tiny", "unetca_tokens4_cls_tiny", "unetca_tokens12_mean_tiny",
    "cfg1_cifar_label", "cfg2_in64_label", "cfg4_voc_clusterlayout",
    "cfg5_coco_stego",
]


def load_unet_case(name):
    z = np.load(os.path.join(GOLDEN, f"unet_{name}.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    arrays = {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}
    return meta, arrays


def load_npz(name):
    z = np.load(os.path.join(GOLDEN, name))
    meta = json.loads(bytes(z["meta"]).decode()) if "meta" in z.files else None
    return meta, {k: z[k] for k in z.files if k != "meta"}


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr_u8(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10 * np.log10(255.0**2 / mse)


def kwargs_from_arrays(arrays):
    return {k[3:]: v for k, v in arrays.items() if k.startswith("kw_")}
