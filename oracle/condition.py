"""Condition lookup (TEST INFRASTRUCTURE).

Restates prepare_condition_kwargs / prepare_denoise_fn_kwargs_4sampling
(dynamic_input/condition.py:5-86,141-157) for the four methods BASELINE.json
names, plus the n-hot construction of `stego_attr` from a one-hot mask
(dataset/transforms/complex_ds_common_util.py:126-133).
"""
import torch


def denoise_kwargs_for_sampling(condition_method, batch, cond_scale, clusterlayout_how="lost"):
    """-> kwargs handed to forward_with_cond_scale (cond_drop_prob removed, :152-155)."""
    if condition_method is None:
        kw = dict(cond=None)
    elif condition_method in ("label", "cluster"):
        kw = dict(cond=batch[condition_method])  # int64 one-hot [B, cond_dim]
    elif condition_method == "clusterlayout":
        key = {"lost": "lostbboxmask", "oracle": "segmask", "stego": "stegomask"}[clusterlayout_how]
        kw = dict(cond=batch["cluster"].float(), layout=batch[key].float())
    elif condition_method == "stegoclusterlayout":
        kw = dict(cond=batch["stego_attr"].float(), layout=batch["stegomask"].float())
    else:
        raise ValueError(condition_method)
    kw["cond_scale"] = cond_scale
    return kw


def stego_attr_nhot(stegomask_onehot):
    """[B,K,H,W] one-hot mask -> [B,K] n-hot of classes present."""
    return (stegomask_onehot.flatten(2).sum(-1) > 0).to(stegomask_onehot.dtype)
