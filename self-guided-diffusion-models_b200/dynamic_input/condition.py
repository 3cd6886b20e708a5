"""Condition lookup: dataset batch -> denoise kwargs (dynamic_input/condition.py:5-86,141-157).

Integer one-hot / n-hot / mask tensors are passed through or cast to float exactly as the
reference does; nothing here touches the GPU kernels (bit-exact by construction).
`pl_module` only needs `.hparams.{cond_dim,condition_method,cond_drop_prob,condition}`,
`.training` and `.device`.
"""


# every method whose condition is one [B, cond_dim] vector per sample passed through unchanged (condition.py:20-36):
# the UNet sees them exactly like `label` / `cluster` (README.md:32,59 runs `attr` on unetca_fast)
_VECTOR_METHODS = ["label", "attr", "feat", "knn_feat", "patchfeat", "centroid", "labelcentroid", "cluster", "clustermix",
                   "clusterrandom", "labelcluster", "patchcluster"]


def prepare_condition_kwargs(pl_module, batch_data):
    condition_method = pl_module.hparams.condition_method
    if condition_method is not None:
        assert pl_module.hparams.cond_drop_prob > 0
        cond_drop_prob = pl_module.hparams.cond_drop_prob if pl_module.training else 1.0
    else:
        cond_drop_prob = 1.0
    result = dict(cond_drop_prob=cond_drop_prob)
    dev = pl_module.device
    if condition_method is None:
        result.update(cond=None)
    elif condition_method in _VECTOR_METHODS:
        result.update(cond=batch_data[condition_method])
    elif condition_method in ["clusterlayout"]:
        how = pl_module.hparams.condition.clusterlayout.how
        key = {"lost": "lostbboxmask", "oracle": "segmask", "stego": "stegomask"}.get(how)
        if key is None:
            raise RuntimeError(how)
        result.update(cond=batch_data["cluster"].float().to(dev), layout=batch_data[key].float().to(dev))
    elif condition_method in ["layout"]:  # layout-only guidance: no condition vector (condition.py:60-76)
        how = pl_module.hparams.condition.layout.how
        key = {"lost": "lostbboxmask", "oracle": "segmask", "stego": "stegomask"}.get(how)
        if key is None:
            raise RuntimeError(how)
        result.update(layout=batch_data[key].float().to(dev))
    elif condition_method in ["stegoclusterlayout"]:
        result.update(cond=batch_data["stego_attr"].float().to(dev), layout=batch_data["stegomask"].float().to(dev))
    else:
        raise ValueError(condition_method)
    return result


def prepare_denoise_fn_kwargs_4sampling(pl_module, batch_data, sampling_kwargs, cond_scale):
    method = pl_module.hparams.condition_method
    if sampling_kwargs.get("random_sample_condition", False):
        if method in ("label", "cluster", "centroid", "knn_feat"):
            batch_data[method] = batch_data[method + "_random"]  # randomsample_cond, condition.py:96-118
        else:
            raise RuntimeError("random_sample_condition is not defined for condition_method=%s (condition.py:120-134)" % method)
    kw = prepare_condition_kwargs(pl_module, batch_data)
    kw.update(dict(cond_scale=cond_scale))
    kw.pop("cond_drop_prob")
    return kw
