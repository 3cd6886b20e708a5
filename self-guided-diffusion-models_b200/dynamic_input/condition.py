"""Condition lookup: dataset batch -> denoise kwargs (dynamic_input/condition.py:5-86,141-157).

Integer one-hot / n-hot / mask tensors are passed through or cast to float exactly as the
reference does; nothing here touches the GPU kernels (bit-exact by construction).
`pl_module` only needs `.hparams.{cond_dim,condition_method,cond_drop_prob,condition}`,
`.training` and `.device`.
"""


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
    elif condition_method in ["label", "cluster"]:
        result.update(cond=batch_data[condition_method])
    elif condition_method in ["clusterlayout"]:
        how = pl_module.hparams.condition.clusterlayout.how
        key = {"lost": "lostbboxmask", "oracle": "segmask", "stego": "stegomask"}.get(how)
        if key is None:
            raise RuntimeError(how)
        result.update(cond=batch_data["cluster"].float().to(dev), layout=batch_data[key].float().to(dev))
    elif condition_method in ["stegoclusterlayout"]:
        result.update(cond=batch_data["stego_attr"].float().to(dev), layout=batch_data["stegomask"].float().to(dev))
    else:
        raise ValueError(condition_method)
    return result


def prepare_denoise_fn_kwargs_4sampling(pl_module, batch_data, sampling_kwargs, cond_scale):
    method = pl_module.hparams.condition_method
    if sampling_kwargs.get("random_sample_condition", False):
        if method in ("label", "cluster"):
            batch_data[method] = batch_data[method + "_random"]  # condition.py:104-110
        else:
            raise RuntimeError("random_sample_condition is only defined for label / cluster")
    kw = prepare_condition_kwargs(pl_module, batch_data)
    kw.update(dict(cond_scale=cond_scale))
    kw.pop("cond_drop_prob")
    return kw
