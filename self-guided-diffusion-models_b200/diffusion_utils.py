"""Small host helpers with the reference's names (diffusion_utils/util.py:70-100,85-92,254-268)."""
import importlib

import torch

from . import _lib


class dict2obj(object):
    """Nested dict -> attribute object (diffusion_utils/util.py:85-92)."""

    def __init__(self, d):
        for a, b in d.items():
            if isinstance(b, (list, tuple)):
                setattr(self, a, [dict2obj(x) if isinstance(x, dict) else x for x in b])
            else:
                setattr(self, a, dict2obj(b) if isinstance(b, dict) else b)


def instantiate_from_config(config):
    """{'target': 'pkg.Class', 'params': {...}} -> object (diffusion_utils/util.py:254-268):
    the reference's plugin mechanism; pointing `target` at sgdm_b200 classes swaps them in."""
    assert "target" in config
    module, cls = config["target"].rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)(**config.get("params", dict()))


def clip_unnormalize_to_zero_to_255(img, clip=True):
    """((img+1)*127.5).clamp(0,255).to(uint8) (diffusion_utils/util.py:99-100) as one kernel."""
    _lib.require_cuda(img, "img")
    src = img.detach().float().contiguous()
    out = torch.empty(src.shape, dtype=torch.uint8, device=src.device)
    _lib.check(_lib.lib().sgdm_to_uint8(_lib.current_stream(src.device), src.data_ptr(), out.data_ptr(), src.numel()))
    return out
