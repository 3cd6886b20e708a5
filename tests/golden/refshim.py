"""Import shims that let the UNMODIFIED reference (/root/reference) be imported
in the authoring container.  Used ONLY by make_golden.py —
never by the product, the oracle, the -m gpu tests, smoke() or bench.py
(/root/reference does not exist on the GPU box).

Missing third-party modules are stubbed with the minimum surface the hot-path
files touch:
  * einops_exts.{repeat_many,rearrange_many,check_shape}  (crossattetion_lr.py:13,95)
  * pytorch_lightning                                      (ddpm_sampler.py:12, unused)
  * matplotlib / matplotlib.pyplot                         (taokit/wandb_utils.py:9)
  * eval.papervis_utils, eval.test_exps.common_stuff       (ddim_plms_sampler.py:20-21)
  * omegaconf.listconfig                                   (openaimodel.py:536, only if context_dim set)
"""
import sys
import types

REFERENCE_ROOT = "/root/reference"


def install():
    from einops import rearrange, repeat

    def _mod(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    if "einops_exts" not in sys.modules:
        _mod(
            "einops_exts",
            repeat_many=lambda ts, p, **k: tuple(repeat(t, p, **k) for t in ts),
            rearrange_many=lambda ts, p, **k: tuple(rearrange(t, p, **k) for t in ts),
            check_shape=lambda *a, **k: None,
        )
    for name in ("pytorch_lightning", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            _mod(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    _flag = lambda o, s: hasattr(o, s) and getattr(o, s)
    ev = _mod("eval")
    ev.__path__ = []
    _mod("eval.papervis_utils", batch_to_conditioninterp_papervis=None)
    te = _mod("eval.test_exps")
    te.__path__ = []
    _mod("eval.test_exps.common_stuff", should_vis=_flag, should_exp=_flag)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    from loguru import logger

    logger.remove()  # the reference logs every block at construction


def condition_obj():
    from diffusion_utils.util import dict2obj

    return dict2obj(
        {
            "scale_type": "imagen",
            "clusterlayout": {"layout_dim": 1},
            "stegoclusterlayout": {"layout_dim": 27},
            "layout": {"layout_dim": 21},
        }
    )
