"""Drop-in for `dynamic.diffusionmodules.openaimodel_ca.UNetModel` (config `unetca_fast`).

Select it with `dynamic.target=sgdm_b200.dynamic.diffusionmodules.openaimodel_ca.UNetModel`;
constructor kwargs, parameter names/shapes, `forward` and `forward_with_cond_scale`
follow the reference (openaimodel_ca.py:449-1033).  All compute runs in libsgdm_b200.so.
"""
from ._unet_base import EngineUNet
from ... import _lib


class UNetModel(EngineUNet):
    _KIND = _lib.KIND_UNETCA_FAST
    _FLOAT_SHORTCUT = False  # isinstance(cond_scale, int) only (openaimodel_ca.py:882,890)

    def __init__(
        self,
        image_size,
        in_channels,
        model_channels,
        out_channels,
        num_res_blocks,
        attention_resolutions,
        dropout=0,
        channel_mult=(1, 2, 4, 8),
        conv_resample=True,
        dims=2,
        num_classes=None,
        use_checkpoint=False,
        use_fp16=False,
        num_heads=-1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
        use_ca_block=False,
        transformer_depth=1,
        context_dim=None,
        n_embed=None,
        legacy=True,
        cond_token_num=0,
        cond_dim=None,
        use_cls_token_as_pooled=None,
        condition=None,
        condition_method=None,
        precision=None,  # sgdm_b200 only: 'fp16' (default) | 'fp16x3' (EngineUNet._build)
    ):
        super().__init__()
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        assert cond_token_num >= 0
        assert isinstance(cond_dim, int)
        unsupported = []
        if dims != 2: unsupported.append("dims != 2")
        if not use_scale_shift_norm: unsupported.append("use_scale_shift_norm=False")
        if not use_ca_block: unsupported.append("use_ca_block=False")
        if context_dim is None: unsupported.append("context_dim=None")
        if num_head_channels != -1: unsupported.append("num_head_channels")
        if resblock_updown: unsupported.append("resblock_updown")
        if not conv_resample: unsupported.append("conv_resample=False")
        if use_fp16: unsupported.append("use_fp16")
        if unsupported:
            raise NotImplementedError(
                "sgdm_b200 unetca_fast covers config/dynamic/unetca_fast.yaml with the README overrides "
                "(any cond_token_num, context_dim=32); not built: " + ", ".join(unsupported))
        if cond_token_num == 0:
            assert cond_dim == 0  # openaimodel_ca.py:562-564: no condition vector (the `layout`-only / unconditional model)
            if condition_method == "clusterlayout":
                raise NotImplementedError  # openaimodel_ca.py:947-948
        if cond_token_num > 1 and condition_method in ("clusterlayout", "stegoclusterlayout", "layout"):
            # openaimodel_ca.py:988-1012 concatenates no layout on this branch (and raises for clusterlayout), while the
            # constructor still widens the first conv: the reference cannot run these combinations either
            raise NotImplementedError("cond_token_num > 1 takes a [B, N, cond_dim] token condition and no layout")
        layout_dim = 0
        if condition_method in ["layout"]:
            layout_dim = condition.layout.layout_dim  # openaimodel_ca.py:634-641
        if condition_method in ["clusterlayout"]:
            layout_dim = condition.clusterlayout.layout_dim  # openaimodel_ca.py:617-624
        if condition_method in ["stegoclusterlayout"]:
            layout_dim = condition.stegoclusterlayout.layout_dim  # :625-632
        self.dropout = dropout
        self.num_heads = num_heads
        self.cond_token_num = cond_token_num
        self.context_dim = context_dim
        self.use_cls_token_as_pooled = use_cls_token_as_pooled
        self._build(
            dict(image_size=image_size, in_channels=in_channels, out_channels=out_channels,
                 model_channels=model_channels, num_res_blocks=num_res_blocks, channel_mult=channel_mult,
                 attention_resolutions=attention_resolutions, num_heads=num_heads, resblock_updown=0,
                 cond_dim=cond_dim, layout_dim=layout_dim, context_dim=context_dim,
                 cond_token_num=cond_token_num,
                 use_cls_token_as_pooled=int(use_cls_token_as_pooled is True or use_cls_token_as_pooled == True)),
            condition, condition_method, precision)

    def forward(self, x, timesteps=None, cond_drop_prob=0.0, cond=None, layout=None):
        if self.cond_token_num == 1:
            assert cond is not None and len(cond.shape) == 2  # openaimodel_ca.py:960-961
        elif self.cond_token_num > 1:
            assert cond is not None and len(cond.shape) == 3  # [B, T, C], openaimodel_ca.py:989
        return self._forward_impl(x, timesteps, cond, layout, cond_drop_prob)

    def forward_with_cond_scale(self, x, t, cond_scale, cond=None, layout=None):
        return self._forward_with_cond_scale_impl(x, t, cond_scale, cond, layout, None)
