"""Drop-in for `dynamic.diffusionmodules.openaimodel.UNetModel` (config `unet_fast`).

Select it with `dynamic.target=sgdm_b200.dynamic.diffusionmodules.openaimodel.UNetModel`;
constructor kwargs, parameter names/shapes, `forward` and `forward_with_cond_scale`
follow the reference (openaimodel.py:466-956).  All compute runs in libsgdm_b200.so.
"""
from ._unet_base import EngineUNet
from ... import _lib


class UNetModel(EngineUNet):
    _KIND = _lib.KIND_UNET_FAST
    _FLOAT_SHORTCUT = True  # is_number(): int or float (openaimodel.py:28-34,868,876)

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
        use_checkpoint=False,
        use_fp16=False,
        num_heads=-1,
        num_head_channels=-1,
        num_heads_upsample=-1,
        use_scale_shift_norm=False,
        resblock_updown=False,
        use_new_attention_order=False,
        use_spatial_transformer=False,
        transformer_depth=1,
        context_dim=None,
        legacy=True,
        cond_dim=None,
        condition=None,
        condition_method=None,
        precision=None,  # sgdm_b200 only: 'fp16' (default) | 'fp16x3' (EngineUNet._build)
    ):
        super().__init__()
        if num_heads == -1:
            assert num_head_channels != -1, "Either num_heads or num_head_channels has to be set"
        unsupported = []
        if dims != 2: unsupported.append("dims != 2")
        if not use_scale_shift_norm: unsupported.append("use_scale_shift_norm=False")
        if use_spatial_transformer or context_dim is not None: unsupported.append("use_spatial_transformer")
        if num_head_channels != -1: unsupported.append("num_head_channels")
        if use_new_attention_order: unsupported.append("use_new_attention_order")
        if not conv_resample: unsupported.append("conv_resample=False")
        if use_fp16: unsupported.append("use_fp16")
        if num_heads_upsample not in (-1, num_heads): unsupported.append("num_heads_upsample")
        if condition_method == "cluster_lookup": unsupported.append("condition_method=cluster_lookup")
        if unsupported:
            raise NotImplementedError(
                "sgdm_b200 unet_fast covers the configurations of config/dynamic/unet_fast.yaml; "
                "not built: " + ", ".join(unsupported))
        cond_dim = 0 if cond_dim is None else cond_dim
        layout_dim = 0
        if condition_method in ["clusterlayout"]:
            layout_dim = condition.clusterlayout.layout_dim  # openaimodel.py:623-630
        self.dropout = dropout  # identity in eval; the sampling path never trains
        self.num_heads = num_heads
        self.use_checkpoint = use_checkpoint
        self._build(
            dict(image_size=image_size, in_channels=in_channels, out_channels=out_channels,
                 model_channels=model_channels, num_res_blocks=num_res_blocks, channel_mult=channel_mult,
                 attention_resolutions=attention_resolutions, num_heads=num_heads,
                 resblock_updown=int(bool(resblock_updown)), cond_dim=cond_dim, layout_dim=layout_dim,
                 context_dim=0, cond_token_num=0),
            condition, condition_method, precision)

    def forward(self, x, timesteps=None, cond=None, layout=None, cond_drop_prob=0.0, image_batch_ids=None):
        return self._forward_impl(x, timesteps, cond, layout, cond_drop_prob)

    def forward_with_cond_scale(self, x, t, cond_scale, cond, layout=None, p0=None, image_batch_ids=None):
        return self._forward_with_cond_scale_impl(x, t, cond_scale, cond, layout, p0)
