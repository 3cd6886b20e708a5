"""Functional fp32 restatement of the reference eps-model (TEST INFRASTRUCTURE).

State-dict in, tensor out; no nn.Module.  Every function cites the reference
lines it restates (paths relative to /root/reference).

`cfg` is a plain dict:
    kind                 'unet_fast' (openaimodel.UNetModel) | 'unetca_fast' (openaimodel_ca.UNetModel)
    image_size, in_channels, out_channels, model_channels, num_res_blocks,
    channel_mult, attention_resolutions, num_heads, resblock_updown,
    cond_dim, condition_method, layout_dim, context_dim, cond_token_num, scale_type

`emu` (optional) switches on bf16-operand emulation at the points where the
B200 kernels round to bf16 (GEMM/conv/attention operands); accumulation stays
fp32.  It exists to forecast the kernels' error budget on CPU — with emu=None
the function is the plain fp32 reference algorithm.
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- helpers
def _emu_kind(emu, kind):
    if not emu:
        return None
    if isinstance(emu, dict):
        return emu.get(kind)
    return "bf16" if emu is True else emu


def _r(x, emu, kind="a"):
    """Round to the kernels' operand precision and back when emulating it.  `emu`: None (plain fp32 reference),
    'bf16' / 'f16' (every operand), or a dict per operand kind {'a': activations, 'w': weights, 'h1': the
    ResBlock-internal conv output the kernels keep in 16 bits, 'x': the image fed to the first conv, 'q': the
    q / k / v tensors as the attention kernel reads them, 'p': attention probabilities, 'o': the attention output as stored}; 'f16x2' = hi + lo
    split (two fp16 values, ~22 bits)."""
    d = _emu_kind(emu, kind)
    if d is None:
        return x
    if d == "bf16":
        return x.bfloat16().float()
    if d == "f16":
        return x.half().float()
    if d == "f16x2":
        hi = x.half().float()
        return hi + (x - hi).half().float()
    raise ValueError(d)


def timestep_embedding(timesteps, dim, max_period=10000):
    """dynamic/diffusionmodules/util.py:151-171 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(
        -math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half
    )
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, p, x):
    """GroupNorm32(32, C): util.py:199-216 — 32 groups, eps 1e-5, fp32."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _conv(sd, p, x, emu, stride=1, padding=1, akind="a"):
    return F.conv2d(_r(x, emu, akind), _r(sd[p + ".weight"], emu, "w"), sd[p + ".bias"], stride=stride, padding=padding)


def _lin(sd, p, x, emu=False, bias=True):
    return F.linear(_r(x, emu), _r(sd[p + ".weight"], emu, "w"), sd[p + ".bias"] if bias else None)


def _mlp2(sd, p, x):
    """Sequential(Linear, SiLU, Linear) with children 0 and 2 (openaimodel.py:570-574)."""
    return _lin(sd, p + ".2", F.silu(_lin(sd, p + ".0", x)))


# --------------------------------------------------------------------------- topology
def topology(cfg):
    """Block list restating the constructors (openaimodel.py:634-835,
    openaimodel_ca.py:645-836).  Returns (input_blocks, middle, output_blocks);
    each block is a list of layer tuples:
        ('conv', prefix)                      first 3x3 conv
        ('res', prefix, cin, cout, up, down)  ResBlock
        ('attn', prefix, ch)                  AttentionBlock / Attention_LR
        ('down', prefix, ch) ('up', prefix, ch)   strided conv / nearest+conv
    """
    mc = cfg["model_channels"]
    ca = cfg["kind"] == "unetca_fast"
    updown = bool(cfg.get("resblock_updown", False))
    attn_res = list(cfg["attention_resolutions"])
    mult = list(cfg["channel_mult"])
    nrb = cfg["num_res_blocks"]

    inp = [[("conv", "input_blocks.0.0")]]
    chans = [mc]
    ch, ds = mc, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            i = len(inp)
            layers = [("res", f"input_blocks.{i}.0", ch, m * mc, False, False)]
            ch = m * mc
            if ds in attn_res:
                layers.append(("attn", f"input_blocks.{i}.1", ch))
            inp.append(layers)
            chans.append(ch)
        if level != len(mult) - 1:
            i = len(inp)
            if updown:
                inp.append([("res", f"input_blocks.{i}.0", ch, ch, False, True)])
            else:
                inp.append([("down", f"input_blocks.{i}.0", ch)])
            chans.append(ch)
            ds *= 2
    mid = [
        ("res", "middle_block.0", ch, ch, False, False),
        ("attn", "middle_block.1", ch),
        ("res", "middle_block.2", ch, ch, False, False),
    ]
    out = []
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            o = len(out)
            layers = [("res", f"output_blocks.{o}.0", ch + ich, mc * m, False, False)]
            ch = mc * m
            if ds in attn_res:
                layers.append(("attn", f"output_blocks.{o}.{len(layers)}", ch))
            if level and i == nrb:
                if updown:
                    layers.append(("res", f"output_blocks.{o}.{len(layers)}", ch, ch, True, False))
                else:
                    layers.append(("up", f"output_blocks.{o}.{len(layers)}", ch))
                ds //= 2
            out.append(layers)
    return inp, mid, out


# --------------------------------------------------------------------------- blocks
def resblock(sd, p, x, emb, up, down, emu):
    """ResBlock._forward, use_scale_shift_norm=True (openaimodel.py:300-320)."""
    h = F.silu(_gn(sd, p + ".in_layers.0", x))
    if up:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    elif down:
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = _r(_conv(sd, p + ".in_layers.2", h, emu), emu, "h1")  # the kernels keep h1 in the 16-bit operand type
    emb_out = _lin(sd, p + ".emb_layers.1", F.silu(emb), emu)[..., None, None]
    scale, shift = torch.chunk(emb_out, 2, dim=1)
    h = _gn(sd, p + ".out_layers.0", h) * (1 + scale) + shift
    h = _conv(sd, p + ".out_layers.3", F.silu(h), emu)  # dropout is identity in eval
    if (p + ".skip_connection.weight") in sd:
        x = _conv(sd, p + ".skip_connection", x, emu, padding=0)
    return x + h


def attention_block(sd, p, x, heads, emu):
    """AttentionBlock._forward + QKVAttentionLegacy (openaimodel.py:365-371,403-420)."""
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_r(_gn(sd, p + ".norm", xf), emu), _r(sd[p + ".qkv.weight"], emu, "w"), sd[p + ".qkv.bias"])
    qkv = _r(qkv, emu, "q")
    ch = c // heads
    q, k, v = qkv.reshape(b * heads, ch * 3, -1).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", _r(w, emu, "p"), v).reshape(b, -1, hh * ww)
    h = F.conv1d(_r(a, emu, "o"), _r(sd[p + ".proj_out.weight"], emu, "w"), sd[p + ".proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def _ln(x, gamma, beta):
    return F.layer_norm(x, x.shape[-1:], gamma, beta)


def attention_lr(sd, p, x, context, heads, emu):
    """Attention_LR.forward (dynamic/crossattetion_lr.py:81-142): multi-query
    attention over [context(16) | null(1) | self(HW)] keys, single shared k/v head."""
    b, c, hh, ww = x.shape
    xt = x.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    xn = _ln(xt, sd[p + ".norm.gamma"], sd[p + ".norm.beta"])
    q = F.linear(_r(xn, emu), _r(sd[p + ".to_q.weight"], emu, "w"))
    kv = F.linear(_r(xn, emu), _r(sd[p + ".to_kv.weight"], emu, "w"))
    k, v = kv.chunk(2, dim=-1)
    d = q.shape[-1] // heads
    q = q.reshape(b, hh * ww, heads, d).permute(0, 2, 1, 3) * d**-0.5
    nk, nv = sd[p + ".null_kv"][0], sd[p + ".null_kv"][1]
    k = torch.cat((nk.expand(b, 1, d), k), dim=-2)
    v = torch.cat((nv.expand(b, 1, d), v), dim=-2)
    ctx = _ln(context, sd[p + ".to_context.0.weight"], sd[p + ".to_context.0.bias"])
    ckv = F.linear(ctx, sd[p + ".to_context.1.weight"], sd[p + ".to_context.1.bias"])
    ck, cv = ckv.chunk(2, dim=-1)
    k = torch.cat((ck, k), dim=-2)
    v = torch.cat((cv, v), dim=-2)
    sim = torch.einsum("bhid,bjd->bhij", _r(q, emu, "q"), _r(k, emu, "q"))
    attn = sim.softmax(dim=-1)
    out = torch.einsum("bhij,bjd->bhid", _r(attn, emu, "p"), _r(v, emu, "q"))
    out = out.permute(0, 2, 1, 3).reshape(b, hh * ww, heads * d)
    out = F.linear(_r(out, emu, "o"), _r(sd[p + ".to_out.0.weight"], emu, "w"))
    out = _ln(out, sd[p + ".to_out.1.gamma"], sd[p + ".to_out.1.beta"])
    return (xt + out).reshape(b, hh, ww, c).permute(0, 3, 1, 2)


def _run_block(sd, cfg, layers, h, emb, context, emu):
    heads = cfg["num_heads"]
    ca = cfg["kind"] == "unetca_fast"
    for layer in layers:
        kind, p = layer[0], layer[1]
        if kind == "conv":  # the first conv: the kernels feed the image as a hi + lo pair (kind "x")
            h = _conv(sd, p, h, emu, akind="x")
        elif kind == "res":
            h = resblock(sd, p, h, emb, layer[4], layer[5], emu)
        elif kind == "attn":
            h = attention_lr(sd, p, h, context, heads, emu) if ca else attention_block(sd, p, h, heads, emu)
        elif kind == "down":  # Downsample(use_conv): conv3x3 stride 2 (openaimodel_ca.py:159-181)
            h = _conv(sd, p + ".op", h, emu, stride=2)
        elif kind == "up":  # Upsample(use_conv): nearest 2x + conv3x3 (openaimodel_ca.py:101-131)
            h = _conv(sd, p + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"), emu)
        else:
            raise ValueError(kind)
    return h


# --------------------------------------------------------------------------- forward
def param_shapes(cfg):
    """(name, shape) inventory of the reference module's state_dict for `cfg`, restating the constructors'
    registration (openaimodel.py:566-835, openaimodel_ca.py:540-836, crossattetion_lr.py:56-79): lets the CPU
    baseline build seeded weights without instantiating anything else.  Checked against the reference-generated
    inventories stored in tests/golden/unet_*.npz (tests/test_oracle_golden.py)."""
    mc, ted = cfg["model_channels"], 4 * cfg["model_channels"]
    cd, L, H = cfg["cond_dim"], cfg.get("layout_dim", 0) or 0, cfg["image_size"]
    ca = cfg["kind"] == "unetca_fast"
    out = []
    add = lambda n, *shape: out.append((n, tuple(shape)))
    has_cond = cd > 0 if not ca else cfg["cond_token_num"] > 0
    if has_cond:
        add("null_cond_emb", cfg["cond_token_num"] if ca and cfg["cond_token_num"] > 1 else 1, cd)
    if L > 0:
        add("null_layout_emb", 1, 1, H, H)
    add("time_embed.0.weight", ted, mc); add("time_embed.0.bias", ted)
    add("time_embed.2.weight", ted, ted); add("time_embed.2.bias", ted)
    if not ca:
        if cd > 0:
            add("mlp_cond.0.weight", ted // 2, cd); add("mlp_cond.0.bias", ted // 2)
            add("mlp_cond.2.weight", ted // 2, ted // 2); add("mlp_cond.2.bias", ted // 2)
        E = ted + (ted // 2 if cd > 0 else 0)
    else:
        ctx = cfg["context_dim"]
        add("norm_cond.weight", ctx); add("norm_cond.bias", ctx)
        add("to_time_tokens.0.weight", mc, mc); add("to_time_tokens.0.bias", mc)
        add("to_time_tokens.2.weight", ctx * 8, mc); add("to_time_tokens.2.bias", ctx * 8)
        E = ted
    if ca and has_cond:
        ctx = cfg["context_dim"]
        add("cond_mlp.0.weight", ted, cd); add("cond_mlp.0.bias", ted)
        add("cond_mlp.2.weight", ted, ted); add("cond_mlp.2.bias", ted)
        add("to_cond_tokens.0.weight", ctx * 8, cd); add("to_cond_tokens.0.bias", ctx * 8)
        mid_d = int(math.sqrt(ctx * cd))  # to_cond_tokens_2d: built for every cond_token_num > 0, used only when > 1
        add("to_cond_tokens_2d.0.weight", mid_d, cd); add("to_cond_tokens_2d.0.bias", mid_d)
        add("to_cond_tokens_2d.2.weight", mid_d, mid_d); add("to_cond_tokens_2d.2.bias", mid_d)
        add("to_cond_tokens_2d.4.weight", mid_d, mid_d); add("to_cond_tokens_2d.4.bias", mid_d)
        add("to_cond_tokens_2d.6.weight", ctx, mid_d); add("to_cond_tokens_2d.6.bias", ctx)
        E = ted
    heads = cfg["num_heads"]
    inp, mid, outb = topology(cfg)
    for layers in inp + [mid] + outb:
        for layer in layers:
            kind, p = layer[0], layer[1]
            if kind == "conv":
                add(p + ".weight", mc, cfg["in_channels"] + L, 3, 3); add(p + ".bias", mc)
            elif kind == "res":
                cin, cout = layer[2], layer[3]
                add(p + ".in_layers.0.weight", cin); add(p + ".in_layers.0.bias", cin)
                add(p + ".in_layers.2.weight", cout, cin, 3, 3); add(p + ".in_layers.2.bias", cout)
                add(p + ".emb_layers.1.weight", 2 * cout, E); add(p + ".emb_layers.1.bias", 2 * cout)
                add(p + ".out_layers.0.weight", cout); add(p + ".out_layers.0.bias", cout)
                add(p + ".out_layers.3.weight", cout, cout, 3, 3); add(p + ".out_layers.3.bias", cout)
                if cin != cout:
                    add(p + ".skip_connection.weight", cout, cin, 1, 1); add(p + ".skip_connection.bias", cout)
            elif kind == "attn" and not ca:
                ch = layer[2]
                add(p + ".norm.weight", ch); add(p + ".norm.bias", ch)
                add(p + ".qkv.weight", 3 * ch, ch, 1); add(p + ".qkv.bias", 3 * ch)
                add(p + ".proj_out.weight", ch, ch, 1); add(p + ".proj_out.bias", ch)
            elif kind == "attn":
                ch, dh, ctx = layer[2], layer[2] // heads, cfg["context_dim"]
                add(p + ".null_kv", 2, dh)
                add(p + ".norm.gamma", ch); add(p + ".norm.beta", ch)
                add(p + ".to_q.weight", dh * heads, ch); add(p + ".to_kv.weight", 2 * dh, ch)
                add(p + ".to_context.0.weight", ctx); add(p + ".to_context.0.bias", ctx)
                add(p + ".to_context.1.weight", 2 * dh, ctx); add(p + ".to_context.1.bias", 2 * dh)
                add(p + ".to_out.0.weight", ch, dh * heads)
                add(p + ".to_out.1.gamma", ch); add(p + ".to_out.1.beta", ch)
            elif kind in ("down", "up"):
                ch, q = layer[2], (".op" if kind == "down" else ".conv")
                add(p + q + ".weight", ch, ch, 3, 3); add(p + q + ".bias", ch)
    add("out.0.weight", mc); add("out.0.bias", mc)
    add("out.2.weight", cfg["out_channels"], mc, 3, 3); add("out.2.bias", cfg["out_channels"])
    return out


def unet_forward(sd, cfg, x, timesteps, cond=None, layout=None, drop_mask=None, emu=None):
    """UNetModel.forward (openaimodel.py:904-956 / openaimodel_ca.py:917-1033).

    `drop_mask` [B] bool replaces prob_mask_like(cond_drop_prob): the sampling
    path only ever uses p in {0,1}, where the mask is deterministic
    (openaimodel.py:462-463,926-928).
    """
    mc = cfg["model_channels"]
    method = cfg.get("condition_method")
    b = x.shape[0]
    if drop_mask is None:
        drop_mask = torch.zeros(b, dtype=torch.bool)
    t_emb = timestep_embedding(timesteps, mc)
    emb = _mlp2(sd, "time_embed", t_emb)
    context = None
    if cfg["kind"] == "unet_fast":
        if cfg["cond_dim"] > 0:
            cond = cond.to(sd["null_cond_emb"].dtype)
            cond_masked = torch.where(drop_mask[:, None], sd["null_cond_emb"], cond)
            if method == "clusterlayout":
                lm = torch.where(drop_mask[:, None, None, None], sd["null_layout_emb"], layout)
                x = torch.cat((x, lm), dim=1)
            emb = torch.cat((emb, _mlp2(sd, "mlp_cond", cond_masked)), dim=-1)
    elif cfg["cond_token_num"] == 0:
        # no condition vector: the context is the time tokens, only the layout is guided (openaimodel_ca.py:944-958)
        tt = _lin(sd, "to_time_tokens.2", F.silu(_lin(sd, "to_time_tokens.0", t_emb)))
        context = tt.reshape(b, 8, cfg["context_dim"])
        if method == "layout":
            lm = torch.where(drop_mask[:, None, None, None], sd["null_layout_emb"], layout)
            x = torch.cat((x, lm), dim=1)
        context = F.layer_norm(context, context.shape[-1:], sd["norm_cond.weight"], sd["norm_cond.bias"])
    elif cfg["cond_token_num"] > 1:
        # token condition [B, N, cond_dim] (openaimodel_ca.py:988-1012): masked against null_cond_emb [N, cond_dim],
        # every token through the to_cond_tokens_2d MLP, pooled (CLS token or mean) into cond_mlp
        assert cond.dim() == 3
        tt = _lin(sd, "to_time_tokens.2", F.silu(_lin(sd, "to_time_tokens.0", t_emb)))
        time_tokens = tt.reshape(b, 8, cfg["context_dim"])
        cond_masked = torch.where(drop_mask[:, None, None], sd["null_cond_emb"], cond.float())
        h2 = cond_masked
        for i in (0, 2, 4):
            h2 = F.silu(_lin(sd, f"to_cond_tokens_2d.{i}", h2))
        cond_tokens = _lin(sd, "to_cond_tokens_2d.6", h2)
        context = torch.cat([time_tokens, cond_tokens], 1)
        pooled = cond_masked[:, 0, :] if cfg.get("use_cls_token_as_pooled", True) == True else cond_masked.mean(dim=1)  # noqa: E712
        emb = emb + _mlp2(sd, "cond_mlp", pooled)
        context = F.layer_norm(context, context.shape[-1:], sd["norm_cond.weight"], sd["norm_cond.bias"])
    else:
        assert cfg["cond_token_num"] == 1 and cond.dim() == 2  # openaimodel_ca.py:960-961
        tt = _lin(sd, "to_time_tokens.2", F.silu(_lin(sd, "to_time_tokens.0", t_emb)))
        time_tokens = tt.reshape(b, 8, cfg["context_dim"])
        cond_masked = torch.where(drop_mask[:, None], sd["null_cond_emb"], cond.float())
        cond_tokens = _lin(sd, "to_cond_tokens.0", cond_masked).reshape(b, 8, cfg["context_dim"])
        context = torch.cat([time_tokens, cond_tokens], 1)
        emb = emb + _mlp2(sd, "cond_mlp", cond_masked)
        if method in ("clusterlayout", "stegoclusterlayout"):
            lm = torch.where(drop_mask[:, None, None, None], sd["null_layout_emb"], layout)
            x = torch.cat((x, lm), dim=1)
        context = F.layer_norm(context, context.shape[-1:], sd["norm_cond.weight"], sd["norm_cond.bias"])

    inp, mid, out = topology(cfg)
    hs = []
    h = x.float()
    for layers in inp:
        h = _run_block(sd, cfg, layers, h, emb, context, emu)
        hs.append(h)
    h = _run_block(sd, cfg, mid, h, emb, context, emu)
    for layers in out:
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, cfg, layers, h, emb, context, emu)
    h = F.silu(_gn(sd, "out.0", h))
    return _conv(sd, "out.2", h, emu)


def guided_score(cfg, eps_u, eps_c, w):
    """get_guided_score (openaimodel.py:853-859)."""
    if cfg.get("scale_type", "imagen") == "imagen":
        return (1 - w) * eps_u + w * eps_c
    if cfg["scale_type"] == "cfg":
        return (1 + w) * eps_c - w * eps_u
    raise ValueError(cfg["scale_type"])


def forward_with_cond_scale(sd, cfg, x, t, cond_scale, cond=None, layout=None, emu=None, return_pair=False):
    """forward_with_cond_scale (openaimodel.py:861-902; openaimodel_ca.py:879-915).

    unet_fast short-circuits on int OR float 1/0; unetca_fast only on int
    (openaimodel_ca.py:882,890) — a float 1.0 there takes the doubled path."""
    b = x.shape[0]
    is_num = isinstance(cond_scale, (int, float)) and not isinstance(cond_scale, bool)
    if cfg["kind"] == "unetca_fast":
        is_num = isinstance(cond_scale, int) and not isinstance(cond_scale, bool)
    if is_num and cond_scale == 1:
        return unet_forward(sd, cfg, x, t, cond, layout, torch.zeros(b, dtype=torch.bool), emu)
    if is_num and cond_scale == 0:
        return unet_forward(sd, cfg, x, t, cond, layout, torch.ones(b, dtype=torch.bool), emu)
    dbl = lambda a: None if a is None else torch.cat((a, a), 0)
    mask = torch.cat((torch.zeros(b, dtype=torch.bool), torch.ones(b, dtype=torch.bool)))
    eps = unet_forward(sd, cfg, dbl(x), dbl(t), dbl(cond), dbl(layout), mask, emu)
    eps_c, eps_u = torch.chunk(eps, 2, dim=0)
    if return_pair:
        return eps_c, eps_u
    return guided_score(cfg, eps_u, eps_c, cond_scale)
