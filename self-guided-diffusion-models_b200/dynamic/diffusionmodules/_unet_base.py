"""Shared host-side machinery of the two drop-in UNet modules.

The nn.Module built here owns the parameters (same names and shapes as the reference
module's state_dict, so `load_state_dict` of a reference checkpoint works and
`named_parameters()` / EMA / optimizers see the usual tree) and forwards every compute
call to the C-ABI engine.  The engine's packed 16-bit weight copies are a CACHE keyed on
each tensor's (data_ptr, _version): in-place updates such as `ema_scope`
(reference dynamic/ema.py:46-53) or an optimizer step are picked up on the next call.
"""
import ctypes as C
import math
import weakref

import torch
import torch.nn as nn

from ... import _lib


def _is_number(x, allow_float):
    if isinstance(x, bool):
        return False
    if isinstance(x, int):
        return True
    return allow_float and isinstance(x, float)


def _destroy_engine(h):
    try:
        _lib.lib().sgdm_destroy(h)
    except Exception:
        pass


class _Node(nn.Module):
    """Anonymous container: lets parameters live at dotted paths like
    `input_blocks.3.0.in_layers.2.weight` without re-creating the layer classes."""

    _root = None  # weakref to the owning EngineUNet (set by _attach)

    def register_parameter(self, name, param):
        super().register_parameter(name, param)
        root = self._root() if self._root is not None else None
        if root is not None:
            root._tensor_list = None  # a Parameter OBJECT was (re)placed: the cached tensor list is stale

    def register_buffer(self, name, tensor, persistent=True):
        super().register_buffer(name, tensor, persistent=persistent)
        root = self._root() if self._root is not None else None
        if root is not None:
            root._tensor_list = None


def _attach(root, dotted, tensor, kind):
    *path, leaf = dotted.split(".")
    mod = root
    for part in path:
        nxt = mod._modules.get(part)
        if nxt is None:
            nxt = _Node()
            object.__setattr__(nxt, "_root", weakref.ref(root))
            mod.add_module(part, nxt)
        mod = nxt
    if kind == "buffer":
        mod.register_buffer(leaf, tensor)
    else:
        mod.register_parameter(leaf, nn.Parameter(tensor, requires_grad=(kind == "param")))


def _init_tensor(name, shape):
    """Reference initialisation (PyTorch defaults of nn.Conv2d/nn.Linear/GroupNorm/LayerNorm,
    zero_module for ResBlock.out_layers[3], AttentionBlock.proj_out and UNetModel.out[2] —
    openaimodel.py:273-276,357,833-834; zeros for the null embeddings :582-600,624-626;
    randn for Attention_LR.null_kv crossattetion_lr.py:69)."""
    leaf = name.rsplit(".", 1)[-1]
    if name.endswith("null_kv"):
        return torch.randn(shape)
    if "null_cond_emb" in name or "null_layout_emb" in name:
        return torch.zeros(shape)
    if ".out_layers.3." in name or ".proj_out." in name or name.startswith("out.2."):
        return torch.zeros(shape)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return (torch.rand(shape) * 2 - 1) / math.sqrt(fan_in)
    if leaf == "gamma" or leaf == "beta":
        return torch.ones(shape) if leaf == "gamma" else torch.zeros(shape)
    if leaf == "weight":
        return torch.ones(shape)  # GroupNorm / LayerNorm scale
    return None  # bias: decided by the caller (needs the fan_in of its weight)


class EngineUNet(nn.Module):
    """Base of the drop-in UNetModel classes. Subclasses set `_KIND` and parse kwargs."""

    _KIND = None
    _FLOAT_SHORTCUT = True  # unet_fast: is_number accepts float; unetca_fast: int only

    def _build(self, cfg_fields, condition, condition_method, precision=None):
        """`precision` (not a reference kwarg; optional): 'fp16' = 16-bit GEMM operands, the throughput path;
        'fp16x3' = split-precision operands (~fp32 products, ~3x the tensor work) for deterministic samplers on
        ill-conditioned networks.  Default: $SGDM_PRECISION, else 'fp16'."""
        import os

        self.condition = condition
        self.condition_method = condition_method
        precision = precision or os.environ.get("SGDM_PRECISION") or "fp16"
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}, got {precision!r}")
        self.precision = precision
        c = _lib.SgdmConfig()
        c.precision = _lib.PRECISIONS[precision]
        for k, v in cfg_fields.items():
            if k in ("channel_mult", "attention_resolutions"):
                vals = [int(a) for a in v]
                if len(vals) > 8:
                    raise ValueError(f"{k}: at most 8 entries")
                setattr(c, "n_" + k, len(vals))
                arr = getattr(c, k)
                for i, a in enumerate(vals):
                    arr[i] = a
            else:
                setattr(c, k, int(v))
        c.kind = self._KIND
        self._cfg = c
        lib = _lib.lib()
        h = C.c_void_p()
        _lib.check(lib.sgdm_create(C.byref(c), C.byref(h)))
        self._h = h
        weakref.finalize(self, _destroy_engine, h)
        dims = (C.c_int64 * _lib.MAX_DIMS)()
        nd = C.c_int()
        inventory = []
        for i in range(lib.sgdm_param_count(h)):
            name = lib.sgdm_param_name(h, i).decode()
            _lib.check(lib.sgdm_param_shape(h, i, dims, C.byref(nd)))
            inventory.append((name, tuple(dims[k] for k in range(nd.value))))
        self._inventory = inventory
        shapes = dict(inventory)
        for name, shape in inventory:
            t = _init_tensor(name, shape)
            if t is None:  # bias of a conv / linear / norm
                wname = name[: -len("bias")] + "weight"
                wshape = shapes.get(wname, ())
                if len(wshape) >= 2:
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    t = (torch.rand(shape) * 2 - 1) / math.sqrt(fan_in)
                    if ".out_layers.3." in name or ".proj_out." in name or name.startswith("out.2."):
                        t = torch.zeros(shape)
                else:
                    t = torch.zeros(shape)
            if name.endswith(".beta"):
                kind = "buffer"  # crossattetion_lr.LayerNorm registers beta as a buffer
            elif "null_cond_emb" in name or "null_layout_emb" in name:
                kind = "frozen"  # nn.Parameter(..., requires_grad=False)
            else:
                kind = "param"
            _attach(self, name, t, kind)
        self._loaded_sig = {}
        self._tensor_list = None
        self._fp_tables, self._fp_last = None, None
        self._freqs_set = False
        self._engine_device = None
        self.image_size = c.image_size
        self.in_channels = c.in_channels
        self.out_channels = c.out_channels
        self.model_channels = c.model_channels
        self.cond_dim = c.cond_dim
        self.dtype = torch.float32

    # ------------------------------------------------------------------ weights
    def invalidate_weight_cache(self):
        """Forget the packed copies: the next call re-packs every tensor (the blunt tool; `refresh_weights`
        finds out which tensors actually changed)."""
        self._loaded_sig = {}
        self._fp_last = None

    # Writes that bypass autograd's version counter (`param.data.copy_(...)`, as the reference's `ema_scope` does,
    # dynamic/ema.py:46-53) are invisible to the (data_ptr, _version) signature.  They are caught by a DEVICE-SIDE
    # FINGERPRINT: one kernel hashes the bits of every parameter tensor (sgdm_fingerprint), 8 bytes per tensor come
    # back, and exactly the tensors whose hash moved are re-packed.  Cost: one ~50 us launch + one small read-back
    # instead of ~300 pack launches, so it runs at the start of EVERY sampler trajectory and (unless
    # `check_weights = False`) in front of every standalone forward call.
    check_weights = True

    def _fingerprint(self):
        lib = _lib.lib()
        sd = self._tensors()
        tensors = [sd[name] for name, _ in self._inventory]
        key = tuple(t.data_ptr() for t in tensors)
        dev = tensors[0].device
        if self._fp_tables is None or self._fp_tables[0] != key:
            for name, t in zip((n for n, _ in self._inventory), tensors):
                _lib.require_cuda(t, f"parameter {name}")
                if t.dtype != torch.float32 or not t.is_contiguous():
                    return None  # exotic storage: fall back to re-packing everything
            ptrs = torch.tensor(key, dtype=torch.int64, device=dev)
            numel = torch.tensor([t.numel() for t in tensors], dtype=torch.int64, device=dev)
            out = torch.zeros(len(tensors), dtype=torch.int64, device=dev)
            self._fp_tables = (key, ptrs, numel, out)
        _, ptrs, numel, out = self._fp_tables
        with torch.cuda.device(dev):
            _lib.check(lib.sgdm_fingerprint(_lib.current_stream(dev), ptrs.data_ptr(), numel.data_ptr(), len(tensors),
                                            out.data_ptr()))
        return out.cpu()  # 8 bytes per tensor; synchronises the stream

    def refresh_weights(self):
        """Bring the engine's packed copies in line with the parameters, whatever way they were written: tensors
        whose (data_ptr, _version) changed are re-packed as always, and so are tensors whose device-side hash moved
        since they were packed (`.data` writes)."""
        fp = self._fingerprint()
        if fp is None:
            self.invalidate_weight_cache()
        elif self._fp_last is not None and len(self._fp_last) == len(fp):
            for i in torch.nonzero(fp != self._fp_last).flatten().tolist():
                self._loaded_sig.pop(self._inventory[i][0], None)
        else:
            self._loaded_sig = {}
        self.sync_weights()
        self._fp_last = fp

    def _tensors(self):
        """{name: tensor} of the inventory.  The module-tree walk (named_parameters over ~300 tensors) is cached: the
        Parameter / buffer OBJECTS only change through register_parameter / register_buffer (hooked in _Node and
        below), while `.to()`, load_state_dict, optimizer steps and EMA swaps update them in place."""
        if self._tensor_list is None:
            sd = dict(self.named_parameters())
            sd.update(dict(self.named_buffers()))
            self._tensor_list = {name: sd[name] for name, _ in self._inventory}
        return self._tensor_list

    def register_parameter(self, name, param):
        super().register_parameter(name, param)
        self.__dict__["_tensor_list"] = None

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._tensor_list = None  # conversions may replace the Parameter objects (torch.__future__ overwrite mode)
        return out

    def sync_weights(self, force=False):
        """(Re)pack every tensor whose storage or version changed since the last call."""
        lib = _lib.lib()
        sd = self._tensors()
        first = sd[self._inventory[0][0]]
        _lib.require_cuda(first, f"parameter {self._inventory[0][0]}")
        with torch.cuda.device(first.device):  # the engine allocates on the CURRENT device: make it the parameters'
            self._sync_weights_on_device(lib, sd, force)

    def _sync_weights_on_device(self, lib, sd, force):
        stream = None
        for name, shape in self._inventory:
            t = sd[name]
            sig = (t.data_ptr(), t._version)
            if not force and self._loaded_sig.get(name) == sig:
                continue
            _lib.require_cuda(t, f"parameter {name}")
            if self._engine_device is None:
                self._engine_device = t.device  # the engine binds to the device of its first parameter
            elif t.device != self._engine_device:
                raise _lib.SgdmError(f"parameter {name} is on {t.device}; the engine was set up on {self._engine_device} "
                                     "(build a new module to change devices)")
            if stream is None:
                stream = _lib.current_stream(t.device)
            src = t.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            dims = (C.c_int64 * len(shape))(*shape)
            _lib.check(lib.sgdm_load_param(self._h, name.encode(), src.data_ptr(), dims, len(shape), stream))
            self._loaded_sig[name] = sig
        if not self._freqs_set:
            half = self.model_channels // 2
            # util.py:160-163, evaluated by torch on the host exactly as the reference does
            freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
            freqs = freqs.contiguous()
            _lib.check(lib.sgdm_set_timestep_freqs(self._h, freqs.data_ptr(), half))
            self._freqs_set = True

    # ------------------------------------------------------------------ helpers
    def _check_x(self, x):
        """The engine's plan is laid out for the configured geometry (it indexes x with its own C/H/W and writes
        B*out_channels*image_size^2 floats): anything else must be an error here, not an out-of-bounds access.
        (The reference module is fully convolutional; this drop-in supports the configured image_size only.)"""
        _lib.require_cuda(x, "x")
        if x.dim() != 4 or x.shape[1] != self.in_channels or x.shape[2] != self.image_size or x.shape[3] != self.image_size:
            raise ValueError(f"x must be [B, {self.in_channels}, {self.image_size}, {self.image_size}] "
                             f"(the engine is built for image_size={self.image_size}), got {tuple(x.shape)}")
        dev = self._engine_device
        if dev is not None and x.device != dev:
            raise ValueError(f"x is on {x.device} but the engine's weights and workspace live on {dev}")

    def _prep_inputs(self, x, t, cond, layout):
        self._check_x(x)
        x = x.detach().float().contiguous()
        t = t.detach().to(device=x.device, dtype=torch.int64).contiguous()
        if self.cond_dim > 0:
            if cond is None:
                raise ValueError("cond is required")
            cond = cond.detach().to(x.device).to(torch.float32).contiguous()  # openaimodel.py:911
            ntok = int(self._cfg.cond_token_num) if self._KIND == _lib.KIND_UNETCA_FAST else 1
            want = (x.shape[0], ntok, self.cond_dim) if ntok > 1 else (x.shape[0], self.cond_dim)
            if tuple(cond.shape) != want:
                raise AssertionError(f"wtf? {tuple(cond.shape)}, expected {want}")
        else:
            cond = None
        if self._cfg.layout_dim > 0:
            if layout is None:
                raise ValueError("layout is required for condition_method=%s" % self.condition_method)
            layout = layout.detach().to(x.device).to(torch.float32)
            L = self._cfg.layout_dim
            if layout.dim() != 4 or layout.shape[1] != L or tuple(layout.shape[2:]) != tuple(x.shape[2:]) \
                    or layout.shape[0] not in (1, x.shape[0]):
                raise ValueError(f"layout must be [B, {L}, {x.shape[2]}, {x.shape[3]}], got {tuple(layout.shape)}")
            layout = layout.expand(x.shape[0], L, x.shape[2], x.shape[3]).contiguous()
        else:
            layout = None
        return x, t, cond, layout

    def _raw_forward(self, x, t, cond, layout, drop_mask):
        """eps for an explicit boolean drop mask [B] (True = null embeddings)."""
        self.refresh_weights() if self.check_weights else self.sync_weights()
        x, t, cond, layout = self._prep_inputs(x, t, cond, layout)
        B = x.shape[0]
        if B == 0:  # empty batch: the reference's torch ops return an empty tensor
            return torch.empty((0, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        drop = drop_mask.to(device=x.device, dtype=torch.uint8).contiguous()
        out = torch.empty((B, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sgdm_forward(self._h, _lib.current_stream(x.device), _lib.ptr(x), _lib.ptr(t),
                                               _lib.ptr(cond), _lib.ptr(layout), _lib.ptr(drop), B, _lib.ptr(out)))
        return out

    def guided_pair_ptrs(self, x, t, cond, layout):
        """Batched cond||uncond pass; returns raw device pointers (eps_c, eps_u) into engine
        memory, valid until the next forward.  Inputs must already be prepared tensors."""
        pc, pu = C.c_void_p(), C.c_void_p()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sgdm_forward_guided(self._h, _lib.current_stream(x.device), _lib.ptr(x), _lib.ptr(t),
                                                      _lib.ptr(cond), _lib.ptr(layout), x.shape[0], C.byref(pc),
                                                      C.byref(pu)))
        return pc.value, pu.value

    def _scale_type(self):
        st = getattr(self.condition, "scale_type", "imagen") if self.condition is not None else "imagen"
        if st not in _lib.SCALE_TYPES:
            raise ValueError(st)  # openaimodel.py:859
        return _lib.SCALE_TYPES[st]

    def get_guided_score(self, z, zc, w):
        """openaimodel.py:853-859 on already-computed scores (kept for API parity)."""
        st = getattr(self.condition, "scale_type", "imagen")
        if st == "imagen":
            return (1 - w) * z + w * zc
        elif st == "cfg":
            return (1 + w) * zc - w * z
        raise ValueError(st)

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def _forward_impl(self, x, timesteps, cond, layout, cond_drop_prob):
        if isinstance(cond_drop_prob, (float, int)):
            cond_drop_prob = torch.full((len(x),), cond_drop_prob, dtype=torch.float, device=x.device)
        assert isinstance(cond_drop_prob, torch.Tensor)
        # prob_mask_like (openaimodel.py:462-463): same draw, same RNG consumption as the reference
        mask = torch.zeros((len(x),), device=x.device).float().uniform_(0, 1) < cond_drop_prob.to(x.device)
        eps = self._raw_forward(x, timesteps, cond, layout, mask)
        return eps, 0.0, dict()

    @torch.no_grad()
    def _forward_with_cond_scale_impl(self, x, t, cond_scale, cond, layout, p0):
        B = x.shape[0]
        if p0 is None and _is_number(cond_scale, self._FLOAT_SHORTCUT) and cond_scale == 1:
            return self._raw_forward(x, t, cond, layout, torch.zeros(B, dtype=torch.bool))
        if _is_number(cond_scale, self._FLOAT_SHORTCUT) and cond_scale == 0:
            return self._raw_forward(x, t, cond, layout, torch.ones(B, dtype=torch.bool))
        if p0 is not None:
            # explicit per-sample keep/drop probabilities for the "conditional" half (chainvis,
            # ddim_plms_sampler.py:170-177): generic doubled-batch path
            dbl = lambda a: None if a is None else torch.cat((a, a), 0)
            p = torch.cat((p0.to(x.device).float(), torch.ones(B, device=x.device)), 0)
            mask = torch.zeros((2 * B,), device=x.device).float().uniform_(0, 1) < p
            eps = self._raw_forward(dbl(x), dbl(t), dbl(cond), dbl(layout), mask)
            eps_c, eps_u = torch.chunk(eps, 2, dim=0)
            return self.get_guided_score(z=eps_u, zc=eps_c, w=cond_scale)
        self.refresh_weights() if self.check_weights else self.sync_weights()
        x, t, cond, layout = self._prep_inputs(x, t, cond, layout)
        if B == 0:
            return torch.empty((0, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        pc, pu = self.guided_pair_ptrs(x, t, cond, layout)
        out = torch.empty((B, self.out_channels, x.shape[2], x.shape[3]), device=x.device, dtype=torch.float32)
        w_ptr, w = None, 0.0
        if torch.is_tensor(cond_scale):
            wt = cond_scale.detach().to(device=x.device, dtype=torch.float32).reshape(-1)
            if wt.numel() == 1:
                wt = wt.expand(B)
            assert wt.numel() == B, "tensor cond_scale must have one entry per sample"
            wt = wt.contiguous()
            w_ptr = wt.data_ptr()
        else:
            w = float(cond_scale)
        per_sample = out.shape[1:].numel()
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().sgdm_mix(_lib.current_stream(x.device), pc, pu, w, w_ptr, self._scale_type(),
                                           _lib.ptr(out), B, per_sample))
        return out

    def convert_to_fp16(self):  # API parity (openaimodel.py:837-851); operand precision is fixed by the engine
        pass

    def convert_to_fp32(self):
        pass
