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


class SampleHandoff:
    """Asynchronous device -> host hand-off of generated uint8 samples for the callers that write PNGs / feed FID
    (eval/eval_fid.py:567-573 loops over `gen_samples` on the GPU and copies every image separately: one blocking
    D2H per image).  `push(samples)` lays the batch out as HWC (what PIL wants), starts ONE asynchronous copy into a
    pinned ring slot on a side stream and returns a ticket; the sampler can start the next batch at once.
    `wait(ticket)` blocks on that copy's event only and returns a numpy view [B, H, W, C] of the pinned slot (valid
    until `depth` further pushes).

        handoff = SampleHandoff(depth=2)
        ticket = None
        for batch in loader:
            samples, _ = diffusion.p_sample_loop(...)           # uint8 [B, 3, H, W] on the GPU
            nxt = handoff.push(samples)
            if ticket is not None:
                for img in handoff.wait(ticket): Image.fromarray(img).save(...)   # overlaps the next trajectory
            ticket = nxt
    """

    def __init__(self, depth=2):
        self.depth = max(1, int(depth))
        self._slots = [None] * self.depth
        self._events = [None] * self.depth
        self._shapes = [None] * self.depth
        self._k = 0
        self._copy_stream = None

    def push(self, samples):
        _lib.require_cuda(samples, "samples")
        if samples.dtype != torch.uint8 or samples.dim() != 4:
            raise ValueError(f"expected uint8 [B, C, H, W] samples, got {samples.dtype} {tuple(samples.shape)}")
        dev = samples.device
        with torch.cuda.device(dev):
            if self._copy_stream is None or self._copy_stream.device != dev:
                self._copy_stream = torch.cuda.Stream(dev)
            hwc = samples.permute(0, 2, 3, 1).contiguous()  # on the producer's stream
            slot = self._k % self.depth
            if self._events[slot] is not None:
                self._events[slot].synchronize()  # the copy that last used this slot (normally long done)
            if self._slots[slot] is None or self._slots[slot].numel() < hwc.numel():
                self._slots[slot] = torch.empty(hwc.numel(), dtype=torch.uint8, pin_memory=True)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            self._copy_stream.wait_event(ready)
            with torch.cuda.stream(self._copy_stream):
                dst = self._slots[slot][: hwc.numel()].view(hwc.shape)
                dst.copy_(hwc, non_blocking=True)
                hwc.record_stream(self._copy_stream)
                done = torch.cuda.Event()
                done.record(self._copy_stream)
            self._events[slot] = done
            self._shapes[slot] = tuple(hwc.shape)
        self._k += 1
        return self._k - 1

    def wait(self, ticket):
        if ticket < self._k - self.depth or ticket >= self._k:
            raise ValueError(f"ticket {ticket} is no longer (or not yet) in the ring of depth {self.depth}")
        slot = ticket % self.depth
        self._events[slot].synchronize()
        n = 1
        for d in self._shapes[slot]:
            n *= d
        return self._slots[slot][:n].view(self._shapes[slot]).numpy()
