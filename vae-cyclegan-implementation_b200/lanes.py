"""Two execution lanes per device: independent network passes of one training step run side by side.

A composite step applies its generators in two independent chains -- x -> G -> F and y -> F -> G in the
cycle models (Networks.py:1909-1914 of the reference), G(x) and G(y) in the paired GANs (:1023-1028) --
and the reference executes them one after the other on one stream.  Here each chain gets its own CUDA
stream ("lane") and half of the SMs for its persistent tensor-core kernels (vcg_set_sm_budget): layers
whose tile count cannot fill 148 SMs (16x16 maps at a per-GPU batch of 8: 32 CTA-pair tiles) run next
to each other instead of in sequence, and the memory-bound transform passes of one chain overlap the
GEMMs of the other.

The backward pass needs no extra wiring: autograd runs every backward node on the stream its forward
ran on and orders cross-stream gradients itself; `dual` only forks the lanes from the calling stream
on entry and joins them on exit, so the region is also valid inside a CUDA-graph capture."""
from __future__ import annotations

import torch

from . import lib as L

_ENABLED = [True]
_LANES = {}
_DEPTH = [0]
_DIRTY = set()        # lane streams that ran plan work since the last join


def set_enabled(flag):
    """Switch lane concurrency off (every pass on the calling stream) or on; returns the previous setting."""
    prev, _ENABLED[0] = _ENABLED[0], bool(flag)
    return prev


def enabled():
    return _ENABLED[0]


def _get(device):
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    s = _LANES.get(device)
    if s is None:
        s = _LANES[device] = (torch.cuda.Stream(device, priority=-1), torch.cuda.Stream(device, priority=-1))
    return s


def note_stream(stream):
    """plan code ran on `stream` (called from PlanFunction): if it is a lane, the end-of-backward flush joins it"""
    for lanes in _LANES.values():
        if stream in lanes:
            _DIRTY.add(stream)


def join_dirty(device=None):
    """make the current stream wait for every lane that ran work since the last join"""
    if _DIRTY:
        cur = torch.cuda.current_stream(device)
        for s in list(_DIRTY):
            if s.device == cur.device:
                cur.wait_stream(s)
                _DIRTY.discard(s)


_WGRAD = {}           # chain stream handle -> low-priority stream for that chain's weight-gradient GEMMs


def wgrad_stream(chain):
    """Side stream for the weight-gradient GEMMs of the backward pass running on `chain` (plan.run_backward): they are
    off the critical path gather -> norm -> data gradient -> next layer, so they run at lower stream priority beside it."""
    key = (chain.device, chain.cuda_stream)
    s = _WGRAD.get(key)
    if s is None:
        s = _WGRAD[key] = torch.cuda.Stream(chain.device, priority=0)
    return s


class _Same:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class dual:
    """`with dual(device, active) as d:` ... `with d.lane(i):` runs the enclosed passes on lane i (0 or 1).
    With active=False (or lanes disabled) everything stays on the calling stream."""

    def __init__(self, device, active=True):
        self.active = bool(active) and _ENABLED[0]
        self.device = device

    def __enter__(self):
        if self.active:
            self.cur = torch.cuda.current_stream(self.device)
            self.lanes = _get(self.cur.device)
            for s in self.lanes:
                s.wait_stream(self.cur)
            _DEPTH[0] += 1
            if _DEPTH[0] == 1:
                sms = torch.cuda.get_device_properties(self.cur.device).multi_processor_count
                L.check(L.load().vcg_set_sm_budget(sms // 2 // 2 * 2), "vcg_set_sm_budget")
        return self

    def lane(self, i):
        if not self.active:
            return _Same()
        return torch.cuda.stream(self.lanes[i])

    def __exit__(self, *exc):
        if self.active:
            _DEPTH[0] -= 1
            if _DEPTH[0] == 0:
                L.check(L.load().vcg_set_sm_budget(0), "vcg_set_sm_budget")
            for s in self.lanes:
                self.cur.wait_stream(s)
                _DIRTY.discard(s)
        return False
