"""Noise source of the stochastic layers.

The reference draws every eps / dropout mask from torch's global generator.  Here
the default source is the device-side Philox stream (seed, layer id, draw counter) — nothing is
materialised in HBM.  For parity tests a queue of *injected* tensors can be installed; layers then
consume them in forward order, exactly like the reference consumes its generator."""
import contextlib

import torch

_state = {"seed": None, "counter": 0, "queue": None, "next_layer_id": 1, "sample_index": None, "sample_batch": None}


def manual_seed(seed):
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state["counter"] = 0


def seed():
    if _state["seed"] is None:
        _state["seed"] = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
    return _state["seed"]


def new_layer_id():
    i = _state["next_layer_id"]
    _state["next_layer_id"] += 1
    return i


def next_draw():
    """A fresh stream_b for one draw (one forward of one layer)."""
    if _state["sample_index"] is not None:
        return int(_state["sample_index"])
    c = _state["counter"]
    _state["counter"] = (c + 1) & 0xFFFFFFFF
    return c


def pop_injected():
    q = _state["queue"]
    if q is None:
        return None
    if not q:
        raise RuntimeError("noise.inject(): queue exhausted — the model drew more noise tensors than were injected")
    return q.pop(0)


@contextlib.contextmanager
def inject(tensors):
    """with noise.inject([eps0, eps1, ...]): y = model(x)   # layers consume in forward order"""
    old = _state["queue"]
    _state["queue"] = list(tensors)
    try:
        yield
    finally:
        _state["queue"] = old


@contextlib.contextmanager
def sample_index(idx):
    """Pin stream_b to a GLOBAL Monte-Carlo sample index: draws become independent of sharding."""
    old = _state["sample_index"]
    _state["sample_index"] = idx
    try:
        yield
    finally:
        _state["sample_index"] = old


@contextlib.contextmanager
def sample_batch(n_samples, sample0, batch, act_bits=8):
    """Run `n_samples` Monte-Carlo samples (global indices sample0 .. sample0+n_samples-1) through ONE forward of an int8
    model: every int8 layer draws n_samples weight tensors (Philox stream_b = global sample index, as under
    `sample_index`) and contracts all samples in one launch; activations carry the samples in their leading dimension
    ([n_samples*batch, ...]; a [batch, ...] input is shared by all samples).  `act_bits` is the width the model clamps its
    activations to (lets the layers pick the tcgen05 kind::i8 kernel, which needs 7-bit inputs)."""
    if _state["queue"] is not None:
        raise RuntimeError("noise.sample_batch(): injected noise is per forward, not per sample batch")
    old = _state["sample_batch"]
    _state["sample_batch"] = (int(n_samples), int(sample0), int(batch), int(act_bits))
    try:
        yield
    finally:
        _state["sample_batch"] = old


def sample_batch_state():
    return _state["sample_batch"]



# ---- per-batch draw offset, kept on the device ---------------------------------------------------------------------------
# The Monte-Carlo engines key Philox by (seed, layer, GLOBAL sample index).  Replaying a captured CUDA graph would hand every
# batch the same S draws; the reference redraws for every batch (experiments/utils.py:342-347).  The samplers therefore add a
# device-side scalar to the sample index (include/qbn.h: qbn_set_sample_base): `set_draw_offset(k * S)` before batch k gives
# batch k the sample indices k*S .. k*S+S-1 with ONE graph, no re-capture and no host synchronisation.
_draw_base = {}


def draw_base(device):
    """The int32 device scalar the samplers read (created and registered with the library on first use)."""
    from . import _lib
    import ctypes
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    t = _draw_base.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", key))
        _draw_base[key] = t
    _lib.call("qbn_set_sample_base", ctypes.c_void_p(t.data_ptr()))
    return t


def set_draw_offset(offset, device="cuda"):
    """All following sampler launches on `device` (eager or graph replays) draw sample indices `offset + ...`."""
    draw_base(device).fill_(int(offset) & 0x7FFFFFFF)
