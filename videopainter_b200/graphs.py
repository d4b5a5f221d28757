"""Whole-forward CUDA graphs ("CUDA-graph the whole step"; PIPE:937-1000 is the loop they serve).

A denoising step of the reference pipeline is two module calls (`self.branch(...)`, `self.transformer(...)`) with Python in
between, so the unit that can be captured behind the reference's API is one forward: ~25 launches for the branch, ~310 for
the backbone (plus the device-side peer barriers and scatters under sequence parallelism).  With graphs enabled
(`VP_B200_GRAPH=1` or `videopainter_b200.enable_graphs()`), the first call with a given signature runs eagerly (it also
warms every cache: packed weights, workspace, tensor maps, RoPE tables), the second is captured, later ones are replayed:

  * small / medium inputs (latents, text states, timestep, masks, branch samples) are copied into the graph's static
    inputs on the caller's stream — 0.3 ms per step for the 2 x 216 MB of branch samples at the full size;
  * large per-window constants (the 42 previous-window hidden states, 9.2 GB; the RoPE tables) are captured BY ADDRESS: the
    signature contains their data pointers, the entry keeps them alive, a new window is a new signature;
  * results are cloned out of the graph's memory pool (fresh tensors, as in eager mode) except the `hidden_states_list`
    (42 x [B, S, D], only handed out when the caller asks for it): its entries are views of the graph's own arena and stay
    valid until the same signature runs again.  The pipeline keeps that list from the LAST step of a window and reads it
    during the next window (PIPE:982-988), whose calls carry `prev_hidden_states` and therefore have another signature;
  * nothing in a captured launch depends on the call history: the peer barrier keeps its epoch on the device
    (vp_peer_barrier epoch 0), tensor maps are kernel parameters, the time-out check of the barriers stays outside.

At most MAX_GRAPHS signatures per module are kept (least recently used first out)."""
import atexit
import gc
import os
import weakref
from collections import OrderedDict
from typing import Any, Callable, Dict, List, Optional

import torch

from . import ops

MAX_GRAPHS = 2
_enabled = os.environ.get("VP_B200_GRAPH", "0") == "1"


def enable_graphs(flag: bool = True) -> None:
    global _enabled
    _enabled = bool(flag)


def graphs_enabled() -> bool:
    return _enabled


class _Entry:
    __slots__ = ("graph", "static", "keep", "result", "launches")


def _sig(t: Optional[torch.Tensor]):
    return None if t is None else (tuple(t.shape), t.dtype, t.device.index)


_registry: List["weakref.ref"] = []        # packed models that hold captured graphs


def clear(pm) -> None:
    """Drop every captured graph of a packed model (their memory pools are released with them)."""
    pm.__dict__.pop("_graphs", None)
    pm.__dict__.pop("_graph_seen", None)


def clear_all() -> None:
    """Drop every captured graph of the process.  REQUIRED before torch.distributed.destroy_process_group(): NCCL does not
    finalise a communicator while CUDA graphs that captured its collectives are alive (the process hangs at exit instead);
    parallel.shutdown() and an atexit hook call this."""
    live = [r() for r in _registry]
    _registry.clear()
    if any(pm is not None and pm.__dict__.get("_graphs") for pm in live):
        torch.cuda.synchronize()
    for pm in live:
        if pm is not None:
            clear(pm)
    gc.collect()


atexit.register(clear_all)


def _guard_process_group_teardown() -> None:
    """torch.distributed.destroy_process_group() with live graphs hangs in NCCL: make it drop them first."""
    import torch.distributed as dist
    if not dist.is_available() or getattr(dist.destroy_process_group, "_vp_b200_guard", False):
        return
    original = dist.destroy_process_group

    def destroy_process_group(*args, **kwargs):
        clear_all()
        return original(*args, **kwargs)

    destroy_process_group._vp_b200_guard = True
    destroy_process_group.__doc__ = original.__doc__
    dist.destroy_process_group = destroy_process_group


def stats(pm) -> Dict[str, int]:
    g = pm.__dict__.get("_graphs") or {}
    return {"graphs": len(g), "replays": pm.__dict__.get("_graph_replays", 0), "captures": pm.__dict__.get("_graph_captures", 0)}


def run(pm, kind: str, fn: Callable[[Dict[str, Any], Dict[str, Any]], Any], copied: Dict[str, Optional[torch.Tensor]],
        pinned: Dict[str, Optional[torch.Tensor]], flags) -> Any:
    """`fn(copied, pinned)` eagerly the first time a signature is seen, captured the second time, replayed afterwards.
    Returns fn's result: on a replay these are the tensors of the capture (the graph's memory) — the caller clones what it
    hands on."""
    key = (kind, flags, tuple((k, _sig(v)) for k, v in copied.items()),
           tuple((k, None if v is None else (v.data_ptr(), tuple(v.shape), v.dtype, v._version if k.startswith("rope") else 0))
                 for k, v in pinned.items()))
    graphs: "OrderedDict[Any, _Entry]" = pm.__dict__.setdefault("_graphs", OrderedDict())
    seen = pm.__dict__.setdefault("_graph_seen", OrderedDict())
    ent = graphs.get(key)
    if ent is None:
        if key not in seen:                                   # first sight: eager (warms every cache the capture relies on)
            seen[key] = True
            while len(seen) > 8:
                seen.popitem(last=False)
            return fn(copied, pinned)
        ent = _Entry()
        ent.static = {k: (None if v is None else v.detach().clone()) for k, v in copied.items()}
        ent.keep = [v for v in pinned.values() if v is not None]
        for ws in pm.workspace.values():
            ws.poll_check()
        torch.cuda.synchronize()
        ent.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count
        with torch.cuda.graph(ent.graph):
            ent.result = fn(ent.static, pinned)
        ent.launches = ops.launch_count - n0
        ops.launch_count = n0                                 # nothing has run yet: the replay below counts them
        while len(graphs) >= MAX_GRAPHS:
            graphs.popitem(last=False)
        graphs[key] = ent
        if not any(r() is pm for r in _registry):
            _registry.append(weakref.ref(pm))
            _guard_process_group_teardown()
        pm.__dict__["_graph_captures"] = pm.__dict__.get("_graph_captures", 0) + 1
    else:
        graphs.move_to_end(key)
        for k, v in copied.items():
            if v is not None:
                ent.static[k].copy_(v, non_blocking=True)
    for ws in pm.workspace.values():
        ws.poll_check()
    ent.graph.replay()
    for ws in pm.workspace.values():
        ws.post_check()
    ops.launch_count += ent.launches
    pm.__dict__["_graph_replays"] = pm.__dict__.get("_graph_replays", 0) + 1
    return ent.result
