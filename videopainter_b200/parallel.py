"""Multi-GPU layout of one denoise step (one process per GPU, torch.distributed / NCCL for the plumbing).

The step has two independent halves — the CFG batch (uncond, cond: PIPE:937-942) — and, inside each half, token-wise
work that shards freely plus one all-to-all pair around every attention (Ulysses).  Layout for N ranks:
    cfg_groups = 2 (N >= 2): ranks [0, N/2) run the uncond sample, ranks [N/2, N) the cond sample; no traffic between
                 the halves except the final noise-pred gather (2.2 MB per sample);
    sp = N / cfg_groups ranks per half share the joint (text ‖ video) sequence: rank r owns rows [r S/sp, (r+1) S/sp).
                 Before attention the QKV GEMM epilogue has already written each destination rank's heads into its own
                 contiguous send block, an all-to-all turns [S/sp, H] into [S, H/sp], attention runs on H/sp heads over
                 the whole sequence, and a second all-to-all returns the output to token sharding, where the out-projection
                 GEMM reads it straight from the receive buffer (K-chunked A operand) — no pack / unpack pass on that side.

Everything here is host-side index arithmetic plus process-group handles; it runs (and is tested) on CPU with gloo.
"""
from __future__ import annotations

import os
import threading
from dataclasses import dataclass
from typing import List, Optional, Tuple


@dataclass(frozen=True)
class Plan:
    world: int
    rank: int
    cfg_groups: int
    sp: int

    @property
    def cfg_index(self) -> int:
        return self.rank // self.sp

    @property
    def sp_rank(self) -> int:
        return self.rank % self.sp

    def sp_ranks(self) -> Tuple[int, ...]:
        """Global ranks of this rank's sequence-parallel group (same CFG half)."""
        base = self.cfg_index * self.sp
        return tuple(range(base, base + self.sp))

    def local_batch(self, global_batch: int) -> int:
        if global_batch % self.cfg_groups:
            raise ValueError(f"batch {global_batch} does not divide over {self.cfg_groups} CFG groups")
        return global_batch // self.cfg_groups

    def batch_slice(self, global_batch: int) -> slice:
        b = self.local_batch(global_batch)
        return slice(self.cfg_index * b, (self.cfg_index + 1) * b)

    def describe(self) -> str:
        if self.world == 1:
            return "single GPU"
        return f"cfg{self.cfg_groups} x ulysses{self.sp}"


def make_plan(world: int, rank: int) -> Plan:
    if world < 1 or (world & (world - 1)):
        raise ValueError("world size must be a power of two")
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    cfg_groups = 2 if world >= 2 else 1
    return Plan(world=world, rank=rank, cfg_groups=cfg_groups, sp=world // cfg_groups)


@dataclass(frozen=True)
class Shard:
    """Rows of the joint (text ‖ video) sequence owned by one sequence-parallel rank."""
    sp: int             # ranks sharing the sequence
    sp_rank: int
    seq: int            # S = text + video rows
    text: int           # text rows in the whole sequence (they come first: AP:2121)
    heads: int

    def __post_init__(self):
        if self.seq % self.sp:
            raise ValueError(f"sequence length {self.seq} does not divide over {self.sp} ranks")
        if self.heads % self.sp:
            raise ValueError(f"{self.heads} heads do not divide over {self.sp} ranks")

    @property
    def rows(self) -> int:              # rows owned by this rank
        return self.seq // self.sp

    @property
    def row0(self) -> int:              # first owned row
        return self.sp_rank * self.rows

    @property
    def text_rows(self) -> int:         # owned rows that are text tokens (all of them sit on the first ranks)
        return min(max(self.text - self.row0, 0), self.rows)

    @property
    def video_rows(self) -> int:
        return self.rows - self.text_rows

    @property
    def video0(self) -> int:            # index of the first owned video row among the video rows
        return max(self.row0 - self.text, 0)

    @property
    def heads_local(self) -> int:
        return self.heads // self.sp


class Runtime:
    """Process-group handles of one rank.  `world_group` spans every rank (final gather), `sp_group` the ranks that share
    this rank's sequence (the per-attention all-to-alls)."""

    def __init__(self, plan: Plan, sp_group=None, world_group=None, p2p: Optional[bool] = None):
        self.plan = plan
        self.sp_group = sp_group
        self.world_group = world_group
        # p2p: fuse the Ulysses all-to-alls into the producing kernels' epilogues as stores into peer memory (NVLink);
        # VP_B200_P2P=0 keeps the NCCL all-to-all path
        self.p2p = (os.environ.get("VP_B200_P2P", "1") != "0") if p2p is None else p2p
        self._shared = {}          # own PeerBuffer + the peers' IPC mappings, by local data pointer (released by release_shared)

    def shard(self, seq: int, text: int, heads: int) -> Shard:
        return Shard(self.plan.sp, self.plan.sp_rank, seq, text, heads)

    # The two collectives of the data path.  Both enqueue NCCL work ordered after the caller's current CUDA stream and
    # make that stream wait for the result; the host does not block.
    def all_to_all(self, out, inp) -> None:
        """Equal-split all-to-all over the sequence-parallel group: chunk d of `inp` goes to peer d, chunk s of `out`
        comes from peer s."""
        import torch.distributed as dist
        dist.all_to_all_single(out.view(-1), inp.view(-1), group=self.sp_group)

    def all_gather(self, out, inp) -> None:
        """`out` = concatenation over all ranks (rank order) of `inp`."""
        import torch.distributed as dist
        dist.all_gather_into_tensor(out.view(-1), inp.view(-1), group=self.world_group)


    # ---- peer memory (CUDA IPC through torch's storage sharing; plumbing only, the kernels do the stores) ----------------
    def alloc_shared(self, nbytes: int, device):
        """Collective over the sequence-parallel group: every rank allocates `nbytes` of zero-filled peer-visible memory.
        Returns (local uint8 tensor, device pointers to every rank's buffer by rank; the local entry is the tensor's)."""
        import torch.distributed as dist
        from . import ops
        try:
            buf, err = ops.PeerBuffer(nbytes, device), None
        except Exception as e:   # noqa: BLE001 - still take part in the collective below, then fail on every rank
            buf, err = None, e
        handles = [None] * self.plan.sp
        dist.all_gather_object(handles, None if buf is None else buf.handle, group=self.sp_group)
        if any(h is None for h in handles):
            raise RuntimeError(f"peer-visible allocation failed on a rank of the group ({err})")
        ptrs = [buf.ptr if r == self.plan.sp_rank else ops.peer_open(h, device) for r, h in enumerate(handles)]
        self._shared[buf.ptr] = (buf, ptrs)
        return buf.tensor, ptrs

    def release_shared(self, allocations) -> None:
        """Collective over the sequence-parallel group: undo alloc_shared for the given (tensor, ptrs) pairs — every rank first
        unmaps its peers' buffers, then (after a group barrier: nobody maps them any more) frees its own."""
        import torch
        import torch.distributed as dist
        from . import ops
        torch.cuda.synchronize()
        mine = []
        for t, _ in allocations:
            rec = self._shared.pop(t.data_ptr(), None)
            if rec is None:
                continue
            buf, ptrs = rec
            for r, ptr in enumerate(ptrs):
                if r != self.plan.sp_rank:
                    ops.peer_close(ptr)
            mine.append(buf)
        dist.barrier(group=self.sp_group)
        for buf in mine:
            buf.free()

    def ready(self) -> None:
        """Collective: everything the ranks did to the shared buffers so far (zeroing the flags) is complete everywhere."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.sp_group)

    def scatter(self, src, peer_ptrs: List[int], bytes_per_peer: int) -> None:
        """Chunk d of `src` -> slot sp_rank of rank d's buffer (include/vp_b200.h vp_peer_scatter)."""
        from . import ops
        ops.peer_scatter(src, peer_ptrs, self.plan.sp_rank, bytes_per_peer)

    def peer_barrier(self, flag_ptrs: List[int], epoch: int) -> None:
        """Device-side barrier of the sequence-parallel group on the current stream (include/vp_b200.h vp_peer_barrier)."""
        from . import ops
        ops.peer_barrier(flag_ptrs, self.plan.sp_rank, epoch)


PEER_ERR_WORD = 8             # 32-bit word of a rank's 64-byte flag buffer that counts device-barrier time-outs (elementwise.cu)

_tls = threading.local()      # per thread, so that tests can run several virtual ranks of one process side by side


def init(world: Optional[int] = None, rank: Optional[int] = None) -> Runtime:
    """Build the plan and the process groups from an initialised torch.distributed (call after init_process_group).
    Every rank must call this (new_group is collective)."""
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    plan = make_plan(world, rank)
    sp_group = None
    if world > 1:
        for g in range(plan.cfg_groups):                      # every rank creates every group, in the same order
            ranks = list(range(g * plan.sp, (g + 1) * plan.sp))
            grp = dist.new_group(ranks) if plan.sp > 1 else None
            if g == plan.cfg_index:
                sp_group = grp
    rt = Runtime(plan, sp_group, None)
    if rt.p2p and plan.sp > 1:
        rt.p2p = _probe_peer_memory(rt)
    return install(rt)


def _probe_peer_memory(rt: Runtime) -> bool:
    """Collective: can every rank of every sequence-parallel group map its peers' memory (CUDA IPC + peer access), store
    into it and pass the device-side barrier?  All ranks get the same answer; on False the data path uses the NCCL
    all-to-all instead (same kernels, same results).  A failure on ONE rank must not leave the others in a group barrier:
    every step that can fail locally is followed by a world-wide agreement (all_reduce MIN) before the next collective."""
    import warnings

    import torch
    import torch.distributed as dist
    from . import ops

    def agree(ok: bool) -> bool:
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item())

    dev = torch.device("cuda", torch.cuda.current_device())
    P, me = rt.plan.sp, rt.plan.sp_rank
    bufs, why = [], None
    try:                                                        # 1. local allocations (data [P][16 B], flags 64 B)
        bufs = [ops.PeerBuffer(16 * P, dev), ops.PeerBuffer(64, dev)]
    except Exception as e:   # noqa: BLE001
        why = e
    handles = [None] * P                                        # 2. exchange the handles (collective, cannot fail locally)
    dist.all_gather_object(handles, None if why else [b.handle for b in bufs], group=rt.sp_group)
    ptrs = [[], []]
    if why is None and all(h is not None for h in handles):     # 3. map the peers' buffers
        try:
            for i in range(2):
                ptrs[i] = [bufs[i].ptr if r == me else ops.peer_open(handles[r][i], dev) for r in range(P)]
        except Exception as e:   # noqa: BLE001
            why = e
    elif why is None:
        why = RuntimeError("a peer could not allocate")
    ok = agree(why is None)
    if ok:                                                      # 4. the real thing: peer stores, device barrier, check
        try:
            torch.cuda.synchronize()
            dist.barrier(group=rt.sp_group)                     # zero fill of every rank's buffers has completed
            src = torch.full((P, 16), me + 1, dtype=torch.uint8, device=dev)
            ops.peer_scatter(src, ptrs[0], me, 16)
            ops.peer_barrier(ptrs[1], me, 1)
            got = bufs[0].tensor.view(P, 16)[:, 0].cpu().tolist()
            errs = int(bufs[1].tensor.view(torch.int32)[PEER_ERR_WORD])
            if got != list(range(1, P + 1)) or errs:
                why = RuntimeError(f"peer stores not visible (slots {got}, barrier time-outs {errs})")
        except Exception as e:   # noqa: BLE001
            why = e
        ok = agree(why is None)
    try:                                                        # 5. release (mappings first, own memory after a barrier)
        torch.cuda.synchronize()
        for i in range(2):
            for r, ptr in enumerate(ptrs[i]):
                if r != me:
                    ops.peer_close(ptr)
    except Exception:   # noqa: BLE001
        pass
    dist.barrier(group=rt.sp_group)
    for b in bufs:
        try:
            b.free()
        except Exception:   # noqa: BLE001
            pass
    if not ok and why is not None:
        warnings.warn(f"videopainter_b200: peer memory unavailable ({why}); using the NCCL all-to-all path")
    return ok


def install(rt: Optional[Runtime]) -> Optional[Runtime]:
    """Make `rt` the runtime of the calling thread (None = single GPU)."""
    _tls.runtime = rt
    return rt


def shutdown() -> None:
    """Forget the runtime of this thread.  Call BEFORE torch.distributed.destroy_process_group(): captured CUDA graphs that
    contain NCCL collectives must be gone by then (graphs.clear_all)."""
    from . import graphs
    graphs.clear_all()
    _tls.runtime = None


def current() -> Optional[Runtime]:
    """The runtime installed by init() / install() on this thread, or None on a single GPU."""
    rt = getattr(_tls, "runtime", None)
    if rt is not None and rt.plan.world > 1:
        return rt
    return None


# ----------------------------------------------------------------------------------------------------------------------
# layout arithmetic shared with the kernels (include/vp_b200.h: vp_gemm_qkv heads_per_dest / vp_a2a_unpack_heads)
# ----------------------------------------------------------------------------------------------------------------------
def send_block_shape(sh: Shard, slots: int) -> Tuple[int, int, int, int, int]:
    """Shape of the QKV send buffer: [destination rank][slot (q, k, v, ...)][head within destination][owned row][64]."""
    return (sh.sp, slots, sh.heads_local, sh.rows, 64)


def qkv_dest_stride(sh: Shard, slots: int) -> int:
    """Elements between two destinations' blocks of the send buffer."""
    return slots * sh.heads_local * sh.rows * 64
