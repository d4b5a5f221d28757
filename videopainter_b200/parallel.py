"""Multi-GPU layout of one denoise step (one process per GPU, torch.distributed / NCCL for the plumbing).

The step has two independent halves — the CFG batch (uncond, cond: PIPE:937-942) — and, inside each half, token-wise
work that shards freely plus one all-to-all pair around every attention (Ulysses).  Layout for N ranks:
    cfg_groups = 2 (N >= 2): ranks [0, N/2) run the uncond sample, ranks [N/2, N) the cond sample, no traffic between the
                 halves except the final noise-pred exchange;
    sp = N / cfg_groups ranks per half share the sequence (Ulysses) when N >= 4.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
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

    def local_batch(self, global_batch: int) -> int:
        if global_batch % self.cfg_groups:
            raise ValueError("CFG batch must divide over the CFG groups")
        return global_batch // self.cfg_groups

    def describe(self) -> str:
        if self.world == 1:
            return "single GPU"
        return f"cfg{self.cfg_groups} x ulysses{self.sp}"


def make_plan(world: int, rank: int) -> Plan:
    if world < 1 or (world & (world - 1)):
        raise ValueError("world size must be a power of two")
    cfg_groups = 2 if world >= 2 else 1
    sp = world // cfg_groups
    if sp > 1:
        raise NotImplementedError("Ulysses sequence parallelism (sp > 1) is not wired into the engine yet")
    return Plan(world=world, rank=rank, cfg_groups=cfg_groups, sp=sp)
