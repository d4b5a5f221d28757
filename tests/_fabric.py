"""Test infrastructure: several virtual ranks of ONE process (threads) exchanging tensors through host-side rendezvous.
Lets the `-m gpu` suite exercise the whole sequence-parallel data path (destination-major QKV epilogue, receive-side unpack,
K-chunked out-projection, final gather, CFG split) on a single GPU; the NCCL collectives themselves are covered by
tests/test_gpu_multi.py on a multi-GPU box and by the gloo test on CPU."""
import threading

import torch

from videopainter_b200 import parallel


class Fabric:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.box = {}


class ThreadRuntime(parallel.Runtime):
    def __init__(self, plan: parallel.Plan, fabric: Fabric):
        super().__init__(plan)
        self.fabric = fabric

    def _sync(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def all_to_all(self, out, inp):
        f, pl = self.fabric, self.plan
        self._sync()
        f.box[("a2a", pl.rank)] = inp
        f.barrier.wait()
        o = out.view(pl.sp, -1)
        for s, src in enumerate(pl.sp_ranks()):
            o[s].copy_(f.box[("a2a", src)].view(pl.sp, -1)[pl.sp_rank])
        self._sync()
        f.barrier.wait()

    def all_gather(self, out, inp):
        f, pl = self.fabric, self.plan
        self._sync()
        f.box[("ag", pl.rank)] = inp
        f.barrier.wait()
        o = out.view(pl.world, -1)
        for r in range(pl.world):
            o[r].copy_(f.box[("ag", r)].reshape(-1))
        self._sync()
        f.barrier.wait()


def run_virtual_ranks(world: int, fn):
    """fn(rank, runtime) on `world` threads; returns the list of results, re-raises the first failure."""
    fab = Fabric(world)
    res, err = [None] * world, [None] * world

    def body(r):
        try:
            rt = ThreadRuntime(parallel.make_plan(world, r), fab)
            parallel.install(rt)
            if torch.cuda.is_available():
                torch.cuda.set_device(0)
            res[r] = fn(r, rt)
        except BaseException as e:   # noqa: BLE001
            err[r] = e
            fab.barrier.abort()
        finally:
            parallel.shutdown()

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for e in err:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in err:
        if e is not None:
            raise e
    return res
