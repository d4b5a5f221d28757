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
    def __init__(self, plan: parallel.Plan, fabric: Fabric, p2p: bool = True):
        super().__init__(plan, p2p=p2p)
        self.fabric = fabric
        self._n_share = 0

    # peer memory: the virtual ranks live in one process on one device, so a "peer pointer" is a plain device pointer
    def alloc_shared(self, nbytes, device):
        f, pl = self.fabric, self.plan
        key = ("share", self._n_share)
        self._n_share += 1
        t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        f.box[(key, pl.rank)] = t
        f.barrier.wait()
        ptrs = [f.box[(key, r)].data_ptr() for r in pl.sp_ranks()]
        f.barrier.wait()
        return t, ptrs

    def ready(self):
        self._sync()
        self.fabric.barrier.wait()

    def release_shared(self, allocations):
        self._sync()
        self.fabric.barrier.wait()

    def peer_barrier(self, flag_ptrs, epoch):
        # kernels of different virtual ranks share one GPU: they must not spin on each other, so the barrier is host-side
        self._sync()
        self.fabric.barrier.wait()

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


def run_virtual_ranks(world: int, fn, p2p: bool = True):
    """fn(rank, runtime) on `world` threads; returns the list of results, re-raises the first failure."""
    fab = Fabric(world)
    res, err = [None] * world, [None] * world

    def body(r):
        try:
            rt = ThreadRuntime(parallel.make_plan(world, r), fab, p2p=p2p)
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
