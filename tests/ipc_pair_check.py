"""Worker of tests/test_gpu_parallel.py::test_two_processes_on_one_gpu_exchange_through_ipc_and_the_device_barrier.

TWO PROCESSES ON ONE GPU: each allocates peer-visible buffers through the C ABI (vp_peer_alloc), maps the other's through CUDA
IPC (vp_peer_open) and runs the data-path kernels of the sequence-parallel exchange against them — vp_peer_scatter (stores
into the peer's buffer) and vp_peer_barrier in its device-counted mode (epoch 0), eagerly and replayed from a CUDA graph.
The GPU time-slices the two contexts, so a barrier completes once both kernels have had a slice; the barrier's own time-out
bounds a failure.  gloo is used for the rendezvous only (NCCL refuses two ranks on one device)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from videopainter_b200 import ops
    ops.peer_set_timeout_ms(15000)
    n = 1 << 16                                                 # bf16 elements per (source, destination) chunk
    data = ops.PeerBuffer(world * n * 2, dev)                   # slot s: what rank s sent to this rank
    flags = ops.PeerBuffer(64, dev)
    handles = [None] * world
    dist.all_gather_object(handles, (data.handle, flags.handle))
    p_data = [data.ptr if r == rank else ops.peer_open(h[0], dev) for r, h in enumerate(handles)]
    p_flag = [flags.ptr if r == rank else ops.peer_open(h[1], dev) for r, h in enumerate(handles)]
    torch.cuda.synchronize()
    dist.barrier()

    src = torch.empty(world, n, dtype=torch.bfloat16, device=dev)
    mine = data.tensor.view(torch.bfloat16).view(world, n)

    def fill(it):
        for d in range(world):
            src[d].fill_(float(16 * rank + 4 * d + it % 4))     # exactly representable; encodes (source, destination, iteration)

    def exchange():
        ops.peer_scatter(src, p_data, rank, n * 2)
        ops.peer_barrier(p_flag, rank, 0)                       # everybody's stores have landed

    def check(it):
        torch.cuda.synchronize()
        for s in range(world):
            want = float(16 * s + 4 * rank + it % 4)
            got = mine[s].float()
            assert float(got.min()) == want and float(got.max()) == want, (rank, s, it, want, float(got.min()), float(got.max()))

    it = 0
    for _ in range(3):                                          # launched from Python
        fill(it)
        exchange()
        check(it)
        ops.peer_barrier(p_flag, rank, 0)                       # everybody has read: the slots may be overwritten
        it += 1
    g = torch.cuda.CUDAGraph()                                  # the same two launches replayed from a CUDA graph
    torch.cuda.synchronize()
    dist.barrier()
    with torch.cuda.graph(g):
        exchange()
    for _ in range(3):
        fill(it)
        g.replay()
        check(it)
        ops.peer_barrier(p_flag, rank, 0)
        it += 1
    torch.cuda.synchronize()
    words = flags.tensor.view(torch.int32)
    assert int(words[8]) == 0, f"rank {rank}: {int(words[8])} barrier time-outs"
    assert int(words[9]) == 12, f"rank {rank}: barrier count {int(words[9])}"
    assert all(int(words[r]) == 12 for r in range(world)), words[:world].tolist()
    dist.barrier()
    del g
    for r in range(world):
        if r != rank:
            ops.peer_close(p_data[r]); ops.peer_close(p_flag[r])
    dist.barrier()
    data.free(); flags.free()
    print(f"rank {rank}: ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
