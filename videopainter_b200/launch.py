"""Run one of the reference's scripts unchanged on the B200 path:

    python -m videopainter_b200.launch infer/inpaint.py --model_path ... --inpainting_branch ... --mask_add
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \
        -m videopainter_b200.launch infer/inpaint.py ...

Patches the forwards of the diffusers fork's `CogVideoXTransformer3DModel` / `CogvideoXBranchModel` (`install()`), sets up the
CFG x Ulysses layout when launched with several ranks, then executes the script as `__main__` (infer/inpaint.py:571 etc.).
The diffusers fork must be importable (the reference's environment); there is no fallback if CUDA or the library is missing."""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    import torch
    from . import install, parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        parallel.init()
    install()
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
