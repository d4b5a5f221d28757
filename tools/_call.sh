cd /root/repo
date +%s
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --breakdown gpurun_out/r2_breakdown_n8_v3.json > gpurun_out/r2_bench_n8_v3.json 2>gpurun_out/err8.txt; echo "rc=$?"
date +%s
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --graphs 0 > gpurun_out/r2_bench_n8_v3_eager.json 2>gpurun_out/err8e.txt; echo "rc=$?"
date +%s
tail -2 gpurun_out/err8.txt
