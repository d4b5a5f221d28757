cd /root/repo
python -m pytest tests/test_gpu_example_loop.py -x -q -m gpu 2>&1 | tail -5
python examples/inpaint_loop.py --steps 50 > gpurun_out/r2_inpaint_loop_n1.json 2>gpurun_out/loop_err.txt; echo "rc=$?"; cat gpurun_out/r2_inpaint_loop_n1.json; tail -2 gpurun_out/loop_err.txt
