cd /root/repo
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_example_loop.py -x -q -m gpu 2>&1 | tail -6
