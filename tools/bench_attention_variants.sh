# development aid: the denoise step (bench.py) with differently-built attention kernels, on one box
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-eager"
pick='import json,sys; d=json.loads(sys.stdin.readline()); print(round(d["ms_per_step"],1), "ms/step; attention", round(d["roofline"]["avg_launch_ms"],3), "ms", round(d["roofline"]["achieved"],1), "TF/s in-step", d["clocks"]["sm_mhz"], d["noise_sha256"][:12])'
for rep in 1 2; do
echo "== default"; $B 2>/dev/null | python -c "$pick"
for v in "$@"; do echo "== $v"; VP_B200_LIB=$PWD/videopainter_b200/csrc/libvp_b200_$v.so $B 2>/dev/null | python -c "$pick"; done
echo "== v1"; VP_B200_ATTN=v1 $B 2>/dev/null | python -c "$pick"
echo "== v4"; VP_B200_ATTN=v4 $B 2>/dev/null | python -c "$pick"
done
