# development aid: time differently-tuned builds of the LayerNorm-modulate kernel
for v in "$@"; do
  echo "== $v"
  VP_B200_LIB=$PWD/videopainter_b200/csrc/libvp_b200_$v.so timeout 120 python tests/prof_kernels.py --only ln_modulate --iters 7 2>&1 | tail -2
done
