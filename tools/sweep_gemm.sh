# development aid: GEMM rasterisation (m-tiles per group) and streaming-store sweep
for lib in gn gs; do for g in 4 8 16 32; do
  echo "== $lib group_m=$g"
  VP_GEMM_GROUP_M=$g VP_B200_LIB=$PWD/videopainter_b200/csrc/libvp_b200_$lib.so timeout 120 python tests/prof_kernels.py --only gemm_ --iters 5 2>&1 | grep -E "qkv|out_gate|ff1|ff2" | sed -E "s/\{'ms': ([0-9.]+), 'tflops': ([0-9.]+)\}/\1 ms \2/" | tr '\n' ';'; echo
done; done
