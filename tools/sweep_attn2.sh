# development aid: correctness + timing of differently-tuned builds of the attention kernel (csrc/libvp_b200_<name>.so)
for v in "$@"; do
  echo "== $v"
  VP_B200_LIB=$PWD/videopainter_b200/csrc/libvp_b200_$v.so timeout 300 python tools/attn_lab.py 2>&1 | tail -1
done
