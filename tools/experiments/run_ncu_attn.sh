set -u
OUT=gpurun_out; T=${1:-r1z}; VARS=${2:-"0 10"}; CTX=${3:-39}
MB=gpt2_image_captioning_b200/csrc/build/microbench
for V in $VARS; do
  timeout 120 $MB 1024 5 $V $CTX > $OUT/${T}_attn_v${V}_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 4 -c 2 -f -o $OUT/${T}_attn_v$V $MB 1024 5 $V $CTX > $OUT/${T}_attn_v${V}_ncu.log 2>&1
  echo "v$V ncu rc=$?"
done
