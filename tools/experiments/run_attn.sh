set -u
OUT=gpurun_out; T=${1:-r1z}
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -k "attn_decode" > $OUT/${T}_ktests.log 2>&1; echo "ktests rc=$?"; tail -5 $OUT/${T}_ktests.log
timeout 300 gpt2_image_captioning_b200/csrc/build/microbench 1024 > $OUT/${T}_microbench.log 2>&1; echo rc=$?; grep -i "attn_decode\|stream_read" $OUT/${T}_microbench.log
