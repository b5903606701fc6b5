set -u
OUT=gpurun_out; T=${1:-r1af}
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cta_pair" > $OUT/${T}_ktests.log 2>&1; echo "ktests rc=$?"; tail -15 $OUT/${T}_ktests.log
