#!/bin/bash
# One GPU-box session.  Usage (repo root, under gpurun):  [STEPS="tests smoke bench ref probes launches full"] bash tools/gpu_round.sh <tag>
# Everything lands in gpurun_out/<tag>_* (keep it well under gpurun's 64 MiB pull limit: the full ncu capture covers ONE
# layer + LM head of one decode step, not the whole step).  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r1}
STEPS=${STEPS:-"tests smoke bench ref launches full"}
OUT=gpurun_out
DTYPE=${DTYPE:-bf16x2}
mkdir -p $OUT
has() { [[ " $STEPS " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if has tests; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log; tail -3 $OUT/${TAG}_tests.log; fi
if has smoke; then timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/${TAG}_smoke.log; fi
if has bench; then timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; head -c 400 $OUT/${TAG}_bench.json; echo; fi
if has ref; then timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"; fi
if has probes; then
  B=gpt2_image_captioning_b200/csrc/build
  timeout 120 $B/tma_probe > $OUT/${TAG}_tma_probe.log 2>&1; echo "tma_probe rc=$?"
  timeout 120 $B/microbench 1024 > $OUT/${TAG}_microbench.log 2>&1; echo "microbench rc=$?"
  timeout 120 $B/microbench 1024 2 > $OUT/${TAG}_mma_probe.log 2>&1; echo "mma_probe rc=$?"
  timeout 120 $B/microbench 1024 3 > $OUT/${TAG}_epi_probe.log 2>&1; echo "epi_probe rc=$?"
fi
if has launches; then
  timeout 300 python tools/profile_step.py --dtype $DTYPE --max-length 6 > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/${TAG}_launches.csv python tools/profile_step.py --dtype $DTYPE --max-length 6 > $OUT/${TAG}_ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
fi
if has full; then
  # kernels matching the regex: prefill = 12 x (ln, gemm, ln, gemm, gemm) ... ; skip the prefill + first decode step's first
  # 11 layers, then capture the last layer (7 launches) + ln_f + LM head + finalize of that step
  SKIP=${NCU_SKIP:-}
  if [ -z "$SKIP" ]; then SKIP=$(DTYPE=$DTYPE python - <<'PY'
# matching launches before the last layer of decode step 1 (fused engines, bf16 / bf16x2: no LayerNorm launches inside the blocks):
# mapper (2 GEMMs) + prefill (12 layers x 4 GEMMs + ln_f + LM head + 2 re-scoring kernels (bf16x2 only) + finalize) + 11 decode layers x 5
import os
x2 = os.environ.get("DTYPE", "bf16x2") == "bf16x2"
print(2 + 12 * 4 + 3 + (2 if x2 else 0) + 11 * 5)
PY
); fi
  timeout 300 python tools/profile_step.py --dtype $DTYPE --max-length ${NCU_MAXLEN:-6} > $OUT/${TAG}_plain2.log 2>&1 &&
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
      -k regex:'gemm_bf16_tcgen05|attn_decode|layernorm|finalize|lm_head' -s $SKIP -c ${NCU_COUNT:-10} -f -o $OUT/${TAG}_step \
      python tools/profile_step.py --dtype $DTYPE --max-length ${NCU_MAXLEN:-6} > $OUT/${TAG}_ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
du -sh $OUT
