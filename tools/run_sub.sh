set -u
OUT=gpurun_out; T=r1x
timeout 600 python -m pytest tests/test_gpu_generate.py -x -q -k "row_group or graph_and_eager or ragged" > $OUT/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${T}_tests.log
for S in 1 2 4 8; do
  GIC_SUBBATCH=$S timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/${T}_bench_s$S.json 2> $OUT/${T}_bench_s$S.err; echo "S=$S rc=$?"; python -c "
import json,sys
d=json.loads(open('$OUT/${T}_bench_s$S.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['decode_step_roofline']['measured_us'])"
done
