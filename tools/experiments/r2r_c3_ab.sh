# c3 (GPT-2 medium, transformer mapper, beam 5) with the round-2 beam changes switched off / on, one B200
mkdir -p gpurun_out/r2r
for dt in bf16x2 bf16; do
  for cfg in "old GIC_BEAM_SHARED_PREFIX=0 GIC_BEAM_LOGITS=1" "attn GIC_BEAM_SHARED_PREFIX=1 GIC_BEAM_LOGITS=1" "head GIC_BEAM_SHARED_PREFIX=0 GIC_BEAM_LOGITS=0" "new GIC_BEAM_SHARED_PREFIX=1 GIC_BEAM_LOGITS=0"; do
    set -- $cfg; name=$1; shift
    env "$@" timeout 600 python tools/bench_configs.py --configs c3 --dtypes $dt --in-flight 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('$dt $name', round(d['ms_per_batch'], 1), 'ms per 1024 images', round(d['captions_per_s']), 'captions/s', d.get('parity', {}).get('c3_beam5_32_rows_vs_hf'))
"
  done
done
