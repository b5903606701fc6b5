mkdir -p gpurun_out/r2o
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-job --no-modes --no-sampling --no-cpu-baseline $EXTRA > gpurun_out/r2o/$name.json 2>gpurun_out/r2o/$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2o/$name.json"))
    print("$name:", round(d["value"]), "e2e", round(d["e2e"]["value"]), "in_flight_1", round(d["in_flight_1"]["value"]), "step_us", round(d["in_graph"]["decode_step_span_us"],1), d["in_graph"]["per_gemm_avg_us"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name: failed", e)
PY
}
EXTRA="" run default X=1
EXTRA="" run nopair GIC_GEMM_PAIR=0
EXTRA="--in-flight 3" run default_f3 X=1
EXTRA="--in-flight 3" run nopair_f3 GIC_GEMM_PAIR=0
EXTRA="--in-flight 4" run nopair_f4 GIC_GEMM_PAIR=0
