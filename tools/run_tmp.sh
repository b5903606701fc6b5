timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm or mlp" 2>&1 | tail -2
timeout 300 gpt2_image_captioning_b200/csrc/build/microbench 1024 2>&1 | grep "fc2  +LN pair M=1024 N=768 K=3072 block_n= 64\|lm_head pair.*256\|prefill gemm fc2  pair.*256"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1as_bench.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r1as_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['decode_step_roofline']['measured_us'])"
