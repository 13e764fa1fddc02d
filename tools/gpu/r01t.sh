# r01t: boundary-condition kernels + staged affine metrics
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --kernel-times gpurun_out/kt_t.json > gpurun_out/b_t.json 2> gpurun_out/b_t.err; tail -2 gpurun_out/b_t.err
python - <<PY
import json
d = json.load(open('gpurun_out/kt_t.json')); b = json.loads(open('gpurun_out/b_t.json').read())
print('affine-staged', round(b['value'],2), round(b['ms_per_step'],3), {k: round(v['ms'],3) for k, v in d['kernels'].items()}, b['compiler'])
PY
