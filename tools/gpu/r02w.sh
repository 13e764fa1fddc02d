# r02w (1 GPU): QS stride without bank conflicts (22 instead of 20 doubles per element)
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity "$@" --kernel-times gpurun_out/r02w_kt_$tag.json > gpurun_out/r02w_b_$tag.json 2> gpurun_out/r02w_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02w_kt_$tag.json')); b = json.loads(open('gpurun_out/r02w_b_$tag.json').read())
    print('$tag', round(b['value'],2), round(b['ms_per_step'],4), b['compiler'], {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02w_b_$tag.err').read()[-1500:])
PY
}
run n64 --n 64
PYFR_B200_PARITY_TAG=r02w timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "affine_mesh or tgv_rhs_matches" 2>&1 | tail -4
