# r02o (1 GPU): gathered common solution with whole rows by TMA bulk copy; attribution; parity
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity "$@" --kernel-times gpurun_out/r02o_kt_$tag.json > gpurun_out/r02o_b_$tag.json 2> gpurun_out/r02o_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02o_kt_$tag.json')); b = json.loads(open('gpurun_out/r02o_b_$tag.json').read())
    print('$tag', round(b['value'],2), round(b['ms_per_step'],4), {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02o_b_$tag.err').read()[-1500:])
PY
}
run n32_default --n 32
run n32_norows --n 32 --opt gather-rows=0
run n32_nofold --n 32 --opt conu-fold=0
run n64_default --n 64
run n64_nofold --n 64 --opt conu-fold=0
PYFR_B200_PARITY_TAG=r02o timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_partitions.py -m gpu -q -x -k "not full_size and not fp32" 2>&1 | tail -12
