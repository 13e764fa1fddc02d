# r02n (1 GPU): intconu folded into the element kernel (gather by cp.async) and the element kernel on half blocks
# (two CTAs per SM); attribution runs; parity of the new paths; the soa16 failure of r02m with its traceback
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity "$@" --kernel-times gpurun_out/r02n_kt_$tag.json > gpurun_out/r02n_b_$tag.json 2> gpurun_out/r02n_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02n_kt_$tag.json')); b = json.loads(open('gpurun_out/r02n_b_$tag.json').read())
    print('$tag', round(b['value'],2), round(b['ms_per_step'],4), {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02n_b_$tag.err').read()[-1500:])
PY
}
run n32_default --n 32
run n32_nofold --n 32 --opt conu-fold=0
run n32_nosplit --n 32 --opt gradflux-split=0
run n32_neither --n 32 --opt conu-fold=0 --opt gradflux-split=0
run n64_default --n 64
run n64_nofold --n 64 --opt conu-fold=0
PYFR_B200_PARITY_TAG=r02n_a timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "soa16 and hllc and curved" --tb=long 2>&1 | grep -v "^$" | tail -60
PYFR_B200_PARITY_TAG=r02n timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_partitions.py -m gpu -q -k "not full_size" 2>&1 | tail -25
