# r02m (1 GPU): gradflux after the register/pipelining work (packed descriptors, no spills, software-pipelined
# line phases, dead-end interpolation skipped, 1/|J| folded into the metric / viscosity); thread-count sweep;
# first device timing of the FP64-mma dense kernel on the mixed mesh; parity of the changed kernels
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity "$@" --kernel-times gpurun_out/r02m_kt_$tag.json > gpurun_out/r02m_b_$tag.json 2> gpurun_out/r02m_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02m_kt_$tag.json')); b = json.loads(open('gpurun_out/r02m_b_$tag.json').read())
    print('$tag', round(b['value'],2), round(b['ms_per_step'],4), {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02m_b_$tag.err').read()[-600:])
PY
}
run n32_default --n 32
run n32_swp0 --n 32 --opt gradflux-swp=0
run n32_t288 --n 32 --opt gradflux-threads=288
run n32_t320 --n 32 --opt gradflux-threads=320
run n32_t384 --n 32 --opt gradflux-threads=384
run n32_t256 --n 32 --opt gradflux-threads=256
run n32_soa4 --n 32 --opt n-soa=4
run n64_default --n 64
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --no-cpu --no-e2e --kernel-times gpurun_out/r02m_kt_mixed.json > gpurun_out/r02m_bench_mixed.json 2> gpurun_out/r02m_bench_mixed.err; tail -2 gpurun_out/r02m_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02m_kt_mixed.json')); b = json.loads(open('gpurun_out/r02m_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b.get('parity')); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:14])
PY
timeout 300 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --no-cpu --no-e2e --no-parity --opt dense-mma=0 --kernel-times gpurun_out/r02m_kt_mixed_fma.json > gpurun_out/r02m_bench_mixed_fma.json 2> gpurun_out/r02m_bench_mixed_fma.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02m_kt_mixed_fma.json')); b = json.loads(open('gpurun_out/r02m_bench_mixed_fma.json').read())
print('mixed dense-mma=0', round(b['value'],2), round(b['ms_per_step'],3)); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:8])
PY
PYFR_B200_PARITY_TAG=r02m timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tgv_rhs or affine or fp32 or mixed" 2>&1 | tail -15
PYFR_B200_PARITY_TAG=r02m_z timeout 600 python -m pytest tests/test_gpu_zlate.py -m gpu -x -q -k "not zz_opt" 2>&1 | tail -5
