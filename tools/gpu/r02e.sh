# r02e: full-size parity with the local criterion, reference host on the device, soa4 / two-CTA variant of
# the sum-factorised kernel, first device runs of configs[4] (p=6 fp32) and configs[3] (mixed), RK45 timings
mkdir -p gpurun_out
PYFR_B200_PARITY_TAG=r02e timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_reference_dropin.py -m gpu -q --durations=3 -k "full_size or reference_host or affine" 2>&1 | tail -14
for o in "gradflux-groups=1" "n-soa=4"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt $o --kernel-times gpurun_out/r02e_kt_$o.json > gpurun_out/r02e_b_$o.json 2> gpurun_out/r02e_b_$o.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02e_kt_$o.json')); b = json.loads(open('gpurun_out/r02e_b_$o.json').read())
print('$o', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
done
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 3 -c 1 -o gpurun_out/gradflux_r02e_soa4 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks --no-parity --opt n-soa=4 > gpurun_out/ncu_r02e.log 2>&1; tail -1 gpurun_out/ncu_r02e.log | cut -c1-200
# configs[4] proxy: p = 6, fp32
for o in "graphs=true" "n-soa=4" "n-soa=8"; do
  timeout 400 python bench.py --n 40 --order 6 --precision single --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt $o --kernel-times gpurun_out/r02e_kt_p6_$o.json > gpurun_out/r02e_b_p6_$o.json 2> gpurun_out/r02e_b_p6_$o.err; tail -2 gpurun_out/r02e_b_p6_$o.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02e_kt_p6_$o.json')); b = json.loads(open('gpurun_out/r02e_b_p6_$o.json').read())
print('p6 $o', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
done
# configs[3]: mixed elements
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --no-clocks --kernel-times gpurun_out/r02e_kt_mixed.json > gpurun_out/r02e_bench_mixed.json 2> gpurun_out/r02e_bench_mixed.err; tail -2 gpurun_out/r02e_bench_mixed.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02e_kt_mixed.json')); b = json.loads(open('gpurun_out/r02e_bench_mixed.json').read())
print('mixed', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step']); print(sorted(((round(x['ms'],3), k) for k, x in d['kernels'].items()), reverse=True)[:12])
PY
# RK45 time stepping: separate stage-update kernels vs the fused epilogue
for f in "" "--fused-update"; do
  timeout 600 python -m pyfr_b200 tgv --n 48 --order 4 --scheme rk45 --cfl 0.3 --dt 1e-3 --steps 40 --every 40 $f 2>&1 | tail -2
done
