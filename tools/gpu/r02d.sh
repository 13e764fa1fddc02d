# r02d: gradflux in two phase-shifted warp groups: parity, timing (groups 1/2, stagger 1/3), ncu
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PYFR_B200_PARITY_TAG=r02d_parity timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_timestep.py tests/test_partitions.py -m gpu -q --durations=4 2>&1 | tail -22
for o in "gradflux-groups=2" "gradflux-groups=1" "gradflux-stagger=3"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt $o --kernel-times gpurun_out/r02d_kt_$o.json > gpurun_out/r02d_b_$o.json 2> gpurun_out/r02d_b_$o.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02d_kt_$o.json')); b = json.loads(open('gpurun_out/r02d_b_$o.json').read())
print('$o', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
done
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 3 -c 1 -o gpurun_out/gradflux_r02d python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks --no-parity > gpurun_out/ncu_r02d.log 2>&1; tail -1 gpurun_out/ncu_r02d.log | cut -c1-200
timeout 600 python bench.py --kernel-times gpurun_out/r02d_kt64.json > gpurun_out/r02d_bench64.json 2> gpurun_out/r02d_bench64.err; tail -c 2600 gpurun_out/r02d_bench64.json; tail -3 gpurun_out/r02d_bench64.err
