# r02c: sum-factorised gradflux (kernels/tensor.py) + provider-private interface ordering on the device:
# parity (partition-boundary path, full suite incl. full-size + p=4 1000 steps), bench vs the table-driven
# kernel, general (non-affine) path, ncu full capture of one RHS
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PYFR_B200_PARITY_TAG=r02c_partitions timeout 900 python -m pytest tests/test_partitions.py -m gpu -q -x 2>&1 | tail -12
PYFR_B200_PARITY_TAG=r02c_parity timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_timestep.py -m gpu -q -x --durations=4 2>&1 | tail -22
timeout 600 python bench.py --no-cpu --kernel-times gpurun_out/r02c_kt64.json > gpurun_out/r02c_bench64.json 2> gpurun_out/r02c_bench64.err; tail -c 1700 gpurun_out/r02c_bench64.json; tail -3 gpurun_out/r02c_bench64.err
for o in "gradflux-tensor=0" "affine-fastpath=0" "kernel-order=host"; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --opt $o --kernel-times gpurun_out/r02c_kt_$o.json > gpurun_out/r02c_b_$o.json 2> gpurun_out/r02c_b_$o.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02c_kt_$o.json')); b = json.loads(open('gpurun_out/r02c_b_$o.json').read())
print('$o', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
done
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"opmul|intconu|gradflux|intcflux" -s 15 -c 5 -o gpurun_out/rhs64_r02c python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_r02c.log 2>&1; tail -1 gpurun_out/ncu_r02c.log | cut -c1-200
