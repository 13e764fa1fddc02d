# r02s (1 GPU): shipped defaults after the late fixes (32-bit view indices at 64^3 again, gather indices fetched a
# whole iteration ahead, row-group rule and fold guard for narrow blocks): bench line, launch list, full ncu
# capture of one RHS, p = 6 fp32, parity of the touched paths
mkdir -p gpurun_out
timeout 900 python bench.py --kernel-times gpurun_out/r02s_kt64.json > gpurun_out/r02s_bench64.json 2> gpurun_out/r02s_bench64.err; tail -c 1200 gpurun_out/r02s_bench64.json; tail -3 gpurun_out/r02s_bench64.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02s_kt64.json')); b = json.loads(open('gpurun_out/r02s_bench64.json').read().strip().splitlines()[-1])
print('n64', round(b['value'],2), round(b['ms_per_step'],4), b['index_bits'], b['compiler'], {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-clocks --no-parity > gpurun_out/r02s_launches.log 2>&1; tail -1 gpurun_out/r02s_launches.log | cut -c1-200
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"opmul|gradflux|intcflux" -s 12 -c 4 -o gpurun_out/rhs64_r02s python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks --no-parity > gpurun_out/ncu_r02s.log 2>&1; tail -1 gpurun_out/ncu_r02s.log | cut -c1-200
timeout 600 python bench.py --n 96 --order 6 --precision single --steps 10 --warmup 3 --no-cpu --no-e2e --no-clocks --kernel-times gpurun_out/r02s_kt_p6.json > gpurun_out/r02s_b_p6.json 2> gpurun_out/r02s_b_p6.err; tail -2 gpurun_out/r02s_b_p6.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02s_kt_p6.json')); b = json.loads(open('gpurun_out/r02s_b_p6.json').read())
print('p6 fp32 96^3', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b.get('parity'), b.get('rhs_model'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
timeout 300 python bench.py --n 64 --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt affine-fastpath=0 --kernel-times gpurun_out/r02s_kt_general.json > gpurun_out/r02s_b_general.json 2> gpurun_out/r02s_b_general.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02s_kt_general.json')); b = json.loads(open('gpurun_out/r02s_b_general.json').read())
print('general geometry', round(b['value'],2), round(b['ms_per_step'],3), {k.split(':')[1]: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
PYFR_B200_PARITY_TAG=r02s timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_partitions.py -m gpu -q -x -k "not full_size" 2>&1 | tail -8
