# r02t (1 GPU): p = 6 fp32 after software pipelining was restricted to where it fits the registers
mkdir -p gpurun_out
timeout 600 python bench.py --n 48 --order 6 --precision single --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --kernel-times gpurun_out/r02t_kt_p6.json > gpurun_out/r02t_b_p6.json 2> gpurun_out/r02t_b_p6.err; tail -2 gpurun_out/r02t_b_p6.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02t_kt_p6.json')); b = json.loads(open('gpurun_out/r02t_b_p6.json').read())
print('p6 fp32 48^3', round(b['value'],2), round(b['ms_per_step'],3), b['launches_per_step'], b.get('parity'), b.get('rhs_model'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
timeout 600 python bench.py --n 48 --order 6 --precision single --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --no-parity --opt gradflux-swp=0 --opt conu-fold=0 --kernel-times gpurun_out/r02t_kt_p6b.json > gpurun_out/r02t_b_p6b.json 2> gpurun_out/r02t_b_p6b.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02t_kt_p6b.json')); b = json.loads(open('gpurun_out/r02t_b_p6b.json').read())
print('p6 fp32 48^3 swp=0 fold=0', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
