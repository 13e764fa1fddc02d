mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --n 64 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt64h.json > gpurun_out/bench64h.json 2> gpurun_out/bench64h.err; tail -2 gpurun_out/bench64h.err
python - <<PY
import json
d = json.load(open('gpurun_out/kt64h.json')); b = json.loads(open('gpurun_out/bench64h.json').read())
print('64^3', b['value'], b['ms_per_step'], b['rhs_model']['frac_of_hbm_3pass'], {k: round(v['ms'],3) for k, v in d['kernels'].items()})
PY
PYFR_B200_FORCE_NVCC=1 ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 2 -c 1 -o gpurun_out/gradflux_r01e python bench.py --n 32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/ncu_gf.log 2>&1; tail -1 gpurun_out/ncu_gf.log
