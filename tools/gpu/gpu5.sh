mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mb in 4 5 6 8; do
python bench.py --n 48 --steps 10 --warmup 3 --no-cpu --no-e2e --opt cflux-minblocks=$mb --kernel-times gpurun_out/kt_mb$mb.json > gpurun_out/b_mb$mb.json 2>gpurun_out/b_mb$mb.err; tail -2 gpurun_out/b_mb$mb.err
python - <<PY
import json
d = json.load(open('gpurun_out/kt_mb$mb.json'))
print('minblocks $mb', round(d['ms_per_step'],3), {k: round(v['ms'],3) for k, v in d['kernels'].items()})
PY
done
python bench.py --n 64 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt64g.json > gpurun_out/bench64g.json 2> gpurun_out/bench64g.err; tail -2 gpurun_out/bench64g.err
python - <<PY
import json
d = json.load(open('gpurun_out/kt64g.json')); b = json.loads(open('gpurun_out/bench64g.json').read())
print('64^3', b['value'], b['ms_per_step'], b['rhs_model']['frac_of_hbm_3pass'], {k: round(v['ms'],3) for k, v in d['kernels'].items()})
PY
