# r01s: constant-Jacobian (affine element) fast path of gradflux
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() {
  tag=$1; shift
  python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks "$@" --kernel-times gpurun_out/kt_$tag.json > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; tail -2 gpurun_out/b_$tag.err
  python - <<PY
import json
d = json.load(open('gpurun_out/kt_$tag.json')); b = json.loads(open('gpurun_out/b_$tag.json').read())
print('$tag', round(b['value'],2), round(b['ms_per_step'],3), {k: round(v['ms'],3) for k, v in d['kernels'].items()}, b['compiler'])
PY
}
run affine
run affine640 --opt gradflux-threads=640
run general --opt affine-fastpath=0
ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 3 -c 1 -o gpurun_out/gradflux_r01s python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_s.log 2>&1; tail -1 gpurun_out/ncu_s.log
