# r02y (1 GPU): the default bench line on the final state of round 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 170 python bench.py --no-cpu --kernel-times gpurun_out/r02y_kt64.json > gpurun_out/r02y_bench64.json 2> gpurun_out/r02y_bench64.err; tail -c 600 gpurun_out/r02y_bench64.json; tail -2 gpurun_out/r02y_bench64.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02y_kt64.json')); b = json.loads(open('gpurun_out/r02y_bench64.json').read().strip().splitlines()[-1])
print('n64', round(b['value'],2), round(b['ms_per_step'],4), b['index_bits'], b['compiler'], b['roofline']['frac'], b['e2e']['value'], b['clocks'], {k.split(':')[1]: round(x['ms'],4) for k, x in d['kernels'].items()})
PY
