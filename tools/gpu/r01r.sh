# r01r: N-GPU parity + weak-scaling bench after anchoring receive-only exchanges late
N=${1:-2}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_parity.py > gpurun_out/mgpu_parity_$N.log 2>&1
echo "parity rc=$?"; grep -c PASS gpurun_out/mgpu_parity_$N.log; grep FAIL gpurun_out/mgpu_parity_$N.log | head -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt_${N}gpu.json > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; python - <<PY
import json
b = json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().split('\n')[-1])
print('${N}gpu', round(b['value'], 2), round(b['ms_per_step'], 3), 'e2e', b['e2e'] and (round(b['e2e']['value'], 2), round(b['e2e']['ms_per_step'], 2)), b['launches_per_step'], b['clocks'])
d = json.load(open('gpurun_out/kt_${N}gpu.json'))
print({k: round(v['ms'], 3) for k, v in d['kernels'].items()}, round(sum(v['ms'] for v in d['kernels'].values()), 3))
PY
wc -l gpurun_out/bench_${N}gpu.json
grep -v Warning gpurun_out/bench_${N}gpu.err | tail -3 | cut -c1-300
