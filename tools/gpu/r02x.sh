# r02x (2 GPUs): weak-scaling point on the final kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu --no-clocks --kernel-times gpurun_out/r02x_kt_weak2.json > gpurun_out/r02x_b_weak2.json 2> gpurun_out/r02x_b_weak2.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02x_kt_weak2.json')); b = json.loads(open('gpurun_out/r02x_b_weak2.json').read().strip().splitlines()[-1])
    print('weak2', round(b['value'],2), round(b['ms_per_step'],3), b.get('parity'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
except Exception as e:
    print('weak2 FAILED', e); print(open('gpurun_out/r02x_b_weak2.err').read()[-2000:])
PY
