# r02r (4 GPUs): weak-scaling point with and without the interior/boundary split of the element kernel
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
run() { # tag, port, extra args
  tag=$1; port=$2; shift; shift
  timeout 600 $TR --master-port $port bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e --no-cpu --no-clocks "$@" --kernel-times gpurun_out/r02r_kt_$tag.json > gpurun_out/r02r_b_$tag.json 2> gpurun_out/r02r_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02r_kt_$tag.json')); b = json.loads(open('gpurun_out/r02r_b_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(b['value'],2), round(b['ms_per_step'],3), b.get('parity'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02r_b_$tag.err').read()[-2000:])
PY
}
run weak4 29521
run weak4_nooverlap 29531 --opt gradflux-overlap=0
