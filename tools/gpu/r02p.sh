# r02p (2 GPUs): element kernel split into interior blocks (behind the exchange of the traces) and boundary blocks,
# intconu folded on partitioned meshes too; A/B against the unsplit path; partition parity over NCCL
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_parity.py 2>&1 | grep -E "rank|Error|PASS|FAIL" | tail -8
run() { # tag, port, extra args
  tag=$1; port=$2; shift; shift
  timeout 600 $TR --master-port $port bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --no-cpu --no-clocks "$@" --kernel-times gpurun_out/r02p_kt_$tag.json > gpurun_out/r02p_b_$tag.json 2> gpurun_out/r02p_b_$tag.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/r02p_kt_$tag.json')); b = json.loads(open('gpurun_out/r02p_b_$tag.json').read().strip().splitlines()[-1])
    print('$tag', round(b['value'],2), round(b['ms_per_step'],3), b.get('parity'), {k: round(x['ms'],3) for k, x in d['kernels'].items()})
except Exception as e:
    print('$tag FAILED', e); print(open('gpurun_out/r02p_b_$tag.err').read()[-2000:])
PY
}
run weak2 29521
run weak2_nooverlap 29531 --opt gradflux-overlap=0
