# r01o: two GPUs -- multi-rank parity (NCCL halo exchange) and weak-scaling bench; every stage bounded
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py > gpurun_out/mgpu_parity.log 2>&1
echo "parity rc=$?"; grep -v Warning gpurun_out/mgpu_parity.log | tail -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt_2gpu.json > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_2gpu.json; grep -v Warning gpurun_out/bench_2gpu.err | tail -3 | cut -c1-300
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/kt_2gpu.json'))
    print({k: round(v['ms'], 3) for k, v in d['kernels'].items()})
except Exception as e:
    print('no kt', e)
PY
