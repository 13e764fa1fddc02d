# r01q: two GPUs -- NCCL transport/bandwidth, parity and bench after copy elision + priority comm stream
mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_info_%h_%p.log timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu/nccl_diag.py > gpurun_out/nccl_diag.out 2>&1
echo "diag rc=$?"; grep -h "p2p 32\|one-directional" gpurun_out/nccl_diag_0.log; grep -h "via \|Channel\|NVLS\|P2P" gpurun_out/nccl_info_*.log | sort | uniq -c | sort -rn | head -12 | cut -c1-220
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_parity.py > gpurun_out/mgpu_parity.log 2>&1
echo "parity rc=$?"; grep -v Warning gpurun_out/mgpu_parity.log | tail -6
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt_2gpu.json > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"; python - <<PY
import json
b = json.loads(open('gpurun_out/bench_2gpu.json').read())
print('2gpu', round(b['value'], 2), round(b['ms_per_step'], 3), 'e2e', b['e2e'] and (round(b['e2e']['value'], 2), round(b['e2e']['ms_per_step'], 2)), b['launches_per_step'])
d = json.load(open('gpurun_out/kt_2gpu.json'))
print({k: round(v['ms'], 3) for k, v in d['kernels'].items()})
PY
grep -v Warning gpurun_out/bench_2gpu.err | tail -3 | cut -c1-300
