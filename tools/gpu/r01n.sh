# r01n: NCCL bring-up diagnostics (2 GPUs), every stage bounded by timeout
mkdir -p gpurun_out
export NCCL_DEBUG=WARN
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu/nccl_diag.py 2>&1 | grep -v "Warning\|warn" | tail -30
echo "diag rc=$?"
cat gpurun_out/nccl_diag_0.log gpurun_out/nccl_diag_1.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tests/mgpu_parity.py 2>&1 | grep -v Warning | tail -12
echo "parity rc=$?"
