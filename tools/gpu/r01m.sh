# r01m: two GPUs -- multi-rank parity (NCCL halo exchange) and weak-scaling bench
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_parity.py 2>&1 | grep -v Warning | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt_2gpu.json > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 2500 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
cat gpurun_out/kt_2gpu.json | tr -d '\n ' | cut -c1-900; echo
python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --kernel-times gpurun_out/kt_1gpu.json > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; python -c "
import json; b=json.loads(open('gpurun_out/bench_1gpu.json').read()); print('1gpu', b['value'], b['ms_per_step'], b['clocks'])"
