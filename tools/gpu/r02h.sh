# r02h (2 GPUs): partition parity over NCCL, weak scaling point, strong scaling point of the 128^3 mesh
# partitioned by the reference's partitioner (64-bit view indices at this size)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tests/mgpu_parity.py 2>&1 | grep -E "rank|Error" | tail -8
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-e2e --kernel-times gpurun_out/r02h_kt_weak2.json > gpurun_out/r02h_b_weak2.json 2> gpurun_out/r02h_b_weak2.err; tail -c 600 gpurun_out/r02h_b_weak2.json; tail -2 gpurun_out/r02h_b_weak2.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02h_kt_weak2.json')); b = json.loads(open('gpurun_out/r02h_b_weak2.json').read())
print('weak2', round(b['value'],2), round(b['ms_per_step'],3), b['parity'], {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
timeout 1500 $TR --master-port 29531 bench.py --gpus 2 --scaling strong --mesh-n 128 --partition reference --steps 10 --warmup 3 --no-e2e --no-clocks --kernel-times gpurun_out/r02h_kt_strong2.json > gpurun_out/r02h_b_strong2.json 2> gpurun_out/r02h_b_strong2.err; tail -3 gpurun_out/r02h_b_strong2.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02h_kt_strong2.json')); b = json.loads(open('gpurun_out/r02h_b_strong2.json').read())
print('strong2', round(b['value'],2), round(b['ms_per_step'],3), b['index_bits'], b['setup_s'], b['parity'], {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
