# r02i (8 GPUs): strong scaling point of the 128^3 mesh (reference partitioner) and the weak-scaling point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 bench.py --gpus 8 --scaling strong --mesh-n 128 --partition reference --steps 20 --warmup 5 --no-e2e --no-clocks --kernel-times gpurun_out/r02i_kt_strong8.json > gpurun_out/r02i_b_strong8.json 2> gpurun_out/r02i_b_strong8.err; tail -3 gpurun_out/r02i_b_strong8.err | cut -c1-300
python - <<PY
import json
d = json.load(open('gpurun_out/r02i_kt_strong8.json')); b = json.loads(open('gpurun_out/r02i_b_strong8.json').read())
print('strong8 ref-part', round(b['value'],2), round(b['ms_per_step'],3), b['index_bits'], round(b['setup_s']), b['parity'], {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
timeout 900 $TR --master-port 29551 bench.py --gpus 8 --scaling strong --mesh-n 128 --steps 20 --warmup 5 --no-e2e --no-clocks --no-parity > gpurun_out/r02i_b_strong8_brick.json 2> gpurun_out/r02i_b_strong8_brick.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02i_b_strong8_brick.json').read())
print('strong8 brick', round(b['value'],2), round(b['ms_per_step'],3), b['roofline']['sum_kernel_ms'])
PY
timeout 900 $TR --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 --partition reference --mesh-n 64 --scaling strong --no-e2e --no-clocks --no-parity > gpurun_out/r02i_b_strong8_64.json 2> gpurun_out/r02i_b_strong8_64.err
python - <<PY
import json
b = json.loads(open('gpurun_out/r02i_b_strong8_64.json').read())
print('strong8 64^3 ref-part', round(b['value'],2), round(b['ms_per_step'],3), b['roofline']['sum_kernel_ms'])
PY
