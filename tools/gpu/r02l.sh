# r02l (4 GPUs): strong-scaling point of the 128^3 mesh (reference partitioner), weak point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 bench.py --gpus 4 --scaling strong --mesh-n 128 --partition reference --steps 20 --warmup 5 --no-e2e --no-clocks --kernel-times gpurun_out/r02l_kt_strong4.json > gpurun_out/r02l_b_strong4.json 2> gpurun_out/r02l_b_strong4.err; tail -3 gpurun_out/r02l_b_strong4.err | cut -c1-300
python - <<PY
import json
d = json.load(open('gpurun_out/r02l_kt_strong4.json')); b = json.loads(open('gpurun_out/r02l_b_strong4.json').read())
print('strong4 ref-part', round(b['value'],2), round(b['ms_per_step'],3), b['index_bits'], round(b['setup_s']), b['parity'], {k: round(x['ms'],3) for k, x in d['kernels'].items()})
PY
