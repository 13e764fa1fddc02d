set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
ncu --set full --clock-control none --import-source on -k regex:"gradflux|intcflux|intconu" -s 6 -c 3 -o gpurun_out/fused_r01c python bench.py --n 32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/ncu_gf.log 2>&1; tail -3 gpurun_out/ncu_gf.log
