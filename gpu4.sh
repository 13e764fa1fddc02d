set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gradflux -s 2 -c 1 -o gpurun_out/gradflux_r01b python bench.py --n 32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/ncu_gf.log 2>&1; tail -3 gpurun_out/ncu_gf.log
ls -la gpurun_out
