mkdir -p gpurun_out
PYFR_B200_FORCE_NVCC=1 ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 2 -c 1 -o gpurun_out/gradflux_r01d python bench.py --n 32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/ncu_gf.log 2>&1; tail -2 gpurun_out/ncu_gf.log
