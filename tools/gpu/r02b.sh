# r02b: device run of the new parity cases (partition-boundary path on one device, full-size RHS and the
# p=4 1000-step run against oracle/crhs) + baseline bench with the run-time scalars moved to device memory
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PYFR_B200_PARITY_TAG=r02b_partitions timeout 900 python -m pytest tests/test_partitions.py -m gpu -q 2>&1 | tail -15
PYFR_B200_PARITY_TAG=r02b_fullsize timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_timestep.py -m gpu -q --durations=5 -k "full_size or 1000_steps_p4 or order=4" 2>&1 | tail -25
timeout 600 python bench.py --no-cpu --kernel-times gpurun_out/r02b_kt64.json > gpurun_out/r02b_bench64.json 2> gpurun_out/r02b_bench64.err; tail -c 2500 gpurun_out/r02b_bench64.json; tail -3 gpurun_out/r02b_bench64.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --opt inters-order=address --kernel-times gpurun_out/r02b_kt_addr.json > gpurun_out/r02b_b_addr.json 2> gpurun_out/r02b_b_addr.err; tail -c 600 gpurun_out/r02b_b_addr.json
nvidia-smi --query-gpu=name,memory.total --format=csv; free -g | head -2; nproc
