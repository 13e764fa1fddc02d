# r01f: re-establish the measured state: GPU parity suite, default bench line,
# ncu launch list and one full capture of every kernel of the RHS at 64^3
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --kernel-times gpurun_out/kt64.json > gpurun_out/bench64.json 2> gpurun_out/bench64.err; tail -c 3500 gpurun_out/bench64.json; tail -3 gpurun_out/bench64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches64.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"opmul|intconu|gradflux|intcflux" -s 15 -c 5 -o gpurun_out/rhs64_r01f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_f.log 2>&1; tail -2 gpurun_out/ncu_f.log
ls -la gpurun_out
