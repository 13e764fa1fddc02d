# r01u: end-of-round state on one GPU: parity suite, bench (+reference arm), ncu launch list and full capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --kernel-times gpurun_out/r01u_kt64.json > gpurun_out/r01u_bench64.json 2> gpurun_out/r01u_bench64.err; tail -c 3800 gpurun_out/r01u_bench64.json; tail -3 gpurun_out/r01u_bench64.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r01u_bench_reference.json 2> gpurun_out/r01u_bench_reference.err; tail -c 1500 gpurun_out/r01u_bench_reference.json | cut -c1-700; tail -2 gpurun_out/r01u_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01u_launches64.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:"opmul|intconu|gradflux|intcflux" -s 15 -c 5 -o gpurun_out/rhs64_r01u python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
