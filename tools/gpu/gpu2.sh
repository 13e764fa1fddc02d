set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --n 32 --steps 10 --warmup 3 --cpu-n 8 --kernel-times gpurun_out/kt32.json > gpurun_out/bench32.json 2> gpurun_out/bench32.err; tail -c 3000 gpurun_out/bench32.json; tail -5 gpurun_out/bench32.err
python bench.py --n 64 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt64.json > gpurun_out/bench64.json 2> gpurun_out/bench64.err; tail -c 3000 gpurun_out/bench64.json; tail -5 gpurun_out/bench64.err
cat gpurun_out/kt64.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches32.csv python bench.py --n 32 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs > gpurun_out/ncu32.log 2>&1; tail -3 gpurun_out/ncu32.log
