set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --n 64 --steps 20 --warmup 5 --no-cpu --kernel-times gpurun_out/kt64f.json > gpurun_out/bench64f.json 2> gpurun_out/bench64f.err; tail -c 2500 gpurun_out/bench64f.json; tail -5 gpurun_out/bench64f.err
cat gpurun_out/kt64f.json
