# r02a (prepared at the end of round 1, not yet run): first GPU job of the next round.
# 1. device parity of everything finished after round 1's GPU budget was spent (tests/test_gpu_zlate.py)
# 2. headline bench + reference arm, unchanged default path (regression check against r01u: 30.6 GDoF/s)
# 3. first timings of the new paths: mixed-element mesh, RK45 with and without the fused stage update,
#    config #5 proxy (p = 6, fp32)
# 4. ncu launch list for the mixed-element case (which dense-operator kernel dominates?)
# Usage: /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu/r02a.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_zlate.py -m gpu -q 2>&1 | tail -25
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zlate.py 2>&1 | tail -4
timeout 600 python bench.py --kernel-times gpurun_out/r02a_kt64.json > gpurun_out/r02a_bench64.json 2> gpurun_out/r02a_bench64.err; tail -c 3800 gpurun_out/r02a_bench64.json; tail -3 gpurun_out/r02a_bench64.err
# gradflux-vec2 (16-byte accesses, two adjacent columns per work item): phase 3 is issue-bound, static SASS count
# of its loop body 58 -> 28 instructions per live output; measure before making it the default
for v in p3 p1,p3 p1,p3,p5; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --opt gradflux-vec2=$v --kernel-times gpurun_out/r02a_kt_vec2_$v.json > gpurun_out/r02a_b_vec2_$v.json 2> gpurun_out/r02a_b_vec2_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02a_kt_vec2_$v.json')); b = json.loads(open('gpurun_out/r02a_b_vec2_$v.json').read())
print('vec2=$v', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()}, b['compiler'])
PY
done
# with half as many work items per phase the best CTA size may move
for t in 384 640; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --opt gradflux-vec2=p1,p3,p5 --opt gradflux-threads=$t --kernel-times gpurun_out/r02a_kt_vec2_t$t.json > gpurun_out/r02a_b_vec2_t$t.json 2> gpurun_out/r02a_b_vec2_t$t.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02a_kt_vec2_t$t.json')); b = json.loads(open('gpurun_out/r02a_b_vec2_t$t.json').read())
print('vec2 threads=$t', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()}, b['compiler'])
PY
done
PYFR_B200_KEEP_SRC=1 ncu --set full --clock-control none --import-source on -k regex:"gradflux" -s 3 -c 1 -o gpurun_out/gradflux_r02a_vec2p3 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-graphs --no-clocks --opt gradflux-vec2=p3 > gpurun_out/ncu_v.log 2>&1; tail -1 gpurun_out/ncu_v.log | cut -c1-200
# interface kernels: points in true address order (-30 % sectors per warp), intconu over pairs (128-bit accesses)
i=0
for o in "--opt inters-order=address" "--opt inters-order=address --opt conu-pairs=1" "--opt inters-order=address --opt conu-pairs=1 --opt gradflux-vec2=p3"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks $o --kernel-times gpurun_out/r02a_kt_int$i.json > gpurun_out/r02a_b_int$i.json 2> gpurun_out/r02a_b_int$i.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02a_kt_int$i.json')); b = json.loads(open('gpurun_out/r02a_b_int$i.json').read())
print('$o', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()}, b['compiler'])
PY
done
# opmul M3 + negdivconf sits at 87 % of the HBM peak with 160 active threads per SM: more row groups = more loads of `out` in flight
for rg in 6 8; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-clocks --opt mul-rowgroups=$rg --kernel-times gpurun_out/r02a_kt_rg$rg.json > gpurun_out/r02a_b_rg$rg.json 2> gpurun_out/r02a_b_rg$rg.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r02a_kt_rg$rg.json')); b = json.loads(open('gpurun_out/r02a_b_rg$rg.json').read())
print('rowgroups=$rg', round(b['value'],2), round(b['ms_per_step'],3), {k: round(x['ms'],3) for k, x in d['kernels'].items()}, b['compiler'])
PY
done
# mixed elements (BASELINE configs[3]): 24^3 cells, p = 3
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --kernel-times gpurun_out/r02a_kt_mixed.json > gpurun_out/r02a_bench_mixed.json 2> gpurun_out/r02a_bench_mixed.err; tail -c 2500 gpurun_out/r02a_bench_mixed.json; tail -3 gpurun_out/r02a_bench_mixed.err
# ... with the dense operators' coefficients in __constant__ memory (LDCU pairs instead of two UMOV per DFMA)
timeout 600 python bench.py --case hex+pri+pyr+tet --n 24 --order 3 --steps 20 --warmup 5 --opt mul-const-table=32 --kernel-times gpurun_out/r02a_kt_mixed_ct.json > gpurun_out/r02a_bench_mixed_ct.json 2> gpurun_out/r02a_bench_mixed_ct.err; tail -c 900 gpurun_out/r02a_bench_mixed_ct.json
timeout 600 python bench.py --case hex+pri --n 32 --order 3 --steps 20 --warmup 5 --kernel-times gpurun_out/r02a_kt_hexpri.json > gpurun_out/r02a_bench_hexpri.json 2> gpurun_out/r02a_bench_hexpri.err; tail -c 1500 gpurun_out/r02a_bench_hexpri.json
# RK45 time stepping: separate stage-update kernels vs the fused epilogue (fixed CFL number => same steps)
for f in "" "--fused-update"; do
  timeout 600 python -m pyfr_b200 tgv --n 48 --order 4 --scheme rk45 --cfl 0.3 --dt 1e-3 --steps 40 --every 40 $f 2>&1 | tail -2
  timeout 600 python -m pyfr_b200 vortex --n 512 --order 3 --scheme rk45 --cfl 0.3 --dt 1e-3 --steps 40 --every 40 $f 2>&1 | tail -2
done
# config #5 proxy: p = 6, fp32 (unfused element chain: the fused kernel's tile does not fit)
timeout 600 python bench.py --n 32 --order 6 --precision single --steps 20 --warmup 5 --no-cpu --no-e2e --kernel-times gpurun_out/r02a_kt_p6.json > gpurun_out/r02a_bench_p6.json 2> gpurun_out/r02a_bench_p6.err; tail -c 1500 gpurun_out/r02a_bench_p6.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_mixed.csv python bench.py --case hex+pri+pyr+tet --n 16 --order 3 --steps 2 --warmup 3 --no-graphs --no-clocks > gpurun_out/ncu_m.log 2>&1; tail -1 gpurun_out/ncu_m.log | cut -c1-200
