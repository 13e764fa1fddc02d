# r02v (1 GPU): the device cases not re-run since the last kernel changes (fp32, late features, time stepping)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
PYFR_B200_PARITY_TAG=r02v timeout 900 python -m pytest tests/test_gpu_zlate.py tests/test_gpu_timestep.py tests/test_gpu_parity.py -m gpu -q -k "not zz_opt and not full_size and not tgv_rhs_matches and not affine_mesh and not boundary_conditions and not 1000_steps_p4" --durations=5 2>&1 | tail -14
