# r01x: GPU suite regression after the warped-mesh generator change (non-degenerate warps on coarse meshes)
python -m pytest tests -m gpu -x -q -k "not full_size and not 1000_steps" 2>&1 | tail -4
