# r01w: regression of the GPU suite after making the fused Euler kernel the default
python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
