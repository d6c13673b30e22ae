# round 2, last call: all GPU tests and smoke() on the final tree
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | tail -6
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
