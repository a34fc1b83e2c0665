#!/bin/bash
# Round-2 GPU session 6 (one B200): the even-batch build — whole -m gpu suite + bench (ours) on the main workload.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_s6}; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
ls -la $O
