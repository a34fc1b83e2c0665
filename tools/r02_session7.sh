#!/bin/bash
# Round-2 GPU session 7 (one B200): new edge-case test + sanitizer passes over the kernels added after session 4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_s7; mkdir -p $O
echo "== pytest ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/pytest_ops.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_ops.log
echo "== sanitizer memcheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_single.txt python tools/sanitize_target.py single > $O/memcheck_single.out 2>&1; echo "rc=$?"; tail -2 $O/memcheck_single.out; tail -2 $O/memcheck_single.txt
echo "== sanitizer racecheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_single.txt python tools/sanitize_target.py single > $O/racecheck_single.out 2>&1; echo "rc=$?"; tail -2 $O/racecheck_single.out; tail -3 $O/racecheck_single.txt
echo "== sanitizer memcheck multi"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_multi.txt python tools/sanitize_target.py multi > $O/memcheck_multi.out 2>&1; echo "rc=$?"; tail -2 $O/memcheck_multi.out; tail -2 $O/memcheck_multi.txt
echo "== sanitizer racecheck multi"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_multi.txt python tools/sanitize_target.py multi > $O/racecheck_multi.out 2>&1; echo "rc=$?"; tail -2 $O/racecheck_multi.out; tail -2 $O/racecheck_multi.txt
ls -la $O
