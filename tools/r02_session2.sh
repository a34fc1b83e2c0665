#!/bin/bash
# Round-2 GPU session 2 (one B200): the whole -m gpu suite (loop-back team, certificate / staircase / assembly behind the C-ABI),
# both bench arms on the BAL-Final-sized workload, the ncu DRAM-traffic capture of the solve kernel, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_s2}; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=15 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 $O/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps ${XM_S2_STEPS:-5} --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-600 $O/bench.json; tail -5 $O/bench.err
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-1500 $O/bench_ref.json; tail -5 $O/bench_ref.err
echo "== ncu traffic"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xm_solve_kernel -c 1 --csv --log-file $O/traffic.csv python tools/ncu_target_big.py 0.25 > $O/traffic_target.log 2>&1; echo "rc=$?"; tail -2 $O/traffic_target.log
python tools/ncu_traffic.py $O/traffic.csv $O/traffic_target.log $O/r02_solve_traffic.json
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
ls -la $O
