#!/bin/bash
# Round-2 final verification on one B200, with the driver's own invocations: DRAM-traffic capture of the final build, the -m gpu
# suite, both bench arms at --steps 20 --warmup 5, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_final}; mkdir -p $O
echo "== ncu traffic"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xm_solve_kernel -c 1 --csv --log-file $O/traffic.csv python tools/ncu_target_big.py 0.25 > $O/traffic_target.log 2>&1; echo "rc=$?"; tail -1 $O/traffic_target.log
python tools/ncu_traffic.py $O/traffic.csv $O/traffic_target.log profiles/r02_solve_traffic.json; cp profiles/r02_solve_traffic.json $O/
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
S=${XM_FINAL_STEPS:-20}; W=${XM_FINAL_WARMUP:-5}
echo "== bench reference"; /usr/bin/time -f "wall %e s" timeout 1500 python bench.py --impl reference --gpus 1 --steps $S --warmup $W > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-400 $O/bench_ref.json; tail -2 $O/bench_ref.err
echo "== bench ours"; /usr/bin/time -f "wall %e s" timeout 1500 python bench.py --gpus 1 --steps $S --warmup $W > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
ls -la $O
