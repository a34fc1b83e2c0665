#!/bin/bash
# Round-2 final verification, part 2: both bench arms with the driver's own invocation (--gpus 1 --steps 20 --warmup 5), then the
# edge-case test added late and the sanitizer passes over the final kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_final2; mkdir -p $O
S=${XM_FINAL_STEPS:-20}; W=${XM_FINAL_WARMUP:-5}
t0=$SECONDS
echo "== bench reference"; timeout 1500 python bench.py --impl reference --gpus 1 --steps $S --warmup $W > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$? wall $((SECONDS-t0)) s"; cut -c1-400 $O/bench_ref.json; tail -2 $O/bench_ref.err
t0=$SECONDS
echo "== bench ours"; timeout 1500 python bench.py --gpus 1 --steps $S --warmup $W > $O/bench.json 2> $O/bench.err; echo "rc=$? wall $((SECONDS-t0)) s"; cut -c1-400 $O/bench.json; tail -2 $O/bench.err
echo "== pytest ops"; timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q > $O/pytest_ops.log 2>&1; echo "rc=$?"; tail -5 $O/pytest_ops.log
echo "== sanitizer memcheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_single.txt python tools/sanitize_target.py single > $O/memcheck_single.out 2>&1; echo "rc=$?"; tail -2 $O/memcheck_single.out; tail -2 $O/memcheck_single.txt
echo "== sanitizer racecheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_single.txt python tools/sanitize_target.py single > $O/racecheck_single.out 2>&1; echo "rc=$?"; tail -2 $O/racecheck_single.out; tail -3 $O/racecheck_single.txt
echo "== sanitizer memcheck multi"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_multi.txt python tools/sanitize_target.py multi > $O/memcheck_multi.out 2>&1; echo "rc=$?"; tail -2 $O/memcheck_multi.out; tail -2 $O/memcheck_multi.txt
ls -la $O
