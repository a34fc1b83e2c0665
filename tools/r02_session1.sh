#!/bin/bash
# Round-2 GPU session 1 (one B200): validate the merged mg-split/tcg2 build, the new parity tests and the loop-back team,
# size the BAL-Final workload (both arms), run the block-CSR occupancy experiment.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_s1; mkdir -p $O
{ nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nproc; free -g | head -2; df -h /dev/shm /tmp | cat; } > $O/env.txt 2>&1
echo "== new parity tests"; timeout 900 python -m pytest tests/test_gpu_parity_r2.py -m gpu -x -q --durations=10 > $O/pytest_parity.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_parity.log
echo "== loop-back multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=10 > $O/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 $O/pytest_multi.log
echo "== erec + solve tests"; timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_gpu_ops.py -m gpu -q --durations=10 > $O/pytest_solve.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_solve.log
echo "== C3 bench erec=0"; timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 2 > $O/bench_c3_erec0.json 2> $O/bench_c3_erec0.err; echo "rc=$?"; cut -c1-300 $O/bench_c3_erec0.json
echo "== C3 bench erec=1"; XM_TUNE_EREC=1 timeout 300 python bench.py --steps 5 --warmup 3 --cpu-seconds 2 > $O/bench_c3_erec1.json 2> $O/bench_c3_erec1.err; echo "rc=$?"; cut -c1-300 $O/bench_c3_erec1.json
echo "== big probe"; timeout 900 python tools/big_probe.py 13682 60 25 > $O/big_probe.jsonl 2> $O/big_probe.err; echo "rc=$?"; tail -c 3000 $O/big_probe.jsonl; tail -5 $O/big_probe.err
echo "== bsr_tune"; for r in 5 10 20; do for pad in 0 1; do timeout 300 tools/bsr_tune 100000 100 $r 10 $pad; done; done > $O/bsr_tune.txt 2>&1; echo "rc=$?"; cat $O/bsr_tune.txt
ls -la $O
