#!/bin/bash
# Round-2 GPU session 4 (one B200): re-run of the fixed tests, block-CSR with the warp-local epilogue (1024 vs 512 threads),
# the standalone v4-operand experiment, sanitizer re-runs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_s4}; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -16 $O/pytest_gpu.log
echo "== bsr ER-100k (1024 threads)"; timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt1024.jsonl > $O/bsr1024.log 2>&1; echo "rc=$?"
echo "== bsr ER-100k (512 threads)"; XM_TUNE_BSR_NT=512 timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt512.jsonl > $O/bsr512.log 2>&1; echo "rc=$?"
echo "== bsr ER-100k (1024 threads, K=4)"; XM_TUNE_BSR_K=4 timeout 600 python tools/bench_bsr.py --out $O/bsr_er100k_nt1024_k4.jsonl > $O/bsr1024k4.log 2>&1; echo "rc=$?"
python - <<PY
import json
for f in ("$O/bsr_er100k_nt1024.jsonl", "$O/bsr_er100k_nt512.jsonl", "$O/bsr_er100k_nt1024_k4.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l); s = d.get("solve", {})
            print(f.split("er100k_")[-1], "r", d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"], 3), round(d["ms_per_product_lockstep"], 3), "frac", round(d["frac"], 3), "solve it/s", round(s.get("tcg_iters_per_s", 0), 1), "ms/prod in solve", round(s.get("ms_per_qy_product_in_solve", 0), 3))
    except Exception as e:
        print(f, e)
PY
echo "== bsr_tune (v4 operand)"; for r in 5 10 20; do timeout 300 tools/bsr_tune 100000 100 $r 10 0 | head -5; done > $O/bsr_tune_v4.txt 2>&1; cat $O/bsr_tune_v4.txt
echo "== ncu full: block-CSR Q.Y (r = 10)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:xm_ops_kernel -s 1 -c 1 -o $O/bsr_qy_full -f python tools/ncu_target_bsr.py > $O/ncu_bsr.log 2>&1; echo "rc=$?"; tail -1 $O/ncu_bsr.log
ncu -i $O/bsr_qy_full.ncu-rep --page raw --csv > $O/bsr_qy_full.raw.csv 2>/dev/null; rm -f $O/bsr_qy_full.ncu-rep
echo "== sanitizer racecheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_single.txt python tools/sanitize_target.py single > $O/racecheck_single.out 2>&1; echo "rc=$?"; tail -2 $O/racecheck_single.out; tail -3 $O/racecheck_single.txt
echo "== sanitizer memcheck multi (loop-back)"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_multi.txt python tools/sanitize_target.py multi > $O/memcheck_multi.out 2>&1; echo "rc=$?"; tail -3 $O/memcheck_multi.out; tail -3 $O/memcheck_multi.txt
echo "== sanitizer racecheck multi (loop-back)"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_multi.txt python tools/sanitize_target.py multi > $O/racecheck_multi.out 2>&1; echo "rc=$?"; tail -3 $O/racecheck_multi.out; tail -3 $O/racecheck_multi.txt
echo "== launch list of the bench command"; XM_BENCH_EXTRAS=0 XM_BENCH_FULL_SOLVE=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:xm_ -c 200 --csv --log-file $O/launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 1 > $O/ncu_bench.log 2>&1; echo "rc=$?"; tail -c 600 $O/ncu_bench.log
ls -la $O
