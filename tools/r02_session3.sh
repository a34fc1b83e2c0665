#!/bin/bash
# Round-2 GPU session 3 (one B200): whole -m gpu suite, block-CSR ER-100k (1024- vs 512-thread CTAs), ncu --set full captures of the
# two top kernels, compute-sanitizer passes, DRAM-traffic capture + bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_s3}; mkdir -p $O
echo "== pytest -m gpu"; XM_ASM_TRACE=1 timeout 1500 python -m pytest tests -m gpu -q --durations=12 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -22 $O/pytest_gpu.log; grep -h "xm_create_matrix\]" $O/pytest_gpu.log | head -12
echo "== bsr ER-100k (1024 threads)"; timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt1024.jsonl > $O/bsr1024.log 2>&1; echo "rc=$?"
echo "== bsr ER-100k (512 threads)"; XM_TUNE_BSR_NT=512 timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt512.jsonl > $O/bsr512.log 2>&1; echo "rc=$?"
python - <<PY
import json
for f in ("$O/bsr_er100k_nt1024.jsonl", "$O/bsr_er100k_nt512.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l); s = d.get("solve", {})
            print(f.split("_")[-1], "r", d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"], 3), round(d["ms_per_product_lockstep"], 3), "frac", round(d["frac"], 3), "solve it/s", round(s.get("tcg_iters_per_s", 0), 1))
    except Exception as e:
        print(f, e)
PY
echo "== ncu full: dense solve kernel (BAL-Final, 0.1 s cap)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:xm_solve_kernel -c 1 -o $O/solve_final_full -f python tools/ncu_target_big.py 0.1 > $O/ncu_solve.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_solve.log
ncu -i $O/solve_final_full.ncu-rep --page raw --csv > $O/solve_final_full.raw.csv 2>/dev/null
echo "== ncu full: block-CSR Q.Y (r = 10)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:xm_ops_kernel -s 1 -c 1 -o $O/bsr_qy_full -f python tools/ncu_target_bsr.py > $O/ncu_bsr.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_bsr.log
ncu -i $O/bsr_qy_full.ncu-rep --page raw --csv > $O/bsr_qy_full.raw.csv 2>/dev/null
echo "== sanitizer memcheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_single.txt python tools/sanitize_target.py single > $O/memcheck_single.out 2>&1; echo "rc=$?"; tail -3 $O/memcheck_single.out; tail -3 $O/memcheck_single.txt
echo "== sanitizer racecheck single"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool racecheck --log-file $O/racecheck_single.txt python tools/sanitize_target.py single > $O/racecheck_single.out 2>&1; echo "rc=$?"; tail -3 $O/racecheck_single.out; tail -3 $O/racecheck_single.txt
echo "== sanitizer memcheck multi (loop-back)"; XM_WATCHDOG_SCALE=500 timeout 900 compute-sanitizer --tool memcheck --log-file $O/memcheck_multi.txt python tools/sanitize_target.py multi > $O/memcheck_multi.out 2>&1; echo "rc=$?"; tail -3 $O/memcheck_multi.out; tail -3 $O/memcheck_multi.txt
echo "== ncu traffic"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xm_solve_kernel -c 1 --csv --log-file $O/traffic.csv python tools/ncu_target_big.py 0.25 > $O/traffic_target.log 2>&1; echo "rc=$?"; tail -1 $O/traffic_target.log
python tools/ncu_traffic.py $O/traffic.csv $O/traffic_target.log profiles/r02_solve_traffic.json; cp profiles/r02_solve_traffic.json $O/
echo "== bench ours"; timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
rm -f $O/*.ncu-rep.tmp; ls -la $O
