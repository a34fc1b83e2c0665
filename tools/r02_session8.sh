#!/bin/bash
# Round-2 GPU session 8 (one B200): block-CSR row product with a by-value device view + explicit global loads — tests, ER-100k, traffic recapture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_s8; mkdir -p $O
echo "== pytest (bsr, multi, ops, parity)"; timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_parity_r2.py tests/test_gpu_multi.py tests/test_certificate_solver.py -m gpu -q > $O/pytest.log 2>&1; echo "rc=$?"; tail -4 $O/pytest.log
echo "== bsr ER-100k (default)"; timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k.jsonl > $O/bsr.log 2>&1; echo "rc=$?"
python - <<PY
import json
for l in open("$O/bsr_er100k.jsonl"):
    d = json.loads(l); s = d.get("solve", {})
    print("r", d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"], 3), round(d["ms_per_product_lockstep"], 3), "frac", round(d["frac"], 3), "solve it/s", round(s.get("tcg_iters_per_s", 0), 1), "ms/prod in solve", round(s.get("ms_per_qy_product_in_solve", 0), 3))
PY
echo "== ncu traffic"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:xm_solve_kernel -c 1 --csv --log-file $O/traffic.csv python tools/ncu_target_big.py 0.25 > $O/traffic_target.log 2>&1; echo "rc=$?"; tail -1 $O/traffic_target.log
python tools/ncu_traffic.py $O/traffic.csv $O/traffic_target.log $O/r02_solve_traffic.json
ls -la $O
