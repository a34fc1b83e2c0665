#!/bin/bash
# Round-2 GPU session 5 (one B200): tests again, block-CSR with the non-inlined row product, dense ranks probe (two cameras per warp).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r02_s5}; mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -14 $O/pytest_gpu.log
echo "== bsr ER-100k (1024 threads)"; timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt1024.jsonl > $O/bsr1024.log 2>&1; echo "rc=$?"
echo "== bsr ER-100k (512 threads)"; XM_TUNE_BSR_NT=512 timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k_nt512.jsonl > $O/bsr512.log 2>&1; echo "rc=$?"
python - <<PY
import json
for f in ("$O/bsr_er100k_nt1024.jsonl", "$O/bsr_er100k_nt512.jsonl"):
    try:
        for l in open(f):
            d = json.loads(l); s = d.get("solve", {})
            print(f.split("er100k_")[-1], "r", d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"], 3), round(d["ms_per_product_lockstep"], 3), "frac", round(d["frac"], 3), "solve it/s", round(s.get("tcg_iters_per_s", 0), 1), "ms/prod in solve", round(s.get("ms_per_qy_product_in_solve", 0), 3))
    except Exception as e:
        print(f, e)
PY
echo "== ncu full: block-CSR Q.Y (r = 10)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:xm_ops_kernel -s 1 -c 1 -o $O/bsr_qy_full -f python tools/ncu_target_bsr.py > $O/ncu_bsr.log 2>&1; echo "rc=$?"; tail -1 $O/ncu_bsr.log
ncu -i $O/bsr_qy_full.ncu-rep --page raw --csv > $O/bsr_qy_full.raw.csv 2>/dev/null; rm -f $O/bsr_qy_full.ncu-rep
echo "== dense ranks probe"; timeout 600 python tools/dense_rank_probe.py > $O/dense_ranks.jsonl 2> $O/dense_ranks.err; echo "rc=$?"; cat $O/dense_ranks.jsonl; tail -3 $O/dense_ranks.err
ls -la $O
