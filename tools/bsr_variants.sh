O=gpurun_out/r01_bsr_v6; mkdir -p $O
for v in 0 1 2 3; do echo "== XM_TUNE_BSR=$v"; XM_TUNE_BSR=$v timeout 300 python tools/bench_bsr.py --out $O/var$v.jsonl > $O/var$v.log 2>&1; python - <<PY
import json
for l in open("$O/var$v.jsonl"):
    d=json.loads(l); print("  r", d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"],3), round(d["ms_per_product_lockstep"],3), "frac", round(d["frac"],3))
PY
done
