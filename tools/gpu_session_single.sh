#!/bin/bash
# One gpurun call on a 1-GPU box: the whole -m gpu suite, both bench arms, the block-CSR bench at Erdos-Renyi scale, ncu captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/${XM_SESSION_TAG:-r01_single}; mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $O/pytest_gpu.log
echo "== bench"; timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; cut -c1-400 $O/bench.json
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; cut -c1-300 $O/bench_ref.json
echo "== bsr ER-100k"; timeout 600 python tools/bench_bsr.py --solve --out $O/bsr_er100k.jsonl > $O/bsr.log 2>&1; echo "rc=$?"; cat $O/bsr_er100k.jsonl | cut -c1-900
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?"; tail -2 $O/smoke.log
echo "== ncu bsr"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:xm_ops_kernel -s 1 -c 2 -o $O/bsr_qy_full -f python tools/ncu_target_bsr.py > $O/ncu_bsr.log 2>&1; echo "rc=$?"; tail -2 $O/ncu_bsr.log
ncu -i $O/bsr_qy_full.ncu-rep --page raw --csv > $O/bsr_qy_full.raw.csv 2>/dev/null
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/ncu_bench.log 2>&1; echo "rc=$?"
ls -la $O
