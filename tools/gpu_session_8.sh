#!/bin/bash
# One gpurun call on an 8-GPU box: the sharded bench at 8 GPUs and the Erdos-Renyi block-CSR config (BASELINE config 5).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
W=${XM_SESSION_WORLD:-8}
O=gpurun_out/${XM_SESSION_TAG:-r01_g8}; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== bench --gpus $W"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 3 --warmup 3 > $O/bench_g$W.log 2>&1; echo "rc=$?"; grep '^{' $O/bench_g$W.log > $O/bench_g$W.json; python - <<PY
import json
for l in open("$O/bench_g$W.json"):
    d = json.loads(l); c = d['config']
    print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'barrier_us', round(c['grid_barrier_us'], 2), 'qy_us free/lock', round(d['roofline']['qy_phase_alone']['us_per_product_free_running'], 2), round(d['roofline']['qy_phase_alone']['us_per_product_lockstep'], 2), 'sync_ms', round(c['in_kernel_ms']['grid_sync_wait'], 1), 'qy_ms', round(c['in_kernel_ms']['qy'], 1), 'its', c['tcg_iters_per_solve'], 'e2e', round(d['e2e']['value']), 'obj', c['final_objective'])
PY
tail -3 $O/bench_g$W.log | cut -c1-300
echo "== bsr ER-100k on $W"; OMP_NUM_THREADS=3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29515 tools/bench_bsr.py --solve --out $O/bsr_er100k_g$W.jsonl > $O/bsr_g$W.log 2>&1; echo "rc=$?"; python - <<PY
import json
for l in open("$O/bsr_er100k_g$W.jsonl"):
    d=json.loads(l); print(d["rank_r"], "ms free/lock", round(d["ms_per_product_free_running"],3), round(d["ms_per_product_lockstep"],3), "frac/gpu", round(d["frac"],3), "solve it/s", round(d["solve"]["tcg_iters_per_s"]), "ms/prod in solve", round(d["solve"]["ms_per_qy_product_in_solve"],3), d["solve"]["exit"])
PY
tail -3 $O/bsr_g$W.log | cut -c1-300
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29516 tools/bench_barrier_multi.py 2>&1 | grep "^{" | tee $O/barrier_w$W.jsonl | cut -c1-400
