#!/bin/bash
# One gpurun call on a W-GPU box: quick single-GPU regression, multi-GPU parity in both set-ups, sharded bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
W=${XM_SESSION_WORLD:-2}
O=gpurun_out/${XM_SESSION_TAG:-r01_multi}; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
echo "== single-GPU regression"; timeout 420 python -m pytest tests/test_recover.py tests/test_gpu_ops.py tests/test_gpu_solve.py -m gpu -x -q -k "not full_size" > $O/single.log 2>&1; echo "rc=$?"; tail -3 $O/single.log
echo "== multi (one process, threads)"; XM_TEST_WORLD=$W timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/multi_threads.log 2>&1; echo "rc=$?"; tail -5 $O/multi_threads.log
echo "== multi (torchrun, IPC)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/multi_torchrun.log 2>&1; echo "rc=$?"; tail -4 $O/multi_torchrun.log
echo "== bench --gpus $W"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 3 --warmup 3 > $O/bench_g$W.log 2>&1; echo "rc=$?"; grep '^{' $O/bench_g$W.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); c = d['config']
    print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 2), 'barrier_us', round(c['grid_barrier_us'], 2), 'qy_us free/lock', round(d['roofline']['qy_phase_alone']['us_per_product_free_running'], 2), round(d['roofline']['qy_phase_alone']['us_per_product_lockstep'], 2), 'sync_ms', round(c['in_kernel_ms']['grid_sync_wait'], 1), 'qy_ms', round(c['in_kernel_ms']['qy'], 1), 'its', c['tcg_iters_per_solve'], 'e2e', round(d['e2e']['value']))
"
if [ -n "$XM_SESSION_BIG" ]; then
  echo "== bench --gpus $W, $XM_SESSION_BIG cameras"; XM_BENCH_CAMERAS=$XM_SESSION_BIG OMP_NUM_THREADS=8 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $W --steps 2 --warmup 3 > $O/bench_big_g$W.log 2>&1; echo "rc=$?"; grep '^{' $O/bench_big_g$W.log | cut -c1-1500
fi
