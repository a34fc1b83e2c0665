#!/bin/bash
# One gpurun call on a 2-GPU box: single-GPU regression (quick subset), multi-GPU parity in both set-ups, sharded bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r01_multi; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv > $O/gpus.txt 2>&1
W=${XM_SESSION_WORLD:-2}
echo "== single-GPU regression"; timeout 420 python -m pytest tests/test_gpu_ops.py tests/test_gpu_solve.py -m gpu -x -q -k "not full_size" > $O/single.log 2>&1; echo "rc=$?"; tail -3 $O/single.log
echo "== multi (one process, threads)"; XM_TEST_WORLD=$W timeout 420 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/multi_threads.log 2>&1; echo "rc=$?"; tail -15 $O/multi_threads.log
echo "== multi (torchrun, IPC)"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29511 -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/multi_torchrun.log 2>&1; echo "rc=$?"; tail -15 $O/multi_torchrun.log
echo "== bench --gpus $W"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $W --steps 3 --warmup 3 > $O/bench_g$W.log 2>&1; echo "rc=$?"; tail -2 $O/bench_g$W.log
echo "== bench --gpus 1"; timeout 300 python bench.py --steps 3 --warmup 3 > $O/bench_g1.log 2>&1; echo "rc=$?"; tail -1 $O/bench_g1.log
