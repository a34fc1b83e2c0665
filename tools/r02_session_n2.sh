#!/bin/bash
# 2-GPU rerun with the even-batch build: multi tests on two real GPUs + bench N=2 on the main workload
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_multi_w2b; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== multi tests (2 real GPUs, threads)"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_multi.log
echo "== multi tests (torchrun, one process per GPU)"; timeout 900 $TR --nproc-per-node 2 --master-port 29531 -m pytest tests/test_gpu_multi.py -m gpu -q -x > $O/pytest_multi_torchrun.log 2>&1; echo "rc=$?"; tail -4 $O/pytest_multi_torchrun.log
echo "== bench N=2"; timeout 900 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"; grep -h '^{' $O/bench_n2.json | cut -c1-300; tail -3 $O/bench_n2.err
ls -la $O
