#!/bin/bash
# Round-2 multi-GPU session: `gpurun --gpus W -- 'XM_W=W bash tools/r02_session_multi.sh'`
#   W = 2: the multi-GPU test file on two real GPUs (one process, a thread per GPU) + bench at N = 2 on the main workload and on C3
#   W = 8: bench at N = 8 and N = 4 (main workload incl. the ER-100k block-CSR sub-record)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
W=${XM_W:-2}
O=gpurun_out/${XM_SESSION_TAG:-r02_multi_w$W}; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$W" = "2" ]; then
  echo "== multi tests (2 real GPUs, threads)"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=8 > $O/pytest_multi.log 2>&1; echo "rc=$?"; tail -14 $O/pytest_multi.log
  echo "== bench N=2"; timeout 900 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"; grep -h '^{' $O/bench_n2.json | cut -c1-400; tail -3 $O/bench_n2.err
  echo "== bench C3 N=2"; XM_BENCH_CAMERAS=1723 XM_BENCH_BSR=0 timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_c3_n2.json 2> $O/bench_c3_n2.err; echo "rc=$?"; grep -h '^{' $O/bench_c3_n2.json | cut -c1-400; tail -3 $O/bench_c3_n2.err
  echo "== bench C3 N=1"; XM_BENCH_CAMERAS=1723 XM_BENCH_BSR=0 timeout 600 python bench.py --steps 5 --warmup 3 --cpu-seconds 2 > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; echo "rc=$?"; grep -h '^{' $O/bench_c3_n1.json | cut -c1-300
else
  for n in 8 4; do
    echo "== bench N=$n"; timeout 900 $TR --nproc-per-node $n --master-port $((29520 + n)) bench.py --gpus $n --steps 2 --warmup 1 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "rc=$?"; grep -h '^{' $O/bench_n$n.json | cut -c1-400; tail -3 $O/bench_n$n.err
  done
fi
ls -la $O
