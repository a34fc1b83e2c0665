#!/bin/bash
# The driver's own N = 8 invocation (both arms): python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 ... bench.py --gpus 8 --steps 20 --warmup 5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02_n8_driver; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
t0=$SECONDS
echo "== bench N=8 ours"; timeout 1200 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$? wall $((SECONDS-t0)) s"; grep -h '^{' $O/bench_n8.json | cut -c1-300; tail -2 $O/bench_n8.err
ls -la $O
