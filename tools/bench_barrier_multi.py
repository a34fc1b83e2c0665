#!/usr/bin/env python
"""Cross-GPU synchronisation micro-benchmark (torchrun --nproc-per-node W tools/bench_barrier_multi.py [cameras] [r]):
average device time of (a) one reduction barrier of the persistent kernel over W GPUs and (b) one operand exchange
(tagged push of the rank's rows to every peer + unpack + local barrier), with thread 0's time per barrier segment
(leader CTA 0 and follower CTA 1; profile timers on, so absolute numbers are a little high)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from xm_code_b200 import capi, dist as xdist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1723
    r = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for profile in (False, True):
        h = capi.Handle(device=local, profile=profile)
        if world > 1:
            xdist.attach(h, N, r)
        rowptr = np.arange(N + 1, dtype=np.int32); col = np.arange(N, dtype=np.int32); vals = np.tile(np.eye(3), (N, 1, 1))
        h.set_q_bsr(rowptr, col, vals, 3)                 # any operator will do: only the barrier / exchange code runs
        iters = 2000
        us = h.bench_barrier(r, iters)
        c = h.debug_counters()
        both = h.bench_barrier(r, -iters)
        if rank == 0:
            seg = [x / iters / 1e3 for x in c]
            line = {"world": world, "cameras": N, "r": r, "profile_timers": profile, "barrier_us": round(us, 3),
                    "operand_exchange_us": round(both - us, 3)}
            if profile:
                line["leader_us"] = {"fence+slot": round(seg[0], 3), "gather_rank_slots": round(seg[1], 3), "send+wait_messages": round(seg[3], 3),
                                     "acquire_fence": round(seg[4], 3)}
                line["follower_us"] = {"fence+slot": round(seg[5], 3), "wait_messages": round(seg[6], 3), "acquire_fence": round(seg[7], 3)}
            print(json.dumps(line), flush=True)
        if world > 1:
            xdist.detach(h)
        else:
            h.close()


if __name__ == "__main__":
    main()
