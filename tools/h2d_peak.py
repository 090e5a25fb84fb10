#!/usr/bin/env python3
"""Host-link ceiling for the end-to-end figure: pinned host -> device copies of one rank's raw clip (1.88 GB = 600 frames of
1080p NV12) on every rank at once, nothing else running.  Run under torchrun with N = 1, 2, 4, 8; rank 0 appends one JSON
line per N to the file given (profiles/r02_h2d_peak.jsonl).  frames/s ceiling = aggregate GB/s / 3.1334 MB per frame.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \\
        tools/h2d_peak.py out.jsonl"""
import json
import os
import sys

import torch
import torch.distributed as dist

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 600 * 1920 * 1080 * 3 // 2
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(7)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
res = {}
for chunk_mb in (0, 3, 32):  # 0 = one copy; 3 = frame-sized copies (what clip_upload issues); 32 MB chunks
    step = nbytes if chunk_mb == 0 else chunk_mb * 1000 * 1000 if chunk_mb != 3 else 1920 * 1080 * 3 // 2
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for o in range(0, nbytes, step):
            dev[o:o + step].copy_(host[o:o + step], non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        best = float(ms.item()) if rep == 0 else min(best, float(ms.item()))
    res["copy_%s" % ("whole" if chunk_mb == 0 else "%dMB" % chunk_mb)] = round(nbytes * world / (best * 1e-3) / 1e9, 2)
if rank == 0:
    line = {"tool": "h2d_peak", "n_gpus": world, "bytes_per_rank": nbytes, "aggregate_GBps": res,
            "frames_per_s_ceiling_1080p_nv12": round(max(res.values()) * 1e9 / (1920 * 1080 * 1.5), 1),
            "cpu_affinity": sorted(os.sched_getaffinity(0))[:4] + ["..."], "host_cores": os.cpu_count()}
    print(json.dumps(line))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "a") as f:
            f.write(json.dumps(line) + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
