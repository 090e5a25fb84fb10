"""Host logic: the synthetic clip generator (torch == C), GOP partitioning, and the N-rank path:
world_size-2 gloo run where each rank encodes its own GOPs and the merged stream must equal the
1-rank stream byte for byte (SURVEY 8e invariant).  On CPU the per-rank encoder is the golden model;
the same merge code is what bench.py / a multi-GPU host uses around the CUDA library."""
import os
import sys

import numpy as np
import pytest
import torch

from cedarx_h264_encoder_b200 import partition, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("w,h,fmt", [(96, 80, 0), (854, 480, 0), (64, 48, 1)])
def test_synth_torch_equals_c(oracle, w, h, fmt):
    for t in (0, 1, 7, 33, 200):
        y, c = oracle.synth_frame(w, h, t, fmt)
        ty, tc = synth.synth_frame(w, h, t, fmt)
        assert np.array_equal(y, ty.numpy()) and np.array_equal(c, tc.numpy())


def test_synth_clip_packing(oracle):
    clip = synth.synth_clip(64, 48, [3, 4]).numpy()
    y, c = oracle.synth_frame(64, 48, 4)
    assert clip.shape == (2, 64 * 48 * 3 // 2)
    assert np.array_equal(clip[1, :64 * 48], y.reshape(-1)) and np.array_equal(clip[1, 64 * 48:], c.reshape(-1))


def test_partition_covers_every_frame_once():
    for nframes, gop, world in [(600, 60, 8), (601, 60, 4), (30, 25, 2), (7, 3, 3), (1200, 60, 8)]:
        seen = []
        for r in range(world):
            seen += partition.frames_for_rank(nframes, gop, r, world)
        assert sorted(seen) == list(range(nframes))
    assert partition.scaling_ceiling(600, 60, 8) == 5.0
    assert abs(partition.scaling_ceiling(1200, 60, 8) - 20 / 3) < 1e-9
    assert partition.scaling_ceiling(960, 60, 8) == 8.0


def _rank_main(rank, world, port, w, h, n, gop, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    stream, sizes = b"", []
    for g in partition.gops_for_rank(n, gop, rank, world):
        # a closed GOP is encoded by a fresh encoder; SPS/PPS only in front of stream frame 0
        enc = O.Encoder(O.make_config(w, h, qp=26, gop=gop, cabac=1, me_range=8))
        for t in partition.frames_of_gop(n, gop, g):
            y, c = O.synth_frame(w, h, t)
            b = enc.encode(y, c)
            if t % gop == 0 and t != 0:
                b = b[b.index(b"\x00\x00\x00\x01\x65"):]  # drop the parameter sets a fresh encoder emits
            stream += b
            sizes.append(len(b))
        enc.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, (stream, sizes))
    if rank == 0:
        merged = partition.merge_rank_streams(n, gop, world, [g[0] for g in gathered], [g[1] for g in gathered])
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gop_parallel_equals_single_stream(oracle):
    import torch.multiprocessing as mp
    w, h, n, gop = 64, 48, 11, 3
    single = oracle.Encoder(oracle.make_config(w, h, qp=26, gop=gop, cabac=1, me_range=8))
    want = b"".join(single.encode(*oracle.synth_frame(w, h, t)) for t in range(n))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, w, h, n, gop, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert merged == want
