"""GOP-parallel partitioning across GPUs.  A closed GOP (IDR + the P frames up to the next
IDR: kernel/cedar.c:1047-1050, 1193-1195) references nothing outside itself, rate control does not
exist (CQP), and SPS/PPS are emitted only before stream frame 0 (cedar.c:1058-1061), so GOP g can
be encoded anywhere and the per-rank bytestreams concatenated in GOP order equal the 1-GPU stream.
No collective is needed on the data path; ranks only exchange their finished bytes."""


def num_gops(nframes, gop):
    return (nframes + gop - 1) // gop


def gops_for_rank(nframes, gop, rank, world):
    """Round-robin: GOP g -> rank g % world."""
    return [g for g in range(num_gops(nframes, gop)) if g % world == rank]


def frames_of_gop(nframes, gop, g):
    return list(range(g * gop, min((g + 1) * gop, nframes)))


def frames_for_rank(nframes, gop, rank, world):
    out = []
    for g in gops_for_rank(nframes, gop, rank, world):
        out.extend(frames_of_gop(nframes, gop, g))
    return out


def scaling_ceiling(nframes, gop, world):
    """Best possible speed-up with whole-GOP granularity: G / ceil(G / N)  (SURVEY M8)."""
    g = num_gops(nframes, gop)
    return g / ((g + world - 1) // world)


def merge_rank_streams(nframes, gop, world, rank_streams, rank_sizes):
    """rank_streams[r] = bytes of rank r's GOPs back to back (in its own GOP order);
    rank_sizes[r] = per-frame byte counts of that stream.  Returns the stream in display order."""
    cursors = [0] * world
    frame_cur = [0] * world
    out = []
    for g in range(num_gops(nframes, gop)):
        r = g % world
        nf = len(frames_of_gop(nframes, gop, g))
        nbytes = int(sum(rank_sizes[r][frame_cur[r]:frame_cur[r] + nf]))
        out.append(rank_streams[r][cursors[r]:cursors[r] + nbytes])
        cursors[r] += nbytes
        frame_cur[r] += nf
    return b"".join(out)
