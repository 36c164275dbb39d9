"""Host-side multi-GPU logic on CPU: the frame-range split (x264dsp_frame_range) and the sharded
lookahead pass with its optional gather, run as two `gloo` ranks.  The per-rank single-device pass
is played by the CPU oracle here (the test checks the SHARDING, not the kernels): stitched results
of the two ranks must equal the one-process result over the whole sequence."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr, i16p, i32p


def test_frame_range_partitions_exactly(pkg):
    for n in (0, 1, 2, 7, 8, 64, 65, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for rank in range(world):
                first, count, need_prev = pkg.frame_range(n, rank, world)
                seen += list(range(first, first + count))
                assert need_prev == (count > 0 and first > 0)
                assert count in (n // world, n // world + 1)
            assert seen == list(range(n)), (n, world)
    with pytest.raises(pkg.X264DspError):
        pkg.frame_range(8, 2, 2)
    with pytest.raises(pkg.X264DspError):
        pkg.frame_range(8, 0, 0)


def oracle_analyse(w, h):
    """stand-in for ctx.lookahead_clip_host on a CPU box: same contract, computed by the oracle"""
    o = cc.oracle()
    g = cc.oracle_geom(w, h)

    def analyse(luma):
        n = luma.shape[0]
        slots = [np.zeros(g.slot_bytes, np.uint8) for _ in range(n)]
        for i in range(n):
            pic = np.concatenate([luma[i], np.full(w * h // 2, 128, np.uint8)])
            o.xo_frame_load_i420(C.byref(g), ptr(pic), ptr(slots[i]))
            o.xo_frame_init_lowres(C.byref(g), ptr(slots[i]))
        mvs = np.zeros((n, g.mb_count, 2), np.int16)
        costs = np.zeros((n, g.mb_count), np.int32)
        sums = np.zeros((n, 8), np.int32)
        for i in range(n):
            o.xo_lookahead_frame_cost(C.byref(g), ptr(slots[i]), ptr(slots[i - 1]) if i else None, 1,
                                      ptr(mvs[i], i16p), ptr(costs[i], i32p), ptr(sums[i], i32p), None)
        return mvs, costs, sums
    return analyse


def _rank_main(rank, world, port, w, h, n, q):
    import torch
    import torch.distributed as dist
    import conftest
    pkg = conftest.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    luma = np.stack([pkg.synth_frame(w, h, i, luma_only=True) for i in range(n)])

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a).astype(np.int32))      # gloo has no int16 all_gather
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64))
        m = int(max(s.item() for s in sizes))
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype)
        pad[: t.shape[0]] = t
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        return np.concatenate([p[: int(s.item())].numpy() for p, s in zip(parts, sizes)]).astype(a.dtype)

    mbc = cc.oracle_geom(w, h).mb_count
    first, mvs, costs, sums = pkg.lookahead_sharded(oracle_analyse(w, h), luma, rank, world, mb_count=mbc)
    _, gm, gc, gs = pkg.lookahead_sharded(oracle_analyse(w, h), luma, rank, world, gather=gather, mb_count=mbc)
    q.put((rank, first, mvs, costs, sums, gm, gc, gs))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("n,world", [(7, 2), (2, 3)])       # (2, 3): more ranks than frames, one rank owns nothing
def test_lookahead_sharded_two_ranks_gloo(pkg, n, world):
    import torch.multiprocessing as mp
    w, h = 176, 144
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_rank_main, args=(r, world, port, w, h, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    luma = np.stack([pkg.synth_frame(w, h, i, luma_only=True) for i in range(n)])
    want_m, want_c, want_s = oracle_analyse(w, h)(luma)
    stitched = [np.concatenate([g[k] for g in got]) for k in (2, 3, 4)]
    assert got[0][1] == 0 and got[1][1] == pkg.frame_range(n, 1, world)[0]
    assert all(g[2].shape[1:] == want_m.shape[1:] and g[3].shape[1:] == want_c.shape[1:] for g in got)     # empty ranks too
    assert np.array_equal(stitched[0], want_m) and np.array_equal(stitched[1], want_c)
    assert np.array_equal(stitched[2][:, :5], want_s[:, :5])
    for g in got:       # gathered view: every rank holds the full-sequence tables
        assert np.array_equal(g[5], want_m) and np.array_equal(g[6], want_c) and np.array_equal(g[7][:, :5], want_s[:, :5])
