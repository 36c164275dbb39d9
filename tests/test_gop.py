"""SURVEY 8(f) N4 -- lookahead-first GOP split.  x264dsp_slicetype_decide restates x264_slicetype_analyse / scenecut / the
key-frame rules of x264_slicetype_decide (encoder/slicetype.c:322-435, 508-537) for a whole sequence; it is pinned against
the frame types the RUNNING reference encoder chose (clips with scene cuts, short key-frame intervals, scenecut on and
off), first from the encoder's own frame costs, then from the costs of our lookahead (CPU oracle here, the device in
tests/test_gpu_gop.py) -- i.e. the chain lookahead -> types -> GOPs equals the reference's.  x264dsp_gop_ranges /
x264dsp_gop_shard turn the types into units of work for the ranks of a box; two gloo ranks check the split."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

import cpu_checkers as cc
from test_oracle_pframe import capture_encode
from test_sharding import oracle_analyse

CASES = [  # w, h, frames, scene cut at, (keyint_max, keyint_min, scenecut threshold)
    (176, 144, 24, 9, (10, 2, 40)), (176, 144, 30, 17, (50, 5, 20)), (208, 160, 20, 6, (8, 8, 40)),
    (176, 144, 26, 11, (12, 3, 0)), (176, 144, 22, 3, (250, 25, 40))]


def reference_types(w, h, n, cut, keyint):
    g, frames, got = capture_encode(w, h, n, cut, 0, 1, 26, 1, keyint=keyint, light=True)
    got = sorted(got, key=lambda d: d["i_frame"])
    assert [d["i_frame"] for d in got] == list(range(n))
    return frames, got


@pytest.mark.parametrize("w,h,n,cut,keyint", CASES)
def test_slicetype_decide_matches_the_running_encoder(pkg, w, h, n, cut, keyint):
    if cc.ref() is None:
        pytest.skip("oracle/_ref not built")
    frames, got = reference_types(w, h, n, cut, keyint)
    want = np.array([d["frame_type"] for d in got], np.uint8)
    assert got[0]["keyint_max"] == keyint[0] and got[0]["scenecut"] == keyint[2]
    kmin = got[0]["keyint_min"]                      # x264's validation may have clamped it
    # (1) from the encoder's own frame costs
    ic = np.array([d["icost"] for d in got], np.int32)
    pc = np.array([d["pcost"] for d in got], np.int32)
    types = pkg.slicetype_decide(ic, pc, keyint[0], kmin, keyint[2])
    assert np.array_equal(types, want), f"types {types} vs the encoder's {want}"
    if keyint[0] < n:
        assert (want != pkg.TYPE_P).sum() >= 2, "the clip never opened a second GOP"
    # (2) from OUR lookahead's costs: every cost the encoder computed is the one our pass computes
    luma = np.stack([f[: w * h] for f in frames])
    _, _, sums = oracle_analyse(w, h)(luma)
    for k in range(n if keyint[2] else 0):          # scenecut off: the encoder never runs its lookahead, there is nothing to compare
        if ic[k] >= 0:
            assert sums[k][pkg.LA_COST_INTRA] == ic[k], f"frame {k}: intra estimate {sums[k][pkg.LA_COST_INTRA]} vs {ic[k]}"
        if k and pc[k] >= 0:
            assert sums[k][pkg.LA_COST_INTER] == pc[k], f"frame {k}: inter estimate {sums[k][pkg.LA_COST_INTER]} vs {pc[k]}"
    types2 = pkg.slicetype_decide(sums[:, pkg.LA_COST_INTRA], sums[:, pkg.LA_COST_INTER], keyint[0], kmin, keyint[2])
    assert np.array_equal(types2, want)


def test_gop_ranges_and_shards_partition_exactly(pkg):
    rng = np.random.default_rng(7)
    for n in (1, 2, 9, 64, 300):
        types = np.where(rng.random(n) < 0.15, pkg.TYPE_IDR, pkg.TYPE_P).astype(np.uint8)
        types[0] = pkg.TYPE_IDR
        gops = pkg.gop_ranges(types)
        assert sum(c for _, c in gops) == n and gops[0][0] == 0
        for (f, c), nxt in zip(gops, gops[1:] + [(n, 0)]):
            assert f + c == nxt[0] and types[f] != pkg.TYPE_P and (types[f + 1: f + c] == pkg.TYPE_P).all()
        longest = max(c for _, c in gops)
        for world in (1, 2, 3, 4, 8):
            seen, loads = [], []
            for rank in range(world):
                mine = pkg.gop_shard(gops, rank, world)
                seen += mine
                loads.append(sum(c for _, c in mine))
            assert seen == gops, (n, world)                               # whole GOPs, in order, nothing twice
            assert max(loads) <= n / world + longest, (loads, n, world)     # within one GOP of the fair share
    with pytest.raises(pkg.X264DspError):
        pkg.gop_shard([(0, 4)], 2, 2)


def _rank_main(rank, world, port, w, h, n, cut, keyint, q):
    import torch
    import torch.distributed as dist
    import conftest
    pkg = conftest.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    luma = np.stack([pkg.synth_frame(w, h, i, cut_frame=cut, luma_only=True) for i in range(n)])

    def gather(a):
        t = torch.from_numpy(np.ascontiguousarray(a).astype(np.int32))
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([t.shape[0]], dtype=torch.int64))
        m = int(max(s.item() for s in sizes))
        pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype)
        pad[: t.shape[0]] = t
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        return np.concatenate([p[: int(s.item())].numpy() for p, s in zip(parts, sizes)]).astype(a.dtype)

    # stage 1: the lookahead of the sequence, sharded by frame range, frame costs gathered (the only collective)
    first, mvs, costs, sums = pkg.lookahead_sharded(oracle_analyse(w, h), luma, rank, world)
    all_sums = gather(sums)
    # stage 2: every rank decides the same types; stage 3: its GOPs
    types = pkg.slicetype_decide(all_sums[:, pkg.LA_COST_INTRA], all_sums[:, pkg.LA_COST_INTER], *keyint)
    gops = pkg.gop_ranges(types)
    q.put((rank, types, gops, pkg.gop_shard(gops, rank, world)))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gop_split_two_ranks_gloo(pkg):
    import torch.multiprocessing as mp
    w, h, n, cut, keyint, world = 176, 144, 18, 9, (10, 2, 40), 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_rank_main, args=(r, world, port, w, h, n, cut, keyint, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    luma = np.stack([pkg.synth_frame(w, h, i, cut_frame=cut, luma_only=True) for i in range(n)])
    _, _, sums = oracle_analyse(w, h)(luma)
    want = pkg.slicetype_decide(sums[:, pkg.LA_COST_INTRA], sums[:, pkg.LA_COST_INTER], *keyint)
    for g in got:
        assert np.array_equal(g[1], want)                  # sharded lookahead + gather -> the single-process decision
    assert got[0][3] + got[1][3] == got[0][2] and len(got[0][3]) > 0 and len(got[1][3]) > 0
    assert want[9] != pkg.TYPE_P, "the scene cut did not open a GOP"
