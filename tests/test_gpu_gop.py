"""SURVEY 8(f) N4 on the device: the lookahead pass of a sequence (x264dsp_lookahead_clip_host) feeds
x264dsp_slicetype_decide, and the frame types -- hence the GOPs the ranks share out -- are the ones the running
reference encoder chose for the same clip."""
import numpy as np
import pytest

import cpu_checkers as cc
from test_gop import CASES, reference_types

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("w,h,n,cut,keyint", CASES[:3])
def test_device_lookahead_decides_the_encoders_gops(pkg, ctx, w, h, n, cut, keyint):
    assert cc.ref() is not None, "oracle/_ref/libx264ref.so must travel to the GPU box"
    frames, got = reference_types(w, h, n, cut, keyint)
    want = np.array([d["frame_type"] for d in got], np.uint8)
    luma = np.stack([f[: w * h] for f in frames])
    _, _, sums = ctx.lookahead_clip_host(w, h, luma)
    ic = np.array([d["icost"] for d in got], np.int32)
    pc = np.array([d["pcost"] for d in got], np.int32)
    for k in range(n):
        if ic[k] >= 0:
            assert sums[k][pkg.LA_COST_INTRA] == ic[k], f"frame {k}: intra estimate"
        if k and pc[k] >= 0:
            assert sums[k][pkg.LA_COST_INTER] == pc[k], f"frame {k}: inter estimate"
    types = pkg.slicetype_decide(sums[:, pkg.LA_COST_INTRA], sums[:, pkg.LA_COST_INTER], keyint[0], got[0]["keyint_min"], keyint[2])
    assert np.array_equal(types, want), f"{types} vs the encoder's {want}"
    gops = pkg.gop_ranges(types)
    assert (len(gops) >= 2 or keyint[0] >= n) and sum(c for _, c in gops) == n
    assert [g for r in range(4) for g in pkg.gop_shard(gops, r, 4)] == gops
