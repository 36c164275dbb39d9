"""The C ABI from its intended host language: examples/lookahead_host.c is a plain C99 program (gcc, no CUDA / C++ /
Python on its side) that links against libx264dsp_b200.so through include/x264dsp_b200.h, runs the lookahead of a
synthetic clip through the host-level call and through the frame-batched device entry points, and prints the frame
costs.  Here it is built (if the in-tree binary is missing), run, and its numbers are compared with the CPU oracle."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import cpu_checkers as cc
from cpu_checkers import ptr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "_build", "lookahead_host")


@pytest.mark.parametrize("w,h,n", [(352, 288, 6), (200, 120, 4)])
def test_c_host_program_matches_oracle(pkg, w, h, n):
    if not os.path.exists(EXE):
        subprocess.run(["make", "-C", ROOT, "examples"], check=True)
    run = subprocess.run([EXE, str(w), str(h), str(n)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, f"rc {run.returncode}: {run.stderr[-400:]} {run.stdout[-400:]}"
    out = json.loads(run.stdout.strip().splitlines()[-1])
    assert out["consistent"] and out["launches"] >= 4 and out["frames"] == n

    o = cc.oracle()
    g = cc.oracle_geom(w, h)
    slots = []
    for i in range(n):
        pic = np.concatenate([pkg.synth_frame(w, h, i, luma_only=True), np.full(w * h // 2, 128, np.uint8)])
        s = np.zeros(g.slot_bytes, np.uint8)
        o.xo_frame_load_i420(C.byref(g), ptr(pic), ptr(s))
        o.xo_frame_init_lowres(C.byref(g), ptr(s))
        slots.append(s)
    for i in range(n):
        mv = np.zeros((g.mb_count, 2), np.int16)
        cost = np.zeros(g.mb_count, np.int32)
        sums = np.zeros(8, np.int32)
        o.xo_lookahead_frame_cost(C.byref(g), ptr(slots[i]), ptr(slots[i - 1]) if i else None, 1,
                                  ptr(mv, cc.i16p), ptr(cost, cc.i32p), ptr(sums, cc.i32p), None)
        got = out["frames_out"][i]
        assert got["cost_intra"] == int(sums[1]), f"frame {i}: intra cost {got['cost_intra']} vs oracle {int(sums[1])}"
        if i:
            assert got["cost_inter"] == int(sums[0]) and got["intra_mbs"] == int(sums[2]), f"frame {i}: {got} vs {sums[:3]}"
            assert got["mv_abs_sum"] == int(np.abs(mv.astype(np.int64)).sum()), f"frame {i}: MVs"
