"""The shippable drop-in: glue/_build/x264ref_gpu is the UNMODIFIED reference CLI (x264.c, input.c, output.c and every
library source compiled where they lie) linked with the three C files of glue/ and libx264dsp_b200.so -- no Python in
the loop.  With the glue installed (the default) every door of glue/x264dsp_doors.c is served by the device:

  x264_frame_init_lowres, the in-loop filter (x264_frame_deblock_row, x264_frame_expand_border, x264_frame_filter,
  x264_frame_expand_border_filtered), x264_slicetype_frame_cost, x264_me_search_ref, x264_mb_mc,
  x264_macroblock_probe_pskip, x264_macroblock_encode

and the bitstream must be byte-identical to the reference CLI's (oracle/_ref/x264ref) on the config-1 clip (CIF, 30
frames, the reference's defaults) and on a 1080p clip; the doors' own counters must show that no eligible call fell back
to the reference's code."""
import json
import os
import subprocess

import numpy as np
import pytest

import cpu_checkers as cc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_CLI = os.path.join(ROOT, "glue", "_build", "x264ref_gpu")


def run_cli(exe, src, out, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    subprocess.run([exe, src, out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env, timeout=1500)
    return np.fromfile(out, np.uint8)


def make_clip(tmp_path, w, h, n, cut):
    clip = np.concatenate([cc.synth_frame(w, h, i, cut_frame=cut) for i in range(n)])
    src = str(tmp_path / f"syn_{w}x{h}.yuv")
    clip.tofile(src)
    return src


def test_glue_cli_with_doors_closed_is_the_reference(tmp_path):
    """X264DSP_GLUE=0: nothing installed, no device opened -- the binary is the reference CLI (runs without a GPU)"""
    if not (os.path.exists(GPU_CLI) and os.path.exists(cc.REF_CLI)):
        pytest.skip("glue/_build/x264ref_gpu or oracle/_ref/x264ref not built")
    src = make_clip(tmp_path, 176, 144, 6, 3)
    want = run_cli(cc.REF_CLI, src, str(tmp_path / "ref.264"))
    got = run_cli(GPU_CLI, src, str(tmp_path / "closed.264"), {"X264DSP_GLUE": "0"})
    assert want.size > 0 and np.array_equal(want, got)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,cut", [(352, 288, 30, 17), (1920, 1080, 3, -1)])
def test_glue_cli_p_slices_on_the_device(tmp_path, w, h, n, cut):
    """the default mode: the macroblock loop of every slice (x264_macroblock_analyse + x264_macroblock_encode) is ONE device
    call per frame -- x264dsp_p_frames_dev for P slices, x264dsp_i_frames_dev for I slices -- and the host keeps the entropy coder"""
    assert os.path.exists(GPU_CLI), "glue/_build/x264ref_gpu must travel to the GPU box (make -C glue)"
    assert os.path.exists(cc.REF_CLI), "oracle/_ref/x264ref must travel to the GPU box (make -C oracle ref)"
    src = make_clip(tmp_path, w, h, n, cut)
    want = run_cli(cc.REF_CLI, src, str(tmp_path / "ref.264"))
    stats_path = str(tmp_path / "stats.json")
    got = run_cli(GPU_CLI, src, str(tmp_path / "gpu.264"), {"X264DSP_GLUE_STATS": stats_path})
    st = json.load(open(stats_path))
    print("GLUESTATS pframe", f"{w}x{h}x{n}", json.dumps(st))
    assert want.size > 0 and got.size == want.size and np.array_equal(want, got), \
        f"bitstreams differ: {got.size} vs {want.size} bytes"
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    seen, served, served_mbs = st["p_slices"]
    assert seen == served == st["p_frames"] >= n - 2, st             # every P slice of the clip, none declined
    assert served_mbs == served * mbs, st
    assert st["me_search"] == 0 and st["probe_pskip"] == 0 and st["mb_mc"] == 0, st    # no per-macroblock round trips left
    iseen, iserved, iserved_mbs = st["i_slices"]
    assert iseen == iserved == st["i_frames"] == n - served >= 1 and iserved_mbs == iserved * mbs, st   # nor for the I slices:
    assert st["macroblock_encode"] == 0, st                    # x264_mb_analyse_intra and the intra coding ran on the device too
    assert st["kernel_launches"] < 40 * n, st                  # a handful of launches per frame
    assert st["lowres"] == n and st["inloop_filter"] == n and st["lookahead_cost"] == n - 1, st


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,cut", [(352, 288, 30, 17), (1920, 1080, 3, -1)])
def test_glue_cli_bitstream_identical(tmp_path, w, h, n, cut):
    assert os.path.exists(GPU_CLI), "glue/_build/x264ref_gpu must travel to the GPU box (make -C glue)"
    assert os.path.exists(cc.REF_CLI), "oracle/_ref/x264ref must travel to the GPU box (make -C oracle ref)"
    src = make_clip(tmp_path, w, h, n, cut)
    want = run_cli(cc.REF_CLI, src, str(tmp_path / "ref.264"))
    stats_path = str(tmp_path / "stats.json")
    got = run_cli(GPU_CLI, src, str(tmp_path / "gpu.264"), {"X264DSP_GLUE_STATS": stats_path, "X264DSP_GLUE_PFRAME": "0"})
    st = json.load(open(stats_path))
    print("GLUESTATS", f"{w}x{h}x{n}", json.dumps(st))
    assert want.size > 0 and got.size == want.size and np.array_equal(want, got), \
        f"bitstreams differ: {got.size} vs {want.size} bytes"
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    assert st["lowres"] == n and st["inloop_filter"] == n and st["lookahead_cost"] == n - 1, st
    assert st["deblocked_frames"] == n, st
    for door in ("door_me", "door_mbenc", "door_pskip", "door_mbmc"):
        entered, eligible, served = st[door]
        assert served == eligible > 0, f"{door}: {st[door]} -- eligible calls fell back to the reference's code"
    assert st["me_search"] == st["door_me"][2] and st["door_me"][0] == st["door_me"][1], st
    assert st["macroblock_encode"] == st["door_mbenc"][2] >= mbs, st
    assert st["probe_pskip"] == st["door_pskip"][2] and st["mb_mc"] == st["door_mbmc"][2], st
    assert st["kernel_launches"] > 3 * n, st


API_GPU = os.path.join(ROOT, "glue", "_build", "x264api_gpu")


def run_api(src, out, w, h, psub, me, subme, qp, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    subprocess.run([API_GPU, src, out, str(w), str(h), str(psub), str(me), str(subme), str(qp)], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env, timeout=1500)
    return np.fromfile(out, np.uint8)


def test_api_example_with_doors_closed_runs_partitions(tmp_path):
    """glue/x264dsp_api_example.c on the CPU alone: X264_ANALYSE_PSUB16x16 reaches the reference's analysis (the stream changes)"""
    if not os.path.exists(API_GPU):
        pytest.skip("glue/_build/x264api_gpu not built")
    src = make_clip(tmp_path, 176, 144, 6, 3)
    a = run_api(src, str(tmp_path / "a.264"), 176, 144, 0, 1, 5, 26, {"X264DSP_GLUE": "0"})
    b = run_api(src, str(tmp_path / "b.264"), 176, 144, 1, 1, 5, 26, {"X264DSP_GLUE": "0"})
    assert a.size > 0 and b.size > 0 and not np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,cut,psub,me,subme,qp", [(352, 288, 12, 7, 1, 1, 5, 26), (352, 288, 8, -1, 1, 0, 1, 30),
                                                       (1920, 1080, 3, -1, 1, 1, 4, 28), (352, 288, 8, 5, 0, 1, 2, 26)])
def test_api_user_with_partitions_on_the_device(tmp_path, w, h, n, cut, psub, me, subme, qp):
    """an application of the library API with analyse.inter = X264_ANALYSE_PSUB16x16: every P slice goes through
    x264dsp_p_frames_part_dev, the bitstream is the reference library's byte for byte"""
    assert os.path.exists(API_GPU), "glue/_build/x264api_gpu must travel to the GPU box (make -C glue)"
    src = make_clip(tmp_path, w, h, n, cut)
    want = run_api(src, str(tmp_path / "ref.264"), w, h, psub, me, subme, qp, {"X264DSP_GLUE": "0"})
    stats_path = str(tmp_path / "stats.json")
    got = run_api(src, str(tmp_path / "gpu.264"), w, h, psub, me, subme, qp, {"X264DSP_GLUE_STATS": stats_path})
    st = json.load(open(stats_path))
    print("GLUESTATS api", f"{w}x{h}x{n} psub={psub}", json.dumps(st))
    assert want.size > 0 and got.size == want.size and np.array_equal(want, got), f"bitstreams differ: {got.size} vs {want.size} bytes"
    mbs = ((w + 15) // 16) * ((h + 15) // 16)
    seen, served, served_mbs = st["p_slices"]
    assert seen == served == st["p_frames"] >= n - 2 and served_mbs == served * mbs, st
    assert st["me_search"] == 0 and st["probe_pskip"] == 0 and st["mb_mc"] == 0 and st["macroblock_encode"] == 0, st
