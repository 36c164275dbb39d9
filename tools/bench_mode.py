#!/usr/bin/env python
"""bench.py with a forced lookahead kernel mapping (x264dsp_lookahead_select_kernel):
   python tools/bench_mode.py MODE [bench.py arguments]   -> prints the bench line"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
mode = int(sys.argv[1])
sys.argv = ["bench.py"] + sys.argv[2:]
import __graft_entry__ as ge   # noqa: E402
import bench                   # noqa: E402

pkg = ge.load_package()
_init = pkg.Context.__init__


def init(self, device=0):
    _init(self, device)
    self.lookahead_select_kernel(mode)


pkg.Context.__init__ = init
sys.exit(bench.main())
