#!/bin/bash
# Where the GPU CLI's time goes, per door: runs glue/_build/x264ref_gpu on a 52-frame 1080p synthetic clip twice and prints the
# glue's own clock (X264DSP_GLUE_STATS: seconds inside each door, context creation, whole process).  Run on a GPU box:
#   gpurun -- 'bash tools/cli_time_split.sh'
python - <<'PY'
import sys, numpy as np
sys.path.insert(0,'tests')
import cpu_checkers as cc
np.concatenate([cc.synth_frame(1920,1080,i) for i in range(52)]).tofile('/tmp/syn_1920x1080.yuv')
PY
for i in 1 2; do
  X264DSP_GLUE_DEBUG=1 X264DSP_GLUE_STATS=/tmp/st.json glue/_build/x264ref_gpu /tmp/syn_1920x1080.yuv /tmp/o.264 2>&1 | tail -12
  cat /tmp/st.json
done
