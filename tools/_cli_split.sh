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
