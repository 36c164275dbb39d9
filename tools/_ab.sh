timeout 300 python -m pytest tests/test_gpu_frame_lookahead.py tests/test_golden.py tests/test_gpu_dropin_drivers.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-me --no-cpu-baseline --steps 10 2>/dev/null | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print('hpel-window', d['value'], d['kernel_ms_per_step'])"
