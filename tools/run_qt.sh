python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/quick_time.py
python tools/gpu_reach_bench.py 2>&1 | tail -4
HAND=adroit_hand python tools/gpu_reach_bench.py 2>&1 | tail -4
