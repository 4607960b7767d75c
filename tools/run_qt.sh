python tools/quick_time.py; python tools/quick_time.py; python -m pytest tests/test_adroit_gpu.py -m gpu -x -q 2>&1 | tail -2
