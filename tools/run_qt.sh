python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_time.py
DEX_WORK_FLAT=1 python tools/quick_time.py
