python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_time.py
DEX_NO_PDL=1 python tools/quick_time.py
python tools/quick_time.py
DEX_NO_PDL=1 python tools/quick_time.py
