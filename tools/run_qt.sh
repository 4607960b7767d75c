python -m pytest tests/test_physics_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_time.py
python tools/quick_time.py variants/libdex_k0h.so
