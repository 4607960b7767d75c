python -m pytest tests/test_physics_gpu.py tests/test_env_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/quick_time.py
DEX_NARROW_ONE_PASS=1 python tools/quick_time.py
for v in nt3 nt8 nb4 nb1; do python tools/quick_time.py variants/libdex_$v.so; done
