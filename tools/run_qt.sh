python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for m in 1 3; do DEX_SOLVER_MODE=$m python tools/quick_time.py; done
for v in s16 s18 s20; do DEX_SOLVER_MODE=3 python tools/quick_time.py variants/libdex_$v.so; done
DEX_SOLVER_MODE=3 DEX_SOLVER_SERIAL=1 python tools/quick_time.py
