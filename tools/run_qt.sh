python tools/quick_time.py
python tools/quick_time.py variants/libdex_mst.so
DEX_SOLVER_MODE=0 python tools/quick_time.py
DEX_SOLVER_MODE=0 python tools/quick_time.py variants/libdex_mst.so
