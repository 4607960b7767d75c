for v in sm0 sm4; do python tools/quick_time.py variants/libdex_$v.so; done
DEX_CHUNKS=2 python tools/quick_time.py
DEX_CHUNKS=2 DEX_SOLVER_MODE=0 python tools/quick_time.py
DEX_CHUNKS=3 DEX_SOLVER_MODE=0 python tools/quick_time.py
DEX_CHUNKS=4 DEX_SOLVER_MODE=0 python tools/quick_time.py
