DEX_SOLVER_MODE=1 python tools/quick_time.py
DEX_SOLVER_MODE=4 python tools/quick_time.py
DEX_SOLVER_MODE=4 python tools/quick_time.py
DEX_SOLVER_MODE=1 python tools/quick_time.py
