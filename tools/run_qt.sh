DEX_SOLVER_MODE=1 python tools/quick_time.py
DEX_SOLVER_MODE=5 python tools/quick_time.py
DEX_SOLVER_MODE=5 DEX_SOLVER_SERIAL=1 python tools/quick_time.py
DEX_SOLVER_MODE=2 python tools/quick_time.py
