import sys, os
sys.path.insert(0, os.getcwd())
from dexterity_b200 import backend
backend.LIB_PATH = sys.argv[1]
sys.argv = sys.argv[:1]
exec(open("tools/profile_target.py").read())
