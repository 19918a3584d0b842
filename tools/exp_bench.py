import os, sys
sys.path.insert(0, '.')
import leaf_pytorch_b200._native as N
e = os.environ.get("EXP", "")
if e:
    N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), "exp" + e, "libleafk.so")
import runpy
sys.argv = ["ab_bench.py"]
runpy.run_path("tools/ab_bench.py", run_name="__main__")
