"""Scratch: run the C3 inflate batch once with the phase-timing build of the library."""
import sys
sys.path.insert(0, ".")
from zipc_b200 import _lib
_lib.LIB_PATH = _lib.LIB_PATH.replace("libzipc_b200.so", "libzipc_b200_timing.so")
sys.argv = [sys.argv[0], sys.argv[1] if len(sys.argv) > 1 else "3000"]
exec(open("tools/inflate_matrix.py").read().split("L.zipc_b200_ctx_profile")[0])
run(); 
import torch; torch.cuda.synchronize()
