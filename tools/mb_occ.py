import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
for w in (24, 25, 26, 27, 28, 29):
    g = ctypes.c_double()
    _cabi.call("dcb_microbench", w, ctypes.byref(g))
    print("microbench %2d %10.1f" % (w, g.value))
