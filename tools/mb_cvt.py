"""Throughput of the two fp32<->fp64 conversion directions alone and mixed with DFMA
(dcb_microbench 30..33, 0, 1): giga-CONVERSIONS per second on the whole GPU (the other operations of a step ride along)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
names = {0: "DFMA", 1: "cvt pair f32->f64->f32 (dependent)",
         30: "widenings   (2 per DFMA + 2 FFMA)", 31: "narrowings  (2 per 2 DFMA + FADD)",
         32: "widenings   (1 per DFMA + FFMA)", 33: "narrowings  (1 per 2 DFMA)",
         34: "STEPS of 1 widening + 4 DFMA", 35: "STEPS of 1 narrowing + 4 DFMA", 36: "STEPS of 4 DFMA alone"}
for w in (0, 1, 30, 31, 32, 33, 34, 35, 36):
    g = ctypes.c_double()
    _cabi.call("dcb_microbench", w, ctypes.byref(g))
    per_clk_sm = g.value * 1e9 / (148 * 1.965e9)
    print("microbench %2d %-40s %9.1f Gops/s = %5.1f ops/clk/SM" % (w, names[w], g.value, per_clk_sm))
