"""Self test of hot_div / hot_sqrt on random operands. usage: [n] [spread]"""
import ctypes, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pumas_b200 as pb
lib = pb.library()
lib.pumas_b200_hot_math_selftest.argtypes = [ctypes.c_long, ctypes.c_ulong, ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 30
for spread in (4, 60, 300, 480):
    bad = (ctypes.c_ulonglong * 2)()
    rc = lib.pumas_b200_hot_math_selftest(n, 7 + spread, spread, bad)
    print(f"selftest n={n} spread=+-{spread}: rc={rc} divisions differ {bad[0]} square roots differ {bad[1]}", flush=True)
