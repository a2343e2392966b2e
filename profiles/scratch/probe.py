import ctypes, sys
sys.path.insert(0, '.')
from ternary_birch_b200 import cabi
L = cabi.lib()
g = ctypes.c_double(); t = ctypes.c_int(); a = ctypes.c_int()
print(L.tb_host_widen_probe(400_000_000, ctypes.byref(g), ctypes.byref(t), ctypes.byref(a)), g.value, t.value, a.value)
