import ctypes as C, numpy as np, time, sys, os
sys.path.insert(0,'/root/repo')
from opm_simulators_legacy_b200 import _lib
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian
n=int(sys.argv[1])
s=synth_blackoil_jacobian(n,n,n,perm="homogeneous")
rp, ci = s.rowptr.numpy(), s.colidx.numpy()
lib=_lib.load()
f=lib.opmgpu_debug_analyse_only
ip=C.POINTER(C.c_int)
f.argtypes=[C.c_int, ip, ip, C.c_int]; f.restype=C.c_double
for _ in range(2):
    t=f(len(rp)-1, rp.ctypes.data_as(ip), ci.ctypes.data_as(ip), 148)
    print("analyse_pattern ms", t)
