import ctypes as C, sys
sys.path.insert(0,'.')
from abrsimulator_b200 import _lib
lib=_lib.load()
for k,n in ((10,'DADD'),(11,'DMUL'),(12,'DADD+mask')):
    g=C.c_double(); t=C.c_float()
    _lib.check(lib.abr_fp64_probe(C.c_int(k),C.c_int(256),C.byref(g),C.byref(t),None))
    print(n,'cycles per dependent op',g.value)
