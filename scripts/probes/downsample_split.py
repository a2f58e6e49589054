import ctypes as C, sys, numpy as np
sys.path.insert(0,'/root/repo')
import kanter_core_b200 as kc
from kanter_core_b200 import ResizeFilter
from kanter_core_b200._lib import call, kc_image
tp=kc.TextureProcessor.new(math_mode=kc.MATH_FAST); ctx=tp._ctx._h
big=kc.SlotImage.from_planes(tp,[np.random.default_rng(6).random((8192,8192),dtype=np.float32)])
def down():
    o=kc_image(); call("kc_resize",ctx,C.byref(big._im),1024,1024,int(ResizeFilter.Lanczos3),C.byref(o)); kc.SlotImage(tp._ctx,o)
for _ in range(3): down()
tp.synchronize()
ms,n=C.c_double(),C.c_uint64()
call("kc_context_set_timing",ctx,1); call("kc_context_timing_read",ctx,-1,C.byref(ms),C.byref(n))
for _ in range(5): down()
for kind,name in ((4,"V"),(5,"H")):
    call("kc_context_timing_read",ctx,kind,C.byref(ms),C.byref(n)); print(name, ms.value/max(1,n.value), n.value)
