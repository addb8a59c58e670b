import time, ctypes, os, sys
t0=time.time()
cu = ctypes.CDLL("libcuda.so.1")
r = cu.cuInit(0); t1=time.time()
print("cuInit %.3f s rc=%d" % (t1-t0, r))
rt = ctypes.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else None
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "lordfast_b200", "liblfgpu.so"))
t2=time.time()
lib.lf_gpu_prewarm.restype=None
lib.lf_gpu_prewarm()
import numpy as np
pac=np.zeros(1000,dtype=np.uint8); ctx=ctypes.c_void_p()
lib.lf_gpu_init.argtypes=[ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
rc=lib.lf_gpu_init(ctypes.byref(ctx), pac.ctypes.data, 3000, None, 0)
t3=time.time()
print("prewarm + lf_gpu_init after cuInit: %.3f s rc=%d" % (t3-t2, rc))
