import time, ctypes, os, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LF_INIT_TRACE"] = "1"
lib = ctypes.CDLL(os.path.join(root, "lordfast_b200", "liblfgpu.so"))
t2 = time.time()
lib.lf_gpu_prewarm.restype = None
lib.lf_gpu_prewarm()
buf = (ctypes.c_uint8 * 1000)(); ctx = ctypes.c_void_p()
lib.lf_gpu_init.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
rc = lib.lf_gpu_init(ctypes.byref(ctx), ctypes.addressof(buf), 3000, None, 0)
print("prewarm + lf_gpu_init (no cuInit before): %.3f s rc=%d CUDA_VISIBLE_DEVICES=%s" % (time.time() - t2, rc, os.environ.get("CUDA_VISIBLE_DEVICES")))
