import ctypes, os, sys
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libvosk.so"))
out = (ctypes.c_double * 4)()
for (M, N, K, bits) in [(256, 96, 1024, 16), (256, 96, 1024, 12), (256, 96, 1024, 20), (256, 96, 1024, 0), (256, 512, 192, 0), (300, 2496, 192, 0), (256, 96, 1024, 8), (256, 96, 64, 8), (256, 96, 4096, 8), (200, 64, 216, 0), (8192, 512, 192, 0), (5000, 2496, 192, 0), (20000, 96, 1024, 0)]:
    rc = lib.vosk_b200_gemm_selftest(M, N, K, bits, out)
    print("M=%d N=%d K=%d bits=%d rc=%d  fp32 max err %.3e  tc max err %.3e  tc rms err %.3e  ref rms %.3f" % (M, N, K, bits, rc, out[0], out[1], out[2], out[3]))
