import ctypes as C, sys
sys.path.insert(0, ".")
import keyhunt_b200 as K
kh = K.KeyHunt(0)
L = kh._lib
L.kh_hash_peak.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
for bps in (2, 4, 8):
    arr = (C.c_double * 2)()
    rc = L.kh_hash_peak(kh._h, bps, arr)
    print("blocks/SM=%d sha256 %.2f G/s  ripemd160 %.2f G/s  -> hash160(1 blk) %.2f G/s" % (bps, arr[0]/1e9, arr[1]/1e9, 1/(1/arr[0]+1/arr[1])/1e9))
