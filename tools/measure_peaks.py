"""Measure the CUDA-core FMA peaks (roofline denominators MEASURED_PEAKS.json lacks) on this GPU."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biped_mpc_py_b200 import _lib

def measure(device=0):
    lib = _lib.load()
    out = {}
    for name, flag in (("fp32_fma_tflops", 0), ("fp64_fma_tflops", 1)):
        v = ctypes.c_double()
        _lib.check(lib.bmpc_measure_fma_peak(device, flag, ctypes.byref(v)))
        out[name] = v.value
    return out

if __name__ == "__main__":
    res = measure()
    print(json.dumps(res))
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"))
