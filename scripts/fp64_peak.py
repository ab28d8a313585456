"""Measure the fp64 FMA issue rate on the GPU (iamrx_debug_fp64_peak) -> profiles/fp64_peak.json (bench.py's second roofline)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import iamr_b200 as ix  # noqa: E402

lib = ix.load()
r = C.c_double()
lib.check(lib.iamrx_debug_fp64_peak(C.byref(r), None))
out = {"dp_ginstr_per_s": r.value, "dp_tflops": 2 * r.value / 1e3,
       "how": "8 independent DFMA chains/thread, 256 threads, 8 CTAs/SM, 32768 iterations, best of 4 (CUDA events)",
       "nominal": "148 SMs x 64 DP lanes x 1.965 GHz = 18.6e3 G instr/s"}
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "fp64_peak.json")
os.makedirs(os.path.dirname(path), exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out))
