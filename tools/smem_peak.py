#!/usr/bin/env python
"""Measured shared-memory read bandwidth of the GPU (tools/smem_bw/smem_bw.cu, built by __graft_entry__.build()): the denominator
of K1's roofline.  Prints one JSON object; `--out profiles/smem_peak.json` stores it where bench.py picks it up."""
import argparse, ctypes as C, json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ap = argparse.ArgumentParser()
ap.add_argument("--out", default="")
a = ap.parse_args()
lib = C.CDLL(os.path.join(HERE, "smem_bw", "libsmem_bw.so"))
gbs, bpc, sms, khz = C.c_double(), C.c_double(), C.c_int(), C.c_int()
rc = lib.smem_bw_measure(10, C.byref(gbs), C.byref(bpc), C.byref(sms), C.byref(khz))
if rc:
    sys.exit(f"smem_bw_measure failed: {rc}")
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.max.sm", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip().splitlines()[0]
except Exception:
    q = ""
res = dict(smem_read_gbs=gbs.value, bytes_per_clk_per_sm_at_max_clock=bpc.value, sm_count=sms.value, max_clock_khz=khz.value, gpu=q,
           how="tools/smem_bw/smem_bw.cu: conflict-free LDS.128 loads, 2 CTAs x 1024 threads per SM, 8 loads in flight per thread, "
               "best of 10 launches of ~50 ms, CUDA events; formula for comparison: sm_count x 128 B/clk x max clock",
           formula_gbs=sms.value * 128 * khz.value * 1e3 / 1e9)
print(json.dumps(res))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
