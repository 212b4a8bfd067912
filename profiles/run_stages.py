"""Launches each whole-stream (materialising) kernel once on device-resident buffers - a short target for ncu.
   ncu --set full --clock-control none --import-source on -k regex:'fir_|raw2iq_store|resample|colsum' -o gpurun_out/prof_stages python profiles/run_stages.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-rtl-sdr-calibration_b200"))
import torch  # noqa: E402
import gsmcal  # noqa: E402
from gsmcal._lib import lib  # noqa: E402

n, cols = 21666667, 4
L = lib()
raw = torch.randint(0, 256, (cols, 2 * n), dtype=torch.uint8, device="cuda")
a = torch.empty((cols, n, 2), dtype=torch.float64, device="cuda")
b = torch.empty((cols, n, 2), dtype=torch.float64, device="cuda")
coef = gsmcal.fir1(46, 200e3 / (gsmcal.api.SYMBOL_RATE * 8))
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    if rep == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for stage, src, dst in ((0, raw, a), (1, raw, a), (2, a, b), (3, raw, b), (4, a, b), (5, a, b), (6, raw, b)):
        rc = L.gsmcal_stage_launch(stage, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), n, cols, coef.ctypes.data_as(C.c_void_p), len(coef), C.c_void_p(st))
        assert rc == 0, L.gsmcal_last_error()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
