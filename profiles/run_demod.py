"""One call of each SURVEY 8(f) demod function on a 10 s stream (for an ncu launch list):
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_demod.csv python profiles/run_demod.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-rtl-sdr-calibration_b200"))
import numpy as np
import gsmcal
from gsmcal import synth

N = 21_666_667
FS = (1625.0 / 6.0) * 1e3 * 8
spec = synth.random_spec(7, N)
spec.tsc = 3
raw = synth.generate_batch([spec], device="cuda").cpu().numpy()
coef, tpl = gsmcal.fir1(46, 200e3 / FS), gsmcal.gsm_SCH_training_sequence_gen(8)
r = gsmcal.fir_filter(coef, gsmcal.raw2iq(raw[0]))
r = np.ascontiguousarray(np.asarray(r).reshape(-1))
pos, _ = gsmcal.FCCH_coarse_position(r[::64], 8)
fpos, r1, _, _ = gsmcal.FCCH_fine_correction(r, pos, 8, 957.4e6)
pinfo, r2, _ = gsmcal.SCH_corr_rate_correction(r1, fpos, tpl, 8)
r3, _ = gsmcal.carrier_correct_post_SCH(r2, pinfo, 8, 957.4e6)
keep = np.array([not (t == 1 and p - 64 + 1552 - 1 > len(r3)) for p, t in pinfo])
pinfo = pinfo[keep]
for name, fn in (("FCCH_demod", lambda: gsmcal.FCCH_demod(r3, pinfo, 8, 957.4e6)),
                 ("BCCH_demod", lambda: gsmcal.BCCH_demod(r3, pinfo, gsmcal.gsm_normal_training_sequence_gen(8), 8, 957.4e6)),
                 ("SCH_demod", lambda: gsmcal.SCH_demod(r3, pinfo, tpl, 8))):
    t = time.perf_counter(); out = fn(); dt = time.perf_counter() - t
    print(name, f"{dt * 1e3:.1f} ms (host call incl. H2D of the {len(r3) * 16 / 1e6:.0f} MB stream)",
          {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()} if isinstance(out, dict) else out[:2])
d = gsmcal.SCH_demod(r3, pinfo, tpl, 8)
print("SCH bursts", d["corr_val"].shape[0], "peak at 42:", float(np.mean(d["corr_val"].argmax(axis=1) == 42)), "median peak", float(np.median(d["corr_val"].max(axis=1))))
