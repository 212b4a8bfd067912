"""Randomised parity sweep (not collected by pytest): N streams with wide impairments through gsmcal.calibrate_batch vs the oracle.
   python tests/stress_parity.py   (needs the B200)"""
import sys, math, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(R,'multi-rtl-sdr-calibration_b200')); sys.path.insert(0, os.path.join(R,'oracle'))
import numpy as np
import gsmcal, gsmcal_oracle as o
from gsmcal import synth
N=1020000; FS=o.SYMBOL_RATE*8
tpl=o.gsm_SCH_training_sequence_gen(8); coef=o.fir1(46,200e3/FS)
rng=np.random.default_rng(7)
specs=[]
for i in range(64):
    specs.append(synth.StreamSpec(seed=3000+i, n_samples=N, sampling_ppm=float(rng.uniform(-45,45)), carrier_ppm=float(rng.uniform(-28,28)),
                                  snr_db=float(rng.uniform(4,25)), phase0=float(rng.uniform(0,6.28)), start_offset=float(rng.integers(0,synth.MULTIFRAME)),
                                  amplitude=float(rng.uniform(8,60))))
raw=synth.generate_batch(specs, device='cuda').cpu().numpy()
got=gsmcal.calibrate_batch(raw,957.4e6,tpl,coef)
bad=0; kinds={}
for d in range(len(specs)):
    ref=o.calibrate_stream(raw[d],957.4e6,tpl,coef)
    ok=(np.array_equal(got[d]['coarse_pos'],ref['coarse_pos']) and np.array_equal(got[d]['fcch_pos'],ref['fcch_pos']) and np.array_equal(got[d]['pos_info'],ref['pos_info']))
    for k in ('sampling_ppm','carrier_ppm'):
        for a,b in zip(got[d][k],ref[k]):
            ok = ok and ((a==b) if math.isinf(b) else abs(a-b)<1e-3)
    key=(len(ref['coarse_pos']) if ref['coarse_pos'][0]!=-1 else -1, ref['fcch_pos'][0]==-1, ref['pos_info'].shape[0])
    kinds[key]=kinds.get(key,0)+1
    if not ok:
        bad+=1; print('MISMATCH stream',d,specs[d].snr_db, got[d]['coarse_pos'][:4],ref['coarse_pos'][:4],got[d]['fcch_pos'][:3],ref['fcch_pos'][:3],got[d]['carrier_ppm'],ref['carrier_ppm'], got[d]['flags'])
print('mismatches',bad,'of',len(specs)); print(kinds)
from gsmcal._lib import lib
print('tier2', lib().gsmcal_debug_get(2), 'tier3', lib().gsmcal_debug_get(1))

# ---- hand-derived known answers through the C ABI (the CPU twins live in tests/test_oracle.py; not yet run on a GPU) --------------
def planted_known_answers():
    osr, frame = 8, 10000
    bad = 0
    for gaps, flagged in (((10, 10, 10, 11, 10, 10), {5}), ((10, 10, 10, 10, 11, 10), {1, 6}), ((10, 10, 10, 10, 10), set())):
        fcch = np.cumsum([2001] + [g * frame for g in gaps]).astype(np.float64)
        n = int(fcch[-1]) + 10336 + 512 + 5 * frame
        s = np.zeros(n, dtype=np.complex128)
        for p in fcch:
            s[int(p) + 10336 - 1:int(p) + 10336 - 1 + 512] = tpl
        pos_info, r, ppm = gsmcal.SCH_corr_rate_correction(s, fcch, tpl, osr)
        rows = []
        for i, p in enumerate(fcch, 1):
            rows += [[p, 0.0], [p + 10000, 1.0]] + ([[p + 10000 + k * frame, 2.0] for k in (1, 2, 3, 4)] if i in flagged else [])
        ok = pos_info.tolist() == rows and ppm == 0.0 and np.array_equal(r, s)
        bad += not ok
        print('planted templates', gaps, 'ok' if ok else 'MISMATCH')
    f_tone = o.SYMBOL_RATE / 4 + 2500.0
    rg = np.random.default_rng(5)
    starts = np.cumsum([30001] + [g * frame for g in (10, 10, 11, 10, 10)])
    n = int(starts[-1]) + 3 * frame
    s = 1e-3 * (rg.standard_normal(n) + 1j * rg.standard_normal(n))
    for p in starts:
        s[p - 1:p - 1 + 1184] += np.exp(2j * np.pi * f_tone * (p - 1 + np.arange(1184)) / FS)
    base = np.round((starts - 1) / osr) + 1 + np.array([3, -7, 0, 11, -20, 5])
    fpos, r, sppm, cppm = gsmcal.FCCH_fine_correction(s, base, osr, 957.4e6)
    ok = fpos.tolist() == starts.astype(float).tolist() and sppm == 0.0 and abs(cppm - 1e6 * 2500.0 / 957.4e6) < 1e-3
    bad += not ok
    print('planted tones', 'ok' if ok else ('MISMATCH', fpos, sppm, cppm))
    return bad


print('known-answer mismatches', planted_known_answers())
