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
