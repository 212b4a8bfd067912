"""Writes the INPUTS of the reference-pinning run (test infrastructure, see oracle/run_reference.m).

    python oracle/export_fixtures.py [out_dir]           (default tests/golden/reference_run)

Neither MATLAB nor Octave exists in the build image, so the reference's function files cannot be executed here and the oracle
is "parity unpinned".  This script + run_reference.m turn pinning into one command for anyone who has GNU Octave (or MATLAB):

    python oracle/export_fixtures.py
    cd tests/golden/reference_run && octave --no-gui --eval "run_reference('/path/to/multi-rtl-sdr-calibration')"
    python -m pytest tests/test_reference_run.py         # oracle == reference at the north-star tolerances

What is written (MAT v5 files Octave and MATLAB both read):
  case_<name>.mat      raw (2N x 1 uint8, what fread(tcp,...,'uint8') delivers), coef (fir1(46, 200e3/fs) taps - passed in because
                       Octave's fir1 is a different algorithm than MATLAB's), tpl (the 512-sample SCH template - an INPUT at the
                       boundary, SCH_corr_rate_correction.m:5, because comm.GMSKModulator is closed source), carrier_freq, osr
  gsm_chn_filter_8x.mat, gsm_chn_filter_4x.mat   variable Num: the numerators chn_filter_8x_4x.m:9 / chn_filter_4x.m:9 load, recovered
                       from the reference's own .fda sessions (tests/golden/chn_filter_taps.json)
  planted_*.mat        the Appendix-A fixtures of tests/appendix_a_fixtures.py (complex128 streams + positions)
  manifest.json        the case list the test walks
run_reference.m is copied next to them so the run is self-contained.
"""
from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np
import scipy.io

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (os.path.join(ROOT, "multi-rtl-sdr-calibration_b200"), HERE, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CARRIER = 957.4e6


def main(out_dir: str, n: int = 1020000) -> dict:
    import gsmcal_oracle as o
    from gsmcal import synth
    import appendix_a_fixtures as fx
    os.makedirs(out_dir, exist_ok=True)
    fs = o.SYMBOL_RATE * 8
    coef, tpl = o.fir1(46, 200e3 / fs), o.gsm_SCH_training_sequence_gen(8)
    cases = []
    specs = {
        "seed1": synth.random_spec(1, n), "seed2": synth.random_spec(2, n), "seed3": synth.random_spec(3, n),
        "noise_only": synth.StreamSpec(seed=21, n_samples=n, noise_only=True),                                   # coarse -1
        "weak_0db": synth.StreamSpec(seed=22, n_samples=n, snr_db=0.0, sampling_ppm=5, carrier_ppm=3),            # sentinel chain
        "dropped_fcch": synth.StreamSpec(seed=23, n_samples=n, sampling_ppm=-12, carrier_ppm=8, drop_fcch=(2,)),  # +11-frame fallback
        "offset_start": synth.StreamSpec(seed=24, n_samples=n, sampling_ppm=10, carrier_ppm=-20, start_offset=123456.0),
    }
    for name, sp in specs.items():
        raw = synth.generate_stream(sp).numpy()
        scipy.io.savemat(os.path.join(out_dir, f"case_{name}.mat"),
                         {"raw": raw.reshape(-1, 1), "coef": coef.reshape(1, -1), "tpl": tpl.reshape(-1, 1),
                          "carrier_freq": float(CARRIER), "osr": 8.0}, do_compression=True)
        cases.append({"name": name, "kind": "capture", "n_iq": n})
    with open(os.path.join(ROOT, "tests", "golden", "chn_filter_taps.json")) as f:
        g = json.load(f)
    for key, fn in (("Num_8x", "gsm_chn_filter_8x.mat"), ("Num_4x", "gsm_chn_filter_4x.mat")):
        scipy.io.savemat(os.path.join(out_dir, fn), {"Num": np.array([float.fromhex(h) for h in g[key]["hex"]]).reshape(1, -1)})
    planted = {"fine_overrun": fx.fine_overrun_drops_to_four()[0], "fine_snr_gate": fx.fine_snr_gate_return()[0]}
    for name, (s, base, osr, cf) in planted.items():
        scipy.io.savemat(os.path.join(out_dir, f"planted_{name}.mat"),
                         {"s": s.reshape(-1, 1), "base_position": np.asarray(base, dtype=np.float64).reshape(1, -1), "osr": float(osr), "carrier_freq": float(cf)},
                         do_compression=True)
        cases.append({"name": name, "kind": "planted_fine"})
    sch = {"sch_e_zero": fx.sch_e_zero_skips_interp1(tpl)[0], "sch_last_slot": fx.sch_last_slot_does_not_fit(tpl)[0],
           "sch_bcch_runout": fx.sch_bcch_rows_run_out(tpl)[0]}
    for name, (s, fcch, t, osr) in sch.items():
        scipy.io.savemat(os.path.join(out_dir, f"planted_{name}.mat"),
                         {"s": s.reshape(-1, 1), "FCCH_pos": fcch.reshape(1, -1), "tpl": np.asarray(t).reshape(-1, 1), "osr": float(osr)}, do_compression=True)
        cases.append({"name": name, "kind": "planted_sch"})
    manifest = {"cases": cases, "carrier_freq": CARRIER, "made_by": "oracle/export_fixtures.py",
                "tolerances": {"positions": "bit-exact", "raw2iq": "bit-exact", "ppm": 1e-3, "streams_rel": 1e-9}}
    with open(os.path.join(out_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    shutil.copy(os.path.join(HERE, "run_reference.m"), os.path.join(out_dir, "run_reference.m"))
    return manifest


if __name__ == "__main__":
    m = main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "reference_run"))
    print("wrote %d cases" % len(m["cases"]))
