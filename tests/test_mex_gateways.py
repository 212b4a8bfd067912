"""The MEX gateways (mex/gsmcal_mex.c) cannot be linked against MATLAB/Octave here (neither is installed), so every
gateway is compiled against the stub mex.h and the host-only ones are driven end to end through mexFunction."""
import os
import subprocess
import textwrap

import numpy as np
import pytest

import gsmcal_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEX = os.path.join(ROOT, "multi-rtl-sdr-calibration_b200", "mex")
FUNCS = ["raw2iq", "fir_filter", "chn_filter_8x_4x", "chn_filter_4x", "move_fft_snr_runtime_avg", "specific_fft_snr_fix_avg",
         "FCCH_coarse_position", "FCCH_fine_correction", "gsm_SCH_training_sequence_gen", "SCH_corr_rate_correction",
         "carrier_correct_post_SCH", "total_ppm_calculation", "gsm_calibrate_batch",
         "gsm_normal_training_sequence_gen", "FCCH_demod", "BCCH_demod", "SCH_demod"]


@pytest.mark.parametrize("fn", FUNCS)
def test_gateway_compiles_against_stub(fn, tmp_path):
    obj = tmp_path / f"{fn}.o"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wno-unused-function", "-c", f"-DGSMCAL_MEX_{fn}", f"-I{MEX}/stub", f"-I{ROOT}/include",
                    os.path.join(MEX, "gsmcal_mex.c"), "-o", str(obj)], check=True)
    sym = subprocess.run(["nm", str(obj)], capture_output=True, text=True).stdout
    assert " T mexFunction" in sym
    assert " U gsmcal_" in sym                  # binds the C ABI, computes nothing itself


def _run_harness(fn, body, built_lib, tmp_path):
    src = tmp_path / "h.c"
    src.write_text('#include "mex.h"\n#include "gsmcal.h"\n' + textwrap.dedent(body))
    exe = tmp_path / "h"
    libdir = os.path.dirname(built_lib)
    subprocess.run(["gcc", "-std=c11", "-Wno-unused-function", f"-DGSMCAL_MEX_{fn}", f"-I{MEX}/stub", f"-I{ROOT}/include", str(src),
                    os.path.join(MEX, "gsmcal_mex.c"), f"-L{libdir}", "-lgsmcal", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout


def test_total_ppm_gateway_roundtrip(built_lib, tmp_path):
    out = _run_harness("total_ppm_calculation", """
        int main(void) {
            mxArray *in = mxCreateDoubleMatrix(1, 2, mxREAL); mxGetPr(in)[0] = -35.0; mxGetPr(in)[1] = 1.25;
            mxArray *out[1]; const mxArray *rhs[1] = {in};
            mexFunction(1, out, 1, rhs);
            printf("%.17g\\n", mxGetPr(out[0])[0]);
            mxGetPr(in)[0] = INFINITY; mxGetPr(in)[1] = INFINITY;
            mexFunction(1, out, 1, rhs);
            printf("%.17g\\n", mxGetPr(out[0])[0]);
            return 0;
        }""", built_lib, tmp_path).split()
    assert float(out[0]) == oracle.total_ppm_calculation([-35.0, 1.25]) and out[1] == "inf"


def test_template_gateway_returns_split_complex_column(built_lib, tmp_path):
    out = _run_harness("gsm_SCH_training_sequence_gen", """
        int main(void) {
            mxArray *in = mxCreateDoubleScalar(8.0);
            mxArray *out[1]; const mxArray *rhs[1] = {in};
            mexFunction(1, out, 1, rhs);
            printf("%zu %zu %d\\n", mxGetM(out[0]), mxGetN(out[0]), mxIsComplex(out[0]));
            for (int i = 0; i < 512; ++i) printf("%.17g %.17g\\n", mxGetPr(out[0])[i], mxGetPi(out[0])[i]);
            return 0;
        }""", built_lib, tmp_path).split("\n")
    assert out[0] == "512 1 1"
    got = np.array([[float(x) for x in ln.split()] for ln in out[1:513]])
    ref = oracle.gsm_SCH_training_sequence_gen(8)
    assert np.max(np.abs(got[:, 0] + 1j * got[:, 1] - ref)) < 1e-12


def test_sentinel_shapes_through_the_gateway(built_lib, tmp_path):
    """FCCH_fine_correction with < 5 hits: FCCH_pos = -1, r = -1, ppm = inf (FCCH_fine_correction.m:8-15), no GPU needed."""
    out = _run_harness("FCCH_fine_correction", """
        int main(void) {
            mxArray *s = mxCreateDoubleMatrix(100, 1, mxCOMPLEX), *base = mxCreateDoubleScalar(-1.0);
            mxArray *osr = mxCreateDoubleScalar(8.0), *cf = mxCreateDoubleScalar(957.4e6);
            mxArray *out[4]; const mxArray *rhs[4] = {s, base, osr, cf};
            mexFunction(4, out, 4, rhs);
            printf("%g %zu %g %zu %g %g\\n", mxGetPr(out[0])[0], mxGetNumberOfElements(out[0]), mxGetPr(out[1])[0], mxGetNumberOfElements(out[1]),
                   mxGetPr(out[2])[0], mxGetPr(out[3])[0]);
            return 0;
        }""", built_lib, tmp_path).split()
    assert out == ["-1", "1", "-1", "1", "inf", "inf"]


def test_normal_training_sequence_gateway(built_lib, tmp_path):
    out = _run_harness("gsm_normal_training_sequence_gen", """
        int main(void) {
            mxArray *in = mxCreateDoubleScalar(4.0);
            mxArray *out[1]; const mxArray *rhs[1] = {in};
            mexFunction(1, out, 1, rhs);
            printf("%zu %zu %d\\n", mxGetM(out[0]), mxGetN(out[0]), mxIsComplex(out[0]));
            for (int i = 0; i < 104 * 8; ++i) printf("%.17g %.17g\\n", mxGetPr(out[0])[i], mxGetPi(out[0])[i]);
            return 0;
        }""", built_lib, tmp_path).split("\n")
    assert out[0] == "104 8 1"
    got = np.array([[float(x) for x in ln.split()] for ln in out[1:104 * 8 + 1]])
    ref = oracle.gsm_normal_training_sequence_gen(4)
    assert np.max(np.abs((got[:, 0] + 1j * got[:, 1]).reshape(8, 104).T - ref)) < 1e-12


def test_demod_gateways_warn_on_invalid_pos_info(built_lib, tmp_path):
    """pos_info == -1: SCH_demod / FCCH_demod print the reference's warning and return (SCH_demod.m:8-11, FCCH_demod.m:7-10)."""
    for fn, extra, msg in (("SCH_demod", "mxCreateDoubleMatrix(512, 1, mxCOMPLEX), mxCreateDoubleScalar(8.0)", "SCH demod: Warning! No valid position information!"),
                           ("FCCH_demod", "mxCreateDoubleScalar(8.0), mxCreateDoubleScalar(957.4e6)", "FCCH demod: Warning! No valid position information!")):
        out = _run_harness(fn, """
            int main(void) {
                mxArray *s = mxCreateDoubleMatrix(100, 1, mxCOMPLEX), *pi = mxCreateDoubleMatrix(1, 2, mxREAL);
                mxGetPr(pi)[0] = -1; mxGetPr(pi)[1] = -1;
                mxArray *out[3]; const mxArray *rhs[4] = {s, pi, %s};
                mexFunction(1, out, 4, rhs);
                printf("n=%%zu\\n", mxGetNumberOfElements(out[0]));
                return 0;
            }""" % extra, built_lib, tmp_path)
        assert msg in out and "n=0" in out
