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


# ---- file-driven harness (tests/mex_harness.c): real arrays through mexFunction, both complex layouts --------------------------
_CLS = {np.dtype(np.float64): 6, np.dtype(np.complex128): 6, np.dtype(np.uint8): 9, np.dtype(np.bool_): 3}


def _write_arrays(path, arrays, nlhs):
    import struct
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(arrays), nlhs))
        for a in arrays:
            a = np.asarray(a)
            if a.ndim == 0:
                a = a.reshape(1, 1)
            if a.ndim == 1:
                a = a.reshape(-1, 1)
            cplx = np.iscomplexobj(a)
            f.write(struct.pack("<iii", _CLS[a.dtype], int(cplx), a.ndim))
            f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            col = np.asfortranarray(a)                              # MATLAB storage order
            if cplx:
                f.write(np.ascontiguousarray(col.real.ravel(order="F")).tobytes())
                f.write(np.ascontiguousarray(col.imag.ravel(order="F")).tobytes())
            else:
                f.write(col.ravel(order="F").tobytes())


def _read_arrays(path):
    import struct
    out = []
    with open(path, "rb") as f:
        (n,) = struct.unpack("<i", f.read(4))
        for _ in range(n):
            cls, cplx, ndim = struct.unpack("<iii", f.read(12))
            dims = struct.unpack("<%dq" % ndim, f.read(8 * ndim))
            cnt = int(np.prod(dims))
            dt = np.float64 if cls == 6 else np.uint8
            re = np.frombuffer(f.read(cnt * np.dtype(dt).itemsize), dtype=dt)
            a = re.astype(np.complex128) if cplx else re.copy()
            if cplx:
                a = a + 1j * np.frombuffer(f.read(cnt * 8), dtype=np.float64)
            out.append(a.reshape(dims, order="F"))
    return out


def run_gateway(fn, arrays, nlhs, built_lib, tmp_path, interleaved=0):
    """compile mex/gsmcal_mex.c (-DGSMCAL_MEX_<fn>) + tests/mex_harness.c against the stub mex.h and call mexFunction once"""
    exe = tmp_path / f"gw_{fn}_{interleaved}"
    libdir = os.path.dirname(built_lib)
    if not exe.exists():
        subprocess.run(["gcc", "-std=c11", "-O1", "-Wno-unused-function", f"-DGSMCAL_MEX_{fn}", f"-DMX_HAS_INTERLEAVED_COMPLEX={interleaved}",
                        f"-I{MEX}/stub", f"-I{ROOT}/include", os.path.join(ROOT, "tests", "mex_harness.c"), os.path.join(MEX, "gsmcal_mex.c"),
                        f"-L{libdir}", "-lgsmcal", f"-Wl,-rpath,{libdir}", "-lm", "-o", str(exe)], check=True)
    fin, fout = tmp_path / f"{fn}_{interleaved}.in", tmp_path / f"{fn}_{interleaved}.out"
    _write_arrays(fin, arrays, nlhs)
    p = subprocess.run([str(exe), str(fin), str(fout)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    return _read_arrays(fout), p.stdout


@pytest.mark.parametrize("interleaved", [0, 1])
def test_file_harness_host_only_gateways(built_lib, tmp_path, interleaved):
    (tot,), _ = run_gateway("total_ppm_calculation", [np.array([[-35.0, 1.25]])], 1, built_lib, tmp_path, interleaved)
    assert tot.shape == (1, 1) and tot[0, 0] == oracle.total_ppm_calculation([-35.0, 1.25])
    (tpl,), _ = run_gateway("gsm_SCH_training_sequence_gen", [np.array(8.0)], 1, built_lib, tmp_path, interleaved)
    assert tpl.shape == (512, 1) and np.max(np.abs(tpl[:, 0] - oracle.gsm_SCH_training_sequence_gen(8))) < 1e-12


def test_fir_filter_gateway_rejects_bad_decim(built_lib, tmp_path):
    """decim 0 / negative / fractional must be a MEX error before any size is derived from it (no SIGFPE inside the interpreter)"""
    exe_dir = tmp_path
    for bad in (0.0, -3.0, 2.5):
        fin, fout = exe_dir / "bad.in", exe_dir / "bad.out"
        coef, s = np.ones((1, 4)), (np.arange(10) + 0j).reshape(-1, 1)
        try:
            run_gateway("fir_filter", [coef, s, np.array(bad)], 1, built_lib, tmp_path)
            raise RuntimeError("gateway accepted decim=%r" % bad)
        except AssertionError as ex:
            assert "decim must be a positive integer" in str(ex)


# ---- GPU work through mexFunction (stub runtime), compared with the oracle ---------------------------------------------------------
FS = oracle.SYMBOL_RATE * 8
CARRIER = 957.4e6


@pytest.fixture(scope="module")
def mex_capture():
    from gsmcal import synth
    raw = synth.generate_batch([synth.random_spec(s, 1020000) for s in (1, 2)]).numpy()
    coef = oracle.fir1(46, 200e3 / FS)
    r = oracle.fir_filter(coef, oracle.raw2iq(raw[0])[:, 0])
    coarse, snr = oracle.FCCH_coarse_position(r[::64], 8)
    return dict(raw=raw, coef=coef, r=r, coarse=coarse, coarse_snr=snr, tpl=oracle.gsm_SCH_training_sequence_gen(8))


@pytest.mark.gpu
def test_mex_raw2iq_uint8_and_double(gpu, built_lib, tmp_path, mex_capture):
    a = np.ascontiguousarray(mex_capture["raw"][:, :2 * 50001].T)           # 2N x 2 uint8, one dongle per column
    ref = oracle.raw2iq(a)
    for interleaved in (0, 1):
        (b,), _ = run_gateway("raw2iq", [a], 1, built_lib, tmp_path, interleaved)
        assert b.shape == ref.shape and np.array_equal(b, ref)
    (b,), _ = run_gateway("raw2iq", [a.astype(np.float64)], 1, built_lib, tmp_path)    # fread(...,'uint8') hands over doubles
    assert np.array_equal(b, ref)


@pytest.mark.gpu
def test_mex_fcch_coarse_position(gpu, built_lib, tmp_path, mex_capture):
    (pos, snr), _ = run_gateway("FCCH_coarse_position", [mex_capture["r"][::64], np.array(8.0)], 2, built_lib, tmp_path)
    assert pos.shape == (1, len(mex_capture["coarse"])) and np.array_equal(pos[0], mex_capture["coarse"])
    assert np.max(np.abs(snr[0] - mex_capture["coarse_snr"])) < 1e-9
    rng = np.random.default_rng(9)
    noise = rng.standard_normal(16000) + 1j * rng.standard_normal(16000)
    (pos, snr), out = run_gateway("FCCH_coarse_position", [noise, np.array(8.0)], 2, built_lib, tmp_path)
    assert pos.shape == (1, 1) and pos[0, 0] == -1 and snr[0, 0] == -1 and "No FCCH found" in out       # FCCH_coarse_position.m:27-30


@pytest.mark.gpu
@pytest.mark.parametrize("interleaved", [0, 1])
def test_mex_fcch_fine_correction_full_outputs(gpu, built_lib, tmp_path, mex_capture, interleaved):
    r = mex_capture["r"]
    ref = oracle.FCCH_fine_correction(r, mex_capture["coarse"], 8, CARRIER)
    (fpos, r1, sppm, cppm), _ = run_gateway("FCCH_fine_correction", [r, mex_capture["coarse"].reshape(1, -1), np.array(8.0), np.array(CARRIER)],
                                            4, built_lib, tmp_path, interleaved)
    assert fpos.shape == (1, len(ref[0])) and np.array_equal(fpos[0], ref[0])
    assert r1.shape == (len(ref[1]), 1) and np.max(np.abs(r1[:, 0] - ref[1])) / np.max(np.abs(ref[1])) < 1e-8
    assert sppm[0, 0] == ref[2] and abs(cppm[0, 0] - ref[3]) < 1e-3


@pytest.mark.gpu
def test_mex_gsm_calibrate_batch(gpu, built_lib, tmp_path, mex_capture):
    raw = mex_capture["raw"]
    a = np.ascontiguousarray(raw.T)                                          # 2N x D uint8
    (sp, cp, nrows, pinfo), _ = run_gateway("gsm_calibrate_batch", [a, np.array(CARRIER), mex_capture["tpl"], mex_capture["coef"].reshape(1, -1)],
                                            4, built_lib, tmp_path)
    assert sp.shape == (3, 2) and cp.shape == (3, 2) and nrows.shape == (1, 2) and pinfo.shape[1:] == (2, 2)
    for d in range(2):
        ref = oracle.calibrate_stream(raw[d], CARRIER, mex_capture["tpl"], mex_capture["coef"])
        n = int(nrows[0, d])
        assert n == len(ref["pos_info"]) and np.array_equal(pinfo[:n, :, d], ref["pos_info"]) and np.isnan(pinfo[n:, :, d]).all()
        assert sp[0, d] == ref["sampling_ppm"][0] and sp[1, d] == ref["sampling_ppm"][1] and abs(sp[2, d] - ref["total_sampling_ppm"]) < 1e-3
        assert abs(cp[0, d] - ref["carrier_ppm"][0]) < 1e-3 and abs(cp[1, d] - ref["carrier_ppm"][1]) < 1e-3 and abs(cp[2, d] - ref["total_carrier_ppm"]) < 1e-3
