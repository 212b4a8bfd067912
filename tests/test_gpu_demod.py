"""SURVEY 8(f) rows 2 and 4 - FCCH_demod / BCCH_demod / SCH_demod / gsm_normal_training_sequence_gen through the C ABI
against the oracle on identical inputs.  Bars: bits, indices and integer correlations exact; frequencies 1e-6 Hz;
SNR 1e-9 dB; ppm 1e-9; correlation magnitudes 1e-10 relative."""
import dataclasses

import numpy as np
import pytest

import gsmcal_oracle as oracle
from gsmcal import synth
from gsmcal._lib import GsmcalError

pytestmark = pytest.mark.gpu

FS = oracle.SYMBOL_RATE * 8
CARRIER = 957.4e6
N_SYNC = 1020000


def _chain(seed, tsc, osr=8):
    spec = dataclasses.replace(synth.random_spec(seed, N_SYNC), tsc=tsc)
    raw = synth.generate_batch([spec]).numpy()[0]
    res = oracle.calibrate_stream(raw, CARRIER, oracle.gsm_SCH_training_sequence_gen(8), oracle.fir1(46, 200e3 / FS))
    return res["r_final"], res["pos_info"]


@pytest.fixture(scope="module")
def fixtures():
    return {seed: _chain(seed, tsc) for seed, tsc in ((1, 0), (2, 5), (3, None))}


def _fit_sch(pinfo, n, osr=8):
    """Drop SCH rows whose equaliser window (1552 samples from sch_pos - 64) overruns the stream (MATLAB would error)."""
    keep = [not (t == 1 and (p - 8 * osr < 1 or p - 8 * osr + 194 * osr - 1 > n)) for p, t in pinfo]
    return pinfo[np.array(keep)]


@pytest.mark.parametrize("osr", [1, 4, 8])
def test_normal_training_sequence_gen(gpu, osr):
    got, ref = gpu.gsm_normal_training_sequence_gen(osr), oracle.gsm_normal_training_sequence_gen(osr)
    assert got.shape == ref.shape == (26 * osr, 8)
    assert np.max(np.abs(got - ref)) < 1e-12


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_fcch_demod_matches_oracle(gpu, fixtures, seed):
    r3, pinfo = fixtures[seed]
    ref = oracle.FCCH_demod(r3, pinfo, 8, CARRIER)
    got = gpu.FCCH_demod(r3, pinfo, 8, CARRIER)
    assert len(got["freq"]) == len(ref["freq"]) == int(np.sum(pinfo[:, 1] == 0)) >= 5
    np.testing.assert_array_equal(got["max_idx"], ref["max_idx"])
    assert np.max(np.abs(got["freq"] - ref["freq"])) < 1e-6
    assert np.max(np.abs(got["snr"] - ref["snr"])) < 1e-9
    assert abs(got["mean_freq"] - ref["mean_freq"]) < 1e-6
    assert abs(got["carrier_ppm"] - ref["carrier_ppm"]) < 1e-9
    # after carrier_correct_post_SCH the mean FCCH tone sits on fs_sym/4 (sanity, not parity)
    assert abs(got["mean_freq"] - oracle.SYMBOL_RATE / 4) < 1e-3


def test_fcch_demod_on_uncorrected_stream_reports_offset(gpu):
    """FCCH_demod on a stream that still carries its carrier error: carrier_ppm agrees with the oracle and with the sign/size
    of the injected offset (the estimator's own bias, SURVEY Appendix B, stays inside a few ppm)."""
    spec = synth.random_spec(4, N_SYNC)
    raw = synth.generate_batch([spec]).numpy()[0]
    r = oracle.fir_filter(oracle.fir1(46, 200e3 / FS), oracle.raw2iq(raw)[:, 0])
    starts = synth.true_fcch_starts(spec)
    starts = starts[(starts > 100) & (starts + 1300 < N_SYNC)]
    pinfo = np.stack([np.round(starts) + 23, np.zeros(len(starts))], axis=1)      # +23: FIR group delay
    ref, got = oracle.FCCH_demod(r, pinfo, 8, CARRIER), gpu.FCCH_demod(r, pinfo, 8, CARRIER)
    np.testing.assert_array_equal(got["max_idx"], ref["max_idx"])
    assert np.max(np.abs(got["freq"] - ref["freq"])) < 1e-6
    assert np.max(np.abs(got["snr"] - ref["snr"])) < 1e-9
    assert abs(got["carrier_ppm"] - spec.carrier_ppm) < 3.0


@pytest.mark.parametrize("seed,tsc", [(1, 0), (2, 5), (3, None)])
def test_bcch_demod_identifies_training_sequence(gpu, fixtures, seed, tsc):
    r3, pinfo = fixtures[seed]
    nts = oracle.gsm_normal_training_sequence_gen(8)
    ref_ppm, ref_idx, ref_mag = oracle.BCCH_demod(r3, pinfo, nts, 8, CARRIER)
    ppm, idx, mag = gpu.BCCH_demod(r3, pinfo, nts, 8, CARRIER)
    assert idx == ref_idx
    if tsc is not None:
        assert idx == tsc + 1
    assert abs(ppm - ref_ppm) < 1e-9
    assert np.max(np.abs(mag - ref_mag)) / np.max(ref_mag) < 1e-10


def test_demod_sentinel_paths(gpu, fixtures):
    r3, pinfo = fixtures[1]
    nts = oracle.gsm_normal_training_sequence_gen(8)
    tpl = oracle.gsm_SCH_training_sequence_gen(8)
    m1 = np.array([[-1.0, -1.0]])
    assert gpu.FCCH_demod(r3, m1, 8, CARRIER) is None and oracle.FCCH_demod(r3, m1, 8, CARRIER) is None
    assert gpu.SCH_demod(r3, m1, tpl, 8) is None and oracle.SCH_demod(r3, m1, tpl, 8) is None
    assert gpu.BCCH_demod(r3, m1, nts, 8, CARRIER)[:2] == (-1.0, -1) == oracle.BCCH_demod(r3, m1, nts, 8, CARRIER)[:2]
    few = pinfo[pinfo[:, 1] != 2]                                   # fewer than 4 BCCH rows (BCCH_demod.m:13-17)
    assert gpu.BCCH_demod(r3, few, nts, 8, CARRIER)[:2] == (-1.0, -1) == oracle.BCCH_demod(r3, few, nts, 8, CARRIER)[:2]
    m3 = -np.ones((30, 2))                                         # the 3H x 2 sentinel of SCH_corr_rate_correction
    assert gpu.SCH_demod(r3, m3, tpl, 8) is None
    no_sch = pinfo[pinfo[:, 1] != 1]
    out = gpu.SCH_demod(r3, no_sch, tpl, 8)
    assert out["demod_bits"].shape == (0, 148) and oracle.SCH_demod(r3, no_sch, tpl, 8)["corr_val"].shape == (0, 85)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sch_demod_matches_oracle(gpu, fixtures, seed):
    r3, pinfo = fixtures[seed]
    tpl = oracle.gsm_SCH_training_sequence_gen(8)
    pi = _fit_sch(pinfo, len(r3))
    ref = oracle.SCH_demod(r3, pi, tpl, 8)
    got = gpu.SCH_demod(r3, pi, tpl, 8)
    H = int(np.sum(pi[:, 1] == 1))
    assert got["demod_bits"].shape == ref["demod_bits"].shape == (H, 148) and H >= 4
    np.testing.assert_array_equal(got["demod_bits"], ref["demod_bits"])
    np.testing.assert_array_equal(got["bits_to_decoder"], ref["bits_to_decoder"])
    np.testing.assert_array_equal(got["corr_val"], ref["corr_val"])
    # sanity: the training sequence is found where the burst format puts it (bit 42); the one-tap-per-bin equaliser of the
    # reference amplifies out-of-band noise, so the peak is well below the ideal 64 at 15-25 dB SNR
    assert np.all(got["corr_val"].argmax(axis=1) == 42) and np.mean(got["corr_val"].max(axis=1) >= 40) >= 0.7


def test_sch_demod_window_overrun_is_a_range_error(gpu, fixtures):
    r3, pinfo = fixtures[1]
    tpl = oracle.gsm_SCH_training_sequence_gen(8)
    bad = np.vstack([pinfo, [[len(r3) - 1000.0, 1.0]]])
    with pytest.raises(IndexError):
        oracle.SCH_demod(r3, bad, tpl, 8)
    with pytest.raises(GsmcalError) as e:
        gpu.SCH_demod(r3, bad, tpl, 8)
    assert e.value.code == -4


def test_sch_demod_recovers_known_bits_osr4(gpu):
    """Clean synthetic SCH burst at osr 4 (generic N = 97 x 8 DFT path): the demodulated burst equals the transmitted bits
    away from the burst edges, and equals the oracle everywhere."""
    osr = 4
    rng = np.random.default_rng(11)
    bits = rng.integers(0, 2, 148 + 90)
    tb = oracle.SCH_TRAINING_BITS
    b0 = 30                                                         # burst starts 30 symbols into the vector
    bits[b0 + 42:b0 + 42 + 64] = tb
    prev = np.concatenate([[1], bits[:-1]])
    tx = oracle.gmsk_modulate((bits == prev).astype(np.int64), osr)
    tx = tx * np.exp(1j * 0.7) + 0.02 * (rng.standard_normal(len(tx)) + 1j * rng.standard_normal(len(tx)))
    tpl = oracle.gsm_SCH_training_sequence_gen(osr)
    pinfo = np.array([[b0 * osr + 1.0, 1.0]])
    ref, got = oracle.SCH_demod(tx, pinfo, tpl, osr), gpu.SCH_demod(tx, pinfo, tpl, osr)
    np.testing.assert_array_equal(got["demod_bits"], ref["demod_bits"])
    np.testing.assert_array_equal(got["corr_val"], ref["corr_val"])
    assert got["corr_val"][0].argmax() == 42 and got["corr_val"][0].max() == 64
