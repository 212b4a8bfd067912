"""Deliberate fixtures for the early-return / partial-output paths of SURVEY.md Appendix A that a random capture rarely hits.

Each builder returns the arguments of ONE drop-in call plus a predicate that says whether a result took the intended branch
(so a test can assert "the oracle went there" and "the CUDA entry point went there" besides comparing the two).  Planted
signals (tones / the SCH template in weak noise) keep every decision margin wide.  Used by tests/test_oracle.py (CPU: the
oracle takes the branch) and tests/test_gpu_parity.py (GPU: C ABI == oracle, same branch).
"""
from __future__ import annotations

import math

import numpy as np

import gsmcal_oracle as oracle

FS = oracle.SYMBOL_RATE * 8
CARRIER = 957.4e6
OSR = 8
FRAME = 10000
N_FCCH = 1184


def _planted_tones(starts, n, amp=1.0, f_off=2500.0, noise=1e-3, seed=5, noisy_burst=None, noisy_sigma=0.0):
    rg = np.random.default_rng(seed)
    s = noise * (rg.standard_normal(n) + 1j * rg.standard_normal(n))
    f_tone = oracle.SYMBOL_RATE / 4 + f_off
    for i, p in enumerate(starts):
        p = int(p)
        m = min(N_FCCH, n - (p - 1))
        s[p - 1:p - 1 + m] += amp * np.exp(2j * np.pi * f_tone * (p - 1 + np.arange(m)) / FS)
        if noisy_burst is not None and i == noisy_burst:
            s[p - 1:p - 1 + m] += noisy_sigma * (rg.standard_normal(m) + 1j * rg.standard_normal(m))
    return s


def fine_overrun_drops_to_four():
    """FCCH_fine_correction.m:135-137,142: five first-round positions, the capture is SHORTER than the regridded last burst needs
    (negative sampling error stretches the ideal grid), so the last position is dropped, num_fcch = 4 < 5 and the function
    returns FCCH_pos 1x4, r = the resampled stream, sampling_ppm valid, carrier_ppm = inf."""
    spacing = 99900                                    # -1000 ppm
    starts = 30001 + spacing * np.arange(5)
    n = int(starts[-1]) + 1300                         # search window of burst 5 fits (base offset -63 symbols), the regridded burst does not
    s = _planted_tones(starts, n)
    base = np.round((starts - 1) / OSR) + 1 + np.array([3, -7, 0, 11, -63])

    def took_branch(res):
        fpos, r, sppm, cppm = res
        return len(fpos) == 4 and fpos[0] != -1 and r is not None and len(r) == n and math.isfinite(sppm) and sppm < -900 and cppm == math.inf
    return (s, base, OSR, CARRIER), took_branch


def fine_snr_gate_return():
    """FCCH_fine_correction.m:192-196: spacing and both ppm estimates are fine, one burst sits in strong in-band noise (gate SNR
    < 5 dB): FCCH_pos = -1, r = resampled AND derotated, both ppm finite."""
    starts = np.cumsum([30001] + [g * FRAME for g in (10, 10, 11, 10, 10)])
    n = int(starts[-1]) + 3 * FRAME
    s = _planted_tones(starts, n, noisy_burst=2, noisy_sigma=1.2)
    base = np.round((starts - 1) / OSR) + 1 + np.array([3, -7, 0, 11, -20, 5])

    def took_branch(res):
        fpos, r, sppm, cppm = res
        return len(fpos) == 1 and fpos[0] == -1 and r is not None and math.isfinite(sppm) and math.isfinite(cppm)
    return (s, base, OSR, CARRIER), took_branch


def _planted_templates(gaps, tail, tpl, first=2001):
    fcch = np.cumsum([first] + [g * FRAME for g in gaps]).astype(np.float64)
    n = int(fcch[-1]) + tail
    s = np.zeros(n, dtype=np.complex128)
    for p in fcch:
        a = int(p) + 10336 - 1
        m = max(0, min(512, n - a))
        s[a:a + m] = tpl[:m]
    return s, fcch


def sch_e_zero_skips_interp1(tpl):
    """SCH_corr_rate_correction.m:120-128: templates exactly on the ideal grid -> e == 0, interp1 is skipped, r is s itself."""
    s, fcch = _planted_templates((10, 10, 10, 11, 10, 10), 10336 + 512 + 5 * FRAME, tpl)

    def took_branch(res):
        pos_info, r, sppm = res
        return sppm == 0.0 and r is not None and len(r) == len(s) and np.array_equal(r, s) and pos_info.shape[0] > 12
    return (s, fcch, tpl, OSR), took_branch, s


def sch_last_slot_does_not_fit(tpl):
    """SCH_corr_rate_correction.m:153-159 (first `break`): the correlation window of the last SCH still fits, its 1250-sample
    slot does not: pos_info ends with that burst's FCCH row and has no SCH row for it."""
    gaps = (10, 10, 10, 10, 10)
    s, fcch = _planted_templates(gaps, 10336 + 64 + 512 + 100, tpl)         # slot would need fcch+10000+1250-1

    def took_branch(res):
        pos_info, r, sppm = res
        return (r is not None and pos_info.shape[0] == 2 * len(fcch) - 1 and pos_info[-1, 1] == 0.0 and pos_info[-1, 0] == fcch[-1]
                and np.count_nonzero(pos_info[:, 1] == 1.0) == len(fcch) - 1)
    return (s, fcch, tpl, OSR), took_branch


def sch_bcch_rows_run_out(tpl):
    """SCH_corr_rate_correction.m:167-178 (second `break`): a flagged burst near the end, only two of its four BCCH slots fit."""
    gaps = (10, 10, 10, 10, 11)                        # b_idx = 5 -> BCCH_flag(6) and BCCH_flag(1)
    s, fcch = _planted_templates(gaps, 10000 + 1250 + 2 * FRAME + 600, tpl)

    def took_branch(res):
        pos_info, r, sppm = res
        last = pos_info[pos_info[:, 0] >= fcch[-1]]
        return r is not None and list(last[:, 1]) == [0.0, 1.0, 2.0, 2.0] and list(pos_info[:6, 1]) == [0.0, 1.0, 2.0, 2.0, 2.0, 2.0]
    return (s, fcch, tpl, OSR), took_branch
