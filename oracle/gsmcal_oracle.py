"""CPU oracle for the GSM sync/calibration hot path (TEST INFRASTRUCTURE - never the product path).

A NumPy/SciPy fp64 restatement of the reference's MATLAB functions, written from the behaviour of the
`.m` files (cited per function as file:line into /root/reference).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module, and only as the checker or
the timed CPU baseline.

PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures for this path, and neither
MATLAB nor Octave exists in this image, so the reference itself cannot be run.  What pins this oracle:
  * the two FIR numerators recovered from the reference's own gsm_chn_filter_{8x,4x}.fda sessions
    (tests/golden/chn_filter_taps.json, made by oracle/make_golden.py);
  * the training-sequence bit patterns and the algorithm constants (thresholds, search ranges, ppm limits), both parsed from
    the text of the .m files by oracle/make_golden.py (tests/golden/training_bits.json, reference_constants.json);
  * the reference's own manual checks as assertions (test_diff_GMSK_mod_demod.m, CW_check.m, the "test on line" blocks);
  * hand-derived known answers for MATLAB semantics (round half away from zero, first-max, 1-based
    indices, toeplitz window ordering, pos_info row rules) in tests/test_oracle.py;
  * MATLAB built-ins are restated by their public definitions: fft -> numpy.fft.fft, filter ->
    scipy.signal.lfilter (direct form II transposed, zero state), fir1 -> scipy.signal.firwin (Hamming
    windowed sinc, unity DC gain), interp1 'linear' on a uniform grid, toeplitz -> sliding windows.
  * comm.GMSKModulator (gsm_SCH_training_sequence_gen.m:14,39) is closed source: the template generator
    here follows GSM 05.04 (BT 0.3, L=4, h=0.5, zero initial phase, no prehistory) and is NOT a bit match
    of MathWorks' output.  The template is an *input* at the boundary (SCH_corr_rate_correction.m:5).
    comm.GMSKDemodulator (SCH_demod.m:63) likewise: restated as a 32-state MLSE over that modulator (gmsk_viterbi_demod).

All positions and indices returned are 1-based doubles exactly as the reference returns them.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.signal import firwin, lfilter

SYMBOL_RATE = (1625.0 / 6.0) * 1e3          # gsm_sync_demod.m:16
LEN_FCCH_CW = 148                           # FCCH_fine_correction.m:20
# Algorithm constants of the reference (SURVEY 8a "must not drift"); tests/golden/reference_constants.json holds the same values
# parsed from the .m files and tests/test_oracle.py compares them.
CONSTANTS = dict(
    coarse_th_db=10.0,            # FCCH_coarse_position.m:21
    coarse_mv_len_factor=10,      # :22  mv_len = 10*fft_len
    coarse_first_frames=23,       # :25
    coarse_max_offset=5,          # :45
    min_bursts=5,                 # FCCH_fine_correction.m:12, SCH_corr_rate_correction.m:11
    fine_max_offset_sym=64,       # FCCH_fine_correction.m:30
    fine_max_ppm=4000,            # :83
    fine_snr_gate_db=5,           # :192
    sch_training_sym=64,          # SCH_corr_rate_correction.m:22
    sch_pre_training_sym=42,      # :24
    sch_max_offset_sym=8,         # :36
    sch_upper_trim_sym=5,         # :46
    sch_max_ppm=400,              # :94
)


def mround(x: float) -> float:
    """MATLAB round(): half away from zero (FCCH_coarse_position.m:35-36 relies on 1562.5 -> 1563)."""
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


def abs2(z):
    """abs(z).^2 as the reference writes it: sqrt first, then square."""
    return np.abs(z) ** 2


# ----------------------------------------------------------------------------------------------------
# K1  raw2iq.m:5-8
# ----------------------------------------------------------------------------------------------------
def raw2iq(a) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim == 1:
        a = a[:, None]
    a = a.astype(np.float64)
    c = a[0::2, :] + 1j * a[1::2, :]
    n = c.shape[0]
    # integer-valued sums are exact in fp64; one divide, one subtract per component
    mean = (c.real.sum(axis=0) / n) + 1j * (c.imag.sum(axis=0) / n)
    return c - mean[None, :]


# ----------------------------------------------------------------------------------------------------
# K2  filter(coef,1,r) call sites gsm_sync_demod.m:34,110 etc.; chn_filter_8x_4x.m:13-15; chn_filter_4x.m:13
# ----------------------------------------------------------------------------------------------------
def fir1(order: int, wn: float) -> np.ndarray:
    """MATLAB fir1(n, Wn): Hamming-windowed sinc low-pass scaled to unity DC gain."""
    return firwin(order + 1, wn)


def fir_filter(coef, s, decim: int = 1) -> np.ndarray:
    """filter(coef, 1, s) column-wise, zero initial state, start-up transient kept; then s(1:decim:end,:)."""
    s = np.asarray(s)
    r = lfilter(np.asarray(coef, dtype=np.float64), 1.0, s, axis=0)
    return r[::decim] if decim != 1 else r


def chn_filter_8x_4x(s, num_8x) -> np.ndarray:
    return fir_filter(num_8x, s, 2)


def chn_filter_4x(s, num_4x) -> np.ndarray:
    return fir_filter(num_4x, s, 1)


# ----------------------------------------------------------------------------------------------------
# K3  move_fft_snr_runtime_avg.m / specific_fft_snr_fix_avg.m
# ----------------------------------------------------------------------------------------------------
def window_snr(win: np.ndarray, fft_len: int) -> float:
    """SNR statistic of one window: move_fft_snr_runtime_avg.m:18-27 == specific_fft_snr_fix_avg.m:11-20."""
    p = abs2(np.fft.fft(win, fft_len))
    k = int(np.argmax(p))                                   # first maximum
    sig = p[(k - 1) % fft_len] + p[k] + p[(k + 1) % fft_len]   # sum(chn_tmp(max_set)), order k-1,k,k+1
    noise = p.sum() - sig
    return 10.0 * math.log10(sig / noise)


def _snr_map(s: np.ndarray, fft_len: int, first: int, last: int) -> np.ndarray:
    """Vectorised window_snr for 1-based window starts first..last (same arithmetic, batched FFT)."""
    idx = np.arange(first - 1, last)[:, None] + np.arange(fft_len)[None, :]
    p = abs2(np.fft.fft(s[idx], fft_len, axis=1))
    k = np.argmax(p, axis=1)
    r = np.arange(p.shape[0])
    sig = p[r, (k - 1) % fft_len] + p[r, k] + p[r, (k + 1) % fft_len]
    noise = p.sum(axis=1) - sig
    with np.errstate(divide="ignore", invalid="ignore"):
        return 10.0 * np.log10(sig / noise)


def move_fft_snr_runtime_avg(s, mv_len: int, fft_len: int, th: float, return_trace: bool = False):
    """move_fft_snr_runtime_avg.m:5-50.  Returns (hit_flag, hit_idx, hit_avg_snr, hit_snr)."""
    s = np.asarray(s).reshape(-1)
    n_win = len(s) - (fft_len - 1)
    fifo = [999.0] * mv_len                 # store_for_moving_avg, newest first
    sum_snr = 999.0 * mv_len                # sum() of identical integers is exact
    snrs = _snr_map(s, fft_len, 1, n_win) if n_win > 0 else np.zeros(0)
    hit = (False, -1.0, math.inf, math.inf)
    head = 0                                # ring buffer: position of the oldest element
    ring = np.full(mv_len, 999.0)
    for i in range(n_win):
        snr = float(snrs[i])
        peak_to_avg = snr - (sum_snr / mv_len)
        if peak_to_avg > th:
            hit = (True, float(i + 1), snr - peak_to_avg, snr)
            break
        sum_snr = sum_snr - ring[head]
        sum_snr = sum_snr + snr
        ring[head] = snr
        head = (head + 1) % mv_len
    if return_trace:
        return hit, snrs
    return hit


def specific_fft_snr_fix_avg(s, target_set, fft_len: int, th: float, avg_snr: float):
    """specific_fft_snr_fix_avg.m:5-34.  target_set = [first, last] (1-based, inclusive)."""
    s = np.asarray(s).reshape(-1)
    for i in range(int(target_set[0]), int(target_set[1]) + 1):
        snr = window_snr(s[i - 1:i - 1 + fft_len], fft_len)
        if snr - avg_snr > th:
            return True, float(i), snr
    return False, -1.0, math.inf


# ----------------------------------------------------------------------------------------------------
# K4  FCCH_coarse_position.m:5-94
# ----------------------------------------------------------------------------------------------------
def FCCH_coarse_position(s, decimation_ratio: int):
    s = np.asarray(s).reshape(-1)
    num_sym_per_frame = (625.0 / 4.0) * 8
    fft_len = 2 ** int(math.floor(math.log2(LEN_FCCH_CW / decimation_ratio)))
    length = len(s)
    th = CONSTANTS["coarse_th_db"]
    mv_len = CONSTANTS["coarse_mv_len_factor"] * fft_len
    n_first = int(math.ceil(CONSTANTS["coarse_first_frames"] * num_sym_per_frame / decimation_ratio))
    if n_first > length:
        raise IndexError("FCCH_coarse_position: stream shorter than 23 frames (reference would error)")
    hit_flag, hit_idx, hit_avg_snr, hit_snr = move_fft_snr_runtime_avg(s[:n_first], mv_len, fft_len, th)
    if not hit_flag:
        return np.array([-1.0]), np.array([-1.0])
    step10 = mround(10 * num_sym_per_frame / decimation_ratio)
    step11 = mround(11 * num_sym_per_frame / decimation_ratio)
    position = [hit_idx]
    snr = [hit_snr]
    max_offset = CONSTANTS["coarse_max_offset"]
    limit = (length - (fft_len - 1)) - max_offset
    while True:
        nxt = position[-1] + step10
        if nxt > limit:
            break
        f, idx, sn = specific_fft_snr_fix_avg(s, (nxt - max_offset, nxt + max_offset), fft_len, th, hit_avg_snr)
        if f:
            position.append(idx)
            snr.append(sn)
            continue
        nxt = position[-1] + step11
        if nxt > limit:
            break
        f, idx, sn = specific_fft_snr_fix_avg(s, (nxt - max_offset, nxt + max_offset), fft_len, th, hit_avg_snr)
        if f:
            position.append(idx)
            snr.append(sn)
        else:
            break
    position = (np.array(position) - 1) * decimation_ratio + 1
    return position.astype(np.float64), np.array(snr, dtype=np.float64)


# ----------------------------------------------------------------------------------------------------
# shared helpers: interp1 'linear' on the uniform grid, derotation, tone estimator, spacing classification
# ----------------------------------------------------------------------------------------------------
def interp1_uniform(v: np.ndarray, e: float, max_len: int) -> np.ndarray:
    """interp1((0:L-1)', v, (0:max_len-1)'.*(1+e), 'linear') - FCCH_fine_correction.m:123-125."""
    xq = np.arange(max_len, dtype=np.float64) * (1.0 + e)
    i0 = np.floor(xq).astype(np.int64)
    i0 = np.minimum(i0, len(v) - 1)
    i1 = np.minimum(i0 + 1, len(v) - 1)
    frac = xq - i0
    return v[i0] + frac * (v[i1] - v[i0])


def derotate(r: np.ndarray, dphi: float) -> np.ndarray:
    """r.*exp(1i.*(0:L-1)'.*dphi) - FCCH_fine_correction.m:165, carrier_correct_post_SCH.m:83."""
    return r * np.exp(1j * (np.arange(len(r), dtype=np.float64) * dphi))


def tone_freq_estimate(r: np.ndarray, pos, fft_len: int, sampling_rate: float):
    """Per-burst tone frequency: FCCH_fine_correction.m:143-155 == carrier_correct_post_SCH.m:58-72.

    Returns (fo per burst, fcch_mat after integer-bin derotation, phase_rotate per burst)."""
    pos = [int(p) for p in pos]
    fcch_mat = np.stack([r[p - 1:p - 1 + fft_len] for p in pos], axis=1)          # fft_len x H
    fd = abs2(np.fft.fft(fcch_mat, fft_len, axis=0))
    fd = np.concatenate([fd[fft_len // 2:, :], fd[:fft_len // 2, :]], axis=0)     # fftshift ordering
    max_idx = np.argmax(fd, axis=0) + 1                                           # 1-based, first max
    int_phase_rotate = 2.0 * np.pi * (max_idx - ((fft_len / 2) + 1)) / fft_len
    n = np.arange(fft_len, dtype=np.float64)[:, None]
    fcch_mat = fcch_mat * np.exp(-1j * (n * int_phase_rotate[None, :]))
    u = np.exp(1j * np.angle(fcch_mat))
    ratio = u[1:, :] / u[:-1, :]
    phase_rotate = np.angle(ratio.sum(axis=0) / ratio.shape[0])
    fo = sampling_rate * (int_phase_rotate + phase_rotate) / (2 * np.pi)
    return fo, fcch_mat, phase_rotate


def matlab_mean(x: np.ndarray) -> float:
    """mean() as sum/n with a left-to-right sum."""
    acc = 0.0
    for v in x:
        acc += float(v)
    return acc / len(x)


def classify_spacing(pos: np.ndarray, osr: int, max_ppm: float):
    """10-frame / 11-frame gap classification: FCCH_fine_correction.m:74-113, SCH_corr_rate_correction.m:89-116."""
    num_sym_per_frame = (625.0 / 4.0) * 8
    d10 = 10 * num_sym_per_frame * osr
    d11 = 11 * num_sym_per_frame * osr
    max_th = math.floor(d10 * max_ppm * 1e-6)
    max_th1 = math.floor(d11 * max_ppm * 1e-6)
    diff_seq = np.diff(pos)
    a_logical = np.abs(diff_seq - d10) < max_th
    b_logical = np.abs(diff_seq - d11) < max_th1
    ok = (int(a_logical.sum()) + int(b_logical.sum())) == len(pos) - 1
    expected = float(a_logical.sum() * d10 + b_logical.sum() * d11)
    return ok, a_logical, b_logical, expected, d10, d11


# ----------------------------------------------------------------------------------------------------
# K5-K8  FCCH_fine_correction.m:5-197
# ----------------------------------------------------------------------------------------------------
def fine_peak_window(s: np.ndarray, sp: int, length: int, fft_len: int, chunk: int = 205):
    """max over bins of |fft|^2 for `length` sliding windows starting at 1-based sp; argmax over windows.

    FCCH_fine_correction.m:48-52 (the toeplitz construction is exactly the sliding-window matrix)."""
    seg = s[sp - 1:sp - 1 + length + fft_len - 1]
    peak = np.empty(length)
    win = np.lib.stride_tricks.sliding_window_view(seg, fft_len)           # length x fft_len
    for c0 in range(0, length, chunk):
        p = abs2(np.fft.fft(win[c0:c0 + chunk], fft_len, axis=1))
        peak[c0:c0 + chunk] = p.max(axis=1)
    return int(np.argmax(peak)) + 1, peak


def FCCH_fine_correction(s, base_position, oversampling_ratio: int, carrier_freq: float, info: dict | None = None):
    """Returns (FCCH_pos, r, sampling_ppm, carrier_ppm); sentinels as Appendix A of SURVEY.md:
    FCCH_pos = array([-1.]) for the scalar -1, r = None for the scalar -1."""
    s = np.asarray(s).reshape(-1)
    base_position = np.asarray(base_position, dtype=np.float64).reshape(-1)
    r = None
    sampling_ppm = math.inf
    carrier_ppm = math.inf
    if len(base_position) < CONSTANTS["min_bursts"]:
        return np.array([-1.0]), r, sampling_ppm, carrier_ppm
    osr = oversampling_ratio
    sampling_rate = SYMBOL_RATE * osr
    fft_len = LEN_FCCH_CW * osr
    half_noise_len = int(math.ceil((fft_len * 200e3 / sampling_rate) / 2))
    len_s = len(s) // osr
    max_offset = CONSTANTS["fine_max_offset_sym"]
    pos_list = []
    margins = []
    for p in base_position:
        p = int(p)
        if (p + max_offset) > (len_s - LEN_FCCH_CW + 1):
            break
        sp = (p - max_offset - 1) * osr + 1
        ep = (p + max_offset - 1) * osr + 1
        length = ep - sp + 1
        max_idx, peak = fine_peak_window(s, sp, length, fft_len)
        if info is not None:
            srt = np.sort(peak)
            margins.append(float((srt[-1] - srt[-2]) / srt[-1]))
        pos_list.append(float(sp + max_idx - 1))
    FCCH_pos = np.array(pos_list, dtype=np.float64)
    last_idx = len(FCCH_pos)
    if info is not None:
        info["fine_first_round"] = FCCH_pos.copy()
        info["fine_margins"] = margins

    if last_idx >= 5:
        r = s
        first = FCCH_pos[0]
        ok, a_l, b_l, expected, d10, d11 = classify_spacing(FCCH_pos, osr, CONSTANTS["fine_max_ppm"])
        if not ok:
            return np.array([-1.0]), r, sampling_ppm, carrier_ppm
        actual = FCCH_pos[-1] - FCCH_pos[0]
        e = (actual - expected) / expected
        sampling_ppm = e * 1e6
        max_len = int(math.floor(len(r) / (1 + e))) if e >= 0 else len(r)
        r = interp1_uniform(r, e, max_len)
        step = np.where(a_l, d10, d11)
        FCCH_pos = np.cumsum(np.concatenate([[1.0], step]))
        first = mround((first - 1) / (1 + e)) + 1
        FCCH_pos = FCCH_pos + first - 1
        if (FCCH_pos[-1] + fft_len - 1) > len(r):
            FCCH_pos = FCCH_pos[:-1]

    num_fcch = len(FCCH_pos)
    if num_fcch >= 5:
        fo, fcch_mat, phase_rotate = tone_freq_estimate(r, FCCH_pos, fft_len, sampling_rate)
        target_freq = SYMBOL_RATE / 4
        fo_mean = matlab_mean(fo)
        carrier_ppm = 1e6 * (fo_mean - target_freq) / carrier_freq
        comp_phase_rotate = (target_freq - fo_mean) * 2 * math.pi / sampling_rate
        r = derotate(r, comp_phase_rotate)
        n = np.arange(fft_len, dtype=np.float64)[:, None]
        fcch_mat = fcch_mat * np.exp(-1j * (n * phase_rotate[None, :]))
        fd = abs2(np.fft.fft(fcch_mat, fft_len, axis=0))
        sig_rows = [0, 1, 2, fft_len - 2, fft_len - 1]                       # [1:3, end-1:end]
        noise_rows = list(range(3, half_noise_len)) + list(range(fft_len - half_noise_len, fft_len - 2))
        signal_power = fd[sig_rows, :].sum(axis=0)
        noise_power = fd[noise_rows, :].sum(axis=0)
        snr = 10.0 * np.log10(signal_power / noise_power)
        if info is not None:
            info["fine_fo"] = fo
            info["fine_gate_snr"] = snr
        if np.sum(snr < CONSTANTS["fine_snr_gate_db"]) > 0:
            return np.array([-1.0]), r, sampling_ppm, carrier_ppm
    return FCCH_pos, r, sampling_ppm, carrier_ppm


# ----------------------------------------------------------------------------------------------------
# T1  gsm_SCH_training_sequence_gen.m:5-45  (GMSK per GSM 05.04; not a bit match of comm.GMSKModulator)
# ----------------------------------------------------------------------------------------------------
SCH_TRAINING_BITS = np.array(
    [1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0,
     0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0,
     1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1], dtype=np.int64)   # gsm_SCH_training_sequence_gen.m:17-19


def gmsk_q(tau):
    """Integrated GMSK frequency pulse, BT=0.3, truncated to L=4 symbols, tau in symbols; q(0)=0, q(4)=1."""
    from scipy.special import erfc
    sigma = math.sqrt(math.log(2.0)) / (2.0 * math.pi * 0.3)

    def big_f(u):      # integral of the normal CDF: u*Phi(u) + phi(u)
        return u * 0.5 * erfc(-u / math.sqrt(2.0)) + np.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)

    def big_g(t):      # integral of the (untruncated, centred) frequency pulse up to t, tends to 1/2
        return (sigma / 2.0) * (big_f((t + 0.5) / sigma) - big_f((t - 0.5) / sigma))

    tau = np.clip(np.asarray(tau, dtype=np.float64), 0.0, 4.0)
    g0 = big_g(np.float64(-2.0))
    g1 = big_g(np.float64(2.0))
    return (big_g(tau - 2.0) - g0) / (g1 - g0)


def differential_encode(bits: np.ndarray) -> np.ndarray:
    """~abs(diff([0; data])) - gsm_SCH_training_sequence_gen.m:32: 1 where a bit equals its predecessor."""
    prev = np.concatenate([[0], bits[:-1]])
    return (bits == prev).astype(np.int64)


def gmsk_modulate(diff_bits: np.ndarray, osr: int) -> np.ndarray:
    """Bit 1 -> +1, bit 0 -> -1 (BitInput); h=0.5; zero initial phase; symbol k's pulse starts at sample k*osr."""
    a = 2.0 * diff_bits.astype(np.float64) - 1.0
    n = np.arange(len(a) * osr, dtype=np.float64)
    phase = np.zeros_like(n)
    for k, ak in enumerate(a):
        phase += ak * gmsk_q((n - k * osr) / osr)
    return np.exp(1j * (math.pi / 2.0) * phase)


def gsm_SCH_training_sequence_gen(oversampling_ratio: int) -> np.ndarray:
    return gmsk_modulate(differential_encode(SCH_TRAINING_BITS), oversampling_ratio)


# ----------------------------------------------------------------------------------------------------
# K9-K10  SCH_corr_rate_correction.m:5-182
# ----------------------------------------------------------------------------------------------------
def SCH_corr_rate_correction(s, FCCH_pos, sch_training_sequence, oversampling_ratio: int, info: dict | None = None):
    """Returns (pos_info [R x 2], r, sampling_ppm); r = None for the scalar -1."""
    FCCH_pos = np.asarray(FCCH_pos, dtype=np.float64).reshape(-1)
    r = None
    sampling_ppm = math.inf
    if len(FCCH_pos) < CONSTANTS["min_bursts"]:
        return np.array([[-1.0, -1.0]]), r, sampling_ppm
    s = np.asarray(s).reshape(-1)
    ts = np.asarray(sch_training_sequence).reshape(-1)
    osr = oversampling_ratio
    slot_ov = int((625 * osr) // 4)                 # num_sym_per_slot_ov (1250 at osr 8)
    frame_ov = slot_ov * 8
    len_ts_ov = CONSTANTS["sch_training_sym"] * osr
    len_pre_ov = CONSTANTS["sch_pre_training_sym"] * osr
    fix_off_ov = frame_ov + len_pre_ov              # (1250+42)*osr
    num_hit = len(FCCH_pos)
    pos_info = -np.ones((3 * num_hit, 2))
    len_s_ov = len(s)
    max_offset = CONSTANTS["sch_max_offset_sym"] * osr
    sch = []
    margins = []
    for p in FCCH_pos:
        training_sp = int(p) + fix_off_ov
        if (training_sp + max_offset) > (len_s_ov - len_ts_ov + 1):
            break
        sp = training_sp - max_offset
        ep = training_sp + max_offset - CONSTANTS["sch_upper_trim_sym"] * osr
        length = ep - sp + 1
        win = np.lib.stride_tricks.sliding_window_view(s[sp - 1:sp - 1 + length + len_ts_ov - 1], len_ts_ov)
        corr_val = abs2(win @ np.conj(ts))
        max_idx = int(np.argmax(corr_val)) + 1
        sch.append(float(sp + max_idx - 1))
        if info is not None:
            srt = np.sort(corr_val)
            margins.append(float((srt[-1] - srt[-2]) / srt[-1]))
        if max_idx == 1 or max_idx == length:
            return np.array([[-1.0, -1.0]]), r, sampling_ppm
    SCH_pos = np.array(sch, dtype=np.float64)
    num_sch = len(SCH_pos)
    if info is not None:
        info["sch_first_round"] = SCH_pos.copy()
        info["sch_margins"] = margins
    if num_sch < 5:
        return pos_info, r, sampling_ppm
    r = s
    first = SCH_pos[0]
    ok, a_l, b_l, expected, d10, d11 = classify_spacing(SCH_pos, osr, CONSTANTS["sch_max_ppm"])
    if not ok:
        return pos_info, r, sampling_ppm
    actual = SCH_pos[-1] - SCH_pos[0]
    e = (actual - expected) / expected
    sampling_ppm = e * 1e6
    if e != 0:
        max_len = int(math.floor(len(r) / (1 + e))) if e > 0 else len(r)
        r = interp1_uniform(r, e, max_len)
    step = np.where(a_l, d10, d11)
    SCH_pos = np.cumsum(np.concatenate([[1.0], step]))
    first = mround((first - 1) / (1 + e)) + 1
    SCH_pos = SCH_pos + first - 1

    bcch_flag = np.zeros(num_sch + 1, dtype=bool)
    b_idx = np.nonzero(b_l)[0] + 1                  # 1-based
    bcch_flag[b_idx + 1 - 1] = True
    bcch_flag[b_idx[b_idx >= 5] - 4 - 1] = True

    rows = []
    len_r = len(r)
    for i in range(num_sch):
        rows.append((SCH_pos[i] - fix_off_ov, 0.0))
        sp = SCH_pos[i] - len_pre_ov
        if sp + slot_ov - 1 <= len_r:
            rows.append((sp, 1.0))
        else:
            break
        if bcch_flag[i]:
            runout = False
            for k in range(1, 5):
                bsp = sp + k * frame_ov
                if bsp + slot_ov - 1 <= len_r:
                    rows.append((bsp, 2.0))
                else:
                    runout = True
                    break
            if runout:
                break
    return np.array(rows, dtype=np.float64).reshape(-1, 2), r, sampling_ppm


# ----------------------------------------------------------------------------------------------------
# K8/K7  carrier_correct_post_SCH.m:5-83
# ----------------------------------------------------------------------------------------------------
def carrier_correct_post_SCH(s, pos_info, oversampling_ratio: int, carrier_freq: float, info: dict | None = None):
    pos_info = np.asarray(pos_info, dtype=np.float64).reshape(-1, 2)
    if pos_info.size > 0 and np.all(pos_info == -1):
        return None, math.inf
    if int(np.sum(pos_info[:, 1] == 2)) < 4:
        return None, math.inf
    s = np.asarray(s).reshape(-1)
    sampling_rate = SYMBOL_RATE * oversampling_ratio
    target_freq = SYMBOL_RATE / 4
    fcch_pos = pos_info[pos_info[:, 1] == 0, 0]
    fft_len = LEN_FCCH_CW * oversampling_ratio
    fo, _, _ = tone_freq_estimate(s, fcch_pos, fft_len, sampling_rate)
    if info is not None:
        info["post_fo"] = fo
    fo_mean = matlab_mean(fo)
    carrier_ppm = 1e6 * (fo_mean - target_freq) / carrier_freq
    comp_phase_rotate = (target_freq - fo_mean) * 2 * math.pi / sampling_rate
    return derotate(s, comp_phase_rotate), carrier_ppm


# ----------------------------------------------------------------------------------------------------
# K11  total_ppm_calculation.m:5-21
# ----------------------------------------------------------------------------------------------------
def total_ppm_calculation(ppm_in) -> float:
    ppm_in = np.asarray(ppm_in, dtype=np.float64).reshape(-1)
    if len(ppm_in) > 0 and np.all(ppm_in == np.inf):
        return math.inf
    acc = 1.0
    for p in ppm_in:                                  # prod() left to right
        acc = acc * (1.0 + p * 1e-6)
    return (acc - 1.0) * 1e6


# ----------------------------------------------------------------------------------------------------
# the consumers of r_correct / pos_info (gsm_sync_demod.m:143-146).  The reference
# functions return nothing (they disp / plot); the restatements return the quantities they compute.
# ----------------------------------------------------------------------------------------------------
NORMAL_TRAINING_BITS = np.array([                      # gsm_normal_training_sequence_gen.m:17-24 (TSC 0..7)
    [0, 0, 1, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 1],
    [0, 0, 1, 0, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 1, 1, 0, 1, 1, 1],
    [0, 1, 0, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 1, 0],
    [0, 1, 0, 0, 0, 1, 1, 1, 1, 0, 1, 1, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1, 0],
    [0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1],
    [0, 1, 0, 0, 1, 1, 1, 0, 1, 0, 1, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 1, 1, 0, 1, 0],
    [1, 0, 1, 0, 0, 1, 1, 1, 1, 1, 0, 1, 1, 0, 0, 0, 1, 0, 1, 0, 0, 1, 1, 1, 1, 1],
    [1, 1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 1, 0, 0, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 0]], dtype=np.int64)


def gsm_normal_training_sequence_gen(oversampling_ratio: int) -> np.ndarray:
    """gsm_normal_training_sequence_gen.m:5-59 -> (26*osr) x 8, one GMSK-modulated normal training sequence per column
    (each column differentially encoded against a leading 0, :38; modulator reset per column, :51)."""
    return np.stack([gmsk_modulate(differential_encode(b), oversampling_ratio) for b in NORMAL_TRAINING_BITS], axis=1)


def _all_minus_one(pos_info: np.ndarray) -> bool:
    return pos_info.size > 0 and bool(np.all(pos_info == -1))


def FCCH_demod(s, pos_info, oversampling_ratio: int, carrier_freq: float):
    """FCCH_demod.m:5-66.  Returns None on the `pos_info==-1` path (:7-10), else a dict with what the function
    displays: freq per burst (:41-42), mean_freq (:43), carrier_ppm (:47), snr per burst (:57-63) and
    max_idx - (fft_len/2+1) (:66)."""
    pos_info = np.asarray(pos_info, dtype=np.float64).reshape(-1, 2)
    if _all_minus_one(pos_info):
        return None
    s = np.asarray(s).reshape(-1)
    osr = oversampling_ratio
    fft_len = LEN_FCCH_CW * osr
    sampling_rate = SYMBOL_RATE * osr
    fcch_pos = [int(p) for p in pos_info[pos_info[:, 1] == 0, 0]]
    fcch_mat = np.stack([s[p - 1:p - 1 + fft_len] for p in fcch_pos], axis=1)
    fd = abs2(np.fft.fft(fcch_mat, fft_len, axis=0))
    fd = np.concatenate([fd[fft_len // 2:, :], fd[:fft_len // 2, :]], axis=0)
    max_idx = np.argmax(fd, axis=0) + 1
    freq, _, _ = tone_freq_estimate(s, fcch_pos, fft_len, sampling_rate)            # :36-41 (same statements)
    mean_freq = matlab_mean(freq)
    carrier_ppm = 1e6 * (mean_freq - SYMBOL_RATE / 4) / carrier_freq
    half_noise_len = math.ceil((fft_len * 200e3 / sampling_rate) / 2)
    sp = (fft_len // 2 + 1) - half_noise_len
    ep = (fft_len // 2 + 1) + half_noise_len - 1
    snr = np.zeros(len(fcch_pos))
    for i in range(len(fcch_pos)):
        sset = (np.arange(max_idx[i] - 2, max_idx[i] + 3) - 1) % fft_len            # 0-based after the mod (:59-60)
        signal_power = float(np.sum(fd[sset, i]))
        noise_power = float(np.sum(fd[sp - 1:ep, i])) - signal_power
        snr[i] = 10.0 * math.log10(signal_power / noise_power) if signal_power / noise_power > 0 else float("nan")
    return dict(freq=freq, mean_freq=mean_freq, carrier_ppm=carrier_ppm, snr=snr,
                max_idx=(max_idx - (fft_len // 2 + 1)).astype(np.float64))


def BCCH_demod(s, pos_info, normal_training_sequence, oversampling_ratio: int, carrier_freq: float):
    """BCCH_demod.m:5-106 (the reference reads `carrier_freq` and `normal_training_sequence` without defining
    them - :68,:91 - so they are arguments here).  Returns (carrier_ppm, normal_training_sequence_idx, |corr_val| 8x4);
    (-1, -1, None) on the early returns (:6-16), idx = -1 when the four bursts disagree (:99-102)."""
    pos_info = np.asarray(pos_info, dtype=np.float64).reshape(-1, 2)
    if _all_minus_one(pos_info) or int(np.sum(pos_info[:, 1] == 2)) < 4:
        return -1.0, -1, None
    osr = oversampling_ratio
    r, carrier_ppm = carrier_correct_post_SCH(s, pos_info, osr, carrier_freq)       # :47-73 are the same statements
    nts = np.asarray(normal_training_sequence)
    bcch_pos = [int(p) for p in pos_info[pos_info[:, 1] == 2, 0]]
    L = 26 * osr
    corr_mat = np.stack([r[p + 61 * osr - 1:p + 61 * osr - 1 + L] for p in bcch_pos[:4]], axis=1)     # :85-89
    corr_val = nts.conj().T @ corr_mat                                               # :91, 8 x 4
    mag = np.abs(corr_val)
    max_idx = np.argmax(mag, axis=0) + 1
    idx = int(max_idx[0]) if np.all(max_idx == max_idx[0]) else -1                   # :94-102
    return carrier_ppm, idx, mag


def gmsk_branch_table(osr: int) -> np.ndarray:
    """Reference waveforms of one symbol interval for the 16 (a_m, a_m-1, a_m-2, a_m-3) combinations (bit 1 -> +1),
    zero accumulated phase: exp(i*pi/2*(a_m q(t) + a_m-1 q(t+1) + a_m-2 q(t+2) + a_m-3 q(t+3))), t = j/osr."""
    tau = np.arange(osr, dtype=np.float64) / osr
    q = [gmsk_q(tau + d) for d in range(4)]
    W = np.zeros((16, osr), dtype=np.complex128)
    for combo in range(16):
        a = [2.0 * ((combo >> (3 - d)) & 1) - 1.0 for d in range(4)]
        W[combo] = np.exp(1j * (math.pi / 2.0) * (a[0] * q[0] + a[1] * q[1] + a[2] * q[2] + a[3] * q[3]))
    return W


def gmsk_viterbi_demod(x, osr: int, traceback: int) -> np.ndarray:
    """MLSE demodulator for the GMSK of gmsk_modulate (BT 0.3, L = 4, h = 1/2, zero phase offset), standing in for
    comm.GMSKDemodulator('BitOutput',true,...,'TracebackDepth',D) of SCH_demod.m:63 (closed source: PARITY UNPINNED).

    32 states (4 accumulated phases x 3 previous symbols), all start metrics 0; per symbol the 16 branch correlations
    sum_j x[m*osr+j]*conj(W[combo][j]) are rotated by the state phase; first maximum wins every comparison; output
    bit m is the decision for symbol m-D traced from the best state after symbol m (0 while m < D)."""
    x = np.asarray(x).reshape(-1)
    nsym = len(x) // osr
    W = gmsk_branch_table(osr).conj()
    D = traceback
    metric = np.zeros(32)
    hist = np.zeros(32, dtype=np.uint64)
    out = np.zeros(nsym, dtype=np.int64)
    for m in range(nsym):
        seg = x[m * osr:(m + 1) * osr]
        c = np.zeros(16, dtype=np.complex128)
        for j in range(osr):                                   # ascending-sample accumulation (the order the kernel uses)
            c = c + seg[j] * W[:, j]
        new_metric = np.empty(32)
        new_hist = np.empty(32, dtype=np.uint64)
        for sn in range(32):
            pn, c1, c2, c3 = sn >> 3, (sn >> 2) & 1, (sn >> 1) & 1, sn & 1
            best, bh = None, None
            for b3 in (0, 1):
                p = (pn - (1 if b3 else -1)) % 4
                pred = p * 8 + (c2 << 2) + (c3 << 1) + b3
                cv = c[(c1 << 3) + (c2 << 2) + (c3 << 1) + b3]
                bm = (cv.real, cv.imag, -cv.real, -cv.imag)[p]
                cand = metric[pred] + bm
                if best is None or cand > best:
                    best, bh = cand, hist[pred]
            new_metric[sn] = best
            new_hist[sn] = ((int(bh) << 1) | c1) & 0xFFFFFFFFFFFFFFFF
        metric, hist = new_metric, new_hist
        if m >= D:
            out[m] = (int(hist[int(np.argmax(metric))]) >> D) & 1
    return out


def SCH_demod(s, pos_info, training_sequence, oversampling_ratio: int):
    """SCH_demod.m:5-121.  Returns None on the `pos_info==-1` path, else a dict of per-SCH-burst arrays:
    demod_bits [H x 148] (:93-94), bits_to_decoder [H x 148] (:97), corr_val [H x 85] (:110)."""
    pos_info = np.asarray(pos_info, dtype=np.float64).reshape(-1, 2)
    if _all_minus_one(pos_info):
        return None
    s = np.asarray(s).reshape(-1)
    ts = np.asarray(training_sequence).reshape(-1)
    osr = oversampling_ratio
    sch_pos = [int(p) for p in pos_info[pos_info[:, 1] == 1, 0]]
    num_ef = int(mround(625 / 4 - 8.25))                                              # 148 (:22)
    L_ts, L_pre, D, ex_len = 64, 42, 30, 8                                           # :25-28,:45,:53
    data = 2 * differential_encode(SCH_TRAINING_BITS) - 1                            # :47-51
    len_fde_ov = (num_ef + 2 * ex_len + D) * osr                                     # :54-55
    sp_tr = (ex_len + L_pre) * osr                                                    # 0-based start of the training part (:56)
    td = np.zeros(len_fde_ov, dtype=np.complex128)
    td[sp_tr:sp_tr + L_ts * osr] = ts
    fd_training = np.fft.fft(td)                                                      # :57-59
    n_lag = num_ef - L_ts + 1                                                         # 85 (:103-105)
    bits_all, dec_all, corr_all = [], [], []
    for p in sch_pos:
        sp = p - ex_len * osr                                                         # :79-81
        if sp < 1 or sp + len_fde_ov - 1 > len(s):
            raise IndexError("SCH_demod: burst window outside the stream (MATLAB: index exceeds matrix dimensions)")
        x = s[sp - 1:sp - 1 + len_fde_ov]
        rt = np.zeros(len_fde_ov, dtype=np.complex128)
        rt[sp_tr:sp_tr + L_ts * osr] = x[sp_tr:sp_tr + L_ts * osr]                    # :83-84
        fd_chn = np.fft.fft(rt) / fd_training                                         # :85-86
        x = np.fft.ifft(np.fft.fft(x) / fd_chn)                                       # :88-90
        bits = gmsk_viterbi_demod(x, osr, D)                                          # :92-93
        bits = bits[D + ex_len:][:num_ef]                                             # :94-95
        nb = 1 - bits
        dec = np.abs(np.diff(np.concatenate([[0], nb])))                              # :97
        pm = 2 * bits - 1
        corr = np.array([int(np.dot(data, pm[k:k + L_ts])) for k in range(n_lag)])    # :103-110
        bits_all.append(bits); dec_all.append(dec); corr_all.append(corr)
    return dict(demod_bits=np.array(bits_all).reshape(-1, num_ef), bits_to_decoder=np.array(dec_all).reshape(-1, num_ef),
                corr_val=np.array(corr_all, dtype=np.float64).reshape(-1, n_lag))


# ----------------------------------------------------------------------------------------------------
# driver restatement: gsm_sync_demod.m:107-124 for one stream; scanners' per-column processing
# ----------------------------------------------------------------------------------------------------
def calibrate_stream(raw_u8: np.ndarray, carrier_freq: float, template: np.ndarray, coef: np.ndarray,
                     osr: int = 8, coarse_dr: int = 8, info: dict | None = None) -> dict:
    """One dongle through gsm_sync_demod.m:107-124.  raw_u8: 2N interleaved uint8."""
    r = raw2iq(raw_u8)[:, 0]
    r = fir_filter(coef, r)
    coarse_pos, coarse_snr = FCCH_coarse_position(r[::osr * coarse_dr], coarse_dr)
    fcch_pos, r1, sppm1, cppm1 = FCCH_fine_correction(r, coarse_pos, osr, carrier_freq, info)
    pos_info, r2, sppm2 = SCH_corr_rate_correction(r1 if r1 is not None else np.array([-1.0]), fcch_pos,
                                                   template, osr, info)
    r3, cppm2 = carrier_correct_post_SCH(r2 if r2 is not None else np.array([-1.0]), pos_info, osr,
                                         carrier_freq, info)
    return dict(coarse_pos=coarse_pos, coarse_snr=coarse_snr, fcch_pos=fcch_pos, pos_info=pos_info,
                sampling_ppm=(sppm1, sppm2), carrier_ppm=(cppm1, cppm2),
                total_sampling_ppm=total_ppm_calculation([sppm1, sppm2]),
                total_carrier_ppm=total_ppm_calculation([cppm1, cppm2]),
                r_final=r3)


def fcch_scan_channel(raw_u8: np.ndarray, coef: np.ndarray, osr: int = 8, coarse_dr: int = 8):
    """One scanned frequency: multi_rtl_sdr_gsm_FCCH_scanner.m:132-135,164-186 -> (snr, num_hit, pos, snrs)."""
    r = fir_filter(coef, raw2iq(raw_u8)[:, 0])
    pos, snrs = FCCH_coarse_position(r[::osr * coarse_dr], coarse_dr)
    snr, num_hit = 0.0, 0
    d = np.diff(pos)
    if len(pos) >= 3:
        a = np.abs(d - 12500) > 50
        if not a.any():
            snr, num_hit = float(np.mean(snrs)), len(pos)
        else:
            b = np.abs(d[a] - 13750) > 50
            if not b.any():
                snr, num_hit = float(np.mean(snrs)), len(pos)
    return snr, num_hit, pos, snrs


def band_power(raw_u8_cols: np.ndarray, coef=None, decim: int = 1) -> np.ndarray:
    """scan_band_power_spectrum.m:80-85 (coef None) / multi_rtl_sdr_split_scanner.m:154-156: linear mean power."""
    r = raw2iq(raw_u8_cols)
    if coef is not None:
        r = fir_filter(coef, r)
    return np.mean(abs2(r[::decim, :]), axis=0)
