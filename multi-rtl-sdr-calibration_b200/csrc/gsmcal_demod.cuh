// gsmcal_demod.cuh - kernels of the SURVEY 8(f) rows 2 and 4 functions (FCCH_demod.m, BCCH_demod.m, SCH_demod.m): the consumers of
// r_correct / pos_info (gsm_sync_demod.m:143-146).  One thread block per burst, materialised complex128 stream in HBM, fp64.
#pragma once
#include "gsmcal_kernels.cuh"

// ===================================================================================================
// FCCH_demod.m:21-66  per FCCH burst: 148*osr-point power spectrum (fftshift order), tone frequency (same statements as K8)
// and the 5-bin-signal / +-half_noise_len-band SNR.  All N bins are needed (the SNR reads 2*55 of them plus 5 around the peak),
// so the spectrum is evaluated once through the 37 x M row FFT.
// ===================================================================================================
#define FD_THREADS 256
__global__ void __launch_bounds__(FD_THREADS) fcch_demod_kernel(const double2 *__restrict__ s, const double *__restrict__ pos, int osr,
                                                               const double2 *__restrict__ tw, double *__restrict__ freq_out,
                                                               double *__restrict__ snr_out, double *__restrict__ idx_out) {
    extern __shared__ double2 sm[];
    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ double red_n[32];
    __shared__ double sh_pr;
    const int burst = blockIdx.x, tid = threadIdx.x;
    const int N = 148 * osr;
    const double sampling_rate = ((1625.0 / 6.0) * 1e3) * (double)osr;
    double2 *u = sm, *A = u + N, *F = A + N;
    double *P = (double *)(F + N);                             // fd_fcch column in fftshift order
    const i64 sp = (i64)pos[burst];
    for (int n = tid; n < N; n += FD_THREADS) u[n] = s[sp - 1 + n];
    __syncthreads();
    const double2 *Tm = fft_rows(u, A, F, N, tw);
    double v = -1.0; int j_best = 0x7fffffff;
    for (int j = tid; j < N; j += FD_THREADS) {
        int k = j + N / 2; if (k >= N) k -= N;
        const double p = abs2_ref(dft_col(Tm, k, N, tw));
        P[j] = p;
        argmax_combine(v, j_best, p, j);
    }
    block_argmax(v, j_best, red_v, red_i);                     // first maximum (:33); also orders the P[] writes
    // SNR (:53-63)
    const int hnl = (int)ceil(((double)N * 200e3 / sampling_rate) / 2.0);
    double sn[2] = {0.0, 0.0};
    if (tid < 5) { int j = (j_best - 2 + tid) % N; if (j < 0) j += N; sn[0] = P[j]; }
    for (int j = N / 2 - hnl + tid; j <= N / 2 + hnl - 1; j += FD_THREADS) sn[1] += P[j];
    block_sum_n<2, false>(sn, red_n);
    if (tid == 0) {
        const double noise = sn[1] - sn[0];
        snr_out[burst] = 10.0 * log10(sn[0] / noise);
        idx_out[burst] = (double)(j_best + 1 - (N / 2 + 1));
    }
    // tone frequency (:35-41): integer-bin derotation, unit phasors, angle of the mean one-sample rotation
    const int jr = j_best + 1 - ((N / 2) + 1);
    const double int_phase_rotate = 2.0 * GSMCAL_PI * (double)jr / (double)N;
    int jm = jr % N; if (jm < 0) jm += N;
    for (int n = tid; n < N; n += FD_THREADS) {
        const double2 w = cmul(u[n], tw[(int)(((i64)n * jm) % N)]);
        const double h = hypot(w.x, w.y);
        A[n] = (h > 0.0) ? make_double2(w.x / h, w.y / h) : make_double2(1.0, 0.0);
    }
    __syncthreads();
    double rri[2] = {0.0, 0.0};
    for (int n = tid; n < N - 1; n += FD_THREADS) {
        const double2 a = A[n + 1], b = A[n];
        const double den = b.x * b.x + b.y * b.y;
        rri[0] += (a.x * b.x + a.y * b.y) / den;
        rri[1] += (a.y * b.x - a.x * b.y) / den;
    }
    block_sum_n<2, false>(rri, red_n);
    if (tid == 0) sh_pr = atan2(rri[1] / (double)(N - 1), rri[0] / (double)(N - 1));
    __syncthreads();
    if (tid == 0) freq_out[burst] = sampling_rate * (int_phase_rotate + sh_pr) / (2 * GSMCAL_PI);
}

// ===================================================================================================
// BCCH_demod.m:85-91  zero-lag correlation of the first 4 BCCH bursts with the 8 normal training sequences, on the carrier-
// corrected stream r = s .* exp(1i*(0:len-1)'*comp) (:72-73) evaluated only where it is read.  One block per burst, one warp
// per training sequence.
// ===================================================================================================
__global__ void __launch_bounds__(256) nts_corr_kernel(const double2 *__restrict__ s, double dphi, const double *__restrict__ bpos, int osr,
                                                       const double2 *__restrict__ nts /* (26*osr) x 8 column-major */, double *__restrict__ corr_abs /* 8 x 4 */) {
    const int burst = blockIdx.x, q = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = 26 * osr;
    const i64 sp0 = (i64)bpos[burst] + 61 * osr - 1;           // 0-based index of the first training sample (:86)
    double ar = 0.0, ai = 0.0;
    for (int n = lane; n < L; n += 32) {
        double sn, cs; sincos((double)(sp0 + n) * dphi, &sn, &cs);
        const double2 r = cmul(s[sp0 + n], make_double2(cs, sn));
        const double2 t = nts[(i64)q * L + n];
        ar += t.x * r.x + t.y * r.y;                             // conj(t) * r
        ai += t.x * r.y - t.y * r.x;
    }
    ar = warp_sum(ar); ai = warp_sum(ai);
    if (lane == 0) corr_abs[burst * 8 + q] = hypot(ar, ai);
}

// ===================================================================================================
// SCH_demod.m:54-110  per SCH burst: frequency-domain equalisation against the known training sequence, MLSE GMSK demodulation,
// +-1 correlation with the 64 training bits.
//   N = (148 + 2*8 + 30)*osr = 194*osr = 97 * (2*osr): two-stage DFT with direct stages (97 is prime), exact twiddle table.
// ===================================================================================================
template <bool INV>
__device__ void dft_pxm(const double2 *in, double2 *tmp, double2 *out, int N, int P, const double2 *__restrict__ tw) {
    const int M = N / P;
    for (int idx = threadIdx.x; idx < N; idx += blockDim.x) {
        const int n2 = idx % M, k1 = idx / M;
        double ar = 0.0, ai = 0.0;
        int t = 0; const int stp = (M * k1) % N;
#pragma unroll 4
        for (int n1 = 0; n1 < P; ++n1) {
            const double2 x = in[M * n1 + n2];
            double2 w = tw[t]; if (INV) w.y = -w.y;
            ar = fma(x.x, w.x, fma(-x.y, w.y, ar));
            ai = fma(x.x, w.y, fma(x.y, w.x, ai));
            t += stp; if (t >= N) t -= N;
        }
        double2 w = tw[(n2 * k1) % N]; if (INV) w.y = -w.y;
        tmp[idx] = cmul(make_double2(ar, ai), w);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < N; idx += blockDim.x) {
        const int k1 = idx % P, k2 = idx / P;
        double ar = 0.0, ai = 0.0;
        int t = 0; const int stp = (P * k2) % N;
        const double2 *a = tmp + M * k1;
#pragma unroll 4
        for (int n2 = 0; n2 < M; ++n2) {
            const double2 x = a[n2];
            double2 w = tw[t]; if (INV) w.y = -w.y;
            ar = fma(x.x, w.x, fma(-x.y, w.y, ar));
            ai = fma(x.x, w.y, fma(x.y, w.x, ai));
            t += stp; if (t >= N) t -= N;
        }
        out[k1 + P * k2] = INV ? make_double2(ar / (double)N, ai / (double)N) : make_double2(ar, ai);
    }
    __syncthreads();
}
__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {    // Smith's algorithm (what the oracle's complex division does)
    if (fabs(b.x) >= fabs(b.y)) {
        const double rat = b.y / b.x, scl = 1.0 / (b.x + b.y * rat);
        return make_double2((a.x + a.y * rat) * scl, (a.y - a.x * rat) * scl);
    }
    const double rat = b.x / b.y, scl = 1.0 / (b.y + b.x * rat);
    return make_double2((a.x * rat + a.y) * scl, (a.y * rat - a.x) * scl);
}

#define SD_THREADS 256
#define SD_NSYM 194            // 148 + 2*ex_len + TracebackDepth (SCH_demod.m:54)
#define SD_TB 30               // TracebackDepth (:45)
#define SD_EX 8                // ex_len (:53)
// fd_training_ov = fft(zero-padded training sequence) (:57-59): once per call
__global__ void __launch_bounds__(SD_THREADS) fde_template_kernel(const double2 *__restrict__ tpl, int osr, const double2 *__restrict__ tw,
                                                                 double2 *__restrict__ Ft) {
    extern __shared__ double2 sm[];
    const int N = SD_NSYM * osr, sp_tr = (SD_EX + 42) * osr, L = 64 * osr;
    double2 *x = sm, *tmp = x + N, *out = tmp + N;
    for (int n = threadIdx.x; n < N; n += SD_THREADS) x[n] = (n >= sp_tr && n < sp_tr + L) ? tpl[n - sp_tr] : make_double2(0.0, 0.0);
    __syncthreads();
    dft_pxm<false>(x, tmp, out, N, 97, tw);
    for (int n = threadIdx.x; n < N; n += SD_THREADS) Ft[n] = out[n];
}

// MLSE over the 32-state trellis of the L=4, h=1/2 GMSK: state = 8*p + 4*a(m-1) + 2*a(m-2) + a(m-3) (bits; p = accumulated quarter turns),
// lane == new state.  cb[m*16 + combo] = sum_j x[m*osr+j] * conj(W[combo][j]), combo = 8*a(m) + 4*a(m-1) + 2*a(m-2) + a(m-3).
__device__ void gmsk_viterbi_warp(const double2 *cb, int nsym, unsigned char *bits_out) {
    const int lane = threadIdx.x & 31;
    const int pn = lane >> 3, c1 = (lane >> 2) & 1, c2 = (lane >> 1) & 1, c3 = lane & 1;
    const int p0 = (pn + 1) & 3, p1 = (pn + 3) & 3;             // predecessor phase for a(m-3) = 0 (-1: p = pn + 1) / 1 (+1: p = pn - 1)
    const int pred0 = p0 * 8 + (c2 << 2) + (c3 << 1), pred1 = p1 * 8 + (c2 << 2) + (c3 << 1) + 1;
    const int combo0 = (c1 << 3) + (c2 << 2) + (c3 << 1);
    double metric = 0.0;
    unsigned hist = 0u;
    for (int m = 0; m < nsym; ++m) {
        const double2 v0 = cb[m * 16 + combo0], v1 = cb[m * 16 + combo0 + 1];
        const double bm0 = (p0 == 0) ? v0.x : (p0 == 1) ? v0.y : (p0 == 2) ? -v0.x : -v0.y;
        const double bm1 = (p1 == 0) ? v1.x : (p1 == 1) ? v1.y : (p1 == 2) ? -v1.x : -v1.y;
        const double m0 = __shfl_sync(0xffffffffu, metric, pred0) + bm0;
        const double m1 = __shfl_sync(0xffffffffu, metric, pred1) + bm1;
        const unsigned h0 = __shfl_sync(0xffffffffu, hist, pred0), h1 = __shfl_sync(0xffffffffu, hist, pred1);
        const bool take1 = m1 > m0;                              // first maximum: a(m-3) = 0 wins ties
        metric = take1 ? m1 : m0;
        hist = ((take1 ? h1 : h0) << 1) | (unsigned)c1;
        if (m >= SD_TB) {
            double bv = metric; int bi = lane;
            for (int o = 16; o > 0; o >>= 1) {
                const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
                if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
            }
            const unsigned hb = __shfl_sync(0xffffffffu, hist, bi);
            if (lane == 0) bits_out[m] = (unsigned char)((hb >> SD_TB) & 1u);
        } else if (lane == 0) bits_out[m] = 0;
    }
}

__global__ void __launch_bounds__(SD_THREADS) sch_demod_kernel(const double2 *__restrict__ s, const double *__restrict__ pos, int osr,
                                                              const double2 *__restrict__ tw, const double2 *__restrict__ Ft,
                                                              const double2 *__restrict__ Wtab /* 16 x osr */, const signed char *__restrict__ data_pm /* 64 */,
                                                              unsigned char *__restrict__ demod_bits /* 148 per burst */,
                                                              unsigned char *__restrict__ dec_bits, double *__restrict__ corr_val /* 85 per burst */) {
    extern __shared__ double2 sm[];
    __shared__ unsigned char bits[SD_NSYM + 2];
    __shared__ signed char dpm[64];
    const int burst = blockIdx.x, tid = threadIdx.x;
    const int N = SD_NSYM * osr, sp_tr = (SD_EX + 42) * osr, L = 64 * osr;
    double2 *x = sm, *tmp = x + N, *fa = tmp + N, *fb = fa + N;
    const i64 sp = (i64)pos[burst] - SD_EX * osr;              // :79
    if (tid < 64) dpm[tid] = data_pm[tid];
    for (int n = tid; n < N; n += SD_THREADS) {
        const double2 v = s[sp - 1 + n];
        x[n] = v;
        fb[n] = (n >= sp_tr && n < sp_tr + L) ? v : make_double2(0.0, 0.0);      // received_training_ov (:83-84)
    }
    __syncthreads();
    dft_pxm<false>(fb, tmp, fa, N, 97, tw);                    // fd_received_training
    for (int n = tid; n < N; n += SD_THREADS) fa[n] = cdiv(fa[n], Ft[n]);       // fd_chn (:86)
    __syncthreads();
    dft_pxm<false>(x, tmp, fb, N, 97, tw);                     // fd_x (:88)
    for (int n = tid; n < N; n += SD_THREADS) fb[n] = cdiv(fb[n], fa[n]);       // :89
    __syncthreads();
    dft_pxm<true>(fb, tmp, x, N, 97, tw);                      // x = ifft(fd_x) (:90)
    // branch correlations of every symbol interval (the trellis then only adds)
    double2 *cb = tmp;                                          // 194*16 <= 2*N entries for osr >= 8; sized by the host for smaller osr
    for (int i = tid; i < SD_NSYM * 16; i += SD_THREADS) {
        const int m = i >> 4, combo = i & 15;
        double ar = 0.0, ai = 0.0;
        for (int j = 0; j < osr; ++j) {
            const double2 v = x[m * osr + j], w = Wtab[combo * osr + j];
            ar += v.x * w.x + v.y * w.y;                         // v * conj(w)
            ai += v.y * w.x - v.x * w.y;
        }
        cb[i] = make_double2(ar, ai);
    }
    __syncthreads();
    if (tid < 32) gmsk_viterbi_warp(cb, SD_NSYM, bits);
    __syncthreads();
    // :94-110
    const unsigned char *db = bits + SD_TB + SD_EX;
    for (int i = tid; i < 148; i += SD_THREADS) {
        demod_bits[(i64)burst * 148 + i] = db[i];
        const int nb = 1 - db[i], pb = (i > 0) ? 1 - db[i - 1] : 0;
        dec_bits[(i64)burst * 148 + i] = (unsigned char)(nb != pb);
    }
    for (int k = tid; k < 148 - 64 + 1; k += SD_THREADS) {
        int acc = 0;
        for (int i = 0; i < 64; ++i) acc += dpm[i] * (2 * (int)db[k + i] - 1);
        corr_val[(i64)burst * 85 + k] = (double)acc;
    }
}
