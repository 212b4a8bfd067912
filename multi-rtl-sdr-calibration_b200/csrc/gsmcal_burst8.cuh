// gsmcal_burst8.cuh - the fine FCCH search specialised for the reference's oversampling ratio 8 on the lazy (uint8) source.
//
// Same arithmetic and the same certificates as fine_peak_core_kernel (gsmcal_kernels.cuh; FCCH_fine_correction.m:32-64), with
// every size a compile-time constant (N = 1184 = 37*32, 1025 window starts, 2208 samples = 69 chunks of 32) and the phases
// rebuilt around what bounds them on an SM:
//   * the 47-tap FIR (56 % of the arithmetic) runs ONCE per burst, 9 outputs per thread from a register window with constant-bank
//     taps (1 LDS.128 per 18 DFMA), in place: staged capture and filtered window share one buffer, so 4 blocks stay resident;
//   * the filtered window goes to an L2/HBM cache with one TMA bulk store, so the two tone stages and the 64-bin tier never filter
//     these samples again;
//   * chunk sums by Horner's rule, 4 bins per thread sharing every sample load (4 DFMA per sample and bin instead of 5, 1 LDS per
//     16 DFMA), on a window layout with one pad slot per chunk so the chunk starts of a quarter-warp fall into different banks;
//   * the slack of the certificate from chunk energies (Cauchy-Schwarz) instead of a per-sample square root.
#pragma once

#define B8_N       1184
#define B8_NWIN    1025
#define B8_CH      32
#define B8_NSEG    32
#define B8_WCH     37
#define B8_THREADS 256
#define B8_R       9
#define B8_BUF     2304                      // staged samples: 2208 + (NT-1 <= 63) + R-1, rounded up
#define B8_CSL     71                        // row stride of the chunk prefix table (odd: bins fall into different banks)
#define B8_SMEM    (B8_BUF * 16 + 8 * B8_CSL * 16 + 72 * 8 + 72 * 8 + 33 * 8 * 8)
#define B8_NPASS   8                         // passes of 8 tracked bins each before a burst is handed to the 64-bin band kernel
// offset from the band centre k0 of tracked bin j (0..7) in pass p: 8 bins around the tone, 8 more on both sides, then 8 at a time
// towards the GMSK data energy (it sits ~37 bins below the FCCH tone)
__device__ __forceinline__ int b8_bin_off(int pass, int j) { return pass == 0 ? j - 3 : (pass == 1 ? (j < 4 ? j - 7 : j + 1) : -8 * pass + 1 + j); }

template <int NT, bool PROF>
__global__ void __maxnreg__(56) fine_core8_kernel(const uint8_t *__restrict__ raw_all, i64 n_iq, const StreamCtl *__restrict__ ctl,
                                                                  const double *__restrict__ base_pos, int cap, const double2 *__restrict__ tw,
                                                                  double *__restrict__ fine_raw, int *__restrict__ need_band, int force_fail,
                                                                  double2 *__restrict__ wcache, int n_pass, unsigned *__restrict__ pass_hist,
                                                                  unsigned long long *__restrict__ prof) {
    extern __shared__ __align__(128) unsigned char b8_sm[];
    double2 *B = reinterpret_cast<double2 *>(b8_sm);             // staged capture, then the filtered window (padded layout)
    double2 *CS = B + B8_BUF;                                    // [8][B8_CSL] prefix of chunk sums, bin-major
    double *PE = reinterpret_cast<double *>(CS + 8 * B8_CSL);    // [70] prefix of chunk energies
    double *E31 = PE + 72;                                       // [69] energy of the first 31 samples of every chunk
    double *T2 = E31 + 72;                                       // [33][8] tracked power of pass 0 per (window, split)
    __shared__ double red_v[8];
    __shared__ int red_i[8];
    __shared__ double scan_sv[16];
    const int burst = blockIdx.x, stream = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // debug hook (gsmcal_debug_set key 13): thread 0 adds the cycles between phase marks to prof[phase]; prof[15] counts blocks
    long long t_prev = PROF ? clock64() : 0;
#define B8_MARK(ph) do { if (PROF && tid == 0) { const long long t_now = clock64(); atomicAdd(prof + (ph), (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
    // the three things the block needs from global memory before it can address the capture are fetched side by side (one latency,
    // not three): the burst's coarse position, the stream's burst count and its mean
    const double pos_d = __ldg(base_pos + (i64)stream * cap + burst);      // (always inside the [D][cap] array; garbage beyond n_coarse is not used)
    const int n_coarse = ctl[stream].n_coarse;
    struct { double mu_re, mu_im; } c = {ctl[stream].mu_re, ctl[stream].mu_im};
    if (n_coarse < 5 || burst >= n_coarse) return;
    const i64 len_s = n_iq / 8;
    const i64 position = (i64)pos_d;
    double *o = fine_raw + (i64)stream * cap + burst;
    if (position + 64 > len_s - 148 + 1) {                       // run out of sampled signal (:35-38)
        if (tid == 0) *o = INFINITY;
        return;
    }
    const i64 sp0 = (position - 65) * 8;                         // 0-based first sample of the first window (sp - 1)
    // ---- stage the DC-removed capture: one 2-byte IQ pair per thread and load, all loads in flight ----
    {
        constexpr int NRAW = B8_NSMP + NT - 1;
        constexpr int NLD = (NRAW + B8_THREADS - 1) / B8_THREADS;
        const unsigned short *rp = reinterpret_cast<const unsigned short *>(raw_all + (i64)stream * 2 * n_iq);
        const i64 j00 = sp0 - (NT - 1);
        const double mur = c.mu_re, mui = c.mu_im;
        unsigned wv[NLD];
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int i = tid + q * B8_THREADS;
            const i64 j = j00 + i;
            wv[q] = (i < NRAW && j >= 0) ? (unsigned)__ldg(rp + j) : 0x10000u;     // zero initial filter state before the first sample
        }
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
            const int i = tid + q * B8_THREADS;
            if (i < NRAW)
                B[i] = (wv[q] & 0x10000u) ? make_double2(0.0, 0.0)
                                          : make_double2(u8_to_f64(wv[q] & 0xffu) - mur, u8_to_f64((wv[q] >> 8) & 0xffu) - mui);
        }
        if (tid < B8_R + 7) B[NRAW + tid] = make_double2(0.0, 0.0);
    }
    __syncthreads();
    B8_MARK(0);                                                  // staging (global-memory latency)
    // ---- FIR, 9 outputs per thread, in place: all reads happen before the barrier, all writes after it ----
    {
        constexpr int NGRP = (B8_NSMP + B8_R - 1) / B8_R;        // 246 <= 256: one round
        double ar[B8_R], ai[B8_R];
        // thread -> output group by an odd multiplier: the loads stay conflict-free (57 = 1 mod 8) and the padded stores of a quarter-warp
        // collide 1.2x instead of 2x (consecutive groups step over a pad slot every 3.6 lanes)
        const int grp = (57 * tid) & (B8_THREADS - 1);
        if (grp < NGRP) fir_taps_const<NT, B8_R>(B + B8_R * grp, ar, ai);
        __syncthreads();
        if (grp < NGRP) {
#pragma unroll
            for (int r = 0; r < B8_R; ++r) {
                const int i = B8_R * grp + r;
                if (i < B8_NSMP) B[B8_WPAD(i)] = make_double2(ar[r], ai[r]);
            }
        }
    }
    __syncthreads();
    B8_MARK(1);                                                  // FIR
    // ---- the filtered window goes to the cache with ONE TMA bulk store (reads shared memory asynchronously: waited for below,
    //      before the window is overwritten by differences) ----
    if (wcache && tid == 0) {
        double2 *dstp = wcache + ((i64)stream * cap + burst) * B8_WLEN;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     ::"l"(dstp), "r"((unsigned)__cvta_generic_to_shared(B)), "r"((unsigned)(B8_WLEN * sizeof(double2))) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    // ---- chunk energies: thread = (chunk, half), 16 samples each, the start rotated per lane so a quarter-warp reads 8 banks ----
    {
        double se = 0.0, e_last = 0.0;
        const int cch = tid >> 1, h = tid & 1;
        if (tid < 2 * B8_NCH) {
            const double2 *bp = B + 33 * cch + 16 * h;
            const int rot = (tid + 1) >> 1;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double2 v = bp[(i + rot) & 15];
                se = fma(v.x, v.x, fma(v.y, v.y, se));
            }
            if (h == 0) { const double2 v = B[33 * cch + 31]; e_last = fma(v.x, v.x, v.y * v.y); }
        }
        if (warp < 5) {                                          // 138 tasks live in warps 0..4; the halves of a chunk are lane neighbours
            const double se_o = __shfl_xor_sync(0xffffffffu, se, 1);
            // E31 = E32 - |s[31]|^2: the rounding of the difference (1e-16 of E32) is far inside the 1e-9 / 1e-6 margins of the certificate
            if (tid < 2 * B8_NCH && h == 0) { const double e32 = se + se_o; PE[cch + 1] = e32; E31[cch] = fmax(e32 - e_last, 0.0) * (1.0 + 1e-12); }
        }
        if (tid == 0) PE[0] = 0.0;
    }
    // ---- band centre from the phase slope of the centre window ----
    {
        double pq[2] = {0.0, 0.0};
        if (5 * tid < B8_N - 1) {                                // 5 consecutive pairs per thread: 6 loads instead of 10
            const int n0 = 512 + 5 * tid;
            double2 prev = B[B8_WPAD(n0)];
#pragma unroll
            for (int k = 1; k <= 5; ++k) {
                const int n = n0 + k;
                if (n < 512 + B8_N) {
                    const double2 v = B[B8_WPAD(n)];
                    const double2 q2 = cmulc(v, prev);
                    pq[0] += q2.x; pq[1] += q2.y;
                    prev = v;
                }
            }
        }
        block_sum_n<2, false>(pq, scan_sv);                      // (its barriers also publish PE / E31)
        // (the band centre only selects WHICH bins are tracked - the certificate decides correctness - so the fp32 atan2 is enough:
        //  it is ~5x shorter on the one thread the whole block waits for)
        if (tid == 0) red_i[0] = (int)floorf(atan2f((float)pq[1], (float)pq[0]) * (float)(B8_N / (2.0 * GSMCAL_PI)) + 0.5f);
    }
    if (warp == 1) warp_scan_smem(PE, B8_NCH + 1, lane);         // PE[i] = sum_{n < 32 i} |s[n]|^2
    __syncthreads();
    const int k0 = red_i[0];
    __syncthreads();
    B8_MARK(2);                                                  // energies, phase slope, k0
    // Up to n_pass passes of 8 tracked bins (b8_bin_off): every pass adds its bins' power to the tracked sums of the certificate, which
    // is re-checked after each; a burst leaves as soon as it is proven.  The 64-bin band kernel only sees what is still open then.
    double g_best = -1.0; int g_bestm = 0x7fffffff;
    int ok = 0, pass = 0;
    bool is_diff = false;                                        // B[0..1024) holds the differences d[m] instead of the samples (block-uniform)
    for (; pass < n_pass && !ok; ++pass) {
        if (is_diff) {                                           // s[m] = s[m+N] - d[m] for m < 1024 (d was stored in place)
            is_diff = false;
            for (int m = tid; m < B8_NWIN - 1; m += B8_THREADS) {
                const int i0 = B8_WPAD(m);
                const double2 dd = B[i0], s_new = B[i0 + 33 * B8_WCH];
                B[i0] = make_double2(s_new.x - dd.x, s_new.y - dd.y);
            }
            __syncthreads();
        }
        // ---- chunk sums sum_{n in chunk} s[n] W^{n k} by Horner's rule from the last sample: thread = (chunk, 4 bins) ----
        if (tid < 2 * B8_NCH) {
            const int cch = tid >> 1, q = tid & 1;
            int kk[4]; double2 z[4], acc[4];
            const double2 *bp = B + 33 * cch;
            const double2 s_last = bp[31];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int k = (k0 + b8_bin_off(pass, 4 * q + b)) % B8_N; if (k < 0) k += B8_N;
                kk[b] = k; z[b] = tw[k]; acc[b] = s_last;
            }
#pragma unroll 31
            for (int i = 30; i >= 0; --i) {
                const double2 s = bp[i];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double nr = fma(acc[b].x, z[b].x, fma(-acc[b].y, z[b].y, s.x));
                    const double ni = fma(acc[b].x, z[b].y, fma(acc[b].y, z[b].x, s.y));
                    acc[b] = make_double2(nr, ni);
                }
            }
            // W^(32 c k) of the thread's four (consecutive) bins from two table entries, W^(32 c (k+1)) = W^(32 c k) W^(32 c): a scattered
            // 16-byte table read costs one L1 wavefront per lane, and the L1/shared data pipe is the busiest unit of this kernel
            double2 rot = tw[(32 * cch * kk[0]) % B8_N];
            const double2 rstep = tw[B8_N + 2 * TW_SEG + cch];    // W^(32 c) = tw[(32 * cch) % B8_N] from the compact copy
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                CS[(4 * q + b) * B8_CSL + cch + 1] = cmul(acc[b], rot);
                rot = cmul(rot, rstep);
            }
        }
        __syncthreads();
        B8_MARK(3);                                              // (restore +) chunk sums
        {   // prefix over chunks: warp b scans bin b (lanes own 3 consecutive chunks)
            double2 *row = CS + warp * B8_CSL;
            constexpr int per = (B8_NCH + 31) / 32;
            const int b0 = 1 + lane * per;
            double lr = 0.0, li = 0.0;
#pragma unroll
            for (int i = 0; i < per; ++i) if (b0 + i <= B8_NCH) { const double2 v = row[b0 + i]; lr += v.x; li += v.y; }
            double ir = lr, ii = li;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double tr = __shfl_up_sync(0xffffffffu, ir, d), ti = __shfl_up_sync(0xffffffffu, ii, d);
                if (lane >= d) { ir += tr; ii += ti; }
            }
            double rr = ir - lr, ri = ii - li;
#pragma unroll
            for (int i = 0; i < per; ++i) if (b0 + i <= B8_NCH) { const double2 v = row[b0 + i]; rr += v.x; ri += v.y; row[b0 + i] = make_double2(rr, ri); }
            if (lane == 0) row[0] = make_double2(0.0, 0.0);
        }
        if (pass == 0 && tid == 0 && wcache) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the bulk store has read the window
        __syncthreads();
        B8_MARK(4);                                              // prefix over chunks
        const int g = tid >> 3, j = tid & 7;
        int k = (k0 + b8_bin_off(pass, j)) % B8_N; if (k < 0) k += B8_N;
        const double2 wk = tw[k];
        double xr, xi;
        {
            const double2 hi = CS[j * B8_CSL + g + B8_WCH], lo = CS[j * B8_CSL + g];
            const double yr = hi.x - lo.x, yi = hi.y - lo.y;
            const double2 t = tw[(32 * g * k) % B8_N];           // X_{m0}[k] = Y_{m0}[k] * exp(+2*pi*i*m0*k/N)
            xr = yr * t.x + yi * t.y;
            xi = yi * t.x - yr * t.y;
        }
        // ---- which (segment, bin) pairs can hold the maximum at all?  |X_{m+1}[k]| <= |X_m[k]| + |s[m]| + |s[m+N]|, so the windows of
        //      segment g stay below |X_{32g}[k]| + slack_g; the largest segment-start power (and the best of earlier passes) is a lower
        //      bound of the final maximum.  Only the few segments around the burst slide; for the bins of the later passes usually none. ----
        const double p0 = fma(xr, xr, xi * xi);
        double gl = p0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gl = fmax(gl, __shfl_xor_sync(0xffffffffu, gl, o));
        if (lane == 0) red_v[warp] = gl;
        __syncthreads();
        gl = fmax(fmax(fmax(red_v[0], red_v[1]), fmax(red_v[2], red_v[3])), fmax(fmax(red_v[4], red_v[5]), fmax(red_v[6], red_v[7])));
        if (pass > 0) gl = fmax(gl, g_best);
        bool need;
        {
            const double slack = (g < B8_NSEG - 1) ? sqrt_ub(31.0 * E31[g]) + sqrt_ub(31.0 * E31[g + B8_WCH])
                                                   : sqrt_ub(32.0 * (PE[g + 1] - PE[g])) + sqrt_ub(32.0 * (PE[g + B8_WCH + 1] - PE[g + B8_WCH]));
            const double bound = (sqrt_ub(p0) + slack) * (1.0 + 1e-9);
            need = !(bound * bound < gl * (1.0 - 1e-6));
        }
        // pass 0 always slides somewhere (the segment that holds the largest start), so only the later passes vote; the barrier inside the
        // branch below (or block_argmax's first one) orders the reads of red_v before it is reused
        const int any_need = (pass == 0) ? 1 : __syncthreads_or(need ? 1 : 0);
        B8_MARK(5);                                              // segment starts, which segments slide
        double best = -1.0; int bestm = 0x7fffffff;
        if (any_need) {
            for (int m = tid; m < B8_NWIN - 1; m += B8_THREADS) {    // d[m] = s[m+N] - s[m] in place
                const int i0 = B8_WPAD(m);
                const double2 s_old = B[i0], s_new = B[i0 + 33 * B8_WCH];
                B[i0] = make_double2(s_new.x - s_old.x, s_new.y - s_old.y);
            }
            is_diff = true;
            __syncthreads();
            if (need) {
                const double wr = wk.x, wi = -wk.y;
                const double2 *dp = B + 33 * g;                      // d[32 g + i] sits at 33 g + i
                int besti = 0;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) {
                    const double p = fma(xr, xr, xi * xi);
                    if (p > best) { best = p; besti = i; }
                    if (i < 31 || g == B8_NSEG - 1) {
                        const double2 d = dp[i];
                        const double tr = xr + d.x, ti = xi + d.y;
                        xr = fma(tr, wr, -(ti * wi));
                        xi = fma(tr, wi, ti * wr);
                    }
                }
                if (g == B8_NSEG - 1) {                              // the last segment also owns window 1024
                    const double p = fma(xr, xr, xi * xi);
                    if (p > best) { best = p; besti = 32; }
                }
                bestm = 32 * g + besti;
            }
        }
        B8_MARK(6);                                              // differences + slide
        block_argmax(best, bestm, red_v, red_i);
        B8_MARK(7);                                              // block argmax
        if (pass == 0) { g_best = best; g_bestm = bestm; }
        else if (best > g_best || (best == g_best && bestm < g_bestm)) break;   // the extra bins would move the argmax: leave it to tier 2
        // ---- certificate at every segment-start window c = 32 g (g = 0..32), straight from the chunk prefix tables: for any split of
        // the window into a part P1 of d samples and the rest P2, every untracked bin obeys
        //   |X_c[k]| <= sqrt(d*E_P1) + sqrt(N*E_P2 - sum_tracked |P2[k']|^2)
        // and the triangle inequality carries the bound to the 31 windows in between with the slack
        //   sum_{i=c}^{c+30} (|s[i]| + |s[i+N]|) <= sqrt(31*E31[g]) + sqrt(31*E31[g+37])        (Cauchy-Schwarz). ----
        ok = 1;
        for (int w0 = 0; w0 <= B8_NSEG; w0 += B8_THREADS / 8) {
            if (w0 > 0 && warp > 0) break;                       // the second round is window 32 alone: 8 threads of warp 0
            const int gq = w0 + (tid >> 3), cand = tid & 7;
            double bnd = INFINITY;
            if (gq <= B8_NSEG && cand < 7) {
                const int cw = gq * B8_CH;
                const bool lead = cw < g_bestm;
                const int dist = lead ? g_bestm - cw : cw - g_bestm;
                const int dch = (cand == 0) ? 0 : (dist + B8_CH - 1) / B8_CH - 3 + cand;
                if (dch == 0 || (cw != g_bestm && dch >= 1 && dch < B8_WCH)) {
                    // P1 = first dch chunks (window starts before the burst) or last dch chunks (window runs past it)
                    const int p1a = lead ? gq : gq + B8_WCH - dch, p1b = p1a + dch;
                    const int p2a = lead ? gq + dch : gq, p2b = lead ? gq + B8_WCH : gq + B8_WCH - dch;
                    const double e1 = PE[p1b] - PE[p1a], e2 = PE[p2b] - PE[p2a];
                    double t2 = (pass == 0) ? 0.0 : T2[gq * 8 + cand];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const double2 hi = CS[jj * B8_CSL + p2b], lo = CS[jj * B8_CSL + p2a];
                        const double yr = hi.x - lo.x, yi = hi.y - lo.y;
                        t2 += yr * yr + yi * yi;
                    }
                    T2[gq * 8 + cand] = t2;
                    const double r2 = (double)B8_N * e2 - t2;
                    bnd = sqrt_ub((double)(dch * B8_CH) * (e1 > 0.0 ? e1 : 0.0)) + sqrt_ub(r2 > 0.0 ? r2 : 0.0);
                }
            }
            bnd = fmin(bnd, __shfl_xor_sync(0xffffffffu, bnd, 1));
            bnd = fmin(bnd, __shfl_xor_sync(0xffffffffu, bnd, 2));
            bnd = fmin(bnd, __shfl_xor_sync(0xffffffffu, bnd, 4));
            if (gq <= B8_NSEG && cand == 0) {
                const double A = (gq == B8_NSEG) ? 0.0 : sqrt_ub(31.0 * E31[gq]) + sqrt_ub(31.0 * E31[gq + B8_WCH]);
                const double bound = (bnd + A) * (1.0 + 1e-9);   // the prefix differences and sums above round to nearest
                if (!(bound * bound < g_best * (1.0 - 1e-6))) ok = 0;
            }
        }
        ok = __syncthreads_and(ok);
        B8_MARK(8);                                              // certificate
    }
    if (PROF && tid == 0) atomicAdd(prof + 15, 1ull);
    if (tid == 0) {
        if (pass_hist) atomicAdd(pass_hist + (ok ? pass : 0), 1u);             // [p] = bursts proven after p passes, [0] = left open
        if (wcache) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the bulk store is complete before the block retires
        *o = (double)(sp0 + 1 + g_bestm);                        // sp + max_idx - 1
        need_band[(i64)stream * cap + burst] = (ok && !force_fail) ? 0 : 1;       // force_fail: test hook, sends every burst to tier 2
    }
}

// ===================================================================================================
// K8 at osr 8 from the filtered-window cache: per-burst tone frequency (+ SNR gate)
//   FCCH_fine_correction.m:143-155,185-189 (which = 1, level 1: resampled stream) and carrier_correct_post_SCH.m:58-72
//   (which = 2, level 3: resampled, derotated, resampled again).
// Same statements as tone_est_kernel, rebuilt for what the cached level-0 window makes possible: no staging, no FIR, two ping-pong
// buffers (39 KB: 4 blocks per SM), and the two partial DFTs (8-bin band around the phase-slope estimate, 5 gate bins) by Horner's
// rule over 10-sample segments with 4 bins (2 gate-bin pairs) per thread sharing every sample load.  A burst whose level-0 range is
// not covered by its cached window, or whose Parseval certificates (integer bin, 5 dB gate) do not hold, is flagged in `need_old`
// and recomputed by tone_est_kernel (all 1184 bins through the 37 x 32 row FFT).
// ===================================================================================================
#define T8_K 5                               // consecutive samples per thread in the fused passes
// interp1 'linear' of consecutive outputs (load_window's arithmetic): the upper neighbour of one output is the lower neighbour of the next
// except where the resampling ratio makes the source index skip or repeat, so it is kept in registers
// sample r of the level-0 range as the TMA bulk copy left it in shared memory: the cached layout (one pad slot per 32 samples of the
// window) starting at window sample o
__device__ __forceinline__ int t8_pad(int r, int o) { return r + ((o + r) >> 5) - (o >> 5); }
struct T8Lerp {
    i64 prev_i1 = -1;
    double2 prev_hi = {0.0, 0.0};
    template <bool PAD>
    __device__ __forceinline__ double2 next(const double2 *S, i64 a_src, i64 last, i64 j, double s, int o = 0) {
        const double xq = (double)j * s;
        i64 i0 = (i64)floor(xq);
        if (i0 > last) i0 = last;
        const i64 i1 = (i0 + 1 > last) ? last : i0 + 1;
        const int r0 = (int)(i0 - a_src), r1 = (int)(i1 - a_src);
        const double2 lo = (i0 == prev_i1) ? prev_hi : S[PAD ? t8_pad(r0, o) : r0];
        const double2 hi = (i1 == i0) ? lo : S[PAD ? t8_pad(r1, o) : r1];
        prev_i1 = i1; prev_hi = hi;
        return lerp_ref(lo, hi, xq - (double)i0);
    }
};
#define T8_THREADS 256
#define T8_SEG     10
#define T8_NSEG    119                       // ceil(1184 / 10); the last segment is padded with zeros
#define T8_BUF     1256                      // 1208 samples of a level + the pad slots of the cached layout they span (one per 32)
#define T8_SMEM    (2 * T8_BUF * 16)
#define T8_LVL     1208                      // samples a resampling level may hold
template <bool PROF>
__global__ void __maxnreg__(56) tone8_kernel(WinSrc src, const StreamCtl *__restrict__ ctl, int which, const double *__restrict__ pos, int cap,
                                                             const double2 *__restrict__ tw, double *__restrict__ fo_out, double *__restrict__ gate_out,
                                                             int *__restrict__ need_old, unsigned long long *__restrict__ prof) {
    extern __shared__ __align__(16) double2 t8_sm[];
    long long t_prev = PROF ? clock64() : 0;                     // debug hook as in fine_core8_kernel: prof[ph] += cycles, prof[15] = blocks
#define T8_MARK(ph) do { if (PROF && threadIdx.x == 0) { const long long t_now = clock64(); atomicAdd(prof + (ph), (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
    double2 *P = t8_sm, *Q = t8_sm + T8_BUF;
    __shared__ double red_n[24];
    __shared__ double2 sh_base, sh_step;
    __shared__ double sh_pb[8], sh_E, sh_pr;
    __shared__ double2 sh_x[8];
    __shared__ int sh_k0, sh_jbest, sh_flag;
    __shared__ __align__(8) unsigned long long t8_bar;
    const int burst = blockIdx.x, stream = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const i64 idx_o = (i64)stream * cap + burst;
    // fetched side by side with the control block (one global-memory latency instead of three); garbage beyond the burst count is not used
    const double pos_d = __ldg(pos + idx_o), wpos_d = src.wcache ? __ldg(src.wc_pos + idx_o) : 0.0;
    const StreamCtl c = ctl[stream];
    constexpr int N = B8_N;
    const int nb = (which == 1) ? (c.tone1_enable ? c.n_fcch : 0) : (c.post_enable ? c.n_post_fcch : 0);
    if (burst >= nb) return;
    const double sampling_rate = ((1625.0 / 6.0) * 1e3) * 8.0;
    const i64 start = (i64)pos_d - 1;
    // ---- index ranges of the levels (load_window's arithmetic) ----
    const i64 n0 = src.n_iq;
    const bool use2 = (src.level == 3) && c.interp2_on;
    const bool use1 = (src.level >= 1) && c.interp1_on;
    const bool derot = (src.level >= 2) && c.derot1_on;
    const double s1 = 1.0 + c.e1, s2 = 1.0 + c.e2;
    const i64 len1 = use1 ? c.len1 : n0;
    i64 a2 = start, b2 = start + N - 1, a1 = a2, b1 = b2, a0, b0;
    if (use2) {
        a1 = (i64)floor((double)a2 * s2);
        b1 = (i64)floor((double)b2 * s2) + 1;
        if (b1 > len1 - 1) b1 = len1 - 1;
        if (a1 > b1) a1 = b1;
    }
    a0 = a1; b0 = b1;
    if (use1) {
        a0 = (i64)floor((double)a1 * s1);
        b0 = (i64)floor((double)b1 * s1) + 1;
        if (b0 > n0 - 1) b0 = n0 - 1;
        if (a0 > b0) a0 = b0;
    }
    if (a0 < 0) a0 = 0;
    if (b0 > n0 - 1) b0 = n0 - 1;
    const int n_l0 = (int)(b0 - a0 + 1), n_l1 = (int)(b1 - a1 + 1);
    const i64 wbase = ((i64)wpos_d - 65) * 8;
    const bool covered = src.wcache && a0 >= wbase && b0 < wbase + B8_NSMP && n_l0 <= T8_LVL && n_l1 <= T8_LVL && start >= 0
                         && (!derot || use1);                     // (derotation without resampling does not occur in the reference flow)
    if (!covered) {                                              // block-uniform
        if (tid == 0) need_old[idx_o] = 1;
        return;
    }
    // ---- level 0 from the cache -> P: ONE TMA bulk copy of the contiguous part of the cached (padded) window that holds the range; it
    //      stays in the padded layout (t8_pad) - no LDG + STS pass, no index arithmetic per sample ----
    const int o = (int)(a0 - wbase);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&t8_bar);
    if (tid == 0) {
        const double2 *wc = src.wcache + idx_o * B8_WLEN + B8_WPAD(o);
        const unsigned bytes = (unsigned)((B8_WPAD(o + n_l0 - 1) - B8_WPAD(o) + 1) * sizeof(double2));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((unsigned)__cvta_generic_to_shared(P)), "l"(wc), "r"(bytes), "r"(bar) : "memory");
        if (derot) { double sn, cs; sincos((double)a1 * c.dphi1, &sn, &cs); sh_base = make_double2(cs, sn); }
    }
    __syncthreads();                                             // the barrier is initialised (and sh_base written) for everyone
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
    T8_MARK(0);                                                  // level 0 from the cache (global-memory latency)
    // The L1/shared-memory data pipe is the busiest unit of this kernel (ncu: 68 % against 34 % for the FP64 pipe), so the passes below
    // give every thread T8_K CONSECUTIVE samples: interp1 re-uses the upper neighbour of one output as the lower neighbour of the next
    // (K+1 loads for K outputs instead of 2K), and the sums over neighbouring samples (energy + phase slope here, the phasor ratios
    // after the band DFT) are taken from registers while the samples are produced - they cost no shared-memory reads of their own.
    // A stride of 5 x 16 bytes keeps the 128-bit accesses of a quarter-warp on different banks.
    const int nf = T8_K * tid;                                   // first sample of the thread (237 threads cover the burst)
    double2 *u = P, *spare = Q;                                  // u: the burst the stage works on; spare: the other buffer
    double epq[3] = {0.0, 0.0, 0.0};                             // energy, sum of x[n+1] * conj(x[n])
    auto energy_pair = [&](int k, int n, const double2 v, double2 &prev) {     // v = final-level sample n (k: its number in the thread)
        if (k < T8_K) epq[0] = fma(v.x, v.x, fma(v.y, v.y, epq[0]));
        if (k > 0) { const double2 q = cmulc(v, prev); epq[1] += q.x; epq[2] += q.y; }
        prev = v;
    };
    if (use1) {                                                  // level 1 (+2): interp1 by (1+e1) [and derotation by dphi1] -> Q
        double2 ph = make_double2(1.0, 0.0), st = make_double2(1.0, 0.0);
        if (derot) {
            double sn, cs;
            sincos((double)nf * c.dphi1, &sn, &cs); ph = cmul(sh_base, make_double2(cs, sn));
            sincos(c.dphi1, &sn, &cs); st = make_double2(cs, sn);
        }
        T8Lerp lc;
        double2 prev = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k <= T8_K; ++k) {                        // (the K+1-th sample only feeds the pair sum when this level is the final one)
            const int i = nf + k;
            if (i < n_l1 && (k < T8_K || !use2)) {
                double2 v = lc.next<true>(P, a0, n0 - 1, a1 + i, s1, o);
                if (derot) { v = cmul(v, ph); ph = cmul(ph, st); }
                if (k < T8_K) Q[i] = v;
                if (!use2 && i < N) energy_pair(k, i, v, prev);
            }
        }
        u = Q; spare = P;
        if (use2) {                                              // level 3: second interp1 by (1+e2) -> P
            __syncthreads();
            T8Lerp l2;
#pragma unroll
            for (int k = 0; k <= T8_K; ++k) {
                const int i = nf + k;
                if (i < N) {
                    const double2 v = l2.next<false>(Q, a1, len1 - 1, a2 + i, s2);
                    if (k < T8_K) P[i] = v;
                    energy_pair(k, i, v, prev);
                }
            }
            u = P; spare = Q;
        }
    } else if (use2) {                                           // e1 path off, second interp1 only (no derotation here, see `covered`)
        T8Lerp l2;
        double2 prev = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k <= T8_K; ++k) {
            const int i = nf + k;
            if (i < N) {
                const double2 v = l2.next<true>(P, a1, len1 - 1, a2 + i, s2, o);
                if (k < T8_K) Q[i] = v;
                energy_pair(k, i, v, prev);
            }
        }
        u = Q; spare = P;
    } else {                                                     // level 0 is the burst: out of the padded layout -> Q
        double2 prev = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k <= T8_K; ++k) {
            const int i = nf + k;
            if (i < N) {
                const double2 v = P[t8_pad(i, o)];
                if (k < T8_K) Q[i] = v;
                energy_pair(k, i, v, prev);
            }
        }
        u = Q; spare = P;
    }
    if (tid < T8_SEG * T8_NSEG - N + 2) u[N + tid] = make_double2(0.0, 0.0);     // zero padding of the last Horner segment
    T8_MARK(1);                                                  // interp1 / derotation levels (+ energy and phase slope on the fly)
    // ---- energy and phase slope -> band centre ----
    block_sum_n<3, false>(epq, red_n);
    if (tid == 0) { sh_E = epq[0]; sh_k0 = (int)floorf(atan2f((float)epq[2], (float)epq[1]) * (float)(N / (2.0 * GSMCAL_PI)) + 0.5f); sh_flag = 0; }   // fp32: see fine_core8_kernel
    __syncthreads();
    T8_MARK(2);                                                  // block sums, band centre
    const int k0 = sh_k0;
    // ---- 8-bin band DFT by Horner's rule: thread = (10-sample segment, 4 bins); partial sums -> spare[bin][120], the rows of the
    //      upper four bins shifted by 4 slots so the two threads of a segment store to different banks ----
    if (tid < 2 * T8_NSEG) {
        const int seg = tid >> 1, q = tid & 1, n0s = T8_SEG * seg;
        double2 z[4], acc[4];
        const double2 s_last = u[n0s + T8_SEG - 1];
        int kf = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int k = (k0 - 3 + 4 * q + b) % N; if (k < 0) k += N;
            if (b == 0) kf = k;
            z[b] = tw[k]; acc[b] = s_last;
        }
#pragma unroll
        for (int i = T8_SEG - 2; i >= 0; --i) {
            const double2 s = u[n0s + i];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const double nr = fma(acc[b].x, z[b].x, fma(-acc[b].y, z[b].y, s.x));
                const double ni = fma(acc[b].x, z[b].y, fma(acc[b].y, z[b].x, s.y));
                acc[b] = make_double2(nr, ni);
            }
        }
        // W^(n0 k) of the four bins from two table entries: W^(n0 (k+1)) = W^(n0 k) W^n0 (scattered 16-byte table reads cost a
        // wavefront per lane on the L1 data pipe)
        double2 rot = tw[(n0s * kf) % N];
        const double2 rstep = tw[N + seg];                       // W^(10 seg) from the compact copy (same value as tw[n0s])
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            spare[(4 * q + b) * 120 + 4 * q + seg] = cmul(acc[b], rot);
            rot = cmul(rot, rstep);
        }
    }
    __syncthreads();
    T8_MARK(3);                                                  // Horner band DFT
    {   // warp b adds the partials of bin b
        double xr = 0.0, xi = 0.0;
        const double2 *row = spare + warp * 120 + 4 * (warp >> 2);
        for (int sgi = lane; sgi < T8_NSEG; sgi += 32) { const double2 v = row[sgi]; xr += v.x; xi += v.y; }
        xr = warp_sum(xr); xi = warp_sum(xi);
        if (lane == 0) sh_x[warp] = make_double2(xr, xi);
    }
    __syncthreads();
    if (warp == 0) {   // first maximum in fftshift-ed order (:149-150) and the Parseval certificate for every bin outside the band;
                       // abs(.)^2 of the 8 bins by 8 lanes of ONE warp (hypot is ~150 instructions: not once per warp)
        double pw = 0.0, v = -1.0; int j = 0x7fffffff;
        if (lane < 8) {
            int k = (k0 - 3 + lane) % N; if (k < 0) k += N;
            j = k - N / 2; if (j < 0) j += N;
            pw = abs2_ref(sh_x[lane]); v = pw;
        }
        double band_sum = pw;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            band_sum += __shfl_down_sync(0xffffffffu, band_sum, o);
            const double v2 = __shfl_down_sync(0xffffffffu, v, o);
            const int j2 = __shfl_down_sync(0xffffffffu, j, o);
            argmax_combine(v, j, v2, j2);
        }
        if (lane == 0) {
            if (!((double)N * sh_E - band_sum < v * (1.0 - 1e-9))) sh_flag = 1;
            sh_jbest = j;
        }
    }
    __syncthreads();
    T8_MARK(4);                                                  // bin totals, argmax, certificate
    if (sh_flag) {                                               // block-uniform: not certified, the row-FFT kernel takes the burst
        if (tid == 0) need_old[idx_o] = 1;
        return;
    }
    const int jr = sh_jbest + 1 - ((N / 2) + 1);                 // max_idx - (fft_len/2 + 1)
    const double int_phase_rotate = 2.0 * GSMCAL_PI * (double)jr / (double)N;
    // ---- integer-bin derotation (:152-153) -> g (the other buffer; the SNR gate reads it), unit phasors and the mean phasor ratio of
    //      neighbours from registers: the twiddle of the thread's first sample from the table, then one step per sample ----
    double2 *g = spare;
    {
        int jm = jr % N; if (jm < 0) jm += N;
        double rri[2] = {0.0, 0.0};
        if (nf < N) {
            double2 ph = tw[(nf * jm) % N];
            const double2 stp = tw[jm];
            double2 eprev = make_double2(1.0, 0.0);
#pragma unroll
            for (int k = 0; k <= T8_K; ++k) {
                const int n = nf + k;
                if (n < N) {
                    const double2 w = cmul(u[n], ph);
                    ph = cmul(ph, stp);
                    if (k < T8_K && which == 1) g[n] = w;
                    const double h2 = fma(w.x, w.x, w.y * w.y);
                    const double inv = rsqrt(h2);
                    const double2 e = (h2 > 0.0) ? make_double2(w.x * inv, w.y * inv) : make_double2(1.0, 0.0);
                    if (k > 0) {
                        rri[0] += e.x * eprev.x + e.y * eprev.y;
                        rri[1] += e.y * eprev.x - e.x * eprev.y;
                    }
                    eprev = e;
                }
            }
        }
        if (which == 1 && tid < T8_SEG * T8_NSEG - N + 2) g[N + tid] = make_double2(0.0, 0.0);
        T8_MARK(5);                                              // integer-bin derotation, unit phasors, ratios
        block_sum_n<2, false>(rri, red_n);
        if (tid == 0) {
            const double pr = atan2(rri[1] / (double)(N - 1), rri[0] / (double)(N - 1));
            sh_pr = pr;
            fo_out[idx_o] = sampling_rate * (int_phase_rotate + pr) / (2 * GSMCAL_PI);
            double sn, cs; sincos(pr, &sn, &cs); sh_step = make_double2(cs, sn);      // exp(+i*phi): one step back in n
        }
    }
    T8_MARK(6);                                                  // block sums, atan2, fo
    if (PROF && which != 1 && tid == 0) atomicAdd(prof + 15, 1ull);
    if (which != 1) return;
    __syncthreads();
    // ---- SNR gate (:185-196): bins 0, +-1, +-2 of the finely derotated burst by Horner's rule, thread = (segment, bin pair);
    //      sig >= 10^0.5 (N*E - sig) proves the burst passes the 5 dB gate (Parseval; derotation keeps the energy).  Partial sums go to
    //      the buffer the burst came from (rows 2.. shifted by 4 slots: different banks for the two threads of a segment) ----
    const double phase_rotate = sh_pr;
    if (tid < 2 * T8_NSEG) {
        const int seg = tid >> 1, q = tid & 1, n0s = T8_SEG * seg, n_last = n0s + T8_SEG - 1;
        const double2 zp = tw[q + 1], zm = make_double2(zp.x, -zp.y);      // W^{+(q+1)}, W^{-(q+1)}
        double sn, cs; sincos((double)n_last * phase_rotate, &sn, &cs);
        double2 ph = make_double2(cs, -sn);                      // exp(-i*n*phi) at the segment's last sample
        const double2 st = sh_step;
        double2 accp, accm, acc0;
        {
            const double2 v = cmul(g[n_last], ph);
            accp = v; accm = v; acc0 = v;
        }
#pragma unroll
        for (int i = T8_SEG - 2; i >= 0; --i) {
            ph = cmul(ph, st);
            const double2 v = cmul(g[n0s + i], ph);
            accp = make_double2(fma(accp.x, zp.x, fma(-accp.y, zp.y, v.x)), fma(accp.x, zp.y, fma(accp.y, zp.x, v.y)));
            accm = make_double2(fma(accm.x, zm.x, fma(-accm.y, zm.y, v.x)), fma(accm.x, zm.y, fma(accm.y, zm.x, v.y)));
            acc0.x += v.x; acc0.y += v.y;
        }
        const double2 t = tw[N + TW_SEG * q + seg];              // W^{+n0 (q+1)} = tw[(n0s * (q + 1)) % N] from the compact copies; its conjugate for the negative bin
        u[(2 * q) * 120 + 4 * q + seg] = cmul(accp, t);
        u[(2 * q + 1) * 120 + 4 * q + seg] = cmul(accm, make_double2(t.x, -t.y));
        if (q == 0) u[4 * 120 + 4 + seg] = acc0;
    }
    __syncthreads();
    T8_MARK(7);                                                  // gate: Horner over the finely derotated burst
    if (warp < 5) {
        double xr = 0.0, xi = 0.0;
        const double2 *row = u + warp * 120 + (warp >= 2 ? 4 : 0);
        for (int sgi = lane; sgi < T8_NSEG; sgi += 32) { const double2 v = row[sgi]; xr += v.x; xi += v.y; }
        xr = warp_sum(xr); xi = warp_sum(xi);
        if (lane == 0) sh_pb[warp] = xr * xr + xi * xi;
    }
    __syncthreads();
    if (tid == 0) {
        const double sig5 = ((sh_pb[0] + sh_pb[1]) + (sh_pb[2] + sh_pb[3])) + sh_pb[4];
        if (sig5 >= 3.16227766016838 * (1.0 + 1e-9) * ((double)N * sh_E - sig5)) gate_out[idx_o] = 99.0;      // "certified above the 5 dB gate"
        else need_old[idx_o] = 1;                                // evaluate the 110 gate bins exactly
    }
    T8_MARK(8);                                                  // gate sums, certificate
    if (PROF && tid == 0) atomicAdd(prof + 15, 1ull);
}
