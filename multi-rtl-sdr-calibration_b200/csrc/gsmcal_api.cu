// gsmcal_api.cu - the C ABI declared in include/gsmcal.h: host-side sequencing of the sm_100a kernels.
// No torch types, no CPU fallback: every compute entry point fails with GSMCAL_ERR_CUDA without a device.
#include "../../include/gsmcal.h"
#include "gsmcal_kernels.cuh"
#include "gsmcal_burst8.cuh"
#include "gsmcal_demod.cuh"

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "chn_filter_taps.inc"

static_assert(sizeof(gsmcal_stream_result) == sizeof(StreamResultDev), "result record layout");

namespace {

std::mutex g_mu;
thread_local std::string g_err;
thread_local int g_device = 0;
std::atomic<long long> g_launches{0};
int *g_last_need_full = nullptr, *g_last_need_band = nullptr; long long g_last_need_full_n = 0;
int g_debug_groups = 4;           // stream groups per batch (1 = strictly sequential stages, enables per-stage timing)
int g_debug_fall_limit = FALL_GRID;  // tier 3 handles work lists up to this length (tests set 0 to exercise the single-block list mode)
int g_debug_fail_tier2 = 0;       // tests: pretend the 64-bin certificate failed
int g_debug_force_full = 0;       // tests: run the all-bin fine search for every burst
int g_debug_submit_groups = 2;    // stream groups inside a submitted batch (debug key 8)
int g_debug_persist_colsum = 0;   // submit/collect: blocks per SM of the persistent high-priority column-sum kernel (0 = per-group launches)
int g_debug_hi_prio = 1;          // burst chain of each stream group on a high-priority CUDA stream
int g_debug_core8_passes = B8_NPASS; // passes of 8 tracked bins in fine_core8_kernel (1..8)
unsigned *g_last_pass_hist = nullptr, *g_prev_pass_hist = nullptr;   // counters of the last / the one-before-last batch (debug_get 50.., 150..)
int g_debug_persist_threads = 256;  // block size of the persistent column-sum kernel (debug key 15)
int g_debug_occupy = 0;            // debug key 23: KB of shared memory per block of an extra occupy_kernel launch behind the chain (0 = none)
int g_debug_chain_serial = 0;      // debug key 21: the burst chains of a batch's stream groups run one after the other
int g_debug_chain_lo = 0;          // debug key 22: ... on normal-priority streams (they only take what the FP64 blocks leave free)
int g_debug_sch56 = 0;             // debug key 19: sch_corr_kernel capped at 56 registers
int g_debug_trickle = 2;           // debug key 17: ring stages (x 4 KB) of colsum_u8_trickle_kernel in submit/collect, 0 = off (plain column-sum launches)
int g_debug_trickle_blocks = 3;    // debug key 18: its blocks per SM (3 x 2 stages: 22 GB in 7 ms under the FP64 stages, profiles/r2q)
int g_debug_gate = 1;              // debug key 16: submit/collect staggers the batches - the FP64 stages of batch k+1 wait for batch k, its front runs under them
int g_debug_timeline = 0;          // submit/collect print the device timeline of every batch to stderr (A/B of overlap)
cudaEvent_t g_tl_base = nullptr;
int g_debug_prof = 0;              // fine_core8_kernel accumulates per-phase cycle counts (debug_get 50..65)
int g_debug_no_tone8 = 0;          // tests / A-B: 1 = generic tone estimator for every burst
int g_debug_no_core8 = 0;         // tests / A-B: 1 = round-1 tier-1 kernel and no filtered-window cache

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) return fail(GSMCAL_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define TRY(expr) do { int rc_ = (expr); if (rc_ != GSMCAL_OK) return rc_; } while (0)
#define LAUNCH(kern, grid, block, smem, st, ...)                                                          \
    do { kern<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__); ++g_launches; CU(cudaGetLastError()); } while (0)

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int get(size_t bytes, void **out) {
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr; cap = 0;
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
            if (e != cudaSuccess) return fail(GSMCAL_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
            cap = want;
        }
        *out = p;
        return GSMCAL_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// one in-flight batch of the submit/collect API: its own workspace, streams and pinned result staging, so that consecutive batches
// overlap on the device (the ~1 ms latency-bound front of batch k+1 - column sums, burst chain - runs under the FP64 kernels of batch k)
constexpr int kSlots = 4;
struct Slot {
    DevBuf work, wc;                                            // wc: filtered-window cache of the osr-8 fast path
    cudaStream_t front = nullptr, front_hi = nullptr;
    std::vector<cudaStream_t> co;                               // debug key 22: normal-priority streams for the burst chains
    std::vector<cudaStream_t> grp, hi;
    cudaEvent_t done = nullptr;
    cudaEvent_t tl[8] = {};                                     // debug key 14: front start, column sums done, burst chain done, FP64 stages done,
                                                                // FP64 stages start, fine search done, fine tone done, SCH done
    bool tl_on = false;
    bool busy = false;
    char *stage = nullptr; size_t stage_cap = 0;               // pinned host staging of the results
    gsmcal_stream_result *results = nullptr; double *coarse_pos = nullptr, *coarse_snr = nullptr, *fcch_pos = nullptr, *pos_info = nullptr;
    size_t n_res = 0, n_per = 0;                                // bytes of the result records / of one D*cap double array
};
#include "gsmcal_hostcopy.inc"

constexpr int kNumStageEvents = 8;
struct Ctx {
    StageRing ring;                  // pinned staging for pageable host buffers
    int last_slot = -1;              // slot of the most recently submitted batch (submit/collect pipeline)
    bool attrs = false;
    cudaEvent_t stage_ev[kNumStageEvents];
    bool stage_ev_ok = false;
    Slot slots[kSlots];
    DevBuf in, out, work, tplbuf, wc;
    std::map<int, double2 *> tw;     // N -> exp(-2*pi*i*j/N)
    std::vector<cudaStream_t> side;  // extra streams: stream groups of a batch overlap their latency-bound stages
    std::vector<cudaStream_t> side_hi; // one high-priority stream per group for the latency-bound burst chain (see gsmcal_calibrate_batch)
};
std::map<int, Ctx> g_ctx;

// per-stage CUDA-event timing of the last gsmcal_calibrate_batch call (events are recorded on the call's stream; they belong to the
// device's Ctx - an event of another device cannot be recorded on this one's streams)
int g_stage_n = 0;
float g_stage_ms[kNumStageEvents];
int stage_mark(Ctx &c, cudaStream_t st) {
    if (!c.stage_ev_ok) {
        for (int i = 0; i < kNumStageEvents; ++i) CU(cudaEventCreate(&c.stage_ev[i]));
        c.stage_ev_ok = true;
    }
    if (g_stage_n < kNumStageEvents) CU(cudaEventRecord(c.stage_ev[g_stage_n++], st));
    return GSMCAL_OK;
}

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(GSMCAL_ERR_CUDA, "no CUDA device (the sm_100a kernels are the only implementation; there is no CPU fallback)"); }
    CU(cudaSetDevice(g_device));
    return GSMCAL_OK;
}

constexpr int kFineThreads = 320;
size_t fine_smem(int osr) { int N = 148 * osr, ns = 2 * 64 * osr + 1 + N - 1; return (size_t)(2 * ns + 8 + GSMCAL_XCAP(ns)) * sizeof(double2); }
size_t fine_band_smem(int osr) {
    // window + scratch; the scratch holds the loader's staging for one third of the window, later the piece / chunk sums
    // and the three 16-sample prefix arrays of the certificate
    const int N = 148 * osr, ns = 2 * 64 * osr + 1 + N - 1, n_chunk = ns / 16, pre = (3 * (n_chunk + 2) + 1) / 2 + 2;
    int x = GSMCAL_XCAP((ns + 2) / 3);
    const int need_band = 8 * FB_BINS + pre, need_core = (ns / (4 * osr) + 1) * FC_BINS + pre + 4 + (2 * 64 * osr / (4 * osr) + 2) * 4;   // + pass-0 tracked powers [n_seg+1][8] doubles
    if (x < need_band) x = need_band;
    if (x < need_core) x = need_core;
    return (size_t)(ns + x) * sizeof(double2);
}
size_t tone_smem(int osr) {
    int N = 148 * osr, work = GSMCAL_XCAP(N) + N + 8;
    if (work < 2 * N) work = 2 * N;
    return (size_t)(N + work) * sizeof(double2);
}
size_t sch_smem(int osr) {
    int L = 64 * osr, ns = 16 * osr - 5 * osr + 1 + L - 1;
    size_t slots = (size_t)(2 * ns + L + 8 + GSMCAL_XCAP(ns));
    if (osr == 8) {     // register-tiled path: padded window + padded template + 96 partial-sum rows of 9 doubles behind X
        size_t need = (size_t)(ns + L + 8) + (size_t)(ns + (ns >> 4) + 2) + (size_t)(L + (L >> 4) + 2) + (96 * 9 + 1) / 2 + 2;
        if (slots < need) slots = need;
    }
    return slots * sizeof(double2);
}

int get_ctx(Ctx **out) {
    TRY(ensure_device());
    Ctx &c = g_ctx[g_device];
    if (!c.attrs) {
        CU(cudaFuncSetAttribute(fine_peak_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(cudaFuncSetAttribute(fine_peak_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(cudaFuncSetAttribute(fine_peak_core_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(cudaFuncSetAttribute(fine_core8_kernel<47, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B8_SMEM));
        CU(cudaFuncSetAttribute(fine_core8_kernel<47, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B8_SMEM));
        CU(cudaFuncSetAttribute(fine_core8_kernel<48, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B8_SMEM));
        CU(cudaFuncSetAttribute(fine_core8_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B8_SMEM));
        CU(cudaFuncSetAttribute(tone8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM));
        CU(cudaFuncSetAttribute(tone8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T8_SMEM));
        CU(cudaFuncSetAttribute(tone_est_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CU(cudaFuncSetAttribute(sch_corr_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(sch_corr_kernel<56>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(colsum_u8_trickle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 24 * (TRK_STAGE + 8)));
        CU(cudaFuncSetAttribute(occupy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        CU(cudaFuncSetAttribute(colsum_u8_trickle_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CU(cudaFuncSetAttribute(materialise_r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fir_full_tma_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fir_full_tma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(coarse_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fine_ppm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fine_carrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(sch_ppm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(post_carrier_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fir_decim_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fir_decim_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        CU(cudaFuncSetAttribute(fcch_demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));      // (attributes are per device:
        CU(cudaFuncSetAttribute(fde_template_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));    //  all of them live here, keyed by
        CU(cudaFuncSetAttribute(sch_demod_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));       //  the device's Ctx)
        c.attrs = true;
    }
    *out = &c;
    return GSMCAL_OK;
}

int get_twiddle(Ctx &c, int N, cudaStream_t st, const double2 **out) {
    auto it = c.tw.find(N);
    if (it == c.tw.end()) {
        double2 *p = nullptr;
        CU(cudaMalloc(&p, sizeof(double2) * (N + TW_EXTRA)));
        LAUNCH(twiddle_init_kernel, (N + TW_EXTRA + 127) / 128, 128, 0, st, p, N);
        c.tw[N] = p;
        *out = p;
    } else *out = it->second;
    return GSMCAL_OK;
}

double g_taps_host[GSMCAL_MAX_TAPS + 8]; int g_taps_n = -1; int g_taps_dev = -1;
bool any_slot_busy() {
    for (auto &kv : g_ctx) for (Slot &sl : kv.second.slots) if (sl.busy) return true;
    return false;
}
int set_taps(const double *coef, int n_taps, cudaStream_t st) {
    if (n_taps < 1 || n_taps > GSMCAL_MAX_TAPS) return fail(GSMCAL_ERR_ARG, "n_taps must be 1..%d", GSMCAL_MAX_TAPS);
    double tmp[GSMCAL_MAX_TAPS + 8];
    memset(tmp, 0, sizeof tmp);                          // zero padding (4 before, the rest after) adds exact zeros
    memcpy(tmp + 4, coef, sizeof(double) * n_taps);
    if (g_taps_n == n_taps && g_taps_dev == g_device && memcmp(tmp, g_taps_host, sizeof tmp) == 0) return GSMCAL_OK;   // the constant table already holds them
    if (any_slot_busy()) CU(cudaDeviceSynchronize());    // submitted batches still read the old table
    CU(cudaMemcpyToSymbolAsync(c_tapsp, tmp, sizeof tmp, 0, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));                       // tmp lives on this stack frame
    memcpy(g_taps_host, tmp, sizeof tmp); g_taps_n = n_taps; g_taps_dev = g_device;
    return GSMCAL_OK;
}

// ---- workspace carve-up ----------------------------------------------------------------------------
struct Work {
    StreamCtl *ctl; StreamResultDev *res;
    double *coarse_pos, *coarse_snr, *fine_raw, *fcch_pos, *fo, *gate, *sch_raw, *sch_pos, *post_pos, *pos_info, *snr_map, *power;
    int *sch_edge, *need_full, *need_band, *tone_need, *fall_list, *fall_count, *fall_m; double *fall_best; unsigned char *kind; double2 *tpl;
    i64 snr_stride;
    unsigned *pass_hist;          // [16] fine_core8_kernel: bursts proven after p passes ([0] = left to the band kernel)
    double2 *wcache;              // [D][cap][B8_WLEN] or nullptr (set by attach_wcache)
    bool wc_valid;                // the fine search of this call filled the cache
};
size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr int kMaxGroups = 32;     // stream groups per batch (tier-3 scratch is per group); ceil(148*8/64) = 19 <= 24 bands

int make_work(DevBuf &wb, i64 D, int cap, i64 snr_len, int tpl_len, Work *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
    size_t o_ctl = take(sizeof(StreamCtl) * D), o_res = take(sizeof(StreamResultDev) * D);
    size_t per = sizeof(double) * D * cap;
    size_t o_cp = take(per), o_cs = take(per), o_fr = take(per), o_fp = take(per), o_fo = take(per), o_g = take(per), o_sr = take(per), o_sp = take(per), o_pp = take(per);
    size_t o_pi = take(per * 12), o_snr = take(sizeof(double) * D * snr_len), o_pw = take(sizeof(double) * D);
    size_t o_se = take(sizeof(int) * D * cap), o_nf = take(sizeof(int) * D * cap), o_nb = take(sizeof(int) * D * cap), o_tn = take(sizeof(int) * D * cap), o_fl = take(sizeof(int) * D * cap), o_fc = take(sizeof(int) * D), o_fm = take(sizeof(int) * (size_t)FALL_GRID * 24 * kMaxGroups), o_fb = take(sizeof(double) * (size_t)FALL_GRID * 24 * kMaxGroups), o_k = take(D * cap), o_tpl = take(sizeof(double2) * (tpl_len > 0 ? tpl_len : 1)), o_ph = take(sizeof(unsigned) * 16 + sizeof(unsigned long long) * 64);
    void *base;
    TRY(wb.get(off, &base));
    char *b = static_cast<char *>(base);
    w->ctl = (StreamCtl *)(b + o_ctl); w->res = (StreamResultDev *)(b + o_res);
    w->coarse_pos = (double *)(b + o_cp); w->coarse_snr = (double *)(b + o_cs); w->fine_raw = (double *)(b + o_fr); w->fcch_pos = (double *)(b + o_fp);
    w->fo = (double *)(b + o_fo); w->gate = (double *)(b + o_g); w->sch_raw = (double *)(b + o_sr); w->sch_pos = (double *)(b + o_sp); w->post_pos = (double *)(b + o_pp);
    w->pos_info = (double *)(b + o_pi); w->snr_map = (double *)(b + o_snr); w->power = (double *)(b + o_pw);
    w->sch_edge = (int *)(b + o_se); w->need_full = (int *)(b + o_nf); w->need_band = (int *)(b + o_nb); w->tone_need = (int *)(b + o_tn); w->fall_list = (int *)(b + o_fl); w->fall_count = (int *)(b + o_fc); w->fall_m = (int *)(b + o_fm); w->fall_best = (double *)(b + o_fb); w->kind = (unsigned char *)(b + o_k); w->tpl = (double2 *)(b + o_tpl);
    w->snr_stride = snr_len;
    w->pass_hist = (unsigned *)(b + o_ph);
    w->wcache = nullptr; w->wc_valid = false;
    return GSMCAL_OK;
}

// filtered-window cache of the osr-8 fast path: 36,432 bytes per (stream, burst); skipped (the stages re-filter) when it would not fit
constexpr size_t kMaxWcacheBytes = (size_t)24 << 30;
int attach_wcache(DevBuf &buf, i64 D, int cap, int osr, Work *w) {
    w->wcache = nullptr;
    if (osr != 8) return GSMCAL_OK;
    const size_t bytes = (size_t)D * cap * B8_WLEN * sizeof(double2);
    if (bytes > kMaxWcacheBytes) return GSMCAL_OK;
    void *p;
    if (buf.get(bytes, &p) != GSMCAL_OK) { cudaGetLastError(); return GSMCAL_OK; }   // no memory for it: run without
    w->wcache = (double2 *)p;
    return GSMCAL_OK;
}

Work sub_work(const Work &w, i64 d0, int cap, int group) {
    Work s = w;
    s.ctl += d0; s.res += d0; s.power += d0;
    s.coarse_pos += d0 * cap; s.coarse_snr += d0 * cap; s.fine_raw += d0 * cap; s.fcch_pos += d0 * cap; s.fo += d0 * cap; s.gate += d0 * cap;
    s.sch_raw += d0 * cap; s.sch_pos += d0 * cap; s.post_pos += d0 * cap; s.pos_info += d0 * cap * 12; s.snr_map += d0 * w.snr_stride;
    s.sch_edge += d0 * cap; s.need_full += d0 * cap; s.need_band += d0 * cap; s.tone_need += d0 * cap; s.kind += d0 * cap;
    s.fall_list += d0 * cap; s.fall_count += d0;
    s.fall_m += (size_t)group * FALL_GRID * 24; s.fall_best += (size_t)group * FALL_GRID * 24;
    if (s.wcache) s.wcache += (size_t)d0 * cap * B8_WLEN;
    return s;
}

WinSrc mat_src(const double2 *base, i64 len, i64 stride, int interp) {
    WinSrc s; memset(&s, 0, sizeof s);
    s.lazy = 0; s.base = base; s.base_len = len; s.base_stride = stride; s.mat_interp = interp; s.dec = 1;
    return s;
}
WinSrc lazy_src(const uint8_t *raw, i64 n_iq, int n_taps, int level, int dec) {
    WinSrc s; memset(&s, 0, sizeof s);
    s.lazy = 1; s.raw = raw; s.n_iq = n_iq; s.n_taps = n_taps; s.level = level; s.dec = dec;
    return s;
}
WinSrc with_cache(WinSrc s, const Work &w, int cap) {            // level-0 samples of cached bursts come from the fine search's window cache
    if (w.wcache && w.wc_valid) { s.wcache = w.wcache; s.wc_pos = w.coarse_pos; s.wc_cap = cap; }
    return s;
}

// ---- stage launchers (device buffers) -------------------------------------------------------------
int run_colsum_u8(const uint8_t *raw, i64 n_iq, i64 D, StreamCtl *ctl, cudaStream_t st) {
    if (((uintptr_t)raw & 1) != 0) return fail(GSMCAL_ERR_ARG, "uint8 capture must start on an even address");
    const i64 chunk = 512 * 1024;
    i64 chunks = (2 * n_iq + chunk - 1) / chunk; if (chunks < 1) chunks = 1;
    if (D > 65535) return fail(GSMCAL_ERR_ARG, "more than 65535 streams per call");
    LAUNCH(colsum_u8_kernel, dim3((unsigned)chunks, (unsigned)D), 256, 0, st, raw, n_iq, chunk, ctl);
    return GSMCAL_OK;
}

int run_colsum_u8_persist(const uint8_t *raw, i64 n_iq, i64 D, StreamCtl *ctl, int blocks, cudaStream_t st) {
    if (((uintptr_t)raw & 1) != 0) return fail(GSMCAL_ERR_ARG, "uint8 capture must start on an even address");
    const i64 chunk = 512 * 1024;
    i64 chunks = (2 * n_iq + chunk - 1) / chunk; if (chunks < 1) chunks = 1;
    LAUNCH(colsum_u8_persist_kernel, (unsigned)blocks, g_debug_persist_threads, 0, st, raw, n_iq, chunk, chunks, D, ctl);
    return GSMCAL_OK;
}

int run_colsum_u8_trickle(const uint8_t *raw, i64 n_iq, i64 D, StreamCtl *ctl, int blocks, int n_stage, cudaStream_t st) {
    if (((uintptr_t)raw & 1) != 0) return fail(GSMCAL_ERR_ARG, "uint8 capture must start on an even address");
    const i64 chunk = 512 * 1024;
    i64 chunks = (2 * n_iq + chunk - 1) / chunk; if (chunks < 1) chunks = 1;
    LAUNCH(colsum_u8_trickle_kernel, (unsigned)blocks, 32, (size_t)n_stage * (TRK_STAGE + 8), st, raw, n_iq, chunk, chunks, D, ctl, n_stage);
    return GSMCAL_OK;
}

struct CoarseParams { int fft_len, mv_len, step10, step11, dr; i64 n_first; double th; };
double host_mround(double x) { return x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5); }
int coarse_params(int dr, CoarseParams *p) {
    if (dr < 1 || dr > 148) return fail(GSMCAL_ERR_ARG, "decimation_ratio out of range");
    const double num_sym_per_frame = (625.0 / 4.0) * 8.0;
    p->fft_len = 1 << (int)floor(log2(148.0 / dr));            // FCCH_coarse_position.m:17
    p->mv_len = 10 * p->fft_len;                                // :22
    p->th = 10.0;                                               // :21
    p->n_first = (i64)ceil(23.0 * num_sym_per_frame / dr);      // :25
    p->step10 = (int)host_mround(10.0 * num_sym_per_frame / dr);   // :35
    p->step11 = (int)host_mround(11.0 * num_sym_per_frame / dr);   // :36
    p->dr = dr;
    return GSMCAL_OK;
}

int run_coarse(WinSrc src, i64 len, const CoarseParams &p, i64 D, int cap, Work &w, cudaStream_t st) {
    const i64 n_win = p.n_first - (p.fft_len - 1);
    size_t smem = sizeof(double2) * (p.fft_len + SNR_THREADS + p.fft_len);
    LAUNCH(snr_map_kernel, dim3((unsigned)((n_win + SNR_THREADS - 1) / SNR_THREADS), (unsigned)D), SNR_THREADS, smem, st,
           src, w.ctl, (i64)0, n_win, p.fft_len, w.snr_map, w.snr_stride);
    LAUNCH(first_hit_scan_kernel, (unsigned)((D + 3) / 4), 128, 0, st, w.snr_map, w.snr_stride, n_win, p.mv_len, p.th, w.ctl, (int)D);
    const size_t chain_smem = sizeof(double2) * (size_t)(p.fft_len + 2 * (2 * 5 + p.fft_len));    // twiddles + the two candidate groups
    LAUNCH(coarse_chain_kernel, (unsigned)D, CHAIN_THREADS, chain_smem, st, src, w.ctl, len, p.fft_len, p.th, p.step10, p.step11, p.dr, cap, w.coarse_pos, w.coarse_snr,
           g_debug_prof ? (unsigned long long *)(w.pass_hist + 16) + 48 : nullptr);
    return GSMCAL_OK;
}

constexpr int kNeedGroup = 8;      // bursts per block of the fallback kernels that run behind a need-mask (tone_est_kernel, tier 2)
int run_fine_peak(Ctx &c, WinSrc src_peak, i64 n_iq, int osr, i64 D, int cap, Work &w, cudaStream_t st) {
    const double2 *tw; TRY(get_twiddle(c, 148 * osr, st, &tw));
    if (g_debug_force_full || (osr % 4) != 0) {      // the band kernel's 16-sample certificate grid needs osr % 4 == 0
        LAUNCH(fine_peak_full_kernel, dim3((unsigned)cap, (unsigned)D), kFineThreads, fine_smem(osr), st, src_peak, w.ctl, w.coarse_pos, cap, osr, n_iq, tw, w.fine_raw, (const int *)nullptr, (const int *)nullptr, 0);
        return GSMCAL_OK;
    }
    CU(cudaMemsetAsync(w.need_full, 0, sizeof(int) * D * cap, st));
    CU(cudaMemsetAsync(w.need_band, 0, sizeof(int) * D * cap, st));
    w.wc_valid = false;
    if (src_peak.lazy && osr == 8 && src_peak.n_taps <= 64 && !g_debug_no_core8) {
        // osr-8 fast path: FIR once per burst, filtered window cached for tier 2 and the tone stages
        const dim3 grid((unsigned)cap, (unsigned)D);
#define CORE8_ARGS src_peak.raw, n_iq, w.ctl, w.coarse_pos, cap, tw, w.fine_raw, w.need_band, g_debug_fail_tier2, w.wcache, g_debug_core8_passes, w.pass_hist, \
                   (unsigned long long *)(w.pass_hist + 16)
        if (src_peak.n_taps == 47 && g_debug_prof) LAUNCH((fine_core8_kernel<47, true>), grid, B8_THREADS, B8_SMEM, st, CORE8_ARGS);
        else if (src_peak.n_taps == 47) LAUNCH((fine_core8_kernel<47, false>), grid, B8_THREADS, B8_SMEM, st, CORE8_ARGS);
        else if (src_peak.n_taps <= 48) LAUNCH((fine_core8_kernel<48, false>), grid, B8_THREADS, B8_SMEM, st, CORE8_ARGS);
        else                            LAUNCH((fine_core8_kernel<64, false>), grid, B8_THREADS, B8_SMEM, st, CORE8_ARGS);
#undef CORE8_ARGS
        w.wc_valid = (w.wcache != nullptr);
        src_peak = with_cache(src_peak, w, cap);
    } else
    LAUNCH(fine_peak_core_kernel, dim3((unsigned)cap, (unsigned)D), FC_THREADS, fine_band_smem(osr), st, src_peak, w.ctl, w.coarse_pos, cap, osr, n_iq, tw, w.fine_raw, w.need_band, g_debug_fail_tier2);
    CU(cudaMemsetAsync(w.fall_count, 0, sizeof(int), st));
    LAUNCH(fine_peak_band_kernel, dim3((unsigned)((cap + kNeedGroup - 1) / kNeedGroup), (unsigned)D), FB_THREADS, fine_band_smem(osr), st, src_peak, w.ctl, w.coarse_pos, cap, osr, n_iq, tw, w.fine_raw,
           (const int *)w.need_band, w.need_full, 0, w.fall_list, w.fall_count, w.fall_best, w.fall_m, g_debug_fail_tier2, kNeedGroup);
    const int nb3 = (148 * osr + FB_BINS - 1) / FB_BINS;
    LAUNCH(fine_peak_band_kernel, dim3((unsigned)nb3, (unsigned)FALL_GRID), FB_THREADS, fine_band_smem(osr), st, src_peak, w.ctl, w.coarse_pos, cap, osr, n_iq, tw, w.fine_raw,
           (const int *)nullptr, (int *)nullptr, 1, w.fall_list, w.fall_count, w.fall_best, w.fall_m, g_debug_fall_limit, 1);
    LAUNCH(fine_fall_combine_kernel, FALL_GRID / 128, 128, 0, st, w.fall_list, w.fall_count, w.fall_best, w.fall_m, nb3, w.coarse_pos, cap, osr, w.fine_raw, w.ctl, g_debug_fall_limit);
    LAUNCH(fine_peak_full_kernel, 296, kFineThreads, fine_smem(osr), st, src_peak, w.ctl, w.coarse_pos, cap, osr, n_iq, tw, w.fine_raw,
           (const int *)w.fall_list, (const int *)w.fall_count, g_debug_fall_limit);
    return GSMCAL_OK;
}
// per-burst tone estimate: with the filtered-window cache (osr 8) tone8_kernel does every burst it can certify and flags the others
// for the generic row-FFT kernel; without a cache the generic kernel does them all
int run_tone(WinSrc src, int which, const double *pos, int osr, i64 D, int cap, const double2 *tw, Work &w, cudaStream_t st) {
    const int *need = nullptr;
    if (src.lazy && src.wcache && osr == 8 && !g_debug_no_tone8) {
        CU(cudaMemsetAsync(w.tone_need, 0, sizeof(int) * D * cap, st));
        unsigned long long *prof = (unsigned long long *)(w.pass_hist + 16) + 16 * which;      // [1]: fine stage, [2]: post stage (debug key 13)
        if (g_debug_prof) LAUNCH(tone8_kernel<true>, dim3((unsigned)cap, (unsigned)D), T8_THREADS, T8_SMEM, st, src, w.ctl, which, pos, cap, tw, w.fo, w.gate, w.tone_need, prof);
        else              LAUNCH(tone8_kernel<false>, dim3((unsigned)cap, (unsigned)D), T8_THREADS, T8_SMEM, st, src, w.ctl, which, pos, cap, tw, w.fo, w.gate, w.tone_need, prof);
        need = w.tone_need;
    }
    const int gsz = need ? kNeedGroup : 1;                       // behind tone8_kernel's mask: groups of bursts per block (few or none are flagged)
    LAUNCH(tone_est_kernel, dim3((unsigned)((cap + gsz - 1) / gsz), (unsigned)D), TONE_THREADS, tone_smem(osr), st, src, w.ctl, which, pos, cap, osr, tw, w.fo, w.gate, need, gsz);
    return GSMCAL_OK;
}

int run_fine_rest(Ctx &c, WinSrc src_tone, i64 n_iq, int osr, double carrier_freq, i64 D, int cap, Work &w, cudaStream_t st) {
    const double2 *tw; TRY(get_twiddle(c, 148 * osr, st, &tw));
    LAUNCH(fine_ppm_kernel, (unsigned)D, PS_THREADS, (size_t)cap * 17 + 16, st, w.ctl, (int)D, cap, osr, n_iq, w.fine_raw, w.fcch_pos, w.kind);
    TRY(run_tone(src_tone, 1, w.fcch_pos, osr, D, cap, tw, w, st));
    LAUNCH(fine_carrier_kernel, (unsigned)D, PS_THREADS, (size_t)cap * 16, st, w.ctl, (int)D, cap, osr, carrier_freq, w.fo, w.gate);
    return GSMCAL_OK;
}

int run_fine(Ctx &c, WinSrc src_peak, WinSrc src_tone, i64 n_iq, int osr, double carrier_freq, i64 D, int cap, Work &w, cudaStream_t st) {
    TRY(run_fine_peak(c, src_peak, n_iq, osr, D, cap, w, st));
    return run_fine_rest(c, src_tone, n_iq, osr, carrier_freq, D, cap, w, st);
}

int run_sch(WinSrc src, int osr, i64 D, int cap, Work &w, cudaStream_t st) {
    if (g_debug_sch56) LAUNCH(sch_corr_kernel<56>, dim3((unsigned)cap, (unsigned)D), SCH_THREADS, sch_smem(osr), st, src, w.ctl, w.fcch_pos, cap, osr, w.tpl, w.sch_raw, w.sch_edge);
    else               LAUNCH(sch_corr_kernel<64>, dim3((unsigned)cap, (unsigned)D), SCH_THREADS, sch_smem(osr), st, src, w.ctl, w.fcch_pos, cap, osr, w.tpl, w.sch_raw, w.sch_edge);
    LAUNCH(sch_ppm_kernel, (unsigned)D, PS_THREADS, (size_t)cap * 21 + 16, st, w.ctl, (int)D, cap, osr, w.sch_raw, w.sch_edge, w.sch_pos, w.kind, w.pos_info, w.post_pos);
    return GSMCAL_OK;
}

int run_post(Ctx &c, WinSrc src, int osr, double carrier_freq, i64 D, int cap, Work &w, bool want_res, cudaStream_t st) {
    const double2 *tw; TRY(get_twiddle(c, 148 * osr, st, &tw));
    TRY(run_tone(src, 2, w.post_pos, osr, D, cap, tw, w, st));
    LAUNCH(post_carrier_kernel, (unsigned)D, PS_THREADS, (size_t)cap * 8, st, w.ctl, (int)D, cap, osr, carrier_freq, w.fo, want_res ? w.res : nullptr);
    return GSMCAL_OK;
}

template <bool U8>
int run_fir(const void *in, i64 n, i64 in_stride, const StreamCtl *ctl, int n_taps, int decim, i64 D, double2 *out, i64 n_out, double *power, cudaStream_t st) {
    if (decim < 1) return fail(GSMCAL_ERR_ARG, "decim must be >= 1");
    if (n <= 0) return GSMCAL_OK;
    if (decim == 1 && !power && !U8 && n_taps <= 64 && in_stride == n && n_out == n && (((uintptr_t)in & 15) == 0)) {
        // complex128 in/out: persistent TMA-pipelined kernel
        const i64 n_tiles = ((n + FT_TILE - 1) / FT_TILE) * D;
        i64 gx = 148 * 4; if (gx > n_tiles) gx = n_tiles;        // 4 resident blocks per SM (122 registers, 38 KB)
        if (n_taps <= 48) LAUNCH((fir_full_tma_kernel<48>), (unsigned)gx, FT_THREADS, sizeof(double2) * 2 * (FT_TILE + 47), st, (const double2 *)in, n, D, out);
        else              LAUNCH((fir_full_tma_kernel<64>), (unsigned)gx, FT_THREADS, sizeof(double2) * 2 * (FT_TILE + 63), st, (const double2 *)in, n, D, out);
        return GSMCAL_OK;
    }
    if (decim == 1 && !power) {
        unsigned gx = (unsigned)((n + FIR_TILE - 1) / FIR_TILE);
        auto smem = [](int NT) { int n_in = FIR_TILE + NT - 1; return sizeof(double2) * (size_t)(n_in + n_in / 8 + 2); };
        if (n_taps <= 32)      LAUNCH((fir_full_kernel<32, U8>), dim3(gx, (unsigned)D), FIR_THREADS, smem(32), st, in, n, in_stride, ctl, out, n_out);
        else if (n_taps <= 48) LAUNCH((fir_full_kernel<48, U8>), dim3(gx, (unsigned)D), FIR_THREADS, smem(48), st, in, n, in_stride, ctl, out, n_out);
        else if (n_taps <= 64) LAUNCH((fir_full_kernel<64, U8>), dim3(gx, (unsigned)D), FIR_THREADS, smem(64), st, in, n, in_stride, ctl, out, n_out);
        else                   LAUNCH((fir_full_kernel<128, U8>), dim3(gx, (unsigned)D), FIR_THREADS, smem(128), st, in, n, in_stride, ctl, out, n_out);
        return GSMCAL_OK;
    }
    if (U8 && decim >= n_taps && (decim % 4) == 0 && n_taps <= 64) {
        unsigned gx = (unsigned)((n_out + FDD_THREADS - 1) / FDD_THREADS);
        LAUNCH(fir_decim_direct_u8_kernel<17>, dim3(gx, (unsigned)D), FDD_THREADS, 0, st, (const uint8_t *)in, n, ctl, n_taps, decim, out, n_out, power);
        return GSMCAL_OK;
    }
    int opb = 2048 / decim; if (opb < 1) opb = 1; if (opb > 1024) opb = 1024;
    int n_in = (opb - 1) * decim + n_taps;
    size_t smem = sizeof(double2) * (size_t)(n_in + n_in / 8 + 2);
    unsigned gx = (unsigned)((n_out + opb - 1) / opb);
    LAUNCH(fir_decim_kernel<U8>, dim3(gx, (unsigned)D), FIRD_THREADS, smem, st, in, n, in_stride, ctl, n_taps, decim, opb, out, n_out, n_out, power);
    return GSMCAL_OK;
}

int run_resample_derotate(const double2 *in, i64 len_in, double e, int do_interp, double dphi, int do_derot, double2 *out, i64 len_out, cudaStream_t st) {
    if (len_out <= 0) return GSMCAL_OK;
    unsigned g = (unsigned)((len_out + 256 * DEROT_K - 1) / (256 * DEROT_K));
    LAUNCH(resample_derotate_kernel, g, 256, 0, st, in, len_in, e, do_interp, dphi, do_derot, out, len_out);
    return GSMCAL_OK;
}

int check_fir_args(const double *coef, int n_taps, const void *s, i64 n, i64 n_col, const void *r) {
    if (!coef || !s || !r || n < 0 || n_col < 1) return fail(GSMCAL_ERR_ARG, "null pointer or negative size");
    if (n_taps < 1 || n_taps > GSMCAL_MAX_TAPS) return fail(GSMCAL_ERR_ARG, "n_taps must be 1..%d", GSMCAL_MAX_TAPS);
    return GSMCAL_OK;
}

void gmsk_template(int osr, double *out);

}  // namespace

// ====================================================================================================
extern "C" {

int gsmcal_abi_version(void) { return GSMCAL_ABI_VERSION; }
const char *gsmcal_last_error(void) { return g_err.c_str(); }
int gsmcal_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int gsmcal_set_device(int device) {
    int n = gsmcal_device_count();
    if (device < 0 || device >= n) return fail(GSMCAL_ERR_CUDA, "device %d not available (%d devices)", device, n);
    g_device = device;
    return GSMCAL_OK;
}
void gsmcal_release(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_last_need_full = nullptr; g_last_need_band = nullptr; g_last_need_full_n = 0; g_last_pass_hist = nullptr; g_prev_pass_hist = nullptr;   // they point into the workspaces freed below
    for (auto &kv : g_ctx) {
        cudaSetDevice(kv.first);
        kv.second.ring.release(); kv.second.in.release(); kv.second.out.release(); kv.second.work.release(); kv.second.tplbuf.release(); kv.second.wc.release();
        for (auto &t : kv.second.tw) cudaFree(t.second);
        kv.second.tw.clear();
        if (kv.second.stage_ev_ok) { for (cudaEvent_t e : kv.second.stage_ev) cudaEventDestroy(e); kv.second.stage_ev_ok = false; }
        for (cudaStream_t s2 : kv.second.side) cudaStreamDestroy(s2);
        kv.second.side.clear();
        for (cudaStream_t s2 : kv.second.side_hi) cudaStreamDestroy(s2);
        kv.second.side_hi.clear();
        for (Slot &sl : kv.second.slots) {
            if (sl.done) { cudaEventSynchronize(sl.done); cudaEventDestroy(sl.done); sl.done = nullptr; }
            for (cudaStream_t s2 : sl.grp) cudaStreamDestroy(s2);
            for (cudaStream_t s2 : sl.hi) cudaStreamDestroy(s2);
            for (cudaStream_t s2 : sl.co) cudaStreamDestroy(s2);
            sl.co.clear();
            if (sl.front) cudaStreamDestroy(sl.front);
            if (sl.front_hi) cudaStreamDestroy(sl.front_hi);
            sl.front_hi = nullptr;
            sl.grp.clear(); sl.hi.clear(); sl.front = nullptr;
            if (sl.stage) cudaFreeHost(sl.stage);
            sl.stage = nullptr; sl.stage_cap = 0; sl.busy = false;
            sl.work.release(); sl.wc.release();
        }
    }
}
int64_t gsmcal_debug_get(int key) {
    // key 1: bursts of the last fine search (this device) that needed the all-bin fallback
    std::lock_guard<std::mutex> lk(g_mu);
    if (key == 40) return (int64_t)g_debug_core8_passes;
    if (key >= 150 && key < 214) {                               // the same counters of the batch submitted before the last one (other slot)
        if (!g_prev_pass_hist) return 0;
        unsigned long long v = 0;
        if (cudaMemcpy(&v, (unsigned long long *)(g_prev_pass_hist + 16) + (key - 150), sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return (int64_t)v;
    }
    if (key >= 50 && key < 114) {                                 // per-phase cycle sums (debug key 13), [15] = blocks: 50.. fine_core8, 66.. tone8 (fine), 82.. tone8 (post)
        if (!g_last_pass_hist) return 0;
        unsigned long long v = 0;
        if (cudaMemcpy(&v, (unsigned long long *)(g_last_pass_hist + 16) + (key - 50), sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return (int64_t)v;
    }
    if (key == 30) return (int64_t)g_staged_bytes.load();           // bytes that went through the pinned staging ring (pageable host buffers)
    if (key >= 10 && key < 26) {                                 // 10 + p: bursts the osr-8 tier-1 kernel proved after p passes (p = 0: left open)
        if (!g_last_pass_hist) return 0;
        unsigned v = 0;
        if (cudaMemcpy(&v, g_last_pass_hist + (key - 10), sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return (int64_t)v;
    }
    if (key == 1 || key == 2) {
        const int *src = (key == 1) ? g_last_need_full : g_last_need_band;
        if (!src || g_last_need_full_n <= 0) return 0;
        std::vector<int> h((size_t)g_last_need_full_n);
        if (cudaMemcpy(h.data(), src, sizeof(int) * h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        int64_t n = 0;
        for (int v : h) n += (v != 0);
        return n;
    }
    return -1;
}
int gsmcal_debug_set(int key, int value) {
    if (key == 0) { g_debug_force_full = value; return GSMCAL_OK; }
    if (key == 4) { g_debug_fail_tier2 = value; return GSMCAL_OK; }
    if (key == 5) { g_debug_fall_limit = value < 0 ? 0 : (value > FALL_GRID ? FALL_GRID : value); return GSMCAL_OK; }
    if (key == 6) { g_debug_hi_prio = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 9) { g_debug_no_core8 = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 11) { g_debug_no_tone8 = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 12) { g_debug_no_staging = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 13) { g_debug_prof = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 14) { g_debug_timeline = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 16) { g_debug_gate = value ? 1 : 0; return GSMCAL_OK; }
    if (key == 23) { g_debug_occupy = value < 0 ? 0 : (value > 96 ? 96 : value); return GSMCAL_OK; }
    if (key == 21) { g_debug_chain_serial = value != 0; return GSMCAL_OK; }
    if (key == 22) { g_debug_chain_lo = value != 0; return GSMCAL_OK; }
    if (key == 19) { g_debug_sch56 = value != 0; return GSMCAL_OK; }
    if (key == 17) { g_debug_trickle = value < 0 ? 0 : (value > 24 ? 24 : value); return GSMCAL_OK; }
    if (key == 18) { g_debug_trickle_blocks = value < 1 ? 1 : (value > 4 ? 4 : value); return GSMCAL_OK; }
    if (key == 15) { g_debug_persist_threads = (value == 64 || value == 128) ? value : 256; return GSMCAL_OK; }
    if (key == 10) { g_debug_core8_passes = value < 1 ? 1 : (value > 8 ? 8 : value); return GSMCAL_OK; }
    if (key == 8) { g_debug_submit_groups = value < 1 ? 1 : (value > kMaxGroups / 2 ? kMaxGroups / 2 : value); return GSMCAL_OK; }
    if (key == 7) { g_debug_persist_colsum = value < 0 ? 0 : (value > 8 ? 8 : value); return GSMCAL_OK; }
    if (key == 3) { g_debug_groups = value < 1 ? 1 : (value > kMaxGroups / 2 ? kMaxGroups / 2 : value); return GSMCAL_OK; }
    return fail(GSMCAL_ERR_ARG, "debug_set: unknown key");
}
int64_t gsmcal_launch_count(int reset) { long long v = g_launches.load(); if (reset) g_launches = 0; return v; }

int64_t gsmcal_max_bursts(int64_t len_decimated, int decimation_ratio) {
    const double num_sym_per_frame = (625.0 / 4.0) * 8.0;
    return (int64_t)ceil((double)len_decimated / (10.0 * num_sym_per_frame / decimation_ratio)) + 2;
}

// ---------------------------------------------------------------------------------------------------
int gsmcal_raw2iq_u8(const uint8_t *a, int64_t n_iq, int64_t n_col, double *b) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!a || !b || n_iq < 0 || n_col < 1) return fail(GSMCAL_ERR_ARG, "raw2iq: bad arguments");
    if (n_iq == 0) return GSMCAL_OK;
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    void *din, *dout; Work w;
    TRY(c->in.get((size_t)2 * n_iq * n_col, &din));
    TRY(c->out.get(sizeof(double2) * (size_t)n_iq * n_col, &dout));
    TRY(make_work(c->work, n_col, 1, 1, 0, &w));
    TRY(copy_h2d(c->ring, g_device, din, a, (size_t)2 * n_iq * n_col, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
    TRY(run_colsum_u8((const uint8_t *)din, n_iq, n_col, w.ctl, st));
    i64 gx = (n_iq + 256 * 8 - 1) / (256 * 8); if (gx > 148 * 16) gx = 148 * 16; if (gx < 1) gx = 1;
    LAUNCH(raw2iq_store_kernel, dim3((unsigned)gx, (unsigned)n_col), 256, 0, st, (const uint8_t *)din, n_iq, w.ctl, (double2 *)dout);
    TRY(copy_d2h(c->ring, g_device, b, dout, sizeof(double2) * (size_t)n_iq * n_col, st));
    return GSMCAL_OK;
}

int gsmcal_raw2iq_f64(const double *a, int64_t n_iq, int64_t n_col, double *b) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!a || !b || n_iq < 0 || n_col < 1) return fail(GSMCAL_ERR_ARG, "raw2iq: bad arguments");
    if (n_iq == 0) return GSMCAL_OK;
    for (int64_t i = 0; i < 2 * n_iq * n_col; ++i)
        if (!(a[i] >= 0.0 && a[i] <= 255.0 && a[i] == floor(a[i])))
            return fail(GSMCAL_ERR_ARG, "raw2iq: double input must hold integers 0..255 (what fread(...,'uint8') returns)");
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    void *din, *dout; Work w;
    TRY(c->in.get(sizeof(double) * (size_t)2 * n_iq * n_col, &din));
    TRY(c->out.get(sizeof(double2) * (size_t)n_iq * n_col, &dout));
    TRY(make_work(c->work, n_col, 1, 1, 0, &w));
    TRY(copy_h2d(c->ring, g_device, din, a, sizeof(double) * (size_t)2 * n_iq * n_col, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
    i64 gx = (n_iq + 256 * 8 - 1) / (256 * 8); if (gx > 148 * 16) gx = 148 * 16; if (gx < 1) gx = 1;
    LAUNCH(colsum_f64_kernel, dim3((unsigned)gx, (unsigned)n_col), 256, 0, st, (const double *)din, n_iq, w.ctl);
    LAUNCH(raw2iq_store_f64_kernel, dim3((unsigned)gx, (unsigned)n_col), 256, 0, st, (const double *)din, n_iq, w.ctl, (double2 *)dout);
    TRY(copy_d2h(c->ring, g_device, b, dout, sizeof(double2) * (size_t)n_iq * n_col, st));
    return GSMCAL_OK;
}

// ---------------------------------------------------------------------------------------------------
int gsmcal_fir1(int order, double wn, double *coef) {
    // MATLAB fir1(n, Wn): low-pass, Hamming window, scaled so the DC gain is exactly 1 (gsm_sync_demod.m:34)
    if (order < 1 || order + 1 > GSMCAL_MAX_TAPS || !(wn > 0.0 && wn < 1.0) || !coef) return fail(GSMCAL_ERR_ARG, "fir1: bad arguments");
    const int n = order + 1;
    const double alpha = 0.5 * order;
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const double m = i - alpha;
        const double x = wn * m;
        const double sinc = (x == 0.0) ? 1.0 : sin(M_PI * x) / (M_PI * x);
        const double win = 0.54 - 0.46 * cos(2.0 * M_PI * i / order);
        coef[i] = wn * sinc * win;
        sum += coef[i];
    }
    for (int i = 0; i < n; ++i) coef[i] /= sum;
    return GSMCAL_OK;
}

int gsmcal_fir_filter(const double *coef, int n_taps, const double *s, int64_t n, int64_t n_col, int decim, double *r) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(check_fir_args(coef, n_taps, s, n, n_col, r));
    if (n == 0) return GSMCAL_OK;
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    const i64 n_out = (n + decim - 1) / decim;
    void *din, *dout;
    TRY(c->in.get(sizeof(double2) * (size_t)n * n_col, &din));
    TRY(c->out.get(sizeof(double2) * (size_t)n_out * n_col, &dout));
    TRY(set_taps(coef, n_taps, st));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)n * n_col, st));
    TRY(run_fir<false>(din, n, n, nullptr, n_taps, decim, n_col, (double2 *)dout, n_out, nullptr, st));
    TRY(copy_d2h(c->ring, g_device, r, dout, sizeof(double2) * (size_t)n_out * n_col, st));
    return GSMCAL_OK;
}

int gsmcal_raw2iq_fir_u8(const uint8_t *a, int64_t n_iq, int64_t n_col, const double *coef, int n_taps, int decim, double *r) {
    std::lock_guard<std::mutex> lk(g_mu);
    TRY(check_fir_args(coef, n_taps, a, n_iq, n_col, r));
    if (n_iq == 0) return GSMCAL_OK;
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    const i64 n_out = (n_iq + decim - 1) / decim;
    void *din, *dout; Work w;
    TRY(c->in.get((size_t)2 * n_iq * n_col, &din));
    TRY(c->out.get(sizeof(double2) * (size_t)n_out * n_col, &dout));
    TRY(make_work(c->work, n_col, 1, 1, 0, &w));
    TRY(set_taps(coef, n_taps, st));
    TRY(copy_h2d(c->ring, g_device, din, a, (size_t)2 * n_iq * n_col, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
    TRY(run_colsum_u8((const uint8_t *)din, n_iq, n_col, w.ctl, st));
    TRY(run_fir<true>(din, n_iq, 2 * n_iq, w.ctl, n_taps, decim, n_col, (double2 *)dout, n_out, nullptr, st));
    TRY(copy_d2h(c->ring, g_device, r, dout, sizeof(double2) * (size_t)n_out * n_col, st));
    return GSMCAL_OK;
}

int gsmcal_chn_filter_taps(int which, double *coef, int *n_taps) {
    if (!coef || !n_taps) return fail(GSMCAL_ERR_ARG, "chn_filter_taps: null");
    if (which == 8) { memcpy(coef, kNum8x, sizeof kNum8x); *n_taps = 60; return GSMCAL_OK; }
    if (which == 4) { memcpy(coef, kNum4x, sizeof kNum4x); *n_taps = 30; return GSMCAL_OK; }
    return fail(GSMCAL_ERR_ARG, "chn_filter_taps: which must be 8 or 4");
}
int gsmcal_chn_filter_8x_4x(const double *s, int64_t n, int64_t n_col, double *r) { return gsmcal_fir_filter(kNum8x, 60, s, n, n_col, 2, r); }
int gsmcal_chn_filter_4x(const double *s, int64_t n, int64_t n_col, double *r) { return gsmcal_fir_filter(kNum4x, 30, s, n, n_col, 1, r); }

int gsmcal_band_power_u8(const uint8_t *a, int64_t n_iq, int64_t n_col, const double *coef, int n_taps, int decim, double *power) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!a || !power || n_iq < 1 || n_col < 1 || decim < 1) return fail(GSMCAL_ERR_ARG, "band_power: bad arguments");
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    void *din; Work w;
    TRY(c->in.get((size_t)2 * n_iq * n_col, &din));
    TRY(make_work(c->work, n_col, 1, 1, 0, &w));
    if (coef) TRY(set_taps(coef, n_taps, st));
    TRY(copy_h2d(c->ring, g_device, din, a, (size_t)2 * n_iq * n_col, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
    CU(cudaMemsetAsync(w.power, 0, sizeof(double) * n_col, st));
    TRY(run_colsum_u8((const uint8_t *)din, n_iq, n_col, w.ctl, st));
    const i64 n_out = (n_iq + decim - 1) / decim;
    if (coef) {
        TRY(run_fir<true>(din, n_iq, 2 * n_iq, w.ctl, n_taps, decim, n_col, nullptr, n_out, w.power, st));
    } else {
        i64 gx = (n_out + 2047) / 2048; if (gx > 1024) gx = 1024;
        LAUNCH(power_u8_kernel, dim3((unsigned)gx, (unsigned)n_col), 256, 0, st, (const uint8_t *)din, n_iq, w.ctl, decim, w.power);
    }
    CU(cudaMemcpyAsync(power, w.power, sizeof(double) * n_col, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_col; ++i) power[i] /= (double)n_out;
    return GSMCAL_OK;
}

int gsmcal_diversity_power_u8(const uint8_t *s_all, int64_t n_iq, int64_t n_freq, int64_t n_dongle, const double *coef, int n_taps, int decim,
                              double *power_spectrum, double *power_spectrum_combine) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!s_all || !coef || !power_spectrum || !power_spectrum_combine || n_iq < 1 || n_freq < 1 || n_dongle < 1 || decim < 1)
        return fail(GSMCAL_ERR_ARG, "diversity_power: bad arguments");
    const i64 n_col = n_freq * n_dongle;
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0;
    void *din; Work w;
    TRY(c->in.get((size_t)2 * n_iq * n_col, &din));
    TRY(make_work(c->work, n_col, 3, 1, 0, &w));                  // cap 3: room for power | per-dongle spectrum | combination in the [D][cap] arrays
    TRY(set_taps(coef, n_taps, st));
    TRY(copy_h2d(c->ring, g_device, din, s_all, (size_t)2 * n_iq * n_col, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
    CU(cudaMemsetAsync(w.power, 0, sizeof(double) * n_col, st));
    TRY(run_colsum_u8((const uint8_t *)din, n_iq, n_col, w.ctl, st));
    const i64 n_out = (n_iq + decim - 1) / decim;
    TRY(run_fir<true>(din, n_iq, 2 * n_iq, w.ctl, n_taps, decim, n_col, nullptr, n_out, w.power, st));
    LAUNCH(diversity_combine_kernel, (unsigned)((n_freq + 127) / 128), 128, 0, st, w.power, n_out, (int)n_freq, (int)n_dongle, w.fo, w.gate);
    CU(cudaMemcpyAsync(power_spectrum, w.fo, sizeof(double) * n_col, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(power_spectrum_combine, w.gate, sizeof(double) * n_freq, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GSMCAL_OK;
}

// ---------------------------------------------------------------------------------------------------
static int snr_windows(const double *s, int64_t len, int fft_len, i64 w0, i64 n_win, Ctx **cout, Work *w, cudaStream_t st) {
    if (!s || len < 1 || fft_len < 1 || fft_len > 128) return fail(GSMCAL_ERR_ARG, "moving fft: bad arguments (fft_len <= 128)");
    Ctx *c; TRY(get_ctx(&c));
    void *din;
    TRY(c->in.get(sizeof(double2) * (size_t)len, &din));
    TRY(make_work(c->work, 1, 1, n_win > 0 ? n_win : 1, 0, w));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)len, st));
    CU(cudaMemsetAsync(w->ctl, 0, sizeof(StreamCtl), st));
    if (n_win > 0) {
        WinSrc src = mat_src((const double2 *)din, len, len, 0);
        size_t smem = sizeof(double2) * (fft_len + SNR_THREADS + fft_len);
        LAUNCH(snr_map_kernel, dim3((unsigned)((n_win + SNR_THREADS - 1) / SNR_THREADS), 1), SNR_THREADS, smem, st, src, w->ctl, w0, n_win, fft_len, w->snr_map, w->snr_stride);
    }
    *cout = c;
    return GSMCAL_OK;
}

int gsmcal_move_fft_snr_runtime_avg(const double *s, int64_t len, int mv_len, int fft_len, double th,
                                    int *hit_flag, double *hit_idx, double *hit_avg_snr, double *hit_snr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!hit_flag || !hit_idx || !hit_avg_snr || !hit_snr || mv_len < 1) return fail(GSMCAL_ERR_ARG, "move_fft_snr_runtime_avg: bad arguments");
    cudaStream_t st = 0; Ctx *c; Work w;
    const i64 n_win = len - (fft_len - 1);
    TRY(snr_windows(s, len, fft_len, 0, n_win, &c, &w, st));
    StreamCtl h; memset(&h, 0, sizeof h); h.first_hit = -1;
    if (n_win > 0) {
        LAUNCH(first_hit_scan_kernel, 1, 32, 0, st, w.snr_map, w.snr_stride, n_win, mv_len, th, w.ctl, 1);
        CU(cudaMemcpyAsync(&h, w.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (h.first_hit > 0) { *hit_flag = 1; *hit_idx = h.first_hit; *hit_avg_snr = h.hit_avg_snr; *hit_snr = h.hit_snr; }
    else { *hit_flag = 0; *hit_idx = -1; *hit_avg_snr = INFINITY; *hit_snr = INFINITY; }
    return GSMCAL_OK;
}

int gsmcal_move_fft_snr_trace(const double *s, int64_t len, int fft_len, double *snr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!snr) return fail(GSMCAL_ERR_ARG, "snr trace: null output");
    cudaStream_t st = 0; Ctx *c; Work w;
    const i64 n_win = len - (fft_len - 1);
    if (n_win < 1) return fail(GSMCAL_ERR_ARG, "snr trace: stream shorter than fft_len");
    TRY(snr_windows(s, len, fft_len, 0, n_win, &c, &w, st));
    CU(cudaMemcpyAsync(snr, w.snr_map, sizeof(double) * n_win, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return GSMCAL_OK;
}

int gsmcal_specific_fft_snr_fix_avg(const double *s, int64_t len, int64_t t0, int64_t t1, int fft_len, double th, double avg_snr,
                                    int *hit_flag, double *hit_idx, double *hit_snr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!hit_flag || !hit_idx || !hit_snr) return fail(GSMCAL_ERR_ARG, "specific_fft_snr_fix_avg: null output");
    *hit_flag = 0; *hit_idx = -1; *hit_snr = INFINITY;
    if (t1 < t0) return GSMCAL_OK;
    if (t0 < 1 || t1 + fft_len - 1 > len) return fail(GSMCAL_ERR_RANGE, "specific_fft_snr_fix_avg: window outside the stream (MATLAB would raise an index error)");
    cudaStream_t st = 0; Ctx *c; Work w;
    const i64 n_win = t1 - t0 + 1;
    TRY(snr_windows(s, len, fft_len, t0 - 1, n_win, &c, &w, st));
    LAUNCH(specific_hit_kernel, 1, 32, 0, st, w.snr_map, n_win, th, avg_snr, w.sch_edge, w.power);
    int off; double v;
    CU(cudaMemcpyAsync(&off, w.sch_edge, sizeof off, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&v, w.power, sizeof v, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (off >= 0) { *hit_flag = 1; *hit_idx = (double)(t0 + off); *hit_snr = v; }
    return GSMCAL_OK;
}

int gsmcal_FCCH_coarse_position(const double *s, int64_t len, int decimation_ratio, double *position, double *snr, int64_t cap_out, int64_t *n_out) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!s || !position || !snr || !n_out || len < 1) return fail(GSMCAL_ERR_ARG, "FCCH_coarse_position: bad arguments");
    CoarseParams p; TRY(coarse_params(decimation_ratio, &p));
    if (p.n_first > len) return fail(GSMCAL_ERR_RANGE, "FCCH_coarse_position: stream shorter than 23 frames (s(1:%lld) would raise an index error)", (long long)p.n_first);
    const int cap = (int)gsmcal_max_bursts(len, decimation_ratio);
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0; void *din; Work w;
    TRY(c->in.get(sizeof(double2) * (size_t)len, &din));
    TRY(make_work(c->work, 1, cap, p.n_first, 0, &w));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)len, st));
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl), st));
    TRY(run_coarse(mat_src((const double2 *)din, len, len, 0), len, p, 1, cap, w, st));
    StreamCtl h;
    CU(cudaMemcpyAsync(&h, w.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h.n_coarse < 0) { *n_out = -1; return GSMCAL_OK; }
    if (h.n_coarse > cap_out) return fail(GSMCAL_ERR_CAPACITY, "FCCH_coarse_position: %d hits, capacity %lld", h.n_coarse, (long long)cap_out);
    CU(cudaMemcpy(position, w.coarse_pos, sizeof(double) * h.n_coarse, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(snr, w.coarse_snr, sizeof(double) * h.n_coarse, cudaMemcpyDeviceToHost));
    *n_out = h.n_coarse;
    return GSMCAL_OK;
}

// ---------------------------------------------------------------------------------------------------
int gsmcal_FCCH_fine_correction(const double *s, int64_t n, const double *base_position, int64_t n_base, int osr, double carrier_freq,
                                double *FCCH_pos, int64_t pos_cap, int64_t *n_pos, double *r, int64_t r_cap, int64_t *r_len,
                                double *sampling_ppm, double *carrier_ppm) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!s || !n_pos || !r_len || !sampling_ppm || !carrier_ppm || n < 1 || n_base < 0 || osr < 1 || osr > 8) return fail(GSMCAL_ERR_ARG, "FCCH_fine_correction: bad arguments (osr 1..8)");
    *n_pos = -1; *r_len = -1; *sampling_ppm = INFINITY; *carrier_ppm = INFINITY;
    if (n_base < 5) return GSMCAL_OK;                                   // FCCH_fine_correction.m:12-15
    if (!base_position || !FCCH_pos) return fail(GSMCAL_ERR_ARG, "FCCH_fine_correction: null positions");
    const i64 len_s = n / osr;
    for (int64_t i = 0; i < n_base; ++i) {
        const double p = base_position[i];
        if (p + 64 > (double)(len_s - 148 + 1)) break;
        if (p != floor(p) || p - 64 < 1) return fail(GSMCAL_ERR_RANGE, "FCCH_fine_correction: base_position(%lld)=%g puts the search window before the first sample", (long long)i + 1, p);
    }
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0; void *din; Work w;
    const int cap = (int)n_base;
    TRY(c->in.get(sizeof(double2) * (size_t)n, &din));
    TRY(make_work(c->work, 1, cap, 1, 0, &w));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)n, st));
    StreamCtl h; memset(&h, 0, sizeof h); h.n_coarse = cap;
    CU(cudaMemcpyAsync(w.ctl, &h, sizeof h, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w.coarse_pos, base_position, sizeof(double) * cap, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    TRY(run_fine(*c, mat_src((const double2 *)din, n, n, 0), mat_src((const double2 *)din, n, n, 1), n, osr, carrier_freq, 1, cap, w, st));
    CU(cudaMemcpyAsync(&h, w.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *sampling_ppm = h.sppm1; *carrier_ppm = h.cppm1; *n_pos = h.n_fcch;
    if (h.n_fcch > 0) {
        if (h.n_fcch > pos_cap) return fail(GSMCAL_ERR_CAPACITY, "FCCH_fine_correction: FCCH_pos capacity");
        CU(cudaMemcpy(FCCH_pos, w.fcch_pos, sizeof(double) * h.n_fcch, cudaMemcpyDeviceToHost));
    }
    *r_len = h.len1;
    if (h.len1 >= 0) {
        if (!r || h.len1 > r_cap) return fail(GSMCAL_ERR_CAPACITY, "FCCH_fine_correction: r capacity %lld < %lld", (long long)r_cap, (long long)h.len1);
        if (!h.interp1_on && !h.derot1_on) memcpy(r, s, sizeof(double2) * (size_t)h.len1);     // r = s (:72)
        else {
            void *dout; TRY(c->out.get(sizeof(double2) * (size_t)h.len1, &dout));
            TRY(run_resample_derotate((const double2 *)din, n, h.e1, h.interp1_on, h.dphi1, h.derot1_on, (double2 *)dout, h.len1, st));
            TRY(copy_d2h(c->ring, g_device, r, dout, sizeof(double2) * (size_t)h.len1, st));
        }
    }
    return GSMCAL_OK;
}

int gsmcal_SCH_training_sequence_gen(int osr, double *s) {
    if (!s || osr < 1 || osr > 64) return fail(GSMCAL_ERR_ARG, "gsm_SCH_training_sequence_gen: bad arguments");
    gmsk_template(osr, s);
    return GSMCAL_OK;
}

int gsmcal_SCH_corr_rate_correction(const double *s, int64_t n, const double *FCCH_pos, int64_t n_fcch, const double *tpl, int osr,
                                    double *pos_info, int64_t rows_cap, int64_t *n_rows, double *r, int64_t r_cap, int64_t *r_len, double *sampling_ppm) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!n_rows || !r_len || !sampling_ppm || osr < 1 || osr > 8) return fail(GSMCAL_ERR_ARG, "SCH_corr_rate_correction: bad arguments (osr 1..8)");
    *n_rows = -1; *r_len = -1; *sampling_ppm = INFINITY;
    if (n_fcch < 5) return GSMCAL_OK;                                    // SCH_corr_rate_correction.m:11-14
    if (!s || !FCCH_pos || !tpl || !pos_info || n < 1) return fail(GSMCAL_ERR_ARG, "SCH_corr_rate_correction: null input");
    const int cap = (int)n_fcch;
    const int L = 64 * osr;
    for (int i = 0; i < cap; ++i) if (FCCH_pos[i] != floor(FCCH_pos[i]) || FCCH_pos[i] + (625 * osr / 4) * 8 + 42 * osr - 8 * osr < 1)
        return fail(GSMCAL_ERR_RANGE, "SCH_corr_rate_correction: FCCH_pos(%d) puts the search window before the first sample", i + 1);
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0; void *din; Work w;
    TRY(c->in.get(sizeof(double2) * (size_t)n, &din));
    TRY(make_work(c->work, 1, cap, 1, L, &w));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)n, st));
    CU(cudaMemcpyAsync(w.tpl, tpl, sizeof(double2) * L, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(w.fcch_pos, FCCH_pos, sizeof(double) * cap, cudaMemcpyHostToDevice, st));
    StreamCtl h; memset(&h, 0, sizeof h); h.n_fcch = cap; h.sch_enable = 1; h.len1 = n;
    CU(cudaMemcpyAsync(w.ctl, &h, sizeof h, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    TRY(run_sch(mat_src((const double2 *)din, n, n, 0), osr, 1, cap, w, st));
    CU(cudaMemcpyAsync(&h, w.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *sampling_ppm = h.sppm2; *n_rows = h.n_pos_info;
    if (h.n_pos_info > 0) {
        if (h.n_pos_info > rows_cap) return fail(GSMCAL_ERR_CAPACITY, "SCH_corr_rate_correction: pos_info capacity %lld < %d", (long long)rows_cap, h.n_pos_info);
        std::vector<double> tmp(2 * (size_t)h.n_pos_info);
        CU(cudaMemcpy(tmp.data(), w.pos_info, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < h.n_pos_info; ++i) { pos_info[i] = tmp[2 * i]; pos_info[h.n_pos_info + i] = tmp[2 * i + 1]; }
    }
    *r_len = h.len2;
    if (h.len2 >= 0) {
        if (!r || h.len2 > r_cap) return fail(GSMCAL_ERR_CAPACITY, "SCH_corr_rate_correction: r capacity");
        if (!h.interp2_on) memcpy(r, s, sizeof(double2) * (size_t)h.len2);                    // r = s (:87,:120)
        else {
            void *dout; TRY(c->out.get(sizeof(double2) * (size_t)h.len2, &dout));
            TRY(run_resample_derotate((const double2 *)din, n, h.e2, 1, 0.0, 0, (double2 *)dout, h.len2, st));
            TRY(copy_d2h(c->ring, g_device, r, dout, sizeof(double2) * (size_t)h.len2, st));
        }
    }
    return GSMCAL_OK;
}

int gsmcal_carrier_correct_post_SCH(const double *s, int64_t n, const double *pos_info, int64_t n_rows, int osr, double carrier_freq,
                                    double *r, int64_t r_cap, int64_t *r_len, double *carrier_ppm) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!r_len || !carrier_ppm || n_rows < 0 || osr < 1 || osr > 8) return fail(GSMCAL_ERR_ARG, "carrier_correct_post_SCH: bad arguments");
    *r_len = -1; *carrier_ppm = INFINITY;
    if (n_rows > 0 && !pos_info) return fail(GSMCAL_ERR_ARG, "carrier_correct_post_SCH: null pos_info");
    bool all_m1 = n_rows > 0;                                            // `if pos_info==-1` is true only when ALL elements are -1 (:10)
    for (int64_t i = 0; i < 2 * n_rows; ++i) if (pos_info[i] != -1.0) { all_m1 = false; break; }
    if (all_m1) return GSMCAL_OK;
    int n_b = 0; std::vector<double> fpos;
    for (int64_t i = 0; i < n_rows; ++i) {
        if (pos_info[n_rows + i] == 2.0) ++n_b;
        if (pos_info[n_rows + i] == 0.0) fpos.push_back(pos_info[i]);
    }
    if (n_b < 4) return GSMCAL_OK;                                       // :15-19
    if (!s || n < 1) return fail(GSMCAL_ERR_ARG, "carrier_correct_post_SCH: null stream");
    const int N = 148 * osr;
    for (double p : fpos) if (p != floor(p) || p < 1 || p + N - 1 > (double)n) return fail(GSMCAL_ERR_RANGE, "carrier_correct_post_SCH: FCCH window outside the stream");
    if (fpos.empty()) return fail(GSMCAL_ERR_RANGE, "carrier_correct_post_SCH: no FCCH rows (mean of empty is NaN in the reference)");
    if (!r || n > r_cap) return fail(GSMCAL_ERR_CAPACITY, "carrier_correct_post_SCH: r capacity");
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = 0; void *din, *dout; Work w;
    const int cap = (int)fpos.size();
    TRY(c->in.get(sizeof(double2) * (size_t)n, &din));
    TRY(c->out.get(sizeof(double2) * (size_t)n, &dout));
    TRY(make_work(c->work, 1, cap, 1, 0, &w));
    TRY(copy_h2d(c->ring, g_device, din, s, sizeof(double2) * (size_t)n, st));
    CU(cudaMemcpyAsync(w.post_pos, fpos.data(), sizeof(double) * cap, cudaMemcpyHostToDevice, st));
    StreamCtl h; memset(&h, 0, sizeof h); h.post_enable = 1; h.n_post_fcch = cap; h.len2 = n;
    h.sppm1 = h.sppm2 = h.cppm1 = INFINITY;
    CU(cudaMemcpyAsync(w.ctl, &h, sizeof h, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    TRY(run_post(*c, mat_src((const double2 *)din, n, n, 0), osr, carrier_freq, 1, cap, w, false, st));
    CU(cudaMemcpyAsync(&h, w.ctl, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *carrier_ppm = h.cppm2;
    TRY(run_resample_derotate((const double2 *)din, n, 0.0, 0, h.dphi2, 1, (double2 *)dout, n, st));
    TRY(copy_d2h(c->ring, g_device, r, dout, sizeof(double2) * (size_t)n, st));
    *r_len = n;
    return GSMCAL_OK;
}

int gsmcal_total_ppm_calculation(const double *ppm_in, int64_t n, double *ppm_out) {
    // total_ppm_calculation.m:5-21 - pure scalar host arithmetic (there is nothing to put on a GPU)
    if (!ppm_out || n < 0 || (n > 0 && !ppm_in)) return fail(GSMCAL_ERR_ARG, "total_ppm_calculation: bad arguments");
    bool all_inf = n > 0;
    for (int64_t i = 0; i < n; ++i) if (!(std::isinf(ppm_in[i]) && ppm_in[i] > 0)) { all_inf = false; break; }
    if (all_inf) { *ppm_out = INFINITY; return GSMCAL_OK; }
    double acc = 1.0;
    for (int64_t i = 0; i < n; ++i) acc = acc * (1.0 + ppm_in[i] * 1e-6);
    *ppm_out = (acc - 1.0) * 1e6;
    return GSMCAL_OK;
}

// ---------------------------------------------------------------------------------------------------
static int calibrate_batch_impl(const uint8_t *raw, int raw_mem, int64_t n_iq, int64_t D, double carrier_freq, const double *tpl,
                                const double *coef, int n_taps, int osr, int coarse_dr, gsmcal_stream_result *results,
                                double *coarse_pos, double *coarse_snr, double *fcch_pos, double *pos_info, void *cuda_stream,
                                double *r_out, int r_mem, int64_t r_stride) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (r_out && r_stride < n_iq) return fail(GSMCAL_ERR_ARG, "calibrate_batch_r: r_stride must be >= n_iq");
    if (!raw || !tpl || !coef || !results || n_iq < 1 || D < 1 || osr < 1 || osr > 8) return fail(GSMCAL_ERR_ARG, "calibrate_batch: bad arguments (osr 1..8)");
    CoarseParams p; TRY(coarse_params(coarse_dr, &p));
    const int dec = osr * coarse_dr;
    const i64 len_dec = (n_iq + dec - 1) / dec;
    if (p.n_first > len_dec) return fail(GSMCAL_ERR_RANGE, "calibrate_batch: capture shorter than 23 frames");
    const int cap = (int)gsmcal_max_bursts(len_dec, coarse_dr);
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Work w;
    TRY(make_work(c->work, D, cap, p.n_first, 64 * osr, &w));
    TRY(attach_wcache(c->wc, D, cap, osr, &w));
    TRY(set_taps(coef, n_taps, st));
    const uint8_t *draw = raw;
    if (raw_mem == GSMCAL_MEM_HOST) {
        void *din; TRY(c->in.get((size_t)2 * n_iq * D, &din));
        draw = (const uint8_t *)din;
    }
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * D, st));
    CU(cudaMemsetAsync(w.need_full, 0, sizeof(int) * D * cap, st));
    CU(cudaMemsetAsync(w.need_band, 0, sizeof(int) * D * cap, st));
    CU(cudaMemcpyAsync(w.tpl, tpl, sizeof(double2) * 64 * osr, cudaMemcpyHostToDevice, st));
    { const double2 *twp; TRY(get_twiddle(*c, 148 * osr, st, &twp)); }        // built on `st` before the groups fork
    g_last_need_full = w.need_full; g_last_need_band = w.need_band; g_last_need_full_n = (long long)D * cap;
    g_last_pass_hist = w.pass_hist;
    CU(cudaMemsetAsync(w.pass_hist, 0, sizeof(unsigned) * 16 + sizeof(unsigned long long) * 64, st));
    g_stage_n = 0;
    // Streams are independent, so the batch is cut into groups that run the stage sequence on their own CUDA
    // streams: the latency-bound stages of one group (burst chain, per-stream solves) overlap the FP64-bound
    // stages of the others, and with host input the H2D copy of group g+1 overlaps the compute of group g.
    int n_groups = g_debug_groups;
    if (raw_mem == GSMCAL_MEM_HOST && n_groups > 1) n_groups *= 2;
    if (D < 2 * n_groups) n_groups = 1;
    const bool timing = (n_groups == 1);
    while ((int)c->side.size() < n_groups + 1) { cudaStream_t s2; CU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); c->side.push_back(s2); }
    if (g_debug_hi_prio) {
        int lo_p = 0, hi_p = 0; CU(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
        while ((int)c->side_hi.size() < n_groups) { cudaStream_t s2; CU(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi_p)); c->side_hi.push_back(s2); }
    }
    cudaStream_t cp = c->side[n_groups];                                         // H2D copies
    cudaEvent_t ev_fork; CU(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    CU(cudaEventRecord(ev_fork, st));
    if (raw_mem == GSMCAL_MEM_HOST) CU(cudaStreamWaitEvent(cp, ev_fork, 0));
    std::vector<cudaEvent_t> ev_done;
    const size_t per = (size_t)2 * n_iq;
    if (timing) TRY(stage_mark(*c, st));
    for (int g = 0; g < n_groups; ++g) {
        const i64 d0 = D * g / n_groups, d1 = D * (g + 1) / n_groups, nd = d1 - d0;
        cudaStream_t sg = (n_groups == 1) ? st : c->side[g];
        if (sg != st) CU(cudaStreamWaitEvent(sg, ev_fork, 0));
        Work ws = sub_work(w, d0, cap, g);
        const uint8_t *graw = draw + d0 * per;
        if (raw_mem == GSMCAL_MEM_HOST) {
            TRY(copy_h2d(c->ring, g_device, (void *)graw, raw + d0 * per, per * nd, cp));
            cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CU(cudaEventRecord(ev, cp)); CU(cudaStreamWaitEvent(sg, ev, 0)); CU(cudaEventDestroy(ev));
            TRY(run_colsum_u8(graw, n_iq, nd, ws.ctl, sg));
        } else if (sg == st) {
            TRY(run_colsum_u8(graw, n_iq, nd, ws.ctl, sg));
        } else if (g_debug_persist_colsum > 0 && g_debug_hi_prio && g > 0) {
            // device input, groups after the first: the column sums as ONE persistent launch of small fixed footprint on the group's
            // high-priority stream, so they run BESIDE the FP64 stages of the previous group (which the gate below keeps ahead)
            cudaStream_t sh = c->side_hi[g];
            int n_sm = 148; CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g_device));
            CU(cudaStreamWaitEvent(sh, ev_fork, 0));
            TRY(run_colsum_u8_persist(graw, n_iq, nd, ws.ctl, n_sm * g_debug_persist_colsum, sh));
            cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CU(cudaEventRecord(ev, sh)); CU(cudaStreamWaitEvent(sg, ev, 0)); CU(cudaEventDestroy(ev));
        } else {
            // device input: the HBM-bound column sums run on the caller's stream (group 0 gets the whole bandwidth first)
            TRY(run_colsum_u8(graw, n_iq, nd, ws.ctl, st));
            cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CU(cudaEventRecord(ev, st)); CU(cudaStreamWaitEvent(sg, ev, 0)); CU(cudaEventDestroy(ev));
        }
        LAUNCH(mean_kernel, (unsigned)((nd + 127) / 128), 128, 0, sg, ws.ctl, (int)nd, n_iq);
        if (timing) TRY(stage_mark(*c, sg));
        if (sg != st && g_debug_hi_prio) {
            // the coarse stage is a handful of small latency-bound kernels (a 0.8 ms dependent burst chain): on the group's own stream
            // their blocks queue behind the thousands of pending blocks of the other groups' FP64 kernels.  A high-priority stream
            // lets the block scheduler place them as soon as a slot frees, so the chain runs underneath the heavy kernels.
            cudaStream_t sh = c->side_hi[g];
            cudaEvent_t e1, e2; CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
            CU(cudaEventRecord(e1, sg)); CU(cudaStreamWaitEvent(sh, e1, 0));
            TRY(run_coarse(lazy_src(graw, n_iq, n_taps, 0, dec), len_dec, p, nd, cap, ws, sh));
            CU(cudaEventRecord(e2, sh)); CU(cudaStreamWaitEvent(sg, e2, 0));
            CU(cudaEventDestroy(e1)); CU(cudaEventDestroy(e2));
        } else
        TRY(run_coarse(lazy_src(graw, n_iq, n_taps, 0, dec), len_dec, p, nd, cap, ws, sg));
        if (timing) TRY(stage_mark(*c, sg));
        // stagger the groups (see gsmcal_calibrate_batch_submit): the FP64 stages of group g wait for group g-1, its front does not
        if (g_debug_gate && sg != st && !ev_done.empty()) CU(cudaStreamWaitEvent(sg, ev_done.back(), 0));
        TRY(run_fine_peak(*c, lazy_src(graw, n_iq, n_taps, 0, 1), n_iq, osr, nd, cap, ws, sg));
        if (timing) TRY(stage_mark(*c, sg));
        TRY(run_fine_rest(*c, with_cache(lazy_src(graw, n_iq, n_taps, 1, 1), ws, cap), n_iq, osr, carrier_freq, nd, cap, ws, sg));
        if (timing) TRY(stage_mark(*c, sg));
        TRY(run_sch(lazy_src(graw, n_iq, n_taps, 2, 1), osr, nd, cap, ws, sg));
        if (timing) TRY(stage_mark(*c, sg));
        TRY(run_post(*c, with_cache(lazy_src(graw, n_iq, n_taps, 3, 1), ws, cap), osr, carrier_freq, nd, cap, ws, true, sg));
        if (timing) TRY(stage_mark(*c, sg));
        if (sg != st) {
            cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            CU(cudaEventRecord(ev, sg));
            ev_done.push_back(ev);
        }
    }
    for (cudaEvent_t ev : ev_done) { CU(cudaStreamWaitEvent(st, ev, 0)); CU(cudaEventDestroy(ev)); }   // join (after ALL groups are enqueued)
    CU(cudaEventDestroy(ev_fork));
    if (r_out) {
        // r_correct (gsm_sync_demod.m:120 -> :145) for every stream whose chain completed: one fused pass from the uint8 capture
        double2 *r_dev = (double2 *)r_out;
        if (r_mem == GSMCAL_MEM_HOST) { void *p; TRY(c->out.get(sizeof(double2) * (size_t)r_stride * D, &p)); r_dev = (double2 *)p; }
        const i64 tiles = (n_iq + MAT_T - 1) / MAT_T;
        i64 gx = (148 * 3 + D - 1) / D; if (gx < 1) gx = 1; if (gx > tiles) gx = tiles;
        const size_t smem = sizeof(double2) * (size_t)(MAT_T + GSMCAL_XCAP(MAT_T + 8) + MAT_T + 16);
        LAUNCH(materialise_r_kernel, dim3((unsigned)gx, (unsigned)D), MAT_THREADS, smem, st, lazy_src(draw, n_iq, n_taps, 3, 1), w.ctl, r_dev, (i64)r_stride, tiles);
        if (r_mem == GSMCAL_MEM_HOST) {                          // only the r_len[2] samples of every completed stream go back (nothing else is written)
            CU(cudaMemcpyAsync(results, w.res, sizeof(StreamResultDev) * D, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            for (i64 d = 0; d < D; ++d)
                if (results[d].r_len[2] > 0)
                    TRY(copy_d2h(c->ring, g_device, r_out + 2 * (size_t)r_stride * d, r_dev + (size_t)r_stride * d, sizeof(double2) * (size_t)results[d].r_len[2], st));
        }
    }
    CU(cudaMemcpyAsync(results, w.res, sizeof(StreamResultDev) * D, cudaMemcpyDeviceToHost, st));
    if (coarse_pos) CU(cudaMemcpyAsync(coarse_pos, w.coarse_pos, sizeof(double) * D * cap, cudaMemcpyDeviceToHost, st));
    if (coarse_snr) CU(cudaMemcpyAsync(coarse_snr, w.coarse_snr, sizeof(double) * D * cap, cudaMemcpyDeviceToHost, st));
    if (fcch_pos) CU(cudaMemcpyAsync(fcch_pos, w.fcch_pos, sizeof(double) * D * cap, cudaMemcpyDeviceToHost, st));
    if (pos_info) CU(cudaMemcpyAsync(pos_info, w.pos_info, sizeof(double) * D * cap * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (timing) for (int i = 0; i + 1 < g_stage_n; ++i) CU(cudaEventElapsedTime(&g_stage_ms[i], c->stage_ev[i], c->stage_ev[i + 1]));
    if (!timing) g_stage_n = 0;
    return GSMCAL_OK;
}

int gsmcal_calibrate_batch(const uint8_t *raw, int raw_mem, int64_t n_iq, int64_t D, double carrier_freq, const double *tpl,
                           const double *coef, int n_taps, int osr, int coarse_dr, gsmcal_stream_result *results,
                           double *coarse_pos, double *coarse_snr, double *fcch_pos, double *pos_info, void *cuda_stream) {
    return calibrate_batch_impl(raw, raw_mem, n_iq, D, carrier_freq, tpl, coef, n_taps, osr, coarse_dr, results, coarse_pos, coarse_snr, fcch_pos, pos_info,
                                cuda_stream, nullptr, GSMCAL_MEM_DEVICE, 0);
}

int gsmcal_calibrate_batch_r(const uint8_t *raw, int raw_mem, int64_t n_iq, int64_t D, double carrier_freq, const double *tpl,
                             const double *coef, int n_taps, int osr, int coarse_dr, gsmcal_stream_result *results,
                             double *coarse_pos, double *coarse_snr, double *fcch_pos, double *pos_info, void *cuda_stream,
                             double *r_correct, int r_mem, int64_t r_stride) {
    if (!r_correct) return fail(GSMCAL_ERR_ARG, "calibrate_batch_r: null r_correct");
    return calibrate_batch_impl(raw, raw_mem, n_iq, D, carrier_freq, tpl, coef, n_taps, osr, coarse_dr, results, coarse_pos, coarse_snr, fcch_pos, pos_info,
                                cuda_stream, r_correct, r_mem, r_stride);
}

// ---- submit / collect: the same pipeline, several batches in flight -----------------------------------------------------------
int gsmcal_calibrate_batch_submit(int slot, const uint8_t *raw_dev, int64_t n_iq, int64_t D, double carrier_freq, const double *tpl,
                                  const double *coef, int n_taps, int osr, int coarse_dr, gsmcal_stream_result *results,
                                  double *coarse_pos, double *coarse_snr, double *fcch_pos, double *pos_info, void *cuda_stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (slot < 0 || slot >= kSlots) return fail(GSMCAL_ERR_ARG, "calibrate_batch_submit: slot must be 0..%d", kSlots - 1);
    if (!raw_dev || !tpl || !coef || !results || n_iq < 1 || D < 1 || osr < 1 || osr > 8) return fail(GSMCAL_ERR_ARG, "calibrate_batch_submit: bad arguments (osr 1..8)");
    CoarseParams p; TRY(coarse_params(coarse_dr, &p));
    const int dec = osr * coarse_dr;
    const i64 len_dec = (n_iq + dec - 1) / dec;
    if (p.n_first > len_dec) return fail(GSMCAL_ERR_RANGE, "calibrate_batch_submit: capture shorter than 23 frames");
    const int cap = (int)gsmcal_max_bursts(len_dec, coarse_dr);
    Ctx *c; TRY(get_ctx(&c));
    Slot &sl = c->slots[slot];
    if (sl.busy) return fail(GSMCAL_ERR_ARG, "calibrate_batch_submit: slot %d still holds an uncollected batch", slot);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int n_groups = g_debug_submit_groups;                       // 2: in the staggered pipeline a batch's stages run alone, so the per-stream kernels and
                                                                // the grid tails of one half overlap the burst kernels of the other (31.15 -> 30.85 ms, profiles/r2x)
    if (D < 2 * n_groups) n_groups = 1;
    if (!sl.front) { CU(cudaStreamCreateWithFlags(&sl.front, cudaStreamNonBlocking)); CU(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming)); }
    int lo_p = 0, hi_p = 0; CU(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    if (!sl.front_hi) CU(cudaStreamCreateWithPriority(&sl.front_hi, cudaStreamNonBlocking, hi_p));
    while ((int)sl.grp.size() < n_groups) { cudaStream_t s2; CU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); sl.grp.push_back(s2); }
    while ((int)sl.hi.size() < n_groups) { cudaStream_t s2; CU(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, hi_p)); sl.hi.push_back(s2); }
    Work w;
    TRY(make_work(sl.work, D, cap, p.n_first, 64 * osr, &w));
    TRY(attach_wcache(sl.wc, D, cap, osr, &w));
    TRY(set_taps(coef, n_taps, st));
    // pinned staging: records | coarse_pos | coarse_snr | fcch_pos | pos_info (x12)
    sl.n_res = sizeof(StreamResultDev) * (size_t)D; sl.n_per = sizeof(double) * (size_t)D * cap;
    const size_t need = align_up(sl.n_res) + 15 * sl.n_per + 64 * osr * sizeof(double2);
    if (need > sl.stage_cap) {
        if (sl.stage) cudaFreeHost(sl.stage);
        sl.stage = nullptr; sl.stage_cap = 0;
        CU(cudaMallocHost((void **)&sl.stage, need));
        sl.stage_cap = need;
    }
    char *h_res = sl.stage, *h_arr = sl.stage + align_up(sl.n_res), *h_tpl = h_arr + 15 * sl.n_per;
    memcpy(h_tpl, tpl, sizeof(double2) * 64 * osr);              // the caller's template may be pageable and short-lived
    { const double2 *twp; TRY(get_twiddle(*c, 148 * osr, st, &twp)); }
    // normal-priority kernels of different streams are dispatched in arrival order: behind the 10^5 queued blocks of the previous batch's fine
    // search even a memset waits milliseconds (device timeline, debug key 14), so the staggered mode runs the whole front at high priority
    cudaStream_t fr = (g_debug_trickle > 0 || g_debug_gate) ? sl.front_hi : sl.front;
    cudaEvent_t ev_in; CU(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming));
    CU(cudaEventRecord(ev_in, st)); CU(cudaStreamWaitEvent(fr, ev_in, 0)); CU(cudaEventDestroy(ev_in));     // the capture is ready on the caller's stream
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * D, fr));
    CU(cudaMemsetAsync(w.need_full, 0, sizeof(int) * D * cap, fr));
    CU(cudaMemsetAsync(w.need_band, 0, sizeof(int) * D * cap, fr));
    CU(cudaMemcpyAsync(w.tpl, h_tpl, sizeof(double2) * 64 * osr, cudaMemcpyHostToDevice, fr));
    g_last_need_full = w.need_full; g_last_need_band = w.need_band; g_last_need_full_n = (long long)D * cap;
    g_prev_pass_hist = g_last_pass_hist; g_last_pass_hist = w.pass_hist;
    CU(cudaMemsetAsync(w.pass_hist, 0, sizeof(unsigned) * 16 + sizeof(unsigned long long) * 64, fr));
    const size_t per = (size_t)2 * n_iq;
    std::vector<cudaEvent_t> ev_done;
    cudaEvent_t e_sum = nullptr;
    sl.tl_on = g_debug_timeline != 0;
    if (sl.tl_on) {
        if (!sl.tl[0]) for (auto &e : sl.tl) CU(cudaEventCreate(&e));
        if (!g_tl_base) { CU(cudaEventCreate(&g_tl_base)); CU(cudaEventRecord(g_tl_base, fr)); }
        CU(cudaEventRecord(sl.tl[0], fr));
    }
    std::vector<cudaEvent_t> e_sum_g;                            // per stream group: its column sums are done
    if (g_debug_persist_colsum > 0 || g_debug_trickle > 0) {
        // the column sums as persistent launches of fixed footprint on the high-priority front stream (see colsum_u8_persist_kernel,
        // colsum_u8_trickle_kernel), one per stream group, back to back: the burst chain of group g starts under the sums of group g+1
        int n_sm = 148; CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g_device));
        cudaEvent_t e_in; CU(cudaEventCreateWithFlags(&e_in, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e_sum, cudaEventDisableTiming));
        CU(cudaEventRecord(e_in, fr)); CU(cudaStreamWaitEvent(sl.front_hi, e_in, 0)); CU(cudaEventDestroy(e_in));
        for (int g = 0; g < n_groups; ++g) {
            const i64 d0 = D * g / n_groups, d1 = D * (g + 1) / n_groups;
            if (g_debug_persist_colsum > 0) TRY(run_colsum_u8_persist(raw_dev + d0 * per, n_iq, d1 - d0, w.ctl + d0, n_sm * g_debug_persist_colsum, sl.front_hi));
            else TRY(run_colsum_u8_trickle(raw_dev + d0 * per, n_iq, d1 - d0, w.ctl + d0, n_sm * g_debug_trickle_blocks, g_debug_trickle, sl.front_hi));
            cudaEvent_t eg; CU(cudaEventCreateWithFlags(&eg, cudaEventDisableTiming));
            CU(cudaEventRecord(eg, sl.front_hi));
            e_sum_g.push_back(eg);
        }
        CU(cudaEventRecord(e_sum, sl.front_hi));
        if (sl.tl_on) CU(cudaEventRecord(sl.tl[1], sl.front_hi));
    }
    if (g_debug_chain_lo) while ((int)sl.co.size() < n_groups) { cudaStream_t s2; CU(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking)); sl.co.push_back(s2); }
    cudaEvent_t e_chain_prev = nullptr;
    for (int g = 0; g < n_groups; ++g) {
        const i64 d0 = D * g / n_groups, d1 = D * (g + 1) / n_groups, nd = d1 - d0;
        cudaStream_t sg = sl.grp[g], sh = g_debug_chain_lo ? sl.co[g] : sl.hi[g];
        Work ws = sub_work(w, d0, cap, g);
        const uint8_t *graw = raw_dev + d0 * per;
        cudaEvent_t e0, e1; CU(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
        if (e_sum) CU(cudaStreamWaitEvent(sh, e_sum_g[g], 0));
        else {
            TRY(run_colsum_u8(graw, n_iq, nd, ws.ctl, fr));     // HBM-bound sums back to back on the front stream (group 0 first)
            CU(cudaEventRecord(e0, fr)); CU(cudaStreamWaitEvent(sh, e0, 0));
            if (sl.tl_on && g == n_groups - 1) CU(cudaEventRecord(sl.tl[1], fr));
        }
        if (g_debug_chain_serial && e_chain_prev) CU(cudaStreamWaitEvent(sh, e_chain_prev, 0));
        LAUNCH(mean_kernel, (unsigned)((nd + 127) / 128), 128, 0, sh, ws.ctl, (int)nd, n_iq);
        TRY(run_coarse(lazy_src(graw, n_iq, n_taps, 0, dec), len_dec, p, nd, cap, ws, sh));      // latency-bound chain: high priority
        if (g_debug_occupy > 0) LAUNCH(occupy_kernel, (unsigned)nd, 64, (size_t)g_debug_occupy * 1024, sh, (long long)4500000, w.fine_raw);   // ~2.3 ms at 1.965 GHz
        CU(cudaEventRecord(e1, sh)); CU(cudaStreamWaitEvent(sg, e1, 0));
        if (e_chain_prev) CU(cudaEventDestroy(e_chain_prev));
        e_chain_prev = e1;
        if (sl.tl_on && g == n_groups - 1) CU(cudaEventRecord(sl.tl[2], sh));
        // Staggered batches (debug key 16, default on).  Ungated, two batches in flight run in lockstep (normal-priority kernels are dispatched
        // in arrival order, so their stages interleave and both finish together) and both fronts execute while nothing else does.  Gated,
        // the FP64 stages of batch k+1 start when batch k is done and the front of batch k+1 - on the high-priority stream, with the
        // column sums as the small-footprint TMA-ring kernel - runs UNDER the FP64 stages of batch k.  Device timelines and the cost of the
        // overlap (the ring's shared-memory traffic, the burst chain): profiles/r2n, r2p, r2q, r2z, r2ak and DESIGN.md section 4.
        if (g_debug_gate && c->last_slot >= 0 && c->last_slot != slot && c->slots[c->last_slot].busy)
            CU(cudaStreamWaitEvent(sg, c->slots[c->last_slot].done, 0));
        CU(cudaEventDestroy(e0));
        const bool tl_g = sl.tl_on && g == n_groups - 1;
        if (tl_g) CU(cudaEventRecord(sl.tl[4], sg));
        TRY(run_fine_peak(*c, lazy_src(graw, n_iq, n_taps, 0, 1), n_iq, osr, nd, cap, ws, sg));
        if (tl_g) CU(cudaEventRecord(sl.tl[5], sg));
        TRY(run_fine_rest(*c, with_cache(lazy_src(graw, n_iq, n_taps, 1, 1), ws, cap), n_iq, osr, carrier_freq, nd, cap, ws, sg));
        if (tl_g) CU(cudaEventRecord(sl.tl[6], sg));
        TRY(run_sch(lazy_src(graw, n_iq, n_taps, 2, 1), osr, nd, cap, ws, sg));
        if (tl_g) CU(cudaEventRecord(sl.tl[7], sg));
        TRY(run_post(*c, with_cache(lazy_src(graw, n_iq, n_taps, 3, 1), ws, cap), osr, carrier_freq, nd, cap, ws, true, sg));
        cudaEvent_t ev; CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, sg));
        if (sl.tl_on && g == n_groups - 1) CU(cudaEventRecord(sl.tl[3], sg));
        ev_done.push_back(ev);
    }
    if (e_sum) CU(cudaEventDestroy(e_sum));
    for (cudaEvent_t eg : e_sum_g) CU(cudaEventDestroy(eg));
    if (e_chain_prev) CU(cudaEventDestroy(e_chain_prev));
    for (cudaEvent_t ev : ev_done) { CU(cudaStreamWaitEvent(fr, ev, 0)); CU(cudaEventDestroy(ev)); }
    CU(cudaMemcpyAsync(h_res, w.res, sl.n_res, cudaMemcpyDeviceToHost, fr));
    if (coarse_pos) CU(cudaMemcpyAsync(h_arr, w.coarse_pos, sl.n_per, cudaMemcpyDeviceToHost, fr));
    if (coarse_snr) CU(cudaMemcpyAsync(h_arr + sl.n_per, w.coarse_snr, sl.n_per, cudaMemcpyDeviceToHost, fr));
    if (fcch_pos) CU(cudaMemcpyAsync(h_arr + 2 * sl.n_per, w.fcch_pos, sl.n_per, cudaMemcpyDeviceToHost, fr));
    if (pos_info) CU(cudaMemcpyAsync(h_arr + 3 * sl.n_per, w.pos_info, 12 * sl.n_per, cudaMemcpyDeviceToHost, fr));
    CU(cudaEventRecord(sl.done, fr));
    sl.results = results; sl.coarse_pos = coarse_pos; sl.coarse_snr = coarse_snr; sl.fcch_pos = fcch_pos; sl.pos_info = pos_info;
    sl.busy = true;
    c->last_slot = slot;
    return GSMCAL_OK;
}

int gsmcal_calibrate_batch_collect(int slot) {
    cudaEvent_t done;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (slot < 0 || slot >= kSlots) return fail(GSMCAL_ERR_ARG, "calibrate_batch_collect: slot must be 0..%d", kSlots - 1);
        Ctx *c; TRY(get_ctx(&c));
        if (!c->slots[slot].busy) return fail(GSMCAL_ERR_ARG, "calibrate_batch_collect: nothing was submitted to slot %d", slot);
        done = c->slots[slot].done;
    }
    CU(cudaEventSynchronize(done));                              // outside the lock: other slots can be submitted meanwhile
    std::lock_guard<std::mutex> lk(g_mu);
    Ctx *c; TRY(get_ctx(&c));
    Slot &sl = c->slots[slot];
    if (sl.tl_on && g_tl_base) {
        float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 8; ++i) cudaEventElapsedTime(&t[i], g_tl_base, sl.tl[i]);
        fprintf(stderr, "[gsmcal timeline] slot %d: front start %.3f ms, column sums done %.3f, burst chain done %.3f, FP64 stages start %.3f, "
                        "fine search +%.3f, fine tone +%.3f, SCH +%.3f, post +%.3f, done %.3f\n",
                slot, t[0], t[1], t[2], t[4], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[3] - t[7], t[3]);
        cudaGetLastError();
    }
    const char *h_res = sl.stage, *h_arr = sl.stage + align_up(sl.n_res);
    memcpy(sl.results, h_res, sl.n_res);
    if (sl.coarse_pos) memcpy(sl.coarse_pos, h_arr, sl.n_per);
    if (sl.coarse_snr) memcpy(sl.coarse_snr, h_arr + sl.n_per, sl.n_per);
    if (sl.fcch_pos) memcpy(sl.fcch_pos, h_arr + 2 * sl.n_per, sl.n_per);
    if (sl.pos_info) memcpy(sl.pos_info, h_arr + 3 * sl.n_per, 12 * sl.n_per);
    sl.busy = false;
    return GSMCAL_OK;
}

int gsmcal_calibrate_batch_cancel(int slot) {
    // gives the slot back without delivering results: waits for the batch (its kernels read the capture and write the slot's own
    // staging, never the caller's result pointers, which are only touched by _collect) and forgets the pointers
    cudaEvent_t done;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (slot < 0 || slot >= kSlots) return fail(GSMCAL_ERR_ARG, "calibrate_batch_cancel: slot must be 0..%d", kSlots - 1);
        Ctx *c; TRY(get_ctx(&c));
        if (!c->slots[slot].busy) return GSMCAL_OK;
        done = c->slots[slot].done;
    }
    CU(cudaEventSynchronize(done));
    std::lock_guard<std::mutex> lk(g_mu);
    Ctx *c; TRY(get_ctx(&c));
    Slot &sl = c->slots[slot];
    sl.results = nullptr; sl.coarse_pos = sl.coarse_snr = sl.fcch_pos = sl.pos_info = nullptr;
    sl.busy = false;
    return GSMCAL_OK;
}

int gsmcal_last_batch_stage_ms(double *ms, int cap_n) {
    // colsum(+H2D), coarse, fine_peak, fine ppm+tone+carrier, SCH, post-SCH - of the last gsmcal_calibrate_batch call
    int n = g_stage_n > 0 ? g_stage_n - 1 : 0;
    for (int i = 0; i < n && i < cap_n; ++i) ms[i] = g_stage_ms[i];
    return n;
}

int gsmcal_fcch_scan(const uint8_t *raw, int raw_mem, int64_t n_iq, int64_t n_chan, const double *coef, int n_taps, int osr, int coarse_dr,
                     double *snr, double *num_hit, double *position, int32_t *n_position, void *cuda_stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!raw || !coef || !snr || !num_hit || n_iq < 1 || n_chan < 1 || osr < 1) return fail(GSMCAL_ERR_ARG, "fcch_scan: bad arguments");
    CoarseParams p; TRY(coarse_params(coarse_dr, &p));
    const int dec = osr * coarse_dr;
    const i64 len_dec = (n_iq + dec - 1) / dec;
    if (p.n_first > len_dec) return fail(GSMCAL_ERR_RANGE, "fcch_scan: capture shorter than 23 frames");
    const int cap = (int)gsmcal_max_bursts(len_dec, coarse_dr);
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    Work w;
    TRY(make_work(c->work, n_chan, cap, p.n_first, 0, &w));
    TRY(set_taps(coef, n_taps, st));
    const uint8_t *draw = raw;
    if (raw_mem == GSMCAL_MEM_HOST) {
        void *din; TRY(c->in.get((size_t)2 * n_iq * n_chan, &din));
        TRY(copy_h2d(c->ring, g_device, din, raw, (size_t)2 * n_iq * n_chan, st));
        draw = (const uint8_t *)din;
    }
    CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_chan, st));
    TRY(run_colsum_u8(draw, n_iq, n_chan, w.ctl, st));
    LAUNCH(mean_kernel, (unsigned)((n_chan + 127) / 128), 128, 0, st, w.ctl, (int)n_chan, n_iq);
    TRY(run_coarse(lazy_src(draw, n_iq, n_taps, 0, dec), len_dec, p, n_chan, cap, w, st));
    LAUNCH(scan_accept_kernel, (unsigned)((n_chan + 63) / 64), 64, 0, st, w.ctl, (int)n_chan, cap, w.coarse_pos, w.coarse_snr, w.fo, w.gate);
    CU(cudaMemcpyAsync(snr, w.fo, sizeof(double) * n_chan, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(num_hit, w.gate, sizeof(double) * n_chan, cudaMemcpyDeviceToHost, st));
    if (position) CU(cudaMemcpyAsync(position, w.coarse_pos, sizeof(double) * n_chan * cap, cudaMemcpyDeviceToHost, st));
    std::vector<StreamCtl> h;
    if (n_position) { h.resize(n_chan); CU(cudaMemcpyAsync(h.data(), w.ctl, sizeof(StreamCtl) * n_chan, cudaMemcpyDeviceToHost, st)); }
    CU(cudaStreamSynchronize(st));
    if (n_position) for (int64_t i = 0; i < n_chan; ++i) n_position[i] = h[i].n_coarse;
    return GSMCAL_OK;
}

int gsmcal_fp64_peak(double *tflops, void *cuda_stream) {
    // measured DFMA throughput of this GPU (2 flop per DFMA), CUDA events around a register-only FMA kernel
    std::lock_guard<std::mutex> lk(g_mu);
    if (!tflops) return fail(GSMCAL_ERR_ARG, "fp64_peak: null output");
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, g_device));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 14;
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(e0, st));
        LAUNCH(fp64_peak_kernel, blocks, 256, 0, st, (double *)nullptr, iters, 0.999999, 1e-9);
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    CU(cudaEventDestroy(e0)); CU(cudaEventDestroy(e1));
    *tflops = best;
    return GSMCAL_OK;
}

int gsmcal_stage_launch(int stage, const void *in, void *out, int64_t n_iq, int64_t n_col, const double *coef, int n_taps, void *cuda_stream) {
    std::lock_guard<std::mutex> lk(g_mu);
    Ctx *c; TRY(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    static Work w; static i64 w_cols = 0;
    if (stage == 0 || w_cols != n_col) { TRY(make_work(c->work, n_col, 1, 1, 0, &w)); w_cols = n_col; }
    if (coef) TRY(set_taps(coef, n_taps, st));
    switch (stage) {
    case 0:
        CU(cudaMemsetAsync(w.ctl, 0, sizeof(StreamCtl) * n_col, st));
        return run_colsum_u8((const uint8_t *)in, n_iq, n_col, w.ctl, st);
    case 1: {
        i64 gx = (n_iq + 256 * 8 - 1) / (256 * 8); if (gx > 148 * 16) gx = 148 * 16;
        LAUNCH(raw2iq_store_kernel, dim3((unsigned)gx, (unsigned)n_col), 256, 0, st, (const uint8_t *)in, n_iq, w.ctl, (double2 *)out);
        return GSMCAL_OK; }
    case 2: return run_fir<false>(in, n_iq, n_iq, nullptr, n_taps, 1, n_col, (double2 *)out, n_iq, nullptr, st);
    case 3: return run_fir<true>(in, n_iq, 2 * n_iq, w.ctl, n_taps, 1, n_col, (double2 *)out, n_iq, nullptr, st);
    case 4: for (i64 d = 0; d < n_col; ++d) TRY(run_resample_derotate((const double2 *)in + d * n_iq, n_iq, -35e-6, 1, 0.0, 0, (double2 *)out + d * n_iq, n_iq, st)); return GSMCAL_OK;
    case 5: for (i64 d = 0; d < n_col; ++d) TRY(run_resample_derotate((const double2 *)in + d * n_iq, n_iq, 0.0, 0, 0.0123, 1, (double2 *)out + d * n_iq, n_iq, st)); return GSMCAL_OK;
    case 6: return run_fir<true>(in, n_iq, 2 * n_iq, w.ctl, n_taps, 64, n_col, (double2 *)out, (n_iq + 63) / 64, nullptr, st);
    default: return fail(GSMCAL_ERR_ARG, "stage_launch: unknown stage %d", stage);
    }
}

}  // extern "C"

// ====================================================================================================
// T1  gsm_SCH_training_sequence_gen.m:17-19,32,39 - GMSK per GSM 05.04 (host code: 512 samples, once per session)
// ====================================================================================================
namespace {
double gq_big_f(double u) { return u * 0.5 * erfc(-u / sqrt(2.0)) + exp(-0.5 * u * u) / sqrt(2.0 * M_PI); }
double gq_big_g(double t) {
    const double sigma = sqrt(log(2.0)) / (2.0 * M_PI * 0.3);
    return (sigma / 2.0) * (gq_big_f((t + 0.5) / sigma) - gq_big_f((t - 0.5) / sigma));
}
double gmsk_q(double tau) {
    if (tau < 0.0) tau = 0.0;
    if (tau > 4.0) tau = 4.0;
    const double g0 = gq_big_g(-2.0), g1 = gq_big_g(2.0);
    return (gq_big_g(tau - 2.0) - g0) / (g1 - g0);
}
void gmsk_bits(const int *bits, int nb, int osr, double *out) {      // differential encoding against a leading 0, bit 1 -> +1, then GMSK
    std::vector<double> a((size_t)nb);
    int prev = 0;
    for (int k = 0; k < nb; ++k) { a[k] = (bits[k] == prev) ? 1.0 : -1.0; prev = bits[k]; }   // ~abs(diff([0;data]))
    for (int n = 0; n < nb * osr; ++n) {
        double phase = 0.0;
        for (int k = 0; k < nb; ++k) phase += a[k] * gmsk_q(((double)n - (double)k * osr) / osr);
        out[2 * n] = cos((M_PI / 2.0) * phase);
        out[2 * n + 1] = sin((M_PI / 2.0) * phase);
    }
}
void gmsk_template(int osr, double *out) {
    static const int bits[64] = {1, 0, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0,
                                 0, 0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 0, 0,
                                 1, 0, 1, 0, 1, 1, 1, 0, 1, 1, 0, 0, 0, 0, 1, 1, 0, 1, 1};
    gmsk_bits(bits, 64, osr, out);
}
}  // namespace

#include "gsmcal_demod_api.inc"
