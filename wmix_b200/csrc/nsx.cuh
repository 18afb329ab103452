// WebRTC FIXED-POINT noise suppressor ("nsx"), one stream-frame per WARP.
//
// What it computes: WebRtcNsx_ProcessCore for one 10 ms mono frame
// (T:webrtc/modules/audio_processing/ns/nsx_core.c:1501-2118 with DataAnalysis :1184, DataSynthesis :1421,
// NoiseEstimation :334, the feature extraction :821-1181 and nsx_core_c.c:26 SpeechNoiseProb), on the SPL 16-bit complex
// FFT (T:webrtc/common_audio/signal_processing/complex_fft.c:27-301 behind real_fft.c:46-103) — the core wmix's ns_process
// runs when its switch is thrown (R:src/webrtc.c:511-523, `#define MAKE_WEBRTC_NSX`).  Integer throughout, so results
// are bit-exact, and every sum the reference forms is associative (wrapping 32-bit adds, max, min): they are warp
// reductions here, there are no in-order chains and no transcendentals — the two things that bound the float kernel.
//
// How it is laid out for the GPU:
//   * per-stream state is one contiguous, 128-byte-aligned record of 32-bit words (Geo<>): nine word arrays over the
//     bins — 16-bit quantities packed in pairs: (log-quantile, density) x 3 estimates, (quantile, filter), previous
//     magnitude; 32-bit: smoothed LRT, pause average, previous noise, start-up magnitude sum — then one line with the
//     Nyquist bin of all nine, one line of scalars, and the analysis history / synthesis tail as int16.  5.1 KB at 16 kHz
//     (2.8 KB at 8 kHz) against the float core's 8 KB.  Lane L owns bins L, L+32, ... so every array moves as whole
//     128-byte lines; lane 0 also owns the Nyquist bin.
//   * the 256 / 128-point COMPLEX radix-2 FFT runs in registers, 8 (4) packed int16 pairs per lane.  Input sample
//     L + 32 r sits at bit-reversed position 8*rev5(L) + rev3(r): the bit reversal is a renaming of lanes and
//     registers, no data moves.  Three (two) passes run per register group, with the reference's 16-bit rounding after
//     every pass, between groups the points cross lanes through a 1 KB tile whose XOR swizzle makes every pattern
//     conflict-free.  After the last group lane L holds points L + 32 h — exactly the bins (forward) and output samples
//     (inverse) it owns.  The inverse transform's per-pass scaling needs max|x| over all points: one warp max per pass.
//   * the body is a sequence of PHASES; lanes talk through the tile and through warp reductions that sit BETWEEN phases
//     (hardware redux on the device, a loop over lanes in the host emulation of tests/emu).
//
// x86 semantics are pinned where the reference leans on undefined behaviour (oracle/orc_nsx.c header): 32-bit shift
// counts are taken modulo 32 (xshr / xshl / xsar), signed overflow wraps.
#pragma once
#include "common.cuh"

namespace wmx {
namespace nsx {

constexpr int kStartupShort = 50;   // END_STARTUP_SHORT
constexpr int kStartupLong = 200;   // END_STARTUP_LONG
constexpr int kHistBins = 1000;     // HIST_PAR_EST
constexpr int kStartBand = 5;       // kStartBand
constexpr uint32_t kSatMax = 1048575u;

// word arrays of the record
enum ArrayId { A_LQD0 = 0, A_LQD1, A_LQD2, A_QF, A_MPREV, A_LRT, A_PAUSE, A_NPREV, A_INIT, kNumArrays };

// scalar words of the record
enum ScalarId {
    S_FRAME_IDX = 0, S_MODEL_COUNT, S_COUNTER0, S_COUNTER1, S_COUNTER2, S_Q_NOISE, S_Q_NOISE_PREV, S_Q_MAGN_PREV,
    S_MIN_NORM, S_PRIOR, S_FEAT_LRT, S_THR_LRT, S_FEAT_FLAT, S_THR_FLAT, S_FEAT_DIFF, S_THR_DIFF,
    S_W_LRT, S_W_FLAT, S_W_DIFF, S_CUR_AVG_E, S_TIME_AVG_E, S_TIME_AVG_ACC, S_WHITE, S_PINK_NUM, S_PINK_EXP,
    S_COUNT = 32
};

template <int ANA>
struct Geo {
    static constexpr int kAna = ANA;
    static constexpr int kBlock = ANA == 256 ? 160 : 80;
    static constexpr int kKeep = ANA - kBlock;               // analysis history / synthesis tail, samples
    static constexpr int kHalf = ANA / 2;
    static constexpr int kBins = kHalf + 1;
    static constexpr int kStages = ANA == 256 ? 8 : 7;
    static constexpr int kR = ANA / 32;                      // FFT points per lane
    static constexpr int kRB = ANA == 256 ? 3 : 2;           // log2(kR): passes per register group
    static constexpr int kGroups = (kStages + kRB - 1) / kRB;
    static constexpr int kK = kHalf / 32;                    // body bins per lane
    // record layout, 32-bit word offsets
    static constexpr int kOffArrays = 0;
    static constexpr int kOffNyq = kNumArrays * kHalf;       // [32], first kNumArrays used
    static constexpr int kOffScal = kOffNyq + 32;            // [32]
    static constexpr int kOffHist = kOffScal + 32;           // int16 [kKeep]
    static constexpr int kOffSyn = kOffHist + kKeep / 2;     // int16 [kKeep]
    static constexpr int kRecWords = (kOffSyn + kKeep / 2 + 31) / 32 * 32;   // 1312 (16 kHz) / 704 (8 kHz)
    // per-warp shared tile, words: the FFT exchange area, then one word per bin for neighbour / mirror look-ups
    static constexpr int kShFft = 0;
    static constexpr int kShBins = ANA;
    static constexpr int kShWords = ANA + kBins + 3;
};

// engine-wide constants (device global memory, copied to shared once per CTA)
struct Tables {
    uint32_t twiddle[128];    // FFT: entry m = (cos << 16) | (sin & 0xffff) of kSinTable1024 at m * 1024 / ANA
    int16_t window[256];      // kBlocks80w128x / kBlocks160w256x
    int16_t log_frac[256];    // WebRtcNsx_kLogTableFrac
    int16_t counter_div[202]; // WebRtcNsx_kCounterDiv
    int16_t log_index[130];   // kLogIndex
    int16_t factor1[258];     // kFactor1Table
    int16_t factor2[258];     // kFactor2Aggressiveness<policy>
    int16_t sigmoid[18];      // kIndicatorTable
    int16_t log_stage[10];    // WebRtcNsx_kLogTable
    // the pink-noise fit's constants for this rate (nsx_core.c:1364-1377, folded on the host)
    int32_t fit_det, fit_sum_log, fit_sum_sq;
    int32_t overdrive, floor_gain, gain_map;     // policy (nsx_core.c:786-814)
    int32_t lrt_max, lrt_min;
    int32_t pad[4];
};

// 32-bit shifts with the count taken modulo 32, as the reference's x86 build executes them
WMX_HD uint32_t xshr(uint32_t x, int c) { return x >> (c & 31); }
WMX_HD uint32_t xshl(uint32_t x, int c) { return x << (c & 31); }
WMX_HD int32_t xsar(int32_t x, int c) { return x >> (c & 31); }
WMX_HD int32_t xshift(int32_t x, int c) { return c >= 0 ? (int32_t)xshl((uint32_t)x, c) : xsar(x, -c); }   // WEBRTC_SPL_SHIFT_W32
// WebRtcSpl_NormW16 (spl_inl.h:140-160)
WMX_HD int norm_w16(int a)
{
    if (a == 0) return 0;
    if (a < 0) a = ~a;
    return a == 0 ? 15 : clz32((uint32_t)a) - 17;
}
WMX_HD int32_t mul_round(int32_t a, int32_t b, int c) { return (a * b + (1 << (c - 1))) >> c; }   // MUL_16_16_RSFT_WITH_ROUND
WMX_HD int32_t iabs(int32_t v) { return v < 0 ? -v : v; }
WMX_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
// log2 in Q8 of a non-zero value (nsx_core.c:361-367 and twice more)
WMX_HD int log2_q8(uint32_t v, const Tables& T)
{
    const int z = clz32(v);
    return ((31 - z) << 8) + T.log_frac[((v << z) & 0x7FFFFFFFu) >> 23];
}
// WebRtcSpl_SqrtFloor (spl_sqrt_floor.c:50-76); its argument is an int32: 2^31 arrives negative and gives 0
WMX_HD uint32_t sqrt_floor(uint32_t v)
{
    if (v & 0x80000000u) return 0;
#if defined(__CUDA_ARCH__)
    uint32_t r = (uint32_t)__fsqrt_rz(__uint2float_rz(v));
#else
    uint32_t r = (uint32_t)__builtin_sqrt((double)v);
#endif
    // the float estimate is within one of the floor: settle it exactly
    if (r * r > v) --r;
    if ((r + 1) * (r + 1) <= v) ++r;
    return r;
}
// packed complex int16 pair: low half real, high half imaginary
WMX_HD int32_t c_re(uint32_t w) { return (int32_t)(int16_t)(w & 0xFFFFu); }
WMX_HD int32_t c_im(uint32_t w) { return (int32_t)w >> 16; }
WMX_HD uint32_t c_pack(int32_t re, int32_t im) { return ((uint32_t)im << 16) | ((uint32_t)re & 0xFFFFu); }

// per-halfword signed max / min of three packed int16 pairs (one VIMNMX3.S16x2 on the device)
WMX_HD uint32_t max3_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    int32_t r = c_re(a), i = c_im(a);
    r = c_re(b) > r ? c_re(b) : r; r = c_re(c) > r ? c_re(c) : r;
    i = c_im(b) > i ? c_im(b) : i; i = c_im(c) > i ? c_im(c) : i;
    return c_pack(r, i);
#endif
}
WMX_HD uint32_t min3_s16x2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimin3_s16x2(a, b, c);
#else
    int32_t r = c_re(a), i = c_im(a);
    r = c_re(b) < r ? c_re(b) : r; r = c_re(c) < r ? c_re(c) : r;
    i = c_im(b) < i ? c_im(b) : i; i = c_im(c) < i ? c_im(c) : i;
    return c_pack(r, i);
#endif
}

// position of FFT point p in the exchange tile: the bits above the bank index are folded into the bank bits so that every
// register-group layout (lanes spanning any five of p's bits) hits 32 different banks
template <int ANA>
WMX_HD int swz(int p)
{
    const int h = p >> 5;
    return ANA == 256 ? p ^ ((h & 1) | (((h >> 1) & 1) * 10) | (((h >> 2) & 1) * 20)) : p ^ (((h & 1) * 5) | (((h >> 1) & 1) * 26));
}
// point index of register r of lane `lane` in the layout whose register bits start at b0
template <int ANA>
WMX_HD int fft_pos(int lane, int r, int b0)
{
    return (lane & ((1 << b0) - 1)) | (r << b0) | ((lane >> b0) << (b0 + Geo<ANA>::kRB));
}
WMX_HD int rev_bits(int v, int bits)
{
#if defined(__CUDA_ARCH__)
    return (int)(__brev((unsigned)v) >> (32 - bits));
#endif
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}

// per-lane values that live across phases
template <int ANA>
struct Lane {
    static constexpr int NS = Geo<ANA>::kK + 1;   // bin slots: body bins + the Nyquist slot (lane 0 only)
    uint32_t x[Geo<ANA>::kR];    // FFT points / samples
    uint32_t spec[NS];           // inst->real / inst->imag of the owned bins, packed
    uint32_t magn[NS], noise[NS], post[NS], prior[NS], near_prev[NS], nonsp[NS], fmodel[NS];
    uint32_t a[4];               // operands of the warp reductions
};

template <int ANA>
struct Warp {
#if defined(__CUDA_ARCH__)
    Lane<ANA> lane_regs;
    int lane_id;
#else
    Lane<ANA> lane_regs[32];
#endif
};

#if defined(__CUDA_ARCH__)
#define WMX_NSX_PHASE_BEGIN { const int lane = W.lane_id; Lane<ANA>& R = W.lane_regs; (void)lane; (void)R;
#define WMX_NSX_PHASE_END } __syncwarp();
#else
#define WMX_NSX_PHASE_BEGIN for (int lane = 0; lane < 32; ++lane) { Lane<ANA>& R = W.lane_regs[lane]; (void)R;
#define WMX_NSX_PHASE_END }
#endif
#define WMX_NSX_BINS(k) _Pragma("unroll") for (int k = 0; k <= K; ++k) if (k < K || lane == 0)

// warp reductions over Lane::a[I]; called between phases, the result is warp-uniform
template <int I, int ANA>
WMX_HD uint32_t warp_add(Warp<ANA>& W)
{
#if defined(__CUDA_ARCH__)
    return __reduce_add_sync(0xffffffffu, W.lane_regs.a[I]);
#else
    uint32_t s = 0;
    for (int l = 0; l < 32; ++l) s += W.lane_regs[l].a[I];
    return s;
#endif
}
template <int I, int ANA>
WMX_HD uint32_t warp_max_u(Warp<ANA>& W)
{
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(0xffffffffu, W.lane_regs.a[I]);
#else
    uint32_t s = 0;
    for (int l = 0; l < 32; ++l) s = W.lane_regs[l].a[I] > s ? W.lane_regs[l].a[I] : s;
    return s;
#endif
}
template <int I, int ANA>
WMX_HD int32_t warp_max_s(Warp<ANA>& W)
{
#if defined(__CUDA_ARCH__)
    return __reduce_max_sync(0xffffffffu, (int32_t)W.lane_regs.a[I]);
#else
    int32_t s = (int32_t)W.lane_regs[0].a[I];
    for (int l = 1; l < 32; ++l) s = (int32_t)W.lane_regs[l].a[I] > s ? (int32_t)W.lane_regs[l].a[I] : s;
    return s;
#endif
}
template <int I, int ANA>
WMX_HD int32_t warp_min_s(Warp<ANA>& W)
{
#if defined(__CUDA_ARCH__)
    return __reduce_min_sync(0xffffffffu, (int32_t)W.lane_regs.a[I]);
#else
    int32_t s = (int32_t)W.lane_regs[0].a[I];
    for (int l = 1; l < 32; ++l) s = (int32_t)W.lane_regs[l].a[I] < s ? (int32_t)W.lane_regs[l].a[I] : s;
    return s;
#endif
}

// histogram bin += 1 (counts stay below 512: no carry into the neighbouring half-word).  On the device a fire-and-forget
// RED on the containing 32-bit word: nothing waits for the histogram line to arrive
WMX_HD void hist_inc(int16_t* h, uint32_t idx)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(reinterpret_cast<unsigned int*>(h) + (idx >> 1), (idx & 1) ? 0x10000u : 1u);
#else
    h[idx]++;
#endif
}
// pull the stream's record towards L1 ahead of the phases that walk it (one 128-byte line per lane and call)
WMX_HD void l1_prefetch(const void* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// state word of bin slot k of array a
template <int ANA>
WMX_HD uint32_t* bin_word(uint32_t* rec, int a, int k, int lane)
{
    typedef Geo<ANA> G;
    return k < G::kK ? rec + G::kOffArrays + a * G::kHalf + 32 * k + lane : rec + G::kOffNyq + a;
}

// ------------------------------------------------------------------------------------------
// the SPL complex FFT on Lane::x (complex_fft.c, mode 1).  INV: the inverse with its data-dependent per-pass scaling;
// returns the number of scale shifts (0 for the forward transform).  Entry: lane L's register q holds point
// 8*rev5(L) + q (the caller has renamed its registers); exit: register h holds point L + 32 h.
// ------------------------------------------------------------------------------------------
template <int ANA, bool INV>
WMX_HD int fft_run(Warp<ANA>& W, uint32_t* tile, const Tables& T)
{
    typedef Geo<ANA> G;
    constexpr int ST = G::kStages, RB = G::kRB, NR = G::kR;
    int scale = 0;
#pragma unroll
    for (int g = 0; g < G::kGroups; ++g) {
        const int b0 = g * RB < ST - RB ? g * RB : ST - RB;
        if (g > 0) {
            WMX_NSX_PHASE_BEGIN
#pragma unroll
            for (int r = 0; r < NR; ++r) R.x[r] = tile[G::kShFft + swz<ANA>(fft_pos<ANA>(lane, r, b0))];
            WMX_NSX_PHASE_END
        }
        const int s_end = (g + 1) * RB < ST ? (g + 1) * RB : ST;
#pragma unroll
        for (int s = g * RB; s < s_end; ++s) {
            int shift = 1, round2 = 16384;
            if (INV) {
                // complex_fft.c:170-187: scale a pass down by one or two bits when the data is large.  max|x| over all
                // points from the per-halfword signed max and min of the packed pairs (|-32768| caps at 32767 there too)
                WMX_NSX_PHASE_BEGIN
                uint32_t mx = R.x[0], mn = R.x[0];
#pragma unroll
                for (int r = 1; r + 1 < NR; r += 2) {
                    mx = max3_s16x2(mx, R.x[r], R.x[r + 1]);
                    mn = min3_s16x2(mn, R.x[r], R.x[r + 1]);
                }
                mx = max3_s16x2(mx, R.x[NR - 1], R.x[NR - 1]);
                mn = min3_s16x2(mn, R.x[NR - 1], R.x[NR - 1]);
                const int32_t hi = c_re(mx) > c_im(mx) ? c_re(mx) : c_im(mx), lo = c_re(mn) < c_im(mn) ? c_re(mn) : c_im(mn);
                R.a[0] = (uint32_t)(hi > -lo ? hi : -lo);
                WMX_NSX_PHASE_END
                const uint32_t peak = umin(warp_max_u<0>(W), 32767u);
                // as arithmetic on the two comparisons, not as branches: one copy of the pass for all three scalings
                shift = (int)(peak > 13573u) + (int)(peak > 27146u);
                round2 = 8192 << shift;
                scale += shift;
            }
            WMX_NSX_PHASE_BEGIN
            const int ll = g == 0 ? rev_bits(lane, 5) : lane;   // the first layout is indexed by the bit-reversed lane
            const int rb = s - b0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                if (r & (1 << rb)) continue;
                const int p = fft_pos<ANA>(ll, r, b0);
                const uint32_t tw = T.twiddle[(p & ((1 << s) - 1)) << (ST - 1 - s)];
                const int32_t wr = c_im(tw), wi = INV ? c_re(tw) : -c_re(tw);
                const uint32_t lo = R.x[r], hi = R.x[r | (1 << rb)];
                const int32_t hr = c_re(hi), hi_ = c_im(hi);
                const int32_t tr = (wr * hr - wi * hi_ + 1) >> 1;
                const int32_t ti = (wr * hi_ + wi * hr + 1) >> 1;
                const int32_t qr = c_re(lo) * 16384, qi = c_im(lo) * 16384;
                R.x[r | (1 << rb)] = c_pack((qr - tr + round2) >> (shift + 14), (qi - ti + round2) >> (shift + 14));
                R.x[r] = c_pack((qr + tr + round2) >> (shift + 14), (qi + ti + round2) >> (shift + 14));
            }
            WMX_NSX_PHASE_END
        }
        if (g + 1 < G::kGroups) {
            WMX_NSX_PHASE_BEGIN
            const int ll = g == 0 ? rev_bits(lane, 5) : lane;
#pragma unroll
            for (int r = 0; r < NR; ++r) tile[G::kShFft + swz<ANA>(fft_pos<ANA>(ll, r, b0))] = R.x[r];
            WMX_NSX_PHASE_END
        }
    }
    return scale;
}

// ------------------------------------------------------------------------------------------
// the two-peak search of the threshold re-learning (nsx_core.c:924-940, :970-986) over a 1000-bin histogram: the
// reference's scan keeps the largest bin and the largest of the rest, the earlier index winning ties.  Every lane
// scans a 32-bin chunk with that rule, the 64 candidates are merged in index order with the same rule.
// ------------------------------------------------------------------------------------------
struct Peaks { uint32_t pos1, pos2; int w1, w2; };
WMX_HD void peak_feed(int v, int idx, int& max1, int& max2, int& i1, int& i2)
{
    if (v > max1) { max2 = max1; i2 = i1; max1 = v; i1 = idx; }
    else if (v > max2) { max2 = v; i2 = idx; }
}
template <int ANA>
WMX_HD Peaks two_peaks(Warp<ANA>& W, const int16_t* h, uint32_t* tile)
{
    typedef Geo<ANA> G;
    uint32_t* cand = tile + G::kShFft;   // [32][4]: max1, idx1, max2, idx2 of every chunk
    WMX_NSX_PHASE_BEGIN
    int max1 = 0, max2 = 0, i1 = -1, i2 = -1;
    for (int i = 32 * lane; i < 32 * lane + 32 && i < kHistBins; ++i) peak_feed(h[i], i, max1, max2, i1, i2);
    cand[4 * lane + 0] = (uint32_t)max1; cand[4 * lane + 1] = (uint32_t)i1;
    cand[4 * lane + 2] = (uint32_t)max2; cand[4 * lane + 3] = (uint32_t)i2;
    WMX_NSX_PHASE_END
    int max1 = 0, max2 = 0, i1 = -1, i2 = -1;
    for (int c = 0; c < 32; ++c) {
        // a chunk's two candidates, earlier index first; idx < 0 = none (the running maxima only move on a larger value)
        const int m1 = (int)cand[4 * c], j1 = (int)cand[4 * c + 1], m2 = (int)cand[4 * c + 2], j2 = (int)cand[4 * c + 3];
        if (j2 >= 0 && j2 < j1) { peak_feed(m2, j2, max1, max2, i1, i2); peak_feed(m1, j1, max1, max2, i1, i2); }
        else { if (j1 >= 0) peak_feed(m1, j1, max1, max2, i1, i2); if (j2 >= 0) peak_feed(m2, j2, max1, max2, i1, i2); }
    }
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
    Peaks pk;
    pk.w1 = max1; pk.w2 = max2;
    pk.pos1 = i1 >= 0 ? (uint32_t)(2 * i1 + 1) : 0u;
    pk.pos2 = i2 >= 0 ? (uint32_t)(2 * i2 + 1) : 0u;
    return pk;
}

// 8192 +/- the interpolated sigmoid table value (nsx_core_c.c:98-110, :131-143, :178-191); past the table the caller's
// saturated value stands
WMX_HD int32_t sigmoid_q14(uint32_t x_q14, bool upper, bool rounded, const Tables& T)
{
    const int idx = (int16_t)(x_q14 >> 14);
    if (idx < 0 || idx >= 16) return upper ? 16384 : 0;
    const int d = (int16_t)(T.sigmoid[idx + 1] - T.sigmoid[idx]);
    const int frac = (int)(x_q14 & 0x3fffu);
    const int v = (int16_t)(T.sigmoid[idx] + (int16_t)(rounded ? mul_round(d, frac, 14) : (d * frac) >> 14));
    return (int16_t)(upper ? 8192 + v : 8192 - v);
}

// nsx_core.c:586-628 (CalcParametricNoiseEstimate)
WMX_HD void pink_estimate(int min_norm, int stages, int frame_idx, int exp_avg, int32_t num_avg, int bin, uint32_t& est, uint32_t& est_avg,
                          const Tables& T)
{
    int32_t t2 = (exp_avg * T.log_index[bin]) >> 15;
    int32_t t1 = num_avg - t2;
    t1 += (min_norm - stages) * 2048;
    if (t1 > 0) {
        const int int_part = (int16_t)(t1 >> 11), frac = t1 & 0x7ff;
        if (frac >> 10) {
            t2 = (2048 - frac) * 1244;
            t2 = 2048 - (t2 >> 10);
        } else {
            t2 = (frac * 804) >> 10;
        }
        t2 = xshift(t2, int_part - 11);
        est_avg = xshl(1u, int_part) + (uint32_t)t2;
        est = est_avg * (uint32_t)(frame_idx + 1);
    }
}

// read-out of the synthesis buffer when nothing is added to it (zero input, nsx_core.c:1439-1452)
template <int ANA>
WMX_HD void synth_read_out_only(Warp<ANA>& W, uint32_t* rec, int16_t* out)
{
    typedef Geo<ANA> G;
    int16_t* syn = reinterpret_cast<int16_t*>(rec + G::kOffSyn);
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < G::kR; ++r) {
        const int i = lane + 32 * r;
        R.x[r] = i < G::kKeep ? (uint32_t)(uint16_t)syn[i] : 0u;
    }
    WMX_NSX_PHASE_END
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < G::kR; ++r) {
        const int i = lane + 32 * r;
        if (i < G::kBlock) out[i] = (int16_t)R.x[r];
        else syn[i - G::kBlock] = (int16_t)R.x[r];
    }
    WMX_NSX_PHASE_END
}

// ------------------------------------------------------------------------------------------
// one frame of one stream.  rec: the stream's record; hist: int16 [3][1000] (LRT, flatness, difference);
// in / out: kBlock samples, may alias; tile: Geo::kShWords words of shared memory owned by this warp.
// ------------------------------------------------------------------------------------------
// HB: also returns the time-domain gain (Q14) of wmix's second band for this frame (nsx_core.c:2054-2107), or -1 when the
// frame was all zeros and the second band passes unscaled (:1577-1592); 0 otherwise.
template <int ANA, bool HB = false>
WMX_HD int frame(Warp<ANA>& W, uint32_t* rec, int16_t* hist, const int16_t* in, int16_t* out, uint32_t* tile, const Tables& T)
{
    typedef Geo<ANA> G;
    constexpr int K = G::kK, NR = G::kR, ST = G::kStages, HALF = G::kHalf;
    int32_t* sc = reinterpret_cast<int32_t*>(rec + G::kOffScal);
    int16_t* hist16 = reinterpret_cast<int16_t*>(rec + G::kOffHist);
    int16_t* syn16 = reinterpret_cast<int16_t*>(rec + G::kOffSyn);
    uint32_t* tbins = tile + G::kShBins;

    // ---- analysis buffer, window, peak (nsx_core.c:524-541, :1221-1227) ----
    WMX_NSX_PHASE_BEGIN
    // the record (the start-up array only while it is in use) is wanted in L1 by the time the transform is done
    for (int line = lane; line < (sc[S_FRAME_IDX] < kStartupShort ? G::kOffHist : (int)A_INIT * HALF) / 32; line += 32) l1_prefetch(rec + 32 * line);
    if (lane < (G::kOffHist - G::kOffNyq) / 32) l1_prefetch(rec + G::kOffNyq + 32 * lane);
    int32_t peak = 0, smax = -1;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i = lane + 32 * r;
        const int32_t v = i < G::kKeep ? hist16[i] : in[i - G::kKeep];
        const int32_t w = (int16_t)mul_round(T.window[i], v, 14);
        R.x[r] = (uint32_t)w;
        const int32_t a = iabs(w);
        peak = a > peak ? a : peak;
        const int32_t a16 = (int16_t)a;   // GetScalingSquare takes |x| in int16: -32768 stays negative and never wins
        smax = a16 > smax ? a16 : smax;
    }
    R.a[0] = (uint32_t)peak;
    R.a[1] = (uint32_t)smax;
    WMX_NSX_PHASE_END
    // the new history is the frame's last kKeep samples (every lane has read the old one)
    WMX_NSX_PHASE_BEGIN
    for (int i = lane; i < G::kKeep; i += 32) hist16[i] = in[G::kBlock - G::kKeep + i];
    WMX_NSX_PHASE_END
    const int32_t peak = (int32_t)umin(warp_max_u<0>(W), 32767u);
    const int32_t smax = warp_max_s<1>(W);
    int scale_in = 0;
    {
        const int nbits = size_in_bits((uint32_t)ANA);
        const int t = norm_w32(smax * smax);
        if (smax != 0) scale_in = t > nbits ? 0 : nbits - t;
    }
    WMX_NSX_PHASE_BEGIN
    uint32_t e = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int32_t w = (int32_t)R.x[r];
        e += (uint32_t)((w * w) >> scale_in);
    }
    R.a[0] = e;
    WMX_NSX_PHASE_END
    int32_t energy_in = (int32_t)warp_add<0>(W);
    const int norm = norm_w16(peak);
    if (peak == 0) {
        // zero input: only the buffers move (nsx_core.c:1228-1232, :1574-1594)
        synth_read_out_only<ANA>(W, rec, out);
        return -1;
    }

    int frame_idx = sc[S_FRAME_IDX];
    int min_norm = sc[S_MIN_NORM];
    const int net_norm = ST - norm;
    int drop_magn = norm - min_norm;
    const int drop_init = -drop_magn > 0 ? -drop_magn : 0;
    min_norm -= drop_init;
    if (drop_magn < 0) drop_magn = 0;
    const bool startup = frame_idx < kStartupShort;   // tested on the index BEFORE this frame's increment (:1261)

    // ---- forward transform of the normalised frame (nsx_core.c:1243-1246) ----
    WMX_NSX_PHASE_BEGIN
    uint32_t t[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) t[r] = c_pack((int16_t)((int32_t)R.x[r] << norm), 0);
#pragma unroll
    for (int q = 0; q < NR; ++q) R.x[q] = t[rev_bits(q, G::kRB)];
    WMX_NSX_PHASE_END
    fft_run<ANA, false>(W, tile, T);

    // ---- magnitudes, their sums, the start-up model (nsx_core.c:1248-1418) ----
    WMX_NSX_PHASE_BEGIN
    uint32_t e_sum = 0, m_sum = 0;
    int32_t slm = 0, slilm = 0;
    WMX_NSX_BINS(k) {
        const int bin = k < K ? 32 * k + lane : HALF;
        const uint32_t xw = k < K ? R.x[k] : R.x[K];   // lane 0 register K (= kR / 2) is point ANA / 2
        int32_t re = c_re(xw), im = (int16_t)-c_im(xw);
        uint32_t e, m;
        if (bin == 0 || bin == HALF) {
            im = 0;
            e = (uint32_t)(re * re);
            m = (uint32_t)iabs(re) & 0xFFFFu;
        } else {
            const int32_t fi = c_im(xw);
            e = (uint32_t)(re * re) + (uint32_t)(fi * fi);
            m = sqrt_floor(e) & 0xFFFFu;
        }
        R.spec[k] = c_pack(re, im);
        R.magn[k] = m;
        e_sum += e;
        m_sum += m;
        if (startup) {
            uint32_t* iw = bin_word<ANA>(rec, A_INIT, k, lane);
            *iw = (*iw >> drop_init) + (m >> drop_magn);
            if (bin >= kStartBand) {
                const int lg = m ? (int16_t)log2_q8(m, T) : 0;
                slm += lg;
                slilm += (T.log_index[bin] * lg) >> 3;
            }
        }
    }
    R.a[0] = e_sum; R.a[1] = m_sum; R.a[2] = (uint32_t)slm; R.a[3] = (uint32_t)slilm;
    WMX_NSX_PHASE_END
    const uint32_t magn_energy = warp_add<0>(W);
    const uint32_t sum_magn = warp_add<1>(W);
    if (startup) {
        const int32_t sum_log_magn = (int32_t)warp_add<2>(W), sum_log_i_log_magn = (int32_t)warp_add<3>(W);
        uint32_t white = (uint32_t)sc[S_WHITE] >> drop_init;
        uint32_t tu32 = sum_magn * (uint32_t)T.overdrive;
        tu32 >>= ST + 8;
        tu32 >>= drop_magn;
        white += tu32;
        int16_t det = (int16_t)T.fit_det;
        const int16_t sum_log_i = (int16_t)T.fit_sum_log, sum_log_i_sq = (int16_t)T.fit_sum_sq;
        int zeros = 16 - norm_w32(sum_log_magn);
        if (zeros < 0) zeros = 0;
        int32_t t1 = sum_log_magn << 1;
        const uint16_t slm16 = (uint16_t)(t1 >> zeros);
        int32_t t2 = (int32_t)sum_log_i_sq * slm16;
        tu32 = (uint32_t)(sum_log_i_log_magn >> 12);
        uint16_t tu16 = (uint16_t)((uint16_t)sum_log_i << 1);
        if ((uint32_t)sum_log_i > tu32) tu16 = (uint16_t)(tu16 >> zeros);
        else tu32 >>= zeros;
        t2 -= (int32_t)(tu32 * (uint32_t)tu16);
        det = (int16_t)(det >> zeros);
        t2 = div_w32_w16(t2, det);
        t2 += net_norm * 2048;
        if (t2 < 0) t2 = 0;
        const int32_t pink_num = sc[S_PINK_NUM] + t2;
        int32_t pink_exp = sc[S_PINK_EXP];
        t2 = (int32_t)sum_log_i * slm16;
        t1 = sum_log_i_log_magn >> (3 + zeros);
        t1 *= G::kBins - kStartBand;
        t2 -= t1;
        if (t2 > 0) {
            t1 = div_w32_w16(t2, det);
            pink_exp += t1 > 16384 ? 16384 : (t1 < 0 ? 0 : t1);
        }
        WMX_NSX_PHASE_BEGIN
        if (lane == 0) { sc[S_WHITE] = (int32_t)white; sc[S_PINK_NUM] = pink_num; sc[S_PINK_EXP] = pink_exp; }
        WMX_NSX_PHASE_END
    }

    // from here on the frame counts (nsx_core.c:1597)
    ++frame_idx;
    const int q_magn = norm - ST;

    // ---- spectral flatness (nsx_core.c:1022-1084) ----
    uint32_t feat_flat = (uint32_t)sc[S_FEAT_FLAT];
    {
        WMX_NSX_PHASE_BEGIN
        uint32_t num = 0, any_zero = 0;
        WMX_NSX_BINS(k) {
            const int bin = k < K ? 32 * k + lane : HALF;
            if (bin >= 1) {
                if (R.magn[k]) num += (uint32_t)log2_q8(R.magn[k], T);
                else any_zero = 1;
            }
        }
        R.a[0] = num; R.a[1] = any_zero;
        WMX_NSX_PHASE_END
        const uint32_t num = warp_add<0>(W);
        const uint32_t zero_bins = warp_add<1>(W);
        if (zero_bins) {
            feat_flat -= (feat_flat * 4915u) >> 14;
        } else {
            WMX_NSX_PHASE_BEGIN
            R.a[0] = lane == 0 ? R.magn[0] : 0u;
            WMX_NSX_PHASE_END
            const uint32_t den = sum_magn - warp_add<0>(W);
            const int z = norm_u32(den);
            int32_t t = ((31 - z) << 8) + T.log_frac[((den << z) & 0x7FFFFFFFu) >> 23];
            int32_t lg = (int32_t)num;
            lg += (int32_t)(ST - 1) << (ST + 7);
            lg -= t << (ST - 1);
            lg = (int32_t)((uint32_t)lg << (10 - ST));
            t = (int32_t)(0x00020000 | (iabs(lg) & 0x0001FFFF));
            const int int_part = (int16_t)(7 - (lg >> 17));
            const int32_t cur = int_part > 0 ? xsar(t, int_part) : (int32_t)xshl((uint32_t)t, -int_part);
            t = cur - (int32_t)feat_flat;
            t *= 4915;
            feat_flat += (uint32_t)(t >> 14);
        }
    }

    // ---- quantile noise estimate (nsx_core.c:334-453) ----
    int q_noise = sc[S_Q_NOISE];
    {
        const int tab = ST - norm;
        const int logval = tab < 0 ? -T.log_stage[-tab] : T.log_stage[tab];
        // log-magnitude in Q8 (natural log), shared by the three estimates; parked in the start-up filter slots
        WMX_NSX_PHASE_BEGIN
        WMX_NSX_BINS(k) {
            int lmagn = logval;
            if (R.magn[k]) lmagn = (int16_t)((int16_t)((log2_q8(R.magn[k], T) * 22713) >> 15) + logval);
            R.fmodel[k] = (uint32_t)lmagn;
        }
        WMX_NSX_PHASE_END
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const int counter = sc[S_COUNTER0 + e];
            const int cdiv = T.counter_div[counter];
            const int cprod = (int16_t)(counter * cdiv);
            const bool wrap = counter >= kStartupLong;
            // the estimate whose counter wraps is published once start-up is over; during start-up the LAST one is, every frame
            const bool publish = (wrap && frame_idx >= kStartupLong) || (e == 2 && frame_idx < kStartupLong);
            WMX_NSX_PHASE_BEGIN
            int32_t lq_max = -32768;
            WMX_NSX_BINS(k) {
                const int lmagn = (int32_t)R.fmodel[k];
                uint32_t* wp = bin_word<ANA>(rec, A_LQD0 + e, k, lane);
                const uint32_t w = *wp;
                int lq = lo16((int32_t)w), dn = hi16((int32_t)w);
                // delta = FACTOR_Q16 >> (14 - NormW16(density)): for density > 512 that is 40 << 16 >> floor(log2(density))
                const uint32_t delta = dn > 512 ? 2621440u >> (31 - clz32((uint32_t)dn)) : (frame_idx < kStartupLong ? 1024u : 5120u);
                const uint32_t step = (delta * (uint32_t)cdiv) >> 14;   // positive, < 2^15: the reference's int16 casts and signed divisions are plain shifts
                if (lmagn > lq) {
                    lq = (int16_t)(lq + (int)((step + 2) >> 2));
                } else {
                    lq = (int16_t)(lq - (int)((((step + 1) >> 1) * 3) >> 1));
                    if (lq < logval) lq = logval;
                }
                if (iabs(lmagn - lq) < 3) dn = (int16_t)((int16_t)mul_round(dn, cprod, 15) + (int16_t)mul_round(21845, cdiv, 15));
                *wp = (uint32_t)pack16((int16_t)lq, (int16_t)dn);
                lq_max = lq > lq_max ? lq : lq_max;
                R.noise[k] = (uint32_t)lq;   // kept for the publication below
            }
            R.a[0] = (uint32_t)lq_max;
            WMX_NSX_PHASE_END
            if (publish) {
                // nsx_core.c:303-331 (UpdateNoiseEstimate)
                const int lq_peak = warp_max_s<0>(W);
                q_noise = 14 - mul_round(11819, lq_peak, 21);
                WMX_NSX_PHASE_BEGIN
                WMX_NSX_BINS(k) {
                    const int32_t p = 11819 * (int32_t)R.noise[k];
                    int32_t v = 0x00200000 | (p & 0x001FFFFF);
                    int sh = (int16_t)(p >> 21);
                    sh = (int16_t)(sh - 21);
                    sh = (int16_t)(sh + (int16_t)q_noise);
                    v = sh < 0 ? xsar(v, -sh) : (int32_t)xshl((uint32_t)v, sh);
                    uint32_t* wp = bin_word<ANA>(rec, A_QF, k, lane);
                    *wp = (*wp & 0xFFFF0000u) | (uint16_t)sat16(v);
                }
                WMX_NSX_PHASE_END
            }
            WMX_NSX_PHASE_BEGIN
            if (lane == 0) sc[S_COUNTER0 + e] = (wrap ? 0 : counter) + 1;
            WMX_NSX_PHASE_END
        }
    }
    int16_t q_noise16 = (int16_t)q_noise;

    // ---- noise for this frame; blended with the white / pink model during start-up (nsx_core.c:1609-1710) ----
    uint32_t time_avg_e = (uint32_t)sc[S_TIME_AVG_E];
    {
        const bool blend = frame_idx < kStartupShort;
        int q_use = 0;
        uint32_t est0 = 0, est_avg0 = 0;
        int exp_avg = 0;
        int32_t num_avg = 0;
        const int32_t pink_exp = sc[S_PINK_EXP];
        if (blend) {
            q_use = q_noise16 < min_norm - ST ? q_noise16 : min_norm - ST;
            if (pink_exp) {
                exp_avg = (int16_t)div_w32_w16(pink_exp, (int16_t)(frame_idx + 1));
                num_avg = div_w32_w16(sc[S_PINK_NUM], (int16_t)(frame_idx + 1));
                pink_estimate(min_norm, ST, frame_idx, exp_avg, num_avg, kStartBand, est0, est_avg0, T);
            } else {
                est0 = (uint32_t)sc[S_WHITE];
                est_avg0 = est0 / (uint32_t)(frame_idx + 1);
            }
        }
        WMX_NSX_PHASE_BEGIN
        WMX_NSX_BINS(k) {
            const int bin = k < K ? 32 * k + lane : HALF;
            uint32_t nz = (uint32_t)(int32_t)lo16((int32_t)*bin_word<ANA>(rec, A_QF, k, lane));   // (uint32_t)(int16 quantile)
            if (blend) {
                uint32_t est = est0, est_avg = est_avg0;
                if (pink_exp && bin >= kStartBand) {
                    est = 0;
                    est_avg = 0;
                    pink_estimate(min_norm, ST, frame_idx, exp_avg, num_avg, bin, est, est_avg, T);
                }
                uint32_t fm = (uint32_t)T.floor_gain;
                const uint32_t init = *bin_word<ANA>(rec, A_INIT, k, lane);
                if (init) {
                    uint32_t numer = init << 8;
                    uint32_t a = est * (uint32_t)T.overdrive;
                    if (numer > a) {
                        numer -= a;
                        int n = norm_u32(numer);
                        if (n > 6) n = 6;
                        numer <<= n;
                        a = init >> (6 - n);
                        if (a == 0) a = 1;
                        const uint32_t b = numer / a;
                        fm = b > 16384u ? 16384u : (b < (uint32_t)T.floor_gain ? (uint32_t)T.floor_gain : b);
                    }
                }
                R.fmodel[k] = fm;
                uint32_t a = xshr(nz, q_noise16 - q_use);
                uint32_t b = xshr(est_avg, min_norm - ST - q_use);
                int shifts = 0;
                if (a & 0xfc000000u) { a >>= 6; b >>= 6; shifts = 6; }
                a *= (uint32_t)frame_idx;
                b *= (uint32_t)(kStartupShort - frame_idx);
                nz = (a + b) / (uint32_t)kStartupShort;
                nz <<= shifts;
            }
            R.noise[k] = nz;
        }
        WMX_NSX_PHASE_END
        if (blend) q_noise16 = (int16_t)q_use;
        if (frame_idx < kStartupLong) {
            const uint32_t acc = (uint32_t)sc[S_TIME_AVG_ACC] + xshr(magn_energy, 2 * norm + ST - 1);
            time_avg_e = acc / (uint32_t)(uint16_t)(frame_idx + 1);
            WMX_NSX_PHASE_BEGIN
            if (lane == 0) sc[S_TIME_AVG_ACC] = (int32_t)acc;
            WMX_NSX_PHASE_END
        }
    }

    // ---- step 1: posterior / prior SNR (nsx_core.c:1724-1785); spectral-difference sums (:1107-1137) ----
    const int q_magn_prev = sc[S_Q_MAGN_PREV], q_noise_prev = sc[S_Q_NOISE_PREV];
    WMX_NSX_PHASE_BEGIN
    const int post_shifts = 6 + q_magn - q_noise16;
    const int shifts = 5 - q_magn_prev + q_noise_prev;
    int32_t p_sum = 0, p_max = 0, p_min = 0x7fffffff;
    WMX_NSX_BINS(k) {
        // (the quotients are formed unconditionally on a safe divisor and selected afterwards: with 32 bins per instruction
        // some lane needs every one of them, and a branch around a division only adds the divergence bookkeeping)
        uint32_t a = R.magn[k] << 6, b;
        b = post_shifts < 0 ? xshr(R.noise[k], -post_shifts) : xshl(R.noise[k], post_shifts);
        const uint32_t q_post = (a << 11) / (b ? b : 1u);
        const uint32_t post = a > b ? (b > 0 ? umin(q_post, kSatMax) : kSatMax) : 2048u;
        const uint32_t filt = *bin_word<ANA>(rec, A_QF, k, lane) >> 16;
        const uint32_t mprev = *bin_word<ANA>(rec, A_MPREV, k, lane) & 0xFFFFu;
        a = (mprev * filt) << 3;
        b = xshr(*bin_word<ANA>(rec, A_NPREV, k, lane), shifts);
        a = b > 0 ? umin(a / (b ? b : 1u), kSatMax) : kSatMax;
        R.near_prev[k] = a;
        R.post[k] = post;
        R.prior[k] = 2048u + ((a * 2007u + (post - 2048u) * 41u + 512u) >> 10);
        const int32_t pa = (int32_t)*bin_word<ANA>(rec, A_PAUSE, k, lane);
        p_sum += pa;
        p_max = pa > p_max ? pa : p_max;
        p_min = pa < p_min ? pa : p_min;
    }
    R.a[0] = (uint32_t)p_sum; R.a[1] = (uint32_t)p_max; R.a[2] = (uint32_t)p_min;
    WMX_NSX_PHASE_END

    // ---- spectral difference feature (nsx_core.c:1091-1181) ----
    uint32_t feat_diff = (uint32_t)sc[S_FEAT_DIFF];
    uint32_t cur_avg_e = (uint32_t)sc[S_CUR_AVG_E];
    {
        int32_t avg_pause = (int32_t)warp_add<0>(W);
        const int32_t max_pause = warp_max_s<1>(W), min_pause = warp_min_s<2>(W);
        avg_pause >>= ST - 1;
        const int32_t avg_magn = (int32_t)(sum_magn >> (ST - 1));
        const int32_t dev = max_pause - avg_pause > avg_pause - min_pause ? max_pause - avg_pause : avg_pause - min_pause;
        int shifts = 10 + ST - norm_w32(dev);
        if (shifts < 0) shifts = 0;
        WMX_NSX_PHASE_BEGIN
        uint32_t var_magn = 0, var_pause = 0;
        int32_t cov = 0;
        WMX_NSX_BINS(k) {
            const int32_t dm = (int16_t)((int32_t)R.magn[k] - avg_magn);
            const int32_t dp = (int32_t)*bin_word<ANA>(rec, A_PAUSE, k, lane) - avg_pause;
            var_magn += (uint32_t)(dm * dm);
            cov = wadd(cov, wmul(dp, dm));
            const int32_t q = dp >> shifts;
            var_pause += (uint32_t)wmul(q, q);
        }
        R.a[0] = var_magn; R.a[1] = var_pause; R.a[2] = (uint32_t)cov;
        WMX_NSX_PHASE_END
        const uint32_t var_magn = warp_add<0>(W);
        uint32_t var_pause = warp_add<1>(W);
        const int32_t cov = (int32_t)warp_add<2>(W);
        cur_avg_e += xshr(magn_energy, 2 * norm + ST - 1);
        uint32_t diff = var_magn;
        if (var_pause && cov) {
            uint32_t u1 = (uint32_t)(cov >= 0 ? cov : -cov);
            const int norm32 = norm_u32(u1) - 16;
            u1 = norm32 > 0 ? u1 << norm32 : u1 >> -norm32;
            const uint32_t u2 = u1 * u1;
            shifts += norm32;
            shifts <<= 1;
            if (shifts < 0) { var_pause = xshr(var_pause, -shifts); shifts = 0; }
            if (var_pause > 0) {
                u1 = u2 / var_pause;
                u1 = xshr(u1, shifts);
                diff -= umin(diff, u1);
            } else {
                diff = 0;
            }
        }
        const uint32_t u = diff >> (2 * norm);
        if (feat_diff > u) feat_diff -= ((feat_diff - u) * 77u) >> 8;
        else feat_diff += ((u - feat_diff) * 77u) >> 8;
    }

    // ---- feature histograms; every 512 frames the thresholds and weights are re-learned (nsx_core.c:821-1016, :1795-1835) ----
    int32_t feat_lrt = sc[S_FEAT_LRT], thr_lrt = sc[S_THR_LRT];
    uint32_t thr_flat = (uint32_t)sc[S_THR_FLAT], thr_diff = (uint32_t)sc[S_THR_DIFF];
    int w_lrt = sc[S_W_LRT], w_flat = sc[S_W_FLAT], w_diff = sc[S_W_DIFF];
    int model_count = sc[S_MODEL_COUNT] + 1;
    int16_t* h_lrt = hist;
    int16_t* h_flat = hist + kHistBins;
    int16_t* h_diff = hist + 2 * kHistBins;
    if (model_count != (1 << 9)) {
        WMX_NSX_PHASE_BEGIN
        if (lane == 0) {
            uint32_t idx = (uint32_t)feat_lrt;
            if (idx < (uint32_t)kHistBins) hist_inc(h_lrt, idx);
            idx = (feat_flat * 5u) >> 8;
            if (idx < (uint32_t)kHistBins) hist_inc(h_flat, idx);
            idx = kHistBins;
            if (time_avg_e > 0) idx = ((feat_diff * 5u) >> ST) / time_avg_e;
            if (idx < (uint32_t)kHistBins) hist_inc(h_diff, idx);
        }
        WMX_NSX_PHASE_END
    } else {
        model_count = 0;
        // LRT histogram moments: bins 0..9 also feed the low-range mean and count
        WMX_NSX_PHASE_BEGIN
        int32_t sq = 0, all = 0;
        for (int i = 32 * lane; i < 32 * lane + 32 && i < kHistBins; ++i) {
            const int j = 2 * i + 1;
            const int32_t t = h_lrt[i] * j;
            all = wadd(all, t);
            sq = wadd(sq, wmul(t, j));
        }
        R.a[0] = (uint32_t)sq; R.a[1] = (uint32_t)all;
        WMX_NSX_PHASE_END
        const int32_t avg_sq = (int32_t)warp_add<0>(W), avg_all = (int32_t)warp_add<1>(W);
        int32_t avg = 0;
        int16_t count = 0;
        for (int i = 0; i < 10; ++i) {
            avg += h_lrt[i] * (2 * i + 1);
            count = (int16_t)(count + h_lrt[i]);
        }
        const int32_t fluct = wsub(wmul(avg_sq, count), wmul(avg, avg_all));
        const int32_t thr_fluct = 10240 * count;
        uint32_t u = 6u * (uint32_t)avg;
        if (fluct < thr_fluct || count == 0 || u > (uint32_t)(100 * count)) {
            thr_lrt = T.lrt_max;
        } else {
            const int32_t t = (int32_t)((u << (9 + ST)) / (uint32_t)(int32_t)count / 25u);
            thr_lrt = t > T.lrt_max ? T.lrt_max : (t < T.lrt_min ? T.lrt_min : t);
        }
        int use_flat = 1, use_diff = fluct < thr_fluct ? 0 : 1;
        Peaks pk = two_peaks<ANA>(W, h_flat, tile);
        if (pk.pos1 - pk.pos2 < 4u && pk.w2 * 2 > pk.w1) {
            pk.w1 += pk.w2;
            pk.pos1 = (pk.pos1 + pk.pos2) >> 1;
        }
        if (pk.w1 < 154 || pk.pos1 < 24u) {
            use_flat = 0;
        } else {
            u = 922u * pk.pos1;
            thr_flat = u > 38912u ? 38912u : (u < 4096u ? 4096u : u);
        }
        if (use_diff) {
            pk = two_peaks<ANA>(W, h_diff, tile);
            if (pk.pos1 - pk.pos2 < 4u && pk.w2 * 2 > pk.w1) {
                pk.w1 += pk.w2;
                pk.pos1 = (pk.pos1 + pk.pos2) >> 1;
            }
            u = 6u * pk.pos1;
            thr_diff = u > 100u ? 100u : (u < 16u ? 16u : u);
            if (pk.w1 < 154) use_diff = 0;
        }
        const int share = 6 / (1 + use_flat + use_diff);
        w_lrt = share;
        w_flat = use_flat * share;
        w_diff = use_diff * share;
        WMX_NSX_PHASE_BEGIN
        uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
        for (int i = lane; i < 3 * kHistBins / 2; i += 32) hw[i] = 0u;
        WMX_NSX_PHASE_END
        // normalisation of the difference feature for the next window (nsx_core.c:1800-1835)
        cur_avg_e >>= 9;
        const uint32_t avg_e = (cur_avg_e + time_avg_e + 1) >> 1;
        if (avg_e != time_avg_e && feat_diff && time_avg_e > 0) {
            uint32_t a = avg_e, b = feat_diff;
            int n = 0;
            while (a & 0xFFFF0000u) { a >>= 1; ++n; }
            while (b & 0xFFFF0000u) { b >>= 1; ++n; }
            a = a * b;
            a /= time_avg_e;
            if (norm_u32(a) < n) feat_diff = 0x007FFFFFu;
            else feat_diff = umin(0x007FFFFFu, xshl(a, n));
        }
        time_avg_e = avg_e;
        cur_avg_e = 0;
    }

    // ---- speech / noise probability (nsx_core_c.c:26-260) ----
    WMX_NSX_PHASE_BEGIN
    int32_t lrt_sum = 0;
    WMX_NSX_BINS(k) {
        const uint32_t post = R.post[k], prior = R.prior[k];
        int32_t bessel = (int32_t)post;
        const int n = norm_u32(post);
        const uint32_t num = post << n;
        const uint32_t den = n > 10 ? xshl(prior, n - 11) : xshr(prior, 11 - n);
        bessel = den > 0 ? bessel - (int32_t)(num / (den ? den : 1u)) : 0;
        const int z = norm_u32(prior);
        int32_t frac32 = (int32_t)(((prior << z) & 0x7FFFFFFFu) >> 19);
        int32_t t = (frac32 * frac32 * -43) >> 19;
        t += ((int16_t)frac32 * 5412) >> 12;
        frac32 = t + 37;
        t = (int32_t)(((31 - z) << 12) + frac32) - (11 << 12);
        const int32_t lg = (t * 178) >> 8;
        uint32_t* lw = bin_word<ANA>(rec, A_LRT, k, lane);
        int32_t lrt = (int32_t)*lw;
        lrt = wadd(lrt, wsub(bessel, wadd(lg, lrt) / 2));
        *lw = (uint32_t)lrt;
        R.post[k] = (uint32_t)lrt;   // the posterior SNR is done with; keep the smoothed LRT for the second half
        lrt_sum = wadd(lrt_sum, lrt);
    }
    R.a[0] = (uint32_t)lrt_sum;
    WMX_NSX_PHASE_END
    int prior_ns = sc[S_PRIOR];
    {
        const int32_t lrt_sum = (int32_t)warp_add<0>(W);
        feat_lrt = wmul(lrt_sum, 10) >> (ST + 11);
        int32_t t = wsub(lrt_sum, thr_lrt);
        int shifts = 7 - ST;
        bool upper = true;
        if (t < 0) { upper = false; t = -t; ++shifts; }
        t = xshift(t, shifts);
        int32_t ind;
        {
            // this branch tests 0 <= index < 16 on the signed shifted value (nsx_core_c.c:98-100)
            const int idx = (int16_t)(t >> 14);
            int v = upper ? 16384 : 0;
            if (idx < 16 && idx >= 0) {
                const int d = (int16_t)(T.sigmoid[idx + 1] - T.sigmoid[idx]);
                const int y = (int16_t)(T.sigmoid[idx] + (int16_t)((d * (t & 0x3fff)) >> 14));
                v = (int16_t)(upper ? 8192 + y : 8192 - y);
            }
            ind = w_lrt * v;
        }
        if (w_flat) {
            const uint32_t f = feat_flat * 400u;
            uint32_t d = thr_flat - f;
            upper = true;
            shifts = 4;
            if (thr_flat < f) { upper = false; d = f - thr_flat; ++shifts; }
            ind += w_flat * sigmoid_q14((d << shifts) / 25u, upper, false, T);
        }
        if (w_diff) {
            uint32_t u1 = 0, u2, u3;
            if (feat_diff) {
                int n = norm_u32(feat_diff);
                if (n > 20 - ST) n = 20 - ST;
                u1 = feat_diff << n;
                u2 = time_avg_e >> (20 - ST - n);
                u1 = u2 > 0 ? u1 / u2 : 0x7fffffffu;
            }
            u3 = (thr_diff << 17) / 25u;
            u2 = u1 - u3;
            shifts = 1;
            upper = true;
            if (u2 & 0x80000000u) { upper = false; u2 = u3 - u1; --shifts; }
            ind += w_diff * sigmoid_q14(u2 >> shifts, upper, true, T);
        }
        const int ind16 = (int16_t)((98307 - ind) / 6);
        const int d16 = (int16_t)(ind16 - prior_ns);
        prior_ns = (int16_t)(prior_ns + (int16_t)((1638 * d16) >> 14));
    }
    WMX_NSX_PHASE_BEGIN
    WMX_NSX_BINS(k) {
        uint32_t ns = 0;
        const int32_t lrt = (int32_t)R.post[k];
        if (prior_ns > 0) {                                            // warp-uniform
            int32_t t = wmul(lrt, 23637) >> 14;
            int int_part = (int16_t)(t >> 12);
            if (int_part < -8) int_part = -8;
            const int frac = t & 0xfff;
            int32_t t2 = (frac * frac * 44) >> 19;
            t2 += (frac * 84) >> 7;
            int32_t inv_lrt = (int32_t)xshl(1u, 8 + int_part) + xshift(t2, int_part - 4);
            const int n1 = norm_w32(inv_lrt), n2 = norm_w16((int16_t)(16384 - prior_ns));
            const int32_t lo = xshift(wmul(xsar(inv_lrt, 15 - n2 - n1), 16384 - prior_ns), 7 - n1 - n2);
            const int32_t hi = wmul(inv_lrt, 16384 - prior_ns) >> 8;
            inv_lrt = n1 + n2 < 15 ? lo : hi;
            const int32_t den = prior_ns + inv_lrt;
            const uint32_t q = (uint32_t)((prior_ns << 8) / (den ? den : 1)) & 0xFFFFu;
            ns = (lrt < 65300 && n1 + n2 >= 7) ? q : 0u;                  // both paths of the reference leave 0 otherwise
        }
        R.nonsp[k] = ns;
        tbins[k < K ? 32 * k + lane : HALF] = ns;
    }
    WMX_NSX_PHASE_END

    // ---- step 2: noise update (nsx_core.c:1840-1946).  The update weight of bin i starts from the one bin i-1 chose ----
    WMX_NSX_PHASE_BEGIN
    const int post_shifts = q_noise_prev - q_magn;
    const int shifts = q_magn_prev - q_magn;
    uint32_t max_noise = 0;
    WMX_NSX_BINS(k) {
        const int bin = k < K ? 32 * k + lane : HALF;
        const uint32_t nprev = *bin_word<ANA>(rec, A_NPREV, k, lane);
        const uint32_t nprev16 = (nprev >> 11) & 0xFFFFu;
        const uint32_t nsp = R.nonsp[k];
        const uint32_t gamma_in = bin == 0 ? 26u : (tbins[bin - 1] < 205u ? 3u : 26u);
        const uint32_t gamma = nsp < 205u ? 3u : 26u;
        const uint32_t m = post_shifts < 0 ? xshr(R.magn[k], -post_shifts) : xshl(R.magn[k], post_shifts);
        const bool up = !(nprev16 > m);
        const uint32_t d = up ? m - nprev16 : nprev16 - m;
        uint32_t upd = nprev, w = 0, step;
        if (d && nsp) {
            w = d * nsp;
            step = (w & 0x7c000000u) ? (w >> 5) * gamma_in : (w * gamma_in) >> 5;
            upd = up ? upd + step : upd - step;
        }
        if (gamma_in != gamma) {
            step = (w & 0x7c000000u) ? (w >> 5) * gamma : (w * gamma) >> 5;
            const uint32_t alt = up ? nprev + step : nprev - step;
            if (upd > alt) upd = alt;
        }
        R.noise[k] = upd;
        max_noise = upd > max_noise ? upd : max_noise;
        uint32_t* pw = bin_word<ANA>(rec, A_PAUSE, k, lane);
        const int32_t pause_old = (int32_t)*pw;
        int32_t pause = xshift(pause_old, -shifts);
        if (nsp > 205u) {
            int32_t dp;
            if (shifts < 0) {
                dp = (int32_t)R.magn[k] - pause;
                dp = wmul(dp, 13);
                dp = wadd(dp, 128) >> 8;
            } else {
                dp = wsub((int32_t)xshl(R.magn[k], shifts), pause_old);
                dp = wmul(dp, 13);
                dp = xsar(wadd(dp, (int32_t)xshl(128u, shifts)), 8 + shifts);
            }
            pause = wadd(pause, dp);
        }
        *pw = (uint32_t)pause;
    }
    R.a[0] = max_noise;
    WMX_NSX_PHASE_END
    const uint32_t max_noise = warp_max_u<0>(W);
    const int norm_noise = norm_u32(max_noise);
    const int q_noise_new = (int16_t)(q_noise_prev + norm_noise - 5);

    // ---- step 3: Wiener gain from the updated noise (nsx_core.c:1948-2030), then the filtered spectrum (:456-474) ----
    WMX_NSX_PHASE_BEGIN
    const int shifts = q_noise_prev + 11 - q_magn;
    WMX_NSX_BINS(k) {
        const uint32_t magn = R.magn[k], noise = R.noise[k];
        uint32_t cur = 0, m, nz;
        if (shifts < 0) { m = magn; nz = xshl(noise, -shifts); }
        else if (shifts > 17) { m = magn << 17; nz = xshr(noise, shifts - 17); }
        else { m = xshl(magn, shifts); nz = noise; }
        {
            uint32_t a = m - nz;                                       // only used when m > nz
            int n = norm_u32(a);
            if (n > 11) n = 11;
            a <<= n;
            const uint32_t b = nz >> (11 - n);
            a /= (b ? b : 1u);                                         // b == 0: the reference skips the division
            cur = m > nz ? umin(a, kSatMax) : 0u;
        }
        const uint32_t prior = R.near_prev[k] * 2007u + cur * 41u;
        uint32_t a = (uint32_t)T.overdrive + ((prior + 8192u) >> 14);
        const uint32_t g = ((prior + a / 2) / a) & 0xFFFFu;
        uint32_t filt = g > 16384u ? 16384u : (g < (uint32_t)T.floor_gain ? (uint32_t)T.floor_gain : g);
        if (frame_idx < kStartupShort) {
            a = filt * (uint32_t)frame_idx;
            a += R.fmodel[k] * (uint32_t)(kStartupShort - frame_idx);
            filt = (a / (uint32_t)kStartupShort) & 0xFFFFu;
        }
        uint32_t* qf = bin_word<ANA>(rec, A_QF, k, lane);
        *qf = (*qf & 0xFFFFu) | (filt << 16);
        *bin_word<ANA>(rec, A_NPREV, k, lane) = norm_noise > 5 ? noise << (norm_noise - 5) : noise >> (5 - norm_noise);
        *bin_word<ANA>(rec, A_MPREV, k, lane) = magn;
        // PrepareSpectrum: both parts scaled by the filter, then the conjugate goes into the transform
        const int32_t re = (int16_t)((c_re(R.spec[k]) * (int32_t)(int16_t)filt) >> 14);
        const int32_t im = (int16_t)((c_im(R.spec[k]) * (int32_t)(int16_t)filt) >> 14);
        tbins[k < K ? 32 * k + lane : HALF] = c_pack(re, (int16_t)-im);
        if (HB) {
            // the second band's gain averages speech probability and filter over the upper quarter of the low band
            const int bin = k < K ? 32 * k + lane : HALF;
            if (k == 0) { R.a[0] = 0; R.a[1] = 0; }
            if (bin >= HALF - (HALF >> 2) && bin < HALF) { R.a[0] += R.nonsp[k]; R.a[1] += filt; }
        }
    }
    WMX_NSX_PHASE_END
    int hb_gain = 0;
    if (HB) {
        const uint32_t sum_prob = warp_add<0>(W) & 0xFFFFu, sum_filter = warp_add<1>(W);
        const int avg_prob = (int16_t)(4096 - (int)(sum_prob >> (ST - 7)));
        const int avg_filter = (int16_t)(sum_filter >> (ST - 3));
        const int gain_mod = avg_prob < 3607 ? avg_prob : 3607;
        int g;
        if (avg_prob < 2048) g = (int16_t)((gain_mod << 1) + (avg_filter >> 1));
        else g = (int16_t)((int16_t)((3 * avg_filter) >> 2) + gain_mod);
        hb_gain = g > 16384 ? 16384 : (g < (int16_t)T.floor_gain ? (int16_t)T.floor_gain : g);
    }

    // ---- scalars of the model are final: store them (lane 0) ----
    WMX_NSX_PHASE_BEGIN
    if (lane == 0) {
        sc[S_FRAME_IDX] = frame_idx; sc[S_MODEL_COUNT] = model_count; sc[S_Q_NOISE] = q_noise;
        sc[S_Q_NOISE_PREV] = q_noise_new; sc[S_Q_MAGN_PREV] = q_magn; sc[S_MIN_NORM] = min_norm; sc[S_PRIOR] = prior_ns;
        sc[S_FEAT_LRT] = feat_lrt; sc[S_THR_LRT] = thr_lrt; sc[S_FEAT_FLAT] = (int32_t)feat_flat; sc[S_THR_FLAT] = (int32_t)thr_flat;
        sc[S_FEAT_DIFF] = (int32_t)feat_diff; sc[S_THR_DIFF] = (int32_t)thr_diff; sc[S_W_LRT] = w_lrt; sc[S_W_FLAT] = w_flat;
        sc[S_W_DIFF] = w_diff; sc[S_CUR_AVG_E] = (int32_t)cur_avg_e; sc[S_TIME_AVG_E] = (int32_t)time_avg_e;
    }
    WMX_NSX_PHASE_END

    // ---- inverse transform: bins above ANA/2 by conjugate symmetry (real_fft.c:75-103), input in bit-reversed order ----
    WMX_NSX_PHASE_BEGIN
    uint32_t t[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int f = lane + 32 * r;   // frequency index held by register r before the renaming
        if (f <= HALF) {
            t[r] = tbins[f];
        } else {
            const uint32_t c = tbins[ANA - f];
            t[r] = c_pack(c_re(c), (int16_t)-c_im(c));
        }
    }
#pragma unroll
    for (int q = 0; q < NR; ++q) R.x[q] = t[rev_bits(q, G::kRB)];
    WMX_NSX_PHASE_END
    const int scale_ifft = fft_run<ANA, true>(W, tile, T);

    // ---- denormalise (nsx_core.c:477-487), output energy and the gain map (:1462-1495) ----
    WMX_NSX_PHASE_BEGIN
    int32_t smax2 = -1;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int32_t v = sat16(xshift(c_re(R.x[r]), scale_ifft - norm));
        R.x[r] = (uint32_t)v;
        const int32_t a16 = (int16_t)iabs(v);
        smax2 = a16 > smax2 ? a16 : smax2;
    }
    R.a[0] = (uint32_t)smax2;
    WMX_NSX_PHASE_END
    int gain = 8192;
    if (T.gain_map == 1 && frame_idx > kStartupLong && energy_in > 0) {
        const int32_t smax_out = warp_max_s<0>(W);
        int scale_out = 0;
        {
            const int nbits = size_in_bits((uint32_t)ANA);
            const int t = norm_w32(smax_out * smax_out);
            if (smax_out != 0) scale_out = t > nbits ? 0 : nbits - t;
        }
        WMX_NSX_PHASE_BEGIN
        uint32_t e = 0;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int32_t w = (int32_t)R.x[r];
            e += (uint32_t)((w * w) >> scale_out);
        }
        R.a[0] = e;
        WMX_NSX_PHASE_END
        int32_t e_out = (int32_t)warp_add<0>(W);
        if (scale_out == 0 && !(e_out & 0x7f800000)) e_out = xshift(e_out, 8 + scale_out - scale_in);
        else energy_in = xsar(energy_in, 8 + scale_out - scale_in);
        // the reference asserts energy_in > 0 here (nsx_core.c:1478): a zero would be a division fault there
        int ratio = energy_in ? (int16_t)((e_out + energy_in / 2) / energy_in) : 256;
        ratio = ratio > 256 ? 256 : (ratio < 0 ? 0 : ratio);
        const int g1 = T.factor1[ratio], g2 = T.factor2[ratio];
        gain = (int16_t)((int16_t)(((16384 - prior_ns) * g1) >> 14) + (int16_t)((prior_ns * g2) >> 14));
    }

    // ---- window, overlap-add, read-out (nsx_core.c:491-521) ----
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i = lane + 32 * r;
        const int32_t a = (int16_t)mul_round(T.window[i], (int32_t)R.x[r], 14);
        const int32_t b = sat16(mul_round(a, gain, 13));
        const int32_t old = i < G::kKeep ? syn16[i] : 0;
        R.x[r] = (uint32_t)(int32_t)sat16(old + b);
    }
    WMX_NSX_PHASE_END
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int i = lane + 32 * r;
        if (i < G::kBlock) out[i] = (int16_t)R.x[r];
        else syn16[i - G::kBlock] = (int16_t)R.x[r];
    }
    WMX_NSX_PHASE_END
    return hb_gain;
}

// wmix's stereo case: the right channel rides along as a second band — delayed by the analysis overlap (dataBufHBFX,
// nsx_core.c:2045-2053) and scaled by the frame's gain (:2111-2116); gain < 0 = unscaled (zero-input frame).
// hb: int16 [kKeep] history of the band; in_hb / out_hb: kBlock samples, may alias.
template <int ANA>
WMX_HD void second_band(Warp<ANA>& W, int16_t* hb, const int16_t* in_hb, int16_t* out_hb, int gain)
{
    typedef Geo<ANA> G;
    // the delay line holds the last kKeep samples: output sample i is hb[i] while i < kKeep, else in_hb[i - kKeep]
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < G::kR; ++r) {
        const int i = lane + 32 * r;
        int32_t v = 0;
        if (i < G::kBlock) v = i < G::kKeep ? hb[i] : in_hb[i - G::kKeep];
        else if (i < G::kBlock + G::kKeep) v = in_hb[i - G::kKeep];          // becomes the new history
        R.x[r] = (uint32_t)v;
    }
    WMX_NSX_PHASE_END
    WMX_NSX_PHASE_BEGIN
#pragma unroll
    for (int r = 0; r < G::kR; ++r) {
        const int i = lane + 32 * r;
        const int32_t v = (int32_t)R.x[r];
        if (i < G::kBlock) out_hb[i] = gain < 0 ? (int16_t)v : (int16_t)((gain * v) >> 14);
        else if (i < G::kBlock + G::kKeep) hb[i - G::kBlock] = (int16_t)v;
    }
    WMX_NSX_PHASE_END
}

// initial record (nsx_core.c:631-784); `lane`/`lanes` stride the words so a warp (or one host thread) can fill it
template <int ANA>
WMX_HD void init_record(uint32_t* rec, int16_t* hist, int lane, int lanes, int32_t thr_lrt)
{
    typedef Geo<ANA> G;
    for (int i = lane; i < G::kRecWords; i += lanes) {
        uint32_t v = 0;
        const int a = i < G::kOffNyq ? i / G::kHalf : (i < G::kOffScal ? i - G::kOffNyq : -1);
        if (a >= A_LQD0 && a <= A_LQD2) v = (uint32_t)pack16(2048, 153);
        else if (a == A_QF) v = 16384u << 16;
        if (i >= G::kOffScal && i < G::kOffHist) {
            switch (i - G::kOffScal) {
            case S_FRAME_IDX: v = (uint32_t)-1; break;
            case S_COUNTER0: v = 66; break;      // END_STARTUP_LONG * (e + 1) / SIMULT
            case S_COUNTER1: v = 133; break;
            case S_COUNTER2: v = 200; break;
            case S_MIN_NORM: v = 15; break;
            case S_PRIOR: v = 8192; break;
            case S_FEAT_LRT: case S_THR_LRT: v = (uint32_t)thr_lrt; break;
            case S_FEAT_FLAT: case S_THR_FLAT: v = 20480; break;
            case S_FEAT_DIFF: case S_THR_DIFF: v = 50; break;
            case S_W_LRT: v = 6; break;
            default: break;
            }
        }
        rec[i] = v;
    }
    uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
    for (int i = lane; i < 3 * kHistBins / 2; i += lanes) hw[i] = 0u;
}

}  // namespace nsx
}  // namespace wmx
