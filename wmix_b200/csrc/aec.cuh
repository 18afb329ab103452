// WebRTC float echo canceller (partitioned-block frequency-domain NLMS + coherence NLP), one
// stream per WARP, as wmix configures and drives it.
//
// What it computes (T: = R:pkg/webrtc_cut.tar.gz, webrtc/modules/audio_processing/aec/):
//   aec_process2 / aec_setFrameFar / aec_process     R:src/webrtc.c:286-483
//   WebRtcAec_BufferFarend, _Process, ProcessNormal, EstBufDelayNormal
//                                                    T:echo_cancellation.c:278-339, :341-409, :599-747, :821-872
//   WebRtcAec_ProcessFrames, ProcessBlock, FilterFar, ScaleErrorSignal, FilterAdaptation,
//   NonLinearProcessing (+SubbandCoherence, SmoothedPSD, OverdriveAndSuppress, ComfortNoise,
//   PartitionDelay), BufferFarendPartition, MoveFarReadPtr
//                                                    T:aec_core.c:148-546, :911-1340, :1690-1860
//   aec_rdft_forward_128 / inverse_128               T:aec_rdft.c:126-557
// for nlpMode = aggressive, no skew / metrics / delay logging, reported-delay mode, 12
// partitions, one band (R:src/webrtc.c:223-228); 8 kHz (mult 1) and 16 kHz (mult 2).
//
// How it is laid out for the GPU:
//   * per-stream state is one contiguous float record in HBM (Geo): the three 12-partition
//     spectrum histories (filter W, far X, windowed far Xw) in Ooura's PACKED real-FFT layout
//     (a[0] = Re0, a[1] = Re64, then re/im pairs), so a partition is exactly 128 floats = four
//     128-byte lines and every lane owns two packed complex elements: the hot 12-partition
//     loops (FilterFar, FilterAdaptation) have no odd Nyquist lane.
//   * the reference keeps TWO 250-deep rings of far-end SPECTRA (130 floats each per partition,
//     260 KB per stream).  Both rings are always moved together and a spectrum is a pure
//     function of 128 consecutive far samples, so here the far end is kept as ONE ring of
//     time-domain partitions (64 floats per partition, `depth` partitions) and the two
//     transforms run when a block is read (the same number of FFTs in steady state).  Read
//     pointer rewinds ("stuffing") re-read older partitions exactly as the reference re-reads
//     older spectra; the logical capacity stays 250 for every clamp, and a rewind or backlog
//     that would need more history than `depth` holds sets a sticky per-stream error flag.
//   * two independent 128-point real transforms run at once, one per half-warp (16 radix-4
//     butterflies each); butterfly operand order is the reference's and the twiddles are its
//     literal table, so spectra are bit-identical.
//   * every float sum the reference accumulates serially is accumulated serially by one lane.
//   * the body is a sequence of PHASES separated by warp barriers; lanes only communicate
//     through shared memory and the record, which makes the same source runnable by the
//     lane-loop emulator in tests/emu (host build).
// powf / cosf / sinf are evaluated in double and rounded to float (glibc's float routines are
// themselves double-based but not always correctly rounded): these only shape the OUTPUT
// (suppression gain, comfort noise), never the adaptive state, so a last-bit difference cannot
// accumulate.  Build with --fmad=false / -ffp-contract=off.
#pragma once
#include <math.h>
#include "common.cuh"
#include "ns.cuh"

namespace wmx {
namespace aec {

constexpr int kPart = 64, kPart2 = 128, kFrame = 80, kNPart = 12;
constexpr int kFarLogical = 250;   // capacity (partitions) of the reference's far rings, aec_core.c:37
constexpr int kOutRing = 144;      // FRAME_LEN + PART_LEN, aec_core.c:1352
constexpr int kTile = 160;         // padded exchange tile of one 64-point complex transform

enum BinArray { B_XPOW = 0, B_DPOW, B_DMIN, B_DINIT, B_SD, B_SE, B_SX, B_SDE_RE, B_SDE_IM, B_SXD_RE, B_SXD_IM, B_COUNT };

// record layout (float offsets), every section on a 128-byte line
struct Geo {
    static constexpr int kOffWf = 0;                              // [12][128] packed
    static constexpr int kOffXf = kNPart * kPart2;                // [12][128] packed, circular (S_XF_POS)
    static constexpr int kOffXfw = 2 * kNPart * kPart2;           // [12][128] packed, circular (S_XF_POS)
    static constexpr int kOffBins = 3 * kNPart * kPart2;          // [B_COUNT][64] bins 0..63
    static constexpr int kOffNyq = kOffBins + B_COUNT * kPart;    // [32] bin 64 of every bin array
    static constexpr int kOffDOld = kOffNyq + 32;                 // [64] previous near block
    static constexpr int kOffEOld = kOffDOld + kPart;             // [64] previous error block
    static constexpr int kOffOutB = kOffEOld + kPart;             // [64] overlap-add tail
    static constexpr int kOffNear = kOffOutB + kPart;             // [64] near samples not yet a block
    static constexpr int kOffOutRing = kOffNear + kPart;          // [160] output ring (144 used)
    static constexpr int kOffScal = kOffOutRing + 160;            // [64] scalar words
    static constexpr int kOffFar = kOffScal + 64;                 // [depth][64] far-end partitions
    static constexpr int kFixedFloats = kOffFar;
    // shared tile per warp (floats)
    static constexpr int kShSp = 0;                               // [7][128] spectra / time buffers
    static constexpr int kShX = 7 * kPart2;                       // [2][kTile] exchange tiles
    static constexpr int kShRow = kShX + 2 * kTile;               // [5][68] per-bin rows
    static constexpr int kShNear = kShRow + 5 * 68;               // [144] fifo ++ frame
    static constexpr int kShScal = kShNear + 144;                 // [96] scalar words + transients
    static constexpr int kShFloats = kShScal + 96;
};
WMX_HD int rec_floats(int depth) { return Geo::kFixedFloats + depth * kPart; }

enum Sp { SP_XF = 0, SP_XFW, SP_DF, SP_DFW, SP_EF, SP_EFW, SP_Y };
enum Row { ROW_HNL = 0, ROW_COHDE, ROW_COHXD, ROW_A, ROW_B };

enum ScalarId {
    // core
    S_KNOWN_DELAY = 0, S_DELAY_EST_CTR, S_DELAY_IDX, S_XF_POS, S_SYSTEM_DELAY, S_NOISE_CTR, S_NOISE_FROM_INIT,
    S_HNL_FB_MIN, S_HNL_FB_LOCAL_MIN, S_HNL_XD_AVG_MIN, S_OVER_DRIVE, S_OVER_DRIVE_SM,       // float
    S_HNL_NEW_MIN, S_HNL_MIN_CTR, S_ST_NEAR, S_ECHO_STATE, S_DIVERGE, S_SEED,
    // API layer
    S_BUF_SIZE_START, S_API_KNOWN_DELAY, S_TIME_FOR_CHANGE, S_STARTUP, S_CHECK_BUFF, S_SUM, S_COUNTER, S_FIRST_VAL,
    S_CHECK_CTR, S_MS_IN_SND, S_FILT_DELAY, S_LAST_DIFF,
    // buffers
    S_FAR_WT, S_FAR_PEND, S_FAR_RD, S_NEAR_CNT, S_OUT_RD, S_OUT_WR, S_ERROR,
    S_COUNT = 64,
    // warp-uniform transients (shared memory only)
    U_RUN = 64, U_NBLOCKS, U_RD_CUR, U_SD_SUM, U_SE_SUM, U_XD_AVG, U_DE_AVG, U_HNL_FB, U_HNL_FB_LOW, U_HNL_MODE,
    U_DIVERGE_NOW, U_RESET_W, U_SCAN
};
enum { ERR_FAR_DEPTH = 1, ERR_FAR_UNDERRUN = 2 };

struct Tables {
    ns::DMath dm;              // table-driven double log (ns.cuh)
    float w[32];               // rdft_w[0..31]   T:aec_rdft.c:32-40
    float c[32];               // rdft_w[32..63]  T:aec_rdft.c:41-49
    float hann[65];            // WebRtcAec_sqrtHanning   T:aec_core.c:49-66
    float weight[65];          // WebRtcAec_weightCurve   T:aec_core.c:71-81
    float over[65];            // WebRtcAec_overDriveCurve T:aec_core.c:86-96
    uint32_t lcg_mul[65];      // 69069^k mod 2^32 (k-step jump of WebRtcSpl_RandU's generator)
    uint32_t lcg_add[65];      // sum_{j<k} 69069^j mod 2^32
    uint32_t pad[2];
};

using ns::Cpx;
using ns::f2i;
using ns::i2f;

struct Lane {
    Cpx f[4];
    Cpx px[4], pw[4];   // NLMS update: far spectrum of the next round / filter taps of this round, fetched ahead of the FFTs
};
struct Warp {
    const float* pf_next = nullptr;   // record the warp will work on next: pulled towards L2 late in the tick (see block())
    uint32_t pf_bytes = 0;
#if defined(__CUDA_ARCH__)
    Lane lane_regs;
    int lane_id;
#else
    Lane lane_regs[32];
    int lane_id;        // unused by the emulation; keeps kernel bodies parseable in nvcc's host pass
#endif
};

#if defined(__CUDA_ARCH__)
#define WMX_AEC_PHASE_BEGIN { const int lane = W.lane_id; Lane& R = W.lane_regs; (void)lane; (void)R;
#define WMX_AEC_PHASE_END } __syncwarp();
#else
#define WMX_AEC_PHASE_BEGIN for (int lane = 0; lane < 32; ++lane) { Lane& R = W.lane_regs[lane]; (void)R;
#define WMX_AEC_PHASE_END }
#endif

// bins of a lane: slot 0,1 -> lane + 32*slot; slot 2 -> bin 64, lane 0 only
#define WMX_AEC_FOR_BINS(s, b)                                                \
    _Pragma("unroll") for (int s = 0; s < 3; ++s)                             \
        if (const int b = (s < 2 ? 32 * s + lane : kPart); s < 2 || lane == 0)

// bin b of a packed spectrum
WMX_HD void bin_get(const float* S, int b, float& re, float& im)
{
    if (b == 0) { re = S[0]; im = 0.f; }
    else if (b == kPart) { re = S[1]; im = 0.f; }
    else { re = S[2 * b]; im = S[2 * b + 1]; }
}
WMX_HD float* bin_ptr(float* rec, int arr, int b)
{
    return b < kPart ? rec + Geo::kOffBins + arr * kPart + b : rec + Geo::kOffNyq + arr;
}
WMX_HD float dpowf(float x, float y) { return (float)pow((double)x, (double)y); }
WMX_HD float dcosf(float x) { return (float)cos((double)x); }
WMX_HD float dsinf(float x) { return (float)sin((double)x); }
WMX_HD float fsqrt(float x)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return sqrtf(x);
#endif
}
WMX_HD int imod(int a, int m) { const int r = a % m; return r < 0 ? r + m : r; }

// ---------------------------------------------------------------------------------------------
// Two 128-point real transforms at once, in place: half-warp h works on buf[h] (time data ->
// packed spectrum for forward, packed spectrum -> unscaled time data for inverse) through its
// own exchange tile.  nh = 1 runs only half 0.
// ---------------------------------------------------------------------------------------------
WMX_HD int rev_pos(int c)   // where element c of the natural sequence sits after bitrv2 (pass-1 order)
{
    return 4 * ns::brev(c & 15, 4) + ns::brev(c >> 4, 2);
}

template <typename WarpT>
WMX_HD void passes64(WarpT& W, float* xt, const float* tw, bool back, int nh)
{
    // pass 1 (cft1st): butterfly l owns elements 4l..4l+3, already in R.f
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            float* xb = xt + h * kTile;
            ns::bfly4(R.f, ns::make_tw(tw, l));
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int p = ns::xpos(4 * l + q); xb[p] = R.f[q].r; xb[p + 1] = R.f[q].i; }
        }
    }
    WMX_AEC_PHASE_END
    // pass 2 (cftmdl, l = 8): group g = l/4 spans 16 elements, stride 4
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            float* xb = xt + h * kTile;
            const int g = l >> 2, q = l & 3;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = ns::xpos(16 * g + q + 4 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
            ns::bfly4(R.f, ns::make_tw(tw, g));
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = ns::xpos(16 * g + q + 4 * r); xb[p] = R.f[r].r; xb[p + 1] = R.f[r].i; }
        }
    }
    WMX_AEC_PHASE_END
    // last pass: untwiddled radix-4 on (c, c+16, c+32, c+48)
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            float* xb = xt + h * kTile;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = ns::xpos(l + 16 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
            ns::bfly4_last(R.f, back);
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = ns::xpos(l + 16 * r); xb[p] = R.f[r].r; xb[p + 1] = R.f[r].i; }
        }
    }
    WMX_AEC_PHASE_END
}

template <typename WarpT>
WMX_HD void rdft_fwd(WarpT& W, float* buf0, float* buf1, float* xt, const Tables& T, int nh)
{
    // bit-reversed gather straight from the time buffer
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            const float* t = h ? buf1 : buf0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = ns::gather_index<128>(l, q);
                R.f[q].r = t[2 * c];
                R.f[q].i = t[2 * c + 1];
            }
        }
    }
    WMX_AEC_PHASE_END
    passes64(W, xt, T.w, false, nh);
    // real split (rftfsub_128, aec_rdft.c:493-512) + a[0]/a[1] head (:539-546), packed output
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            const float* xb = xt + h * kTile;
            float* o = h ? buf1 : buf0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = l + 16 * r;
                float re, im;
                if (c == 0) {
                    const float a0 = xb[ns::xpos(0)], a1 = xb[ns::xpos(0) + 1];
                    re = a0 + a1;                                   // a[0] += a[1]
                    im = a0 - a1;                                   // a[1] = a[0] - a[1]
                } else if (c == 32) {
                    re = xb[ns::xpos(32)];
                    im = xb[ns::xpos(32) + 1];
                } else {
                    const bool low = c < 32;
                    const int cj = low ? c : 64 - c, ck = 64 - cj;
                    const float jr = xb[ns::xpos(cj)], ji = xb[ns::xpos(cj) + 1];
                    const float kr = xb[ns::xpos(ck)], ki = xb[ns::xpos(ck) + 1];
                    const float wkr = 0.5f - T.c[32 - cj], wki = T.c[cj];
                    const float xr = jr - kr, xi = ji + ki;
                    const float yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
                    if (low) { re = jr - yr; im = ji - yi; }
                    else { re = kr + yr; im = ki - yi; }
                }
                o[2 * c] = re;
                o[2 * c + 1] = im;
            }
        }
    }
    WMX_AEC_PHASE_END
}

template <typename WarpT>
WMX_HD void rdft_inv(WarpT& W, float* buf0, float* buf1, float* xt, const Tables& T, int nh)
{
    // head + rftbsub_128 (aec_rdft.c:514-537, :549-551), written where pass 1 will pick it up
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            const float* a = h ? buf1 : buf0;
            float* xb = xt + h * kTile;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = l + 16 * r;
                float vr, vi;
                if (c == 0) {
                    const float a0 = a[0], a1 = a[1];
                    const float h1 = 0.5f * (a0 - a1);
                    vr = a0 - h1;
                    vi = -h1;
                } else if (c == 32) {
                    vr = a[64];
                    vi = -a[65];
                } else {
                    const bool low = c < 32;
                    const int cj = low ? c : 64 - c, ck = 64 - cj;
                    const float jr = a[2 * cj], ji = a[2 * cj + 1];
                    const float kr = a[2 * ck], ki = a[2 * ck + 1];
                    const float wkr = 0.5f - T.c[32 - cj], wki = T.c[cj];
                    const float xr = jr - kr, xi = ji + ki;
                    const float yr = wkr * xr + wki * xi, yi = wkr * xi - wki * xr;
                    if (low) { vr = jr - yr; vi = yi - ji; }
                    else { vr = kr + yr; vi = yi - ki; }
                }
                const int p = ns::xpos(rev_pos(c));
                xb[p] = vr;
                xb[p + 1] = vi;
            }
        }
    }
    WMX_AEC_PHASE_END
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            const float* xb = xt + h * kTile;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int p = ns::xpos(4 * l + q); R.f[q].r = xb[p]; R.f[q].i = xb[p + 1]; }
        }
    }
    WMX_AEC_PHASE_END
    passes64(W, xt, T.w, true, nh);
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        if (h < nh) {
            const float* xb = xt + h * kTile;
            float* o = h ? buf1 : buf0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = l + 16 * r;
                o[2 * c] = xb[ns::xpos(c)];
                o[2 * c + 1] = xb[ns::xpos(c) + 1];
            }
        }
    }
    WMX_AEC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// far-end ring of time-domain partitions: logical read/write counters with the reference's
// clamps (ring_buffer.c:192-224 for a 250-element ring), physical depth `depth`
// ---------------------------------------------------------------------------------------------
WMX_HD int far_move(float* sc, int n, int depth)     // WebRtc_MoveReadPtr on both far rings; returns elements moved
{
    const int wt = f2i(sc[S_FAR_WT]);
    int rd = f2i(sc[S_FAR_RD]);
    const int avail = wt - rd, fr = kFarLogical - avail;
    if (n > avail) n = avail;
    if (n < -fr) n = -fr;
    rd += n;
    sc[S_FAR_RD] = i2f(rd);
    (void)depth;
    return n;
}
WMX_HD int far_move_delay(float* sc, int n, int depth)   // WebRtcAec_MoveFarReadPtr, aec_core.c:1709-1717
{
    const int moved = far_move(sc, n, depth);
    sc[S_SYSTEM_DELAY] = i2f(f2i(sc[S_SYSTEM_DELAY]) - moved * kPart);
    return moved;
}

// ---------------------------------------------------------------------------------------------
// one 64-sample block: ProcessBlock + NonLinearProcessing.  d_new = the block's near samples
// (shared memory).  Appends 64 output samples to the output ring.
// ---------------------------------------------------------------------------------------------
template <typename WarpT>
WMX_HD void block(WarpT& W, float* rec, int depth, int mult, const float* d_new, float* sh, const Tables& T)
{
    float* sp = sh + Geo::kShSp;
    float* xt = sh + Geo::kShX;
    float* rows = sh + Geo::kShRow;
    float* sc = sh + Geo::kShScal;
    float* far = rec + Geo::kOffFar;
    const float kScale = 2.0f / kPart2;

    // ---- B0 (lane 0): take the next far block, advance the circular history position ----
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        const int wt = f2i(sc[S_FAR_WT]), rd = f2i(sc[S_FAR_RD]);
        sc[U_RD_CUR] = i2f(rd);
        // the block at rd is built from partitions rd-1 and rd; the slot being filled is partition wt's
        if (wt - (rd - 1) > depth - 1) sc[S_ERROR] = i2f(f2i(sc[S_ERROR]) | ERR_FAR_DEPTH);
        if (wt - rd > 0) sc[S_FAR_RD] = i2f(rd + 1);
        else sc[S_ERROR] = i2f(f2i(sc[S_ERROR]) | ERR_FAR_UNDERRUN);   // the reference asserts this away
        int pos = f2i(sc[S_XF_POS]) - 1;
        if (pos == -1) pos = kNPart - 1;
        sc[S_XF_POS] = i2f(pos);
    }
    WMX_AEC_PHASE_END

    // ---- B1: far block [part(rd-1) | part(rd)], raw and sqrt-Hann windowed (TimeToFrequency :831) ----
    WMX_AEC_PHASE_BEGIN
    {
        const int rd = f2i(sc[U_RD_CUR]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane + 32 * r;
            const int part = i < kPart ? rd - 1 : rd;
            const float v = part < 0 ? 0.f : far[imod(part, depth) * kPart + (i & 63)];
            sp[SP_XF * kPart2 + i] = v;
            sp[SP_XFW * kPart2 + i] = v * T.hann[i < kPart ? i : kPart2 - i];
        }
    }
    WMX_AEC_PHASE_END
    rdft_fwd(W, sp + SP_XF * kPart2, sp + SP_XFW * kPart2, xt, T, 2);

    // ---- B2: spectra into the circular histories; near block [d_old | d_new] raw and windowed ----
    WMX_AEC_PHASE_BEGIN
    {
        const int pos = f2i(sc[S_XF_POS]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = lane + 32 * r;
            rec[Geo::kOffXf + pos * kPart2 + i] = sp[SP_XF * kPart2 + i];
            rec[Geo::kOffXfw + pos * kPart2 + i] = sp[SP_XFW * kPart2 + i];
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = lane + 32 * r;
            const float dold = rec[Geo::kOffDOld + i], dn = d_new[i];
            sp[SP_DF * kPart2 + i] = dold;
            sp[SP_DF * kPart2 + kPart + i] = dn;
            sp[SP_DFW * kPart2 + i] = dold * T.hann[i];
            sp[SP_DFW * kPart2 + kPart + i] = dn * T.hann[kPart - i];
            rec[Geo::kOffDOld + i] = dn;                              // aec_core.c:1128
        }
    }
    WMX_AEC_PHASE_END
    rdft_fwd(W, sp + SP_DF * kPart2, sp + SP_DFW * kPart2, xt, T, 2);

    // ---- B3: power smoothing and noise floor (aec_core.c:1195-1238) ----
    WMX_AEC_PHASE_BEGIN
    {
        const int noise_ctr = f2i(sc[S_NOISE_CTR]);
        const float g1n = 0.1f * (float)kNPart;
        WMX_AEC_FOR_BINS(s, b)
        {
            float xr, xi, dr, di;
            bin_get(sp + SP_XF * kPart2, b, xr, xi);
            bin_get(sp + SP_DF * kPart2, b, dr, di);
            const float far_spec = (xr * xr) + (xi * xi);
            const float near_spec = dr * dr + di * di;
            float* xp = bin_ptr(rec, B_XPOW, b);
            float* dp = bin_ptr(rec, B_DPOW, b);
            float* dm = bin_ptr(rec, B_DMIN, b);
            float* di0 = bin_ptr(rec, B_DINIT, b);
            *xp = 0.9f * *xp + g1n * far_spec;
            const float dpow = 0.9f * *dp + 0.1f * near_spec;
            *dp = dpow;
            float dmin = *dm;
            if (noise_ctr > 50) {
                if (dpow < dmin) dmin = (dpow + 0.1f * (dmin - dpow)) * 1.0002f;
                else dmin *= 1.0002f;
                *dm = dmin;
            }
            if (noise_ctr < 500 * mult) {
                const float dinit = *di0;
                *di0 = dmin > dinit ? 0.999f * dinit + 0.001f * dmin : dmin;
            }
        }
    }
    WMX_AEC_PHASE_END

    // ---- B4: Y = sum_p X_p W_p (FilterFar :148-170), packed; noise counter (lane 0) ----
    WMX_AEC_PHASE_BEGIN
    {
        const int pos = f2i(sc[S_XF_POS]);
        float yr[2] = {0.f, 0.f}, yi[2] = {0.f, 0.f};
        // the 12 partitions in two batches of 6: all 24 eight-byte loads of a batch are in flight before the first
        // multiply, so the warp pays two memory round trips for the whole filter instead of twelve
        for (int p0 = 0; p0 < kNPart; p0 += kNPart / 2) {
            Cpx xa[kNPart / 2][2], wa[kNPart / 2][2];
#pragma unroll
            for (int q = 0; q < kNPart / 2; ++q) {
                int xp = p0 + q + pos;
                if (xp >= kNPart) xp -= kNPart;
                const Cpx* X = reinterpret_cast<const Cpx*>(rec + Geo::kOffXf + xp * kPart2);
                const Cpx* Wf = reinterpret_cast<const Cpx*>(rec + Geo::kOffWf + (p0 + q) * kPart2);
#pragma unroll
                for (int s = 0; s < 2; ++s) { xa[q][s] = X[lane + 32 * s]; wa[q][s] = Wf[lane + 32 * s]; }
            }
#pragma unroll
            for (int q = 0; q < kNPart / 2; ++q) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int c = lane + 32 * s;
                    const float a0 = xa[q][s].r, a1 = xa[q][s].i, b0 = wa[q][s].r, b1 = wa[q][s].i;
                    if (c == 0) {            // bins 0 and 64: purely real operands (imaginary parts are +0)
                        yr[s] += a0 * b0 - 0.f * 0.f;
                        yi[s] += a1 * b1 - 0.f * 0.f;
                    } else {
                        yr[s] += a0 * b0 - a1 * b1;
                        yi[s] += a0 * b1 + a1 * b0;
                    }
                }
            }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int c = lane + 32 * s;
            sp[SP_Y * kPart2 + 2 * c] = yr[s];
            sp[SP_Y * kPart2 + 2 * c + 1] = yi[s];
        }
        if (lane == 0) {
            const int noise_ctr = f2i(sc[S_NOISE_CTR]);
            if (noise_ctr < 500 * mult) {
                sc[S_NOISE_CTR] = i2f(noise_ctr + 1);
                sc[S_NOISE_FROM_INIT] = i2f(1);
            } else {
                sc[S_NOISE_FROM_INIT] = i2f(0);
            }
        }
    }
    WMX_AEC_PHASE_END
    rdft_inv(W, sp + SP_Y * kPart2, nullptr, xt, T, 1);

    // ---- B5: e = d - y; error block zero-padded (for the NLMS) and windowed (for the NLP) ----
    WMX_AEC_PHASE_BEGIN
    {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = lane + 32 * r;
            const float y = sp[SP_Y * kPart2 + kPart + i] * kScale;
            const float e = d_new[i] - y;
            const float eold = rec[Geo::kOffEOld + i];
            sp[SP_EF * kPart2 + i] = 0.f;
            sp[SP_EF * kPart2 + kPart + i] = e;
            sp[SP_EFW * kPart2 + i] = eold * T.hann[i];
            sp[SP_EFW * kPart2 + kPart + i] = e * T.hann[kPart - i];
            rec[Geo::kOffEOld + i] = e;                               // aec_core.c:1129
        }
    }
    WMX_AEC_PHASE_END
    rdft_fwd(W, sp + SP_EF * kPart2, sp + SP_EFW * kPart2, xt, T, 2);

    // ---- B6: ScaleErrorSignal (:172-194), in place on the packed error spectrum ----
    WMX_AEC_PHASE_BEGIN
    {
        const float mu = mult == 1 ? 0.6f : 0.5f;
        const float thr = mult == 1 ? 2e-6f : 1.5e-6f;
        float* E = sp + SP_EF * kPart2;
        WMX_AEC_FOR_BINS(s, b)
        {
            float er, ei;
            bin_get(E, b, er, ei);
            const float xp = *bin_ptr(rec, B_XPOW, b);
            er /= (xp + 1e-10f);
            ei /= (xp + 1e-10f);
            float mag = fsqrt(er * er + ei * ei);
            if (mag > thr) {
                mag = thr / (mag + 1e-10f);
                er *= mag;
                ei *= mag;
            }
            er *= mu;
            ei *= mu;
            if (b == 0) E[0] = er;
            else if (b == kPart) E[1] = er;
            else { E[2 * b] = er; E[2 * b + 1] = ei; }
        }
    }
    WMX_AEC_PHASE_END

    // ---- B7: constrained NLMS update of the 12 partitions, two per round (FilterAdaptation :222-270) ----
    // Global operands are fetched a stage ahead: the far spectrum of round r+1 and the filter taps of round r are
    // requested before round r's two transforms, so their latency hides behind the FFTs instead of stalling the warp
    // twice per round.
    WMX_AEC_PHASE_BEGIN
    {
        const int h = lane >> 4, l = lane & 15;
        int xp = h + f2i(sc[S_XF_POS]);
        if (xp >= kNPart) xp -= kNPart;
        const Cpx* X = reinterpret_cast<const Cpx*>(rec + Geo::kOffXf + xp * kPart2);
#pragma unroll
        for (int r = 0; r < 4; ++r) R.px[r] = X[l + 16 * r];
    }
    WMX_AEC_PHASE_END
    for (int rnd = 0; rnd < kNPart / 2; ++rnd) {
        float* g0 = sp + SP_XF * kPart2;
        float* g1 = sp + SP_DF * kPart2;
        WMX_AEC_PHASE_BEGIN
        {
            const int h = lane >> 4, l = lane & 15;
            const int p = 2 * rnd + h;
            const float* E = sp + SP_EF * kPart2;
            float* G = h ? g1 : g0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = l + 16 * r;
                const float xr = R.px[r].r, x1 = R.px[r].i, er = E[2 * c], e1 = E[2 * c + 1];
                if (c == 0) {
                    // bin 0 and bin 64: conj(x) e with x = (xr, +0), e = (er, +0): xr*er - (-0)*(+0)
                    G[0] = xr * er - (-0.f) * 0.f;
                    G[1] = x1 * e1 - (-0.f) * 0.f;
                } else {
                    const float xi = -x1;
                    G[2 * c] = xr * er - xi * e1;
                    G[2 * c + 1] = xr * e1 + xi * er;
                }
            }
            const Cpx* Wf = reinterpret_cast<const Cpx*>(rec + Geo::kOffWf + p * kPart2);
#pragma unroll
            for (int r = 0; r < 4; ++r) R.pw[r] = Wf[l + 16 * r];
            if (rnd + 1 < kNPart / 2) {
                int xp = p + 2 + f2i(sc[S_XF_POS]);
                if (xp >= kNPart) xp -= kNPart;
                const Cpx* X = reinterpret_cast<const Cpx*>(rec + Geo::kOffXf + xp * kPart2);
#pragma unroll
                for (int r = 0; r < 4; ++r) R.px[r] = X[l + 16 * r];
            }
        }
        WMX_AEC_PHASE_END
        rdft_inv(W, g0, g1, xt, T, 2);
        WMX_AEC_PHASE_BEGIN
        {
            const int h = lane >> 4, l = lane & 15;
            float* G = h ? g1 : g0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = l + 16 * r;
                G[i] *= kScale;
                G[kPart + i] = 0.f;
            }
        }
        WMX_AEC_PHASE_END
        rdft_fwd(W, g0, g1, xt, T, 2);
        WMX_AEC_PHASE_BEGIN
        {
            const int h = lane >> 4, l = lane & 15;
            const int p = 2 * rnd + h;
            const float* G = h ? g1 : g0;
            Cpx* Wf = reinterpret_cast<Cpx*>(rec + Geo::kOffWf + p * kPart2);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int c = l + 16 * r;
                Cpx w = R.pw[r];
                w.r += G[2 * c];
                w.i += G[2 * c + 1];
                Wf[c] = w;
            }
        }
        WMX_AEC_PHASE_END
    }

    // The NLMS half of the block is done: ask for the record this warp turns to next.  Requested earlier (a whole tick
    // ahead) the in-flight records of all resident warps outgrow the L2 and are evicted before they are used.
    WMX_AEC_PHASE_BEGIN
    if (lane == 0 && W.pf_next) l2_prefetch_bulk(W.pf_next, W.pf_bytes);
    WMX_AEC_PHASE_END

    // ======================= NonLinearProcessing (:911-1141) =======================
    // ---- N0 (lane 0): delay-estimation counter ----
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        int ctr = f2i(sc[S_DELAY_EST_CTR]) + 1;
        if (ctr == 10 * mult) ctr = 0;
        sc[S_DELAY_EST_CTR] = i2f(ctr);
        sc[U_SCAN] = i2f(ctr == 0);
    }
    WMX_AEC_PHASE_END
    if (f2i(sc[U_SCAN])) {
        // PartitionDelay (:295-321): lane p sums partition p's energy in bin order
        WMX_AEC_PHASE_BEGIN
        if (lane < kNPart) {
            const float* Wf = rec + Geo::kOffWf + lane * kPart2;
            float en = 0.f;
            en += Wf[0] * Wf[0] + 0.f * 0.f;
            for (int j = 1; j < kPart; ++j) en += Wf[2 * j] * Wf[2 * j] + Wf[2 * j + 1] * Wf[2 * j + 1];
            en += Wf[1] * Wf[1] + 0.f * 0.f;
            rows[ROW_A * 68 + lane] = en;
        }
        WMX_AEC_PHASE_END
        WMX_AEC_PHASE_BEGIN
        if (lane == 0) {
            float best = 0.f;
            int delay = 0;
            for (int p = 0; p < kNPart; ++p)
                if (rows[ROW_A * 68 + p] > best) { best = rows[ROW_A * 68 + p]; delay = p; }
            sc[S_DELAY_IDX] = i2f(delay);
        }
        WMX_AEC_PHASE_END
    }

    // ---- N1: smoothed PSDs (SmoothedPSD :333-395) ----
    WMX_AEC_PHASE_BEGIN
    {
        const float g0c = mult == 1 ? 0.9f : 0.93f, g1c = mult == 1 ? 0.1f : 0.07f;
        int xp = f2i(sc[S_DELAY_IDX]) + f2i(sc[S_XF_POS]);
        if (xp >= kNPart) xp -= kNPart;
        const float* Xw = rec + Geo::kOffXfw + xp * kPart2;
        WMX_AEC_FOR_BINS(s, b)
        {
            float dr, di, er, ei, xr, xi;
            bin_get(sp + SP_DFW * kPart2, b, dr, di);
            bin_get(sp + SP_EFW * kPart2, b, er, ei);
            bin_get(Xw, b, xr, xi);
            float* psd = bin_ptr(rec, B_SD, b);
            float* pse = bin_ptr(rec, B_SE, b);
            float* psx = bin_ptr(rec, B_SX, b);
            const float sd = g0c * *psd + g1c * (dr * dr + di * di);
            const float se = g0c * *pse + g1c * (er * er + ei * ei);
            const float px = xr * xr + xi * xi;
            *psd = sd;
            *pse = se;
            *psx = g0c * *psx + g1c * (px > 15.f ? px : 15.f);
            float* q = bin_ptr(rec, B_SDE_RE, b);
            *q = g0c * *q + g1c * (dr * er + di * ei);
            q = bin_ptr(rec, B_SDE_IM, b);
            *q = g0c * *q + g1c * (dr * ei - di * er);
            q = bin_ptr(rec, B_SXD_RE, b);
            *q = g0c * *q + g1c * (dr * xr + di * xi);
            q = bin_ptr(rec, B_SXD_IM, b);
            *q = g0c * *q + g1c * (dr * xi - di * xr);
            rows[ROW_A * 68 + b] = sd;
            rows[ROW_B * 68 + b] = se;
        }
    }
    WMX_AEC_PHASE_END
    WMX_AEC_PHASE_BEGIN
    if (lane < 2) {
        const float* row = rows + (lane == 0 ? ROW_A : ROW_B) * 68;
        float acc = 0.f;
        for (int i = 0; i <= kPart; ++i) acc += row[i];
        sc[lane == 0 ? U_SD_SUM : U_SE_SUM] = acc;
    }
    WMX_AEC_PHASE_END
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        const float sd_sum = sc[U_SD_SUM], se_sum = sc[U_SE_SUM];
        const int div = (f2i(sc[S_DIVERGE]) ? 1.05f : 1.0f) * se_sum > sd_sum;
        sc[S_DIVERGE] = i2f(div);
        sc[U_DIVERGE_NOW] = i2f(div);
        sc[U_RESET_W] = i2f(se_sum > (19.95f * sd_sum));
    }
    WMX_AEC_PHASE_END

    // ---- N2: divergence guard, coherences (:386-395, :441-449) ----
    WMX_AEC_PHASE_BEGIN
    {
        if (f2i(sc[U_DIVERGE_NOW])) {
#pragma unroll
            for (int r = 0; r < 4; ++r) sp[SP_EFW * kPart2 + lane + 32 * r] = sp[SP_DFW * kPart2 + lane + 32 * r];
        }
        if (f2i(sc[U_RESET_W]))
            for (int i = lane; i < kNPart * kPart2; i += 32) rec[Geo::kOffWf + i] = 0.f;
        WMX_AEC_FOR_BINS(s, b)
        {
            const float sd = *bin_ptr(rec, B_SD, b), se = *bin_ptr(rec, B_SE, b), sx = *bin_ptr(rec, B_SX, b);
            const float a0 = *bin_ptr(rec, B_SDE_RE, b), a1 = *bin_ptr(rec, B_SDE_IM, b);
            const float c0 = *bin_ptr(rec, B_SXD_RE, b), c1 = *bin_ptr(rec, B_SXD_IM, b);
            rows[ROW_COHDE * 68 + b] = (a0 * a0 + a1 * a1) / (sd * se + 1e-10f);
            rows[ROW_COHXD * 68 + b] = (c0 * c0 + c1 * c1) / (sx * sd + 1e-10f);
        }
    }
    WMX_AEC_PHASE_END
    // preferred-band averages, in band order (:963-973)
    WMX_AEC_PHASE_BEGIN
    if (lane < 2) {
        const int pref_size = 24 / mult, pref_min = 4 / mult;
        const float* row = rows + (lane == 0 ? ROW_COHXD : ROW_COHDE) * 68;
        float acc = 0.f;
        for (int i = pref_min; i < pref_size + pref_min; ++i) acc += row[i];
        acc /= pref_size;
        if (lane == 0) acc = 1 - acc;
        sc[lane == 0 ? U_XD_AVG : U_DE_AVG] = acc;
    }
    WMX_AEC_PHASE_END
    // near/echo state machine (:975-1013)
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        const float xd = sc[U_XD_AVG], de = sc[U_DE_AVG];
        float xd_min = sc[S_HNL_XD_AVG_MIN];
        int st_near = f2i(sc[S_ST_NEAR]);
        if (xd < 0.75f && xd < xd_min) xd_min = xd;
        if (de > 0.98f && xd > 0.9f) st_near = 1;
        else if (de < 0.95f || xd < 0.8f) st_near = 0;
        int mode;      // 0: hNl = cohde, 1: hNl = 1 - cohxd, 2: min of both + order statistic
        if (xd_min == 1) {
            sc[S_ECHO_STATE] = i2f(0);
            sc[S_OVER_DRIVE] = 5.0f;                                  // kNormalMinOverDrive[aggressive]
            mode = st_near == 1 ? 0 : 1;
        } else if (st_near == 1) {
            sc[S_ECHO_STATE] = i2f(0);
            mode = 0;
        } else {
            sc[S_ECHO_STATE] = i2f(1);
            mode = 2;
        }
        sc[S_HNL_XD_AVG_MIN] = xd_min;
        sc[S_ST_NEAR] = i2f(st_near);
        sc[U_HNL_MODE] = i2f(mode);
        sc[U_HNL_FB] = mode == 0 ? de : xd;
        sc[U_HNL_FB_LOW] = mode == 0 ? de : xd;
    }
    WMX_AEC_PHASE_END
    WMX_AEC_PHASE_BEGIN
    {
        const int mode = f2i(sc[U_HNL_MODE]);
        WMX_AEC_FOR_BINS(s, b)
        {
            const float cde = rows[ROW_COHDE * 68 + b], alt = 1 - rows[ROW_COHXD * 68 + b];
            rows[ROW_HNL * 68 + b] = mode == 0 ? cde : (mode == 1 ? alt : (cde < alt ? cde : alt));
        }
    }
    WMX_AEC_PHASE_END
    if (f2i(sc[U_HNL_MODE]) == 2) {
        // order statistics of the preferred bands (the reference qsorts a copy, :1003-1010): the value
        // of sorted rank k is the element that has exactly k elements ordered before it
        WMX_AEC_PHASE_BEGIN
        {
            const int pref_size = 24 / mult, pref_min = 4 / mult;
            if (lane < pref_size) {
                const float v = rows[ROW_HNL * 68 + pref_min + lane];
                int rank = 0;
                for (int j = 0; j < pref_size; ++j) {
                    const float u = rows[ROW_HNL * 68 + pref_min + j];
                    rank += (u < v) || (u == v && j < lane);
                }
                if (rank == (int)floor(0.75f * (pref_size - 1))) sc[U_HNL_FB] = v;
                if (rank == (int)floor(0.5f * (pref_size - 1))) sc[U_HNL_FB_LOW] = v;
            }
        }
        WMX_AEC_PHASE_END
    }
    // overdrive tracking (:1015-1047)
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        const float fb_low = sc[U_HNL_FB_LOW];
        float local_min = sc[S_HNL_FB_LOCAL_MIN];
        int new_min = f2i(sc[S_HNL_NEW_MIN]), min_ctr = f2i(sc[S_HNL_MIN_CTR]);
        if (fb_low < 0.6f && fb_low < local_min) {
            local_min = fb_low;
            sc[S_HNL_FB_MIN] = fb_low;
            new_min = 1;
            min_ctr = 0;
        }
        float v = local_min + 0.0008f / mult;
        sc[S_HNL_FB_LOCAL_MIN] = v < 1 ? v : 1;
        v = sc[S_HNL_XD_AVG_MIN] + 0.0006f / mult;
        sc[S_HNL_XD_AVG_MIN] = v < 1 ? v : 1;
        if (new_min == 1) min_ctr++;
        float over = sc[S_OVER_DRIVE];
        if (min_ctr == 2) {
            new_min = 0;
            min_ctr = 0;
            const float cand = -18.4f / (ns::log_f(sc[S_HNL_FB_MIN] + 1e-10f, T.dm) + 1e-10f);   // kTargetSupp[aggressive]
            over = cand > 5.0f ? cand : 5.0f;
        }
        sc[S_HNL_NEW_MIN] = i2f(new_min);
        sc[S_HNL_MIN_CTR] = i2f(min_ctr);
        sc[S_OVER_DRIVE] = over;
        float sm = sc[S_OVER_DRIVE_SM];
        if (over < sm) sm = 0.99f * sm + 0.01f * over;
        else sm = 0.9f * sm + 0.1f * over;
        sc[S_OVER_DRIVE_SM] = sm;
    }
    WMX_AEC_PHASE_END

    // ---- N3: OverdriveAndSuppress (:272-293) + ComfortNoise (:462-546), into the inverse's input ----
    WMX_AEC_PHASE_BEGIN
    {
        const float fb = sc[U_HNL_FB], sm = sc[S_OVER_DRIVE_SM];
        const uint32_t seed = (uint32_t)f2i(sc[S_SEED]);
        const int from_init = f2i(sc[S_NOISE_FROM_INIT]);
        const float* E = sp + SP_EFW * kPart2;
        float* O = sp + SP_Y * kPart2;
        WMX_AEC_FOR_BINS(s, b)
        {
            float hnl = rows[ROW_HNL * 68 + b];
            if (hnl > fb) hnl = T.weight[b] * fb + (1 - T.weight[b]) * hnl;
            hnl = dpowf(hnl, sm * T.over[b]);
            float er, ei;
            bin_get(E, b, er, ei);
            er *= hnl;
            ei *= hnl;
            ei *= -1;
            float ur = 0.f, ui = 0.f;
            if (b >= 1) {
                const uint32_t sd = (T.lcg_mul[b] * seed + T.lcg_add[b]) & 0x7fffffffu;   // b-th draw
                const float rnd = ((float)(int16_t)(sd >> 16)) / 32768;
                const float ang = 6.28318530717959f * rnd;
                const float amp = fsqrt(*bin_ptr(rec, from_init ? B_DINIT : B_DMIN, b));
                ur = amp * dcosf(ang);
                ui = b == kPart ? 0.f : -amp * dsinf(ang);
            }
            const float rest = 1 - hnl * hnl;
            const float wgt = fsqrt(rest > 0 ? rest : 0);
            er += wgt * ur;
            ei += wgt * ui;
            if (b == 0) O[0] = er;
            else if (b == kPart) O[1] = er;
            else { O[2 * b] = er; O[2 * b + 1] = -ei; }
        }
    }
    WMX_AEC_PHASE_END
    rdft_inv(W, sp + SP_Y * kPart2, nullptr, xt, T, 1);

    // ---- N4: window, overlap-add, saturate, append to the output ring (:1066-1088, :1325-1327) ----
    WMX_AEC_PHASE_BEGIN
    {
        const int wr = f2i(sc[S_OUT_WR]);
        const float* t = sp + SP_Y * kPart2;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int i = lane + 32 * r;
            float v = t[i] * kScale;
            v = v * T.hann[i] + rec[Geo::kOffOutB + i];
            const float tail = t[kPart + i] * kScale;
            rec[Geo::kOffOutB + i] = tail * T.hann[kPart - i];
            v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
            rec[Geo::kOffOutRing + imod(wr + i, kOutRing)] = v;
        }
    }
    WMX_AEC_PHASE_END
    WMX_AEC_PHASE_BEGIN
    if (lane == 0) {
        sc[S_OUT_WR] = i2f(f2i(sc[S_OUT_WR]) + kPart);
        const uint32_t seed = (uint32_t)f2i(sc[S_SEED]);              // 64 draws were consumed by N3
        sc[S_SEED] = i2f((int32_t)((T.lcg_mul[kPart] * seed + T.lcg_add[kPart]) & 0x7fffffffu));
    }
    WMX_AEC_PHASE_END
}

// ---------------------------------------------------------------------------------------------
// One call of the wmix wrapper for one stream: optional BufferFarend(far, n), optional
// Process(near, n) -> out.  n = 80 or 160 samples; mult = 1 (8 kHz) or 2 (16 kHz).
// ---------------------------------------------------------------------------------------------
template <typename WarpT>
WMX_HD void tick(WarpT& W, float* rec, int depth, int mult, int n, const int16_t* far_in, const int16_t* near_in,
                 int16_t* out, int delay_ms, float* sh, const Tables& T)
{
    float* sc = sh + Geo::kShScal;
    float* vn = sh + Geo::kShNear;

    WMX_AEC_PHASE_BEGIN
    for (int i = lane; i < S_COUNT; i += 32) sc[i] = rec[Geo::kOffScal + i];
    WMX_AEC_PHASE_END

    if (far_in) {
        // WebRtcAec_BufferFarend (echo_cancellation.c:278-339): the far-end stream, cut in 64-sample partitions
        WMX_AEC_PHASE_BEGIN
        {
            const int wt = f2i(sc[S_FAR_WT]), pend = f2i(sc[S_FAR_PEND]);
            for (int i = lane; i < n; i += 32) {
                const int pos = pend + i;
                rec[Geo::kOffFar + imod(wt + (pos >> 6), depth) * kPart + (pos & 63)] = (float)far_in[i];
            }
        }
        WMX_AEC_PHASE_END
        WMX_AEC_PHASE_BEGIN
        if (lane == 0) {
            int wt = f2i(sc[S_FAR_WT]), pend = f2i(sc[S_FAR_PEND]) + n;
            sc[S_SYSTEM_DELAY] = i2f(f2i(sc[S_SYSTEM_DELAY]) + n);
            while (pend >= kPart) {
                // BufferFarendPartition (aec_core.c:1690-1707): flush the oldest block when the ring is full
                if (wt - f2i(sc[S_FAR_RD]) >= kFarLogical) far_move_delay(sc, 1, depth);
                wt++;
                sc[S_FAR_WT] = i2f(wt);
                pend -= kPart;
            }
            sc[S_FAR_PEND] = i2f(pend);
        }
        WMX_AEC_PHASE_END
    }

    if (near_in) {
        // WebRtcAec_Process + ProcessNormal scalars (echo_cancellation.c:341-409, :599-747)
        WMX_AEC_PHASE_BEGIN
        if (lane == 0) {
            int ms = delay_ms < 0 ? 0 : delay_ms;
            ms = ms > 500 ? 500 : ms;
            ms += 10;
            sc[S_MS_IN_SND] = i2f(ms);
            const int blocks10 = n / (kFrame * mult);
            if (f2i(sc[S_STARTUP])) {
                sc[U_RUN] = i2f(0);
                if (f2i(sc[S_CHECK_BUFF])) {
                    int check_ctr = f2i(sc[S_CHECK_CTR]) + 1, counter = f2i(sc[S_COUNTER]), sum = f2i(sc[S_SUM]);
                    if (counter == 0) {
                        sc[S_FIRST_VAL] = i2f(ms);
                        sum = 0;
                    }
                    const int first = f2i(sc[S_FIRST_VAL]);
                    const double lim = 0.2 * ms > 8 ? 0.2 * ms : 8;
                    const int dist = first - ms < 0 ? ms - first : first - ms;
                    if (dist < lim) { sum += ms; counter++; }
                    else counter = 0;
                    if (counter * blocks10 >= 6) {
                        const int v = (3 * sum * mult * 8) / (4 * counter * kPart);
                        sc[S_BUF_SIZE_START] = i2f(v < 62 ? v : 62);
                        sc[S_CHECK_BUFF] = i2f(0);
                    }
                    if (check_ctr * blocks10 > 50) {
                        const int v = (ms * mult * 3) / 40;
                        sc[S_BUF_SIZE_START] = i2f(v < 62 ? v : 62);
                        sc[S_CHECK_BUFF] = i2f(0);
                    }
                    sc[S_CHECK_CTR] = i2f(check_ctr);
                    sc[S_COUNTER] = i2f(counter);
                    sc[S_SUM] = i2f(sum);
                }
                if (!f2i(sc[S_CHECK_BUFF])) {
                    const int overhead = f2i(sc[S_SYSTEM_DELAY]) / kPart - f2i(sc[S_BUF_SIZE_START]);
                    if (overhead == 0) {
                        sc[S_STARTUP] = i2f(0);
                    } else if (overhead > 0) {
                        far_move_delay(sc, overhead, depth);
                        sc[S_STARTUP] = i2f(0);
                    }
                }
            } else {
                sc[U_RUN] = i2f(1);
                // EstBufDelayNormal (:821-872)
                int cur = ms * 8 * mult - f2i(sc[S_SYSTEM_DELAY]);
                cur += kFrame * mult;
                if (cur < kPart) cur += far_move_delay(sc, 1, depth) * kPart;
                int filt = f2i(sc[S_FILT_DELAY]);
                filt = filt < 0 ? 0 : filt;
                const short f = (short)(0.8 * filt + 0.2 * cur);
                filt = f > 0 ? f : 0;
                sc[S_FILT_DELAY] = i2f(filt);
                int known = f2i(sc[S_API_KNOWN_DELAY]), tfc = f2i(sc[S_TIME_FOR_CHANGE]);
                const int last = f2i(sc[S_LAST_DIFF]);
                const int diff = filt - known;
                if (diff > 224) tfc = last < 96 ? 0 : tfc + 1;
                else if (diff < 96 && known > 0) tfc = last > 224 ? 0 : tfc + 1;
                else tfc = 0;
                sc[S_LAST_DIFF] = i2f((int)(short)diff);
                if (tfc > 25) {
                    const int v = filt - 160;
                    known = v > 0 ? v : 0;
                }
                sc[S_TIME_FOR_CHANGE] = i2f(tfc);
                sc[S_API_KNOWN_DELAY] = i2f(known);
            }
        }
        WMX_AEC_PHASE_END

        if (!f2i(sc[U_RUN])) {
            // start-up: the near end passes through (echo_cancellation.c:646-652)
            WMX_AEC_PHASE_BEGIN
            for (int i = lane; i < n; i += 32) out[i] = near_in[i];
            WMX_AEC_PHASE_END
        } else {
            // WebRtcAec_ProcessFrames (aec_core.c:1719-1860), reported-delay branch
            for (int j = 0; j < n; j += kFrame) {
                WMX_AEC_PHASE_BEGIN
                {
                    const int cnt = f2i(sc[S_NEAR_CNT]);
                    for (int i = lane; i < cnt; i += 32) vn[i] = rec[Geo::kOffNear + i];
                    for (int i = lane; i < kFrame; i += 32) vn[cnt + i] = (float)near_in[j + i];
                }
                WMX_AEC_PHASE_END
                WMX_AEC_PHASE_BEGIN
                if (lane == 0) {
                    if (f2i(sc[S_SYSTEM_DELAY]) < kFrame) far_move_delay(sc, -(mult + 1), depth);
                    const int want = (f2i(sc[S_KNOWN_DELAY]) - f2i(sc[S_API_KNOWN_DELAY]) - 32) / kPart;
                    const int moved = far_move(sc, want, depth);
                    sc[S_KNOWN_DELAY] = i2f(f2i(sc[S_KNOWN_DELAY]) - moved * kPart);
                    sc[U_NBLOCKS] = i2f((f2i(sc[S_NEAR_CNT]) + kFrame) / kPart);
                }
                WMX_AEC_PHASE_END
                const int nb = f2i(sc[U_NBLOCKS]);
                for (int k = 0; k < nb; ++k) block(W, rec, depth, mult, vn + k * kPart, sh, T);
                WMX_AEC_PHASE_BEGIN
                {
                    const int total = f2i(sc[S_NEAR_CNT]) + kFrame, left = total - nb * kPart;
                    for (int i = lane; i < left; i += 32) rec[Geo::kOffNear + i] = vn[nb * kPart + i];
                }
                WMX_AEC_PHASE_END
                WMX_AEC_PHASE_BEGIN
                if (lane == 0) {
                    sc[S_NEAR_CNT] = i2f(f2i(sc[S_NEAR_CNT]) + kFrame - nb * kPart);
                    sc[S_SYSTEM_DELAY] = i2f(f2i(sc[S_SYSTEM_DELAY]) - kFrame);
                    int rd = f2i(sc[S_OUT_RD]), wr = f2i(sc[S_OUT_WR]);
                    const int have = wr - rd;
                    if (have < kFrame) rd += have - kFrame;          // stuff: re-expose what was read before
                    sc[S_OUT_RD] = i2f(rd);
                }
                WMX_AEC_PHASE_END
                WMX_AEC_PHASE_BEGIN
                {
                    const int rd = f2i(sc[S_OUT_RD]);
                    for (int i = lane; i < kFrame; i += 32) out[j + i] = (int16_t)rec[Geo::kOffOutRing + imod(rd + i, kOutRing)];
                }
                WMX_AEC_PHASE_END
                WMX_AEC_PHASE_BEGIN
                if (lane == 0) {
                    int rd = f2i(sc[S_OUT_RD]) + kFrame, wr = f2i(sc[S_OUT_WR]);
                    if (rd >= 2 * kOutRing) { rd -= kOutRing; wr -= kOutRing; }   // keep the counters small
                    sc[S_OUT_RD] = i2f(rd);
                    sc[S_OUT_WR] = i2f(wr);
                }
                WMX_AEC_PHASE_END
            }
        }
    }

    WMX_AEC_PHASE_BEGIN
    for (int i = lane; i < S_COUNT; i += 32) rec[Geo::kOffScal + i] = sc[i];
    WMX_AEC_PHASE_END
}

// Record initialisation: WebRtcAec_InitAec (aec_core.c:1509-1688) + WebRtcAec_Init
// (echo_cancellation.c:196-276) + set_config(aggressive)
WMX_HD void init_record(float* rec, int depth, int lane, int nlanes)
{
    const int total = rec_floats(depth);
    for (int i = lane; i < total; i += nlanes) rec[i] = 0.f;
}
WMX_HD void init_record_values(float* rec, int lane, int nlanes)
{
    for (int i = lane; i < kPart; i += nlanes) {
        rec[Geo::kOffBins + B_DMIN * kPart + i] = 1.0e6f;
        rec[Geo::kOffBins + B_SD * kPart + i] = 1.f;
        rec[Geo::kOffBins + B_SX * kPart + i] = 1.f;
    }
    if (lane == 0) {
        rec[Geo::kOffNyq + B_DMIN] = 1.0e6f;
        rec[Geo::kOffNyq + B_SD] = 1.f;
        rec[Geo::kOffNyq + B_SX] = 1.f;
        float* sc = rec + Geo::kOffScal;
        sc[S_NOISE_FROM_INIT] = i2f(1);
        sc[S_HNL_FB_MIN] = 1.f;
        sc[S_HNL_FB_LOCAL_MIN] = 1.f;
        sc[S_HNL_XD_AVG_MIN] = 1.f;
        sc[S_OVER_DRIVE] = 2.f;
        sc[S_OVER_DRIVE_SM] = 2.f;
        sc[S_SEED] = i2f(777);
        sc[S_CHECK_BUFF] = i2f(1);
        sc[S_STARTUP] = i2f(1);
        sc[S_FILT_DELAY] = i2f(-1);
    }
}

}  // namespace aec
}  // namespace wmx
