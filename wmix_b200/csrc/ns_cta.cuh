// WebRTC float noise suppressor, CTA-cooperative form: W worker warps (one stream-frame each) + ONE reducer warp.
//
// Same arithmetic as ns::frame (ns.cuh) — WebRtcNs_AnalyzeCore + WebRtcNs_ProcessCore,
// T:webrtc/modules/audio_processing/ns/ns_core.c:1043-1415 — re-distributed so that the work a single warp can only do
// with one to four active lanes is done ONCE for all the streams of the CTA:
//
//   * the in-order float sums (129- and 256-term chains the reference accumulates serially: signal energy, sum of
//     magnitudes, flatness numerator, pause average, the four spectral-difference sums, the two energies of the gain
//     map) — in ns::frame four lanes of every warp walk them while 28 idle; here lane 4j+k of the reducer walks sum k
//     of worker j, 32 chains per instruction;
//   * the per-stream scalar model (start-up white/pink fit, flatness / difference features, histogram re-learning,
//     the tanh indicators, prior update, gain-map factor): one lane per stream, eight streams per instruction;
//   * the Nyquist bin (the 129th bin that does not fit 4 bins x 32 lanes): its tracker, SNR, probability and filter
//     updates ran as a fifth pass of every per-bin phase with one lane active; the reducer runs them for all streams.
// ncu on ns::frame (profiles/r1_g): those three groups were 1 230 + ~380 of 4 587 warp instructions per stream-frame at
// 1.7 - 5 active lanes.
//
// Worker and reducer hand over through the worker's shared tile and named barriers (bar.arrive / bar.sync), three
// round trips per frame:
//     worker:  P0 load, window, forward FFT, real split, |X|, log, quantile trackers   --1-->  reducer: Nyquist tracker,
//              4 sums, start-up model, flatness                                          <--2--
//     worker:  noise blend, DD SNR, LRT average, difference terms                        --3-->  reducer: Nyquist SNR/LRT,
//              4 sums, features, histograms, tanh x3, prior, Nyquist probability + filter <--4--
//     worker:  probability, noise update, Wiener gain, inverse split, IFFT, scale        --5-->  reducer: two 256-term
//              energies, gain-map factor                                                 <--6--
//     worker:  window, overlap-add, saturate, emit
// Other changes against ns::frame, all value-preserving: the last FFT pass leaves its results in registers and the
// real split fetches the mirrored bin with warp shuffles (no exchange-tile round trip on either side of the split);
// the filtered spectrum and the scaled IFFT output stay in registers; PCM moves as 32-bit pairs and the history /
// synthesis tail as 8-byte pairs; the tail loops are fully unrolled.
//
// The same source runs under the lane-loop emulator (tests/emu): there the segments of the W workers and of the
// reducer are simply called in order.
#pragma once
#include "ns.cuh"

namespace wmx {
namespace ns {

// slots 32..63 of the worker tile's scalar area (0..31 hold the record's scalar line, see ScalarId)
enum CtaScal {
    C_ACTIVE = 32,     // 1.f: this worker holds a live, non-zero frame (the reducer has work for it)
    C_X0R, C_X0I,      // element 0 of the complex transform (DC and Nyquist before the real split)
    C_AVGMAGN, C_AVGPAUSE, C_PNUM, C_PEXP, C_USE_PINK,
    C_GAIN_PRIOR, C_RELEARNED, C_NYQ_FRE, C_WANT_E, C_FACTOR,
    C_COUNT_
};
static_assert(C_COUNT_ <= 64, "the scalar tile has 64 slots");

template <int ANA>
struct WLane {
    static constexpr int NS = Geo<ANA>::kSlots;
    Cpx f[4];                      // transform data: element lane + 32 r (after a last pass) or the pass's operands
    Cpx m[4];                      // mirrored elements fetched by shuffle (real split)
    float st[kNumRegArrays][NS];   // state arrays, bin = 32*slot + lane
    float mag[NS], noise[NS], prev[NS], prob[NS];
    int flag;
};
template <int ANA>
struct WWarp {
#if defined(__CUDA_ARCH__)
    WLane<ANA> lane_regs;
    int lane_id;
#else
    WLane<ANA> lane_regs[32];
#endif
};
// reducer lane state that lives across its three segments
struct RLane {
    float re, mag, lm, noise, prev, prob127;
    float lrt_prev, gain_prior;       // carried from segment 2a to the deferred Nyquist part 2b
    float mag0;                       // carried from segment 1 to the deferred part 1b
    int frame_idx, active;
};
struct RWarp {
#if defined(__CUDACC__)   // (not __CUDA_ARCH__: a non-template type must look the same in nvcc's host and device passes)
    RLane lane_regs;
    int lane_id;
#else
    RLane lane_regs[32];
#endif
};

#if defined(__CUDA_ARCH__)
#define WMX_CTA_PHASE_BEGIN(LaneT) { const int lane = W.lane_id; LaneT& R = W.lane_regs; (void)lane; (void)R;
#define WMX_CTA_PHASE_END } __syncwarp();
// dst (an lvalue naming R) = value of `val` as evaluated by lane `src`
#define WMX_CTA_SHFL(LaneT, dst, val, src) { const int lane = W.lane_id; LaneT& R = W.lane_regs; (void)lane; const float v_ = (val); dst = __shfl_sync(0xffffffffu, v_, (src)); }
#else
#define WMX_CTA_PHASE_BEGIN(LaneT) for (int lane = 0; lane < 32; ++lane) { LaneT& R = W.lane_regs[lane]; (void)R;
#define WMX_CTA_PHASE_END }
#define WMX_CTA_SHFL(LaneT, dst, val, src) { float pub_[32]; for (int lane = 0; lane < 32; ++lane) { LaneT& R = W.lane_regs[lane]; (void)R; pub_[lane] = (val); } \
                                            for (int lane = 0; lane < 32; ++lane) { LaneT& R = W.lane_regs[lane]; (void)R; dst = pub_[(src)]; } }
#endif

struct alignas(8) F2 { float x, y; };
struct alignas(4) S2 { int16_t x, y; };

// ---------------------------------------------------------------------------------------------------------------
// per-bin bodies, shared by the workers (body bins, state in registers) and the reducer (Nyquist bin, state in the
// tile's Nyquist line).  Each is the corresponding stretch of ns::frame, operand order unchanged.
// ---------------------------------------------------------------------------------------------------------------

// |X| + 1 and its log; rows of the first group of sums; the three quantile trackers (ns_core.c:217-283)
template <int ANA>
WMX_HD void bin_analyze(const Tables<ANA>& T, float* sv, int b, float re, float im, bool startup, const float cf[3], const float rcf[3],
                        const float cfm1[3], int quant_from, float& d0, float& d1, float& d2, float& q0, float& q1, float& q2,
                        float& quant, float& mag_out, float& noise_out)
{
    typedef Geo<ANA> G;
    const float mag = (b == 0 || b == G::kBody) ? (float)(fabs((double)re) + 1.0) : sqrtf(re * re + im * im) + 1.f;
    mag_out = mag;
    const float lm = log_f(mag, T.dm);
    sv[0 * G::kSumStride + b] = re * re + im * im;            // signalEnergy terms
    sv[1 * G::kSumStride + b] = mag;                          // sumMagn
    sv[2 * G::kSumStride + b] = (b >= 1) ? lm : 0.f;          // flatness numerator (bins 1..)
    if (startup) {                                            // start-up regressors (bins 5..)
        sv[4 * G::kSumStride + b] = (b >= 5) ? lm : 0.f;
        sv[5 * G::kSumStride + b] = (b >= 5) ? T.log_i[b] * lm : 0.f;
    }
    float* dens[3] = {&d0, &d1, &d2};
    float* lqs[3] = {&q0, &q1, &q2};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        float d = *dens[t], lq = *lqs[t];
        const float step = (d > 1.0f) ? fdiv(40.f * 1.f, d) : 40.f;
        const bool up = lm > lq;
        const float move = div_by_counter(up ? 0.25f * step : (1.f - 0.25f) * step, cf[t], rcf[t]);
        lq = up ? lq + move : lq - move;
        const float d_new = div_by_counter(cfm1[t] * d + 1.f / (2.f * 0.01f), cf[t], rcf[t]);
        d = (fabs(lm - lq) < 0.01f) ? d_new : d;
        *dens[t] = d;
        *lqs[t] = lq;
    }
    if (quant_from >= 0) {
        const float lq = quant_from == 0 ? q0 : (quant_from == 1 ? q1 : q2);
        quant = exp_f(lq, T.dm);
    }
    noise_out = quant;
}

// start-up noise blend, decision-directed SNR, rows of the second group of sums, LRT average (ns_core.c:1109-1162, :566-589,
// :617-622, :679-687).  param_noise is written only during start-up.
template <int ANA>
WMX_HD void bin_snr(const Tables<ANA>& T, float* sv, int b, int frame_idx, bool use_pink, float white, float pnum, float pexp,
                    float avg_magn, float avg_pause, float mag, float& noise, float magn_prev, float noise_prev, float smooth,
                    float pause, float& lrt, float& prev_out, float& param_noise)
{
    typedef Geo<ANA> G;
    if (frame_idx < kStartupShort) {
        float pn;
        if (!use_pink) {
            pn = white;
        } else {
            const float band = (float)(b < 5 ? 5 : b);
            pn = (float)((double)pnum / pow((double)band, (double)pexp));
        }
        param_noise = pn;
        noise *= (frame_idx);
        const float f2 = pn * (kStartupShort - frame_idx);
        noise += (f2 / (float)(frame_idx + 1));
        noise /= kStartupShort;
    }
    const float prev = fdiv(magn_prev, noise_prev + 0.0001f) * smooth;
    float post = 0.f;
    if (mag > noise) post = fdiv(mag, noise + 0.0001f) - 1.f;
    const float prior = 0.98f * prev + (1.f - 0.98f) * post;
    prev_out = prev;
    sv[0 * G::kSumStride + b] = (mag - avg_magn) * (pause - avg_pause);
    sv[1 * G::kSumStride + b] = (pause - avg_pause) * (pause - avg_pause);
    sv[2 * G::kSumStride + b] = (mag - avg_magn) * (mag - avg_magn);
    const float a = 1.f + 2.f * prior;
    const float bb = fdiv(2.f * prior, a + 0.0001f);
    const float bessel = (post + 1.f) * bb;
    lrt += 0.5f * (bessel - log_f(a, T.dm) - lrt);
    sv[3 * G::kSumStride + b] = lrt;
}

// speech probability of one bin (ns_core.c:741-747)
template <int ANA>
WMX_HD float bin_prob(const Tables<ANA>& T, float lrt, float gain_prior)
{
    float inv = exp_f(-lrt, T.dm);
    inv = (float)gain_prior * inv;
    return fdiv(1.f, 1.f + inv);
}

// noise update + Wiener gain of one bin (ns_core.c:800-846, :985-1010, :1276-1307); returns the gain
template <int ANA>
WMX_HD float bin_filter(const Tables<ANA>& T, int frame_idx, float mag, float ps, bool prev_bin_speech, float& nprev, float& pause,
                        float prev, float& init_magn, float param_noise)
{
    const float pn = 1.f - ps;
    const float gamma_old = prev_bin_speech ? 0.99f : 0.9f;
    const float prov = gamma_old * nprev + (1.f - gamma_old) * (pn * mag + ps * nprev);
    const float gamma = (ps > 0.2f) ? 0.99f : 0.9f;
    if (ps < 0.2f) pause += 0.05f * (mag - pause);
    float noise;
    if (gamma == gamma_old) {
        noise = prov;
    } else {
        noise = gamma * nprev + (1.f - gamma) * (pn * mag + ps * nprev);
        if (prov < noise) noise = prov;
    }
    const bool startup = frame_idx < kStartupShort;
    if (startup) init_magn += mag;
    float cur = 0.f;
    if (mag > noise) cur = fdiv(mag, noise + 0.0001f) - 1.f;
    const float snr = 0.98f * prev + (1.f - 0.98f) * cur;
    float h = fdiv(snr, T.overdrive + snr);
    if (h < T.floor_gain) h = T.floor_gain;
    if (h > 1.f) h = 1.f;
    if (startup) {
        float h0 = (init_magn - T.overdrive * param_noise);
        h0 /= (init_magn + 0.0001f);
        if (h0 < T.floor_gain) h0 = T.floor_gain;
        if (h0 > 1.f) h0 = 1.f;
        h *= (frame_idx);
        h0 *= (kStartupShort - frame_idx);
        h += h0;
        h /= (kStartupShort);
    }
    nprev = noise;
    return h;
}

// tracker bookkeeping every lane derives from the record's scalar line (ns_core.c:224-283)
struct TrackerCtl {
    int counter[3], updates, frame_idx, quant_from;
    bool startup;
    float cf[3], rcf[3], cfm1[3];
};
WMX_HD TrackerCtl tracker_ctl(const float* sc)
{
    TrackerCtl c;
    c.frame_idx = f2i(sc[S_FRAME_IDX]) + 1;     // blockInd after this frame's ++
    c.counter[0] = f2i(sc[S_COUNTER0]);
    c.counter[1] = f2i(sc[S_COUNTER1]);
    c.counter[2] = f2i(sc[S_COUNTER2]);
    c.updates = f2i(sc[S_UPDATES]);
    if (c.updates < kStartupLong) c.updates++;
    c.quant_from = -1;
#pragma unroll
    for (int t = 0; t < 3; ++t)
        if (c.counter[t] >= kStartupLong && c.updates >= kStartupLong) c.quant_from = t;
    if (c.updates < kStartupLong) c.quant_from = 2;
    c.startup = c.frame_idx < kStartupShort;
#pragma unroll
    for (int t = 0; t < 3; ++t) { c.cfm1[t] = (float)c.counter[t]; c.cf[t] = (float)(c.counter[t] + 1); c.rcf[t] = 1.f / c.cf[t]; }
    return c;
}

// ---------------------------------------------------------------------------------------------------------------
// transforms: the passes of ns::complex_passes, except that the LAST pass leaves its results in registers —
// f[r] = element lane + 32 r of the natural-order sequence (256-point; for 128 the tile is used as before and the
// elements are re-read in that distribution).
// ---------------------------------------------------------------------------------------------------------------
template <int ANA, typename WarpT>
WMX_HD void cta_passes(WarpT& W, float* sh, const float* tw, bool back)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    float* xb = sh + G::kShX;
    WMX_CTA_PHASE_BEGIN(L)
    if (lane < G::kBfly) {
        bfly4(R.f, make_tw(tw, lane));
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int p = xpos(4 * lane + q); xb[p] = R.f[q].r; xb[p + 1] = R.f[q].i; }
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(L)
    if (lane < G::kBfly) {
        const int g = lane >> 2, q = lane & 3;
#pragma unroll
        for (int r = 0; r < 4; ++r) { const int p = xpos(16 * g + q + 4 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
        bfly4(R.f, make_tw(tw, g));
#pragma unroll
        for (int r = 0; r < 4; ++r) { const int p = xpos(16 * g + q + 4 * r); xb[p] = R.f[r].r; xb[p + 1] = R.f[r].i; }
    }
    WMX_CTA_PHASE_END
    if (ANA == 256) {
        WMX_CTA_PHASE_BEGIN(L)
        {
            const int g = lane >> 4, q = lane & 15;
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = xpos(64 * g + q + 16 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
            bfly4(R.f, make_tw(tw, g));
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = xpos(64 * g + q + 16 * r); xb[p] = R.f[r].r; xb[p + 1] = R.f[r].i; }
        }
        WMX_CTA_PHASE_END
        // last pass: radix-2 on (c, c+64); lane does c = lane and c = lane+32 and KEEPS the four elements lane + 32 r
        WMX_CTA_PHASE_BEGIN(L)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = xpos(lane + 32 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
            bfly2_last(R.f[0], R.f[2], back);
            bfly2_last(R.f[1], R.f[3], back);
        }
        WMX_CTA_PHASE_END
    } else {
        WMX_CTA_PHASE_BEGIN(L)
        if (lane < 16) {
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = xpos(lane + 16 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
            bfly4_last(R.f, back);
#pragma unroll
            for (int r = 0; r < 4; ++r) { const int p = xpos(lane + 16 * r); xb[p] = R.f[r].r; xb[p + 1] = R.f[r].i; }
        }
        WMX_CTA_PHASE_END
        // 64 complex points: f[0], f[1] = elements lane, lane + 32
        WMX_CTA_PHASE_BEGIN(L)
        {
#pragma unroll
            for (int r = 0; r < 2; ++r) { const int p = xpos(lane + 32 * r); R.f[r].r = xb[p]; R.f[r].i = xb[p + 1]; }
        }
        WMX_CTA_PHASE_END
    }
}

// Mirror fetch for the real split: m[r] <- element (kNc - (lane + 32 r)) mod kNc, which lives in lane (32 - lane) & 31 as
// its register kNcR - 1 - r (kNcR - r in lane 0, whose own mirror is itself).  Two shuffles per element.
template <int ANA, typename WarpT>
WMX_HD void cta_fetch_mirrors(WarpT& W)
{
    typedef WLane<ANA> L;
    constexpr int NR = Geo<ANA>::kNc / 32;     // elements per lane: 4 (256-point) or 2
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        // the value lane `src` publishes for its reader: register NR-1-r, or (NR-r) mod NR when src is lane 0 (its reader is lane 0)
        WMX_CTA_SHFL(L, R.m[r].r, (lane == 0 ? R.f[(NR - r) % NR].r : R.f[NR - 1 - r].r), (32 - lane) & 31)
        WMX_CTA_SHFL(L, R.m[r].i, (lane == 0 ? R.f[(NR - r) % NR].i : R.f[NR - 1 - r].i), (32 - lane) & 31)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// worker segments
// ---------------------------------------------------------------------------------------------------------------

// segment 1: load, window, forward transform, real split, magnitudes, trackers.  Returns false for a zero frame (the
// worker has then already produced its output; the caller still walks the barriers).
template <int ANA, typename WarpT>
WMX_HD bool w_seg1(WarpT& W, float* rec, const int16_t* in, int16_t* out, float* sh, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    float* tb = sh + G::kShTime;
    float* sv = sh + G::kShSum;
    float* sc = sh + G::kShScal;
    float* nq = sh + G::kShNyq;
    float* syn = sh + G::kShSynth;
    float* sq = sh + G::kShSq;
    constexpr int HP = G::kOverlap / 2, BP = G::kBlock / 2;          // float2 / int16x2 pairs
    constexpr int NHP = (HP + 31) / 32, NBP = (BP + 31) / 32;

    WMX_CTA_PHASE_BEGIN(L)
    {
        const F2* rec2 = reinterpret_cast<const F2*>(rec);
        const S2* in2 = reinterpret_cast<const S2*>(in);
        F2 old_hist[NHP], tail[NHP];
        S2 carry[NHP], smp[NBP];
#pragma unroll
        for (int k = 0; k < NHP; ++k) {
            const int i = lane + 32 * k;
            if (i < HP) {
                old_hist[k] = rec2[G::kOffHist / 2 + i];
                tail[k] = rec2[G::kOffSynth / 2 + i];
                carry[k] = in2[(G::kBlock - G::kOverlap) / 2 + i];      // new history = last OVERLAP samples of this frame
            }
        }
#pragma unroll
        for (int k = 0; k < NBP; ++k) {
            const int i = lane + 32 * k;
            if (i < BP) smp[k] = in2[i];
        }
#pragma unroll
        for (int a = 0; a < kNumRegArrays; ++a) {
            if (!early_array(a)) continue;
#pragma unroll
            for (int s = 0; s < G::kSlots; ++s) R.st[a][s] = rec[G::kOffArrays + a * G::kBody + 32 * s + lane];
        }
        float pause_row[G::kSlots];
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) pause_row[s] = rec[G::kOffArrays + A_PAUSE * G::kBody + 32 * s + lane];
        const float nyq = rec[G::kOffNyq + lane], scal = rec[G::kOffScal + lane];
        F2* tb2 = reinterpret_cast<F2*>(tb);
        F2* syn2 = reinterpret_cast<F2*>(syn);
        F2* rec2w = reinterpret_cast<F2*>(rec);
#pragma unroll
        for (int k = 0; k < NHP; ++k) {
            const int i = lane + 32 * k;
            if (i < HP) {
                tb2[i] = old_hist[k];
                syn2[i] = tail[k];
                F2 c;
                c.x = (float)carry[k].x;
                c.y = (float)carry[k].y;
                rec2w[G::kOffHist / 2 + i] = c;
            }
        }
#pragma unroll
        for (int k = 0; k < NBP; ++k) {
            const int i = lane + 32 * k;
            if (i < BP) {
                F2 v;
                v.x = (float)smp[k].x;
                v.y = (float)smp[k].y;
                tb2[HP + i] = v;
            }
        }
        nq[lane] = nyq;
        sc[lane] = scal;
        sc[32 + lane] = 0.f;                                          // C_ACTIVE = 0 until this frame proves non-zero
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) sv[3 * G::kSumStride + 32 * s + lane] = pause_row[s];
        if (lane == A_PAUSE) sv[3 * G::kSumStride + G::kBody] = nyq;
    }
    WMX_CTA_PHASE_END

    // window, bit-reversed gather for pass 1; squares parked for the gain map's input energy
    WMX_CTA_PHASE_BEGIN(L)
    R.flag = 0;
    if (lane < G::kBfly) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = gather_index<ANA>(lane, q);
            const float a = T.window[2 * c] * tb[2 * c];
            const float b = T.window[2 * c + 1] * tb[2 * c + 1];
            R.f[q].r = a;
            R.f[q].i = b;
            const float aa = a * a, bb = b * b;
            sq[2 * c] = aa;
            sq[2 * c + 1] = bb;
            R.flag |= (aa != 0.f) | (bb != 0.f);
        }
    }
    WMX_CTA_PHASE_END

    const bool zero_frame = !WMX_NS_VOTE_ANY(W);
    if (zero_frame) {
        // ns_core.c:1072-1082 + :1239-1263: only the history (done above) and the synthesis tail change
        WMX_CTA_PHASE_BEGIN(L)
        for (int i = lane; i < G::kBlock; i += 32) {
            const float v = (i < G::kOverlap) ? syn[i] : 0.f;
            const float s = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
            out[i] = (int16_t)s;
        }
        for (int i = lane; i < G::kOverlap; i += 32) rec[G::kOffSynth + i] = 0.f;
        WMX_CTA_PHASE_END
        return false;
    }

    cta_passes<ANA>(W, sh, T.w, false);
    cta_fetch_mirrors<ANA>(W);

    // real split (rftfsub, fft4g.c:1234-1256), |X|+1, log|X|, quantile trackers — body bins only
    WMX_CTA_PHASE_BEGIN(L)
    {
        const TrackerCtl c = tracker_ctl(sc);
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) {
            const int b = 32 * s + lane;
            float re, im;
            if (b == 0) {
                re = R.f[0].r + R.f[0].i;                              // a[0] += a[1]
                im = 0.f;
            } else if (b == G::kNc / 2) {
                re = R.f[s].r;
                im = R.f[s].i;
            } else {
                const bool low = b < G::kNc / 2;
                const int cj = low ? b : G::kNc - b;
                const float jr = low ? R.f[s].r : R.m[s].r, ji = low ? R.f[s].i : R.m[s].i;
                const float kr = low ? R.m[s].r : R.f[s].r, ki = low ? R.m[s].i : R.f[s].i;
                const float wkr = 0.5f - T.c[ANA / 4 - cj], wki = T.c[cj];
                const float xr = jr - kr, xi = ji + ki;
                const float yr = wkr * xr - wki * xi, yi = wkr * xi + wki * xr;
                if (low) { re = jr - yr; im = ji - yi; }
                else { re = kr + yr; im = ki - yi; }
            }
            // the spectrum waits in the (now idle) time tile for the filter: re at [b], im of bins 1 .. N/2-1 behind them
            tb[b] = re;
            if (b >= 1) tb[G::kBody + b] = im;
            bin_analyze<ANA>(T, sv, b, re, im, c.startup, c.cf, c.rcf, c.cfm1, c.quant_from, R.st[A_DENS0][s], R.st[A_DENS1][s],
                             R.st[A_DENS2][s], R.st[A_LQ0][s], R.st[A_LQ1][s], R.st[A_LQ2][s], R.st[A_QUANT][s], R.mag[s], R.noise[s]);
        }
        // trackers are final for this frame: back to the record (whole lines); the filter-side arrays are requested
#pragma unroll
        for (int a = 0; a < kNumRegArrays; ++a) {
            if (!tracker_array(a)) continue;
            if (a == A_QUANT && c.quant_from < 0) continue;
#pragma unroll
            for (int s = 0; s < G::kSlots; ++s) rec[G::kOffArrays + a * G::kBody + 32 * s + lane] = R.st[a][s];
        }
#pragma unroll
        for (int a = 0; a < kNumRegArrays; ++a) {
            if (early_array(a)) continue;
#pragma unroll
            for (int s = 0; s < G::kSlots; ++s) R.st[a][s] = rec[G::kOffArrays + a * G::kBody + 32 * s + lane];
        }
        if (lane == 0) {
            sc[C_X0R] = R.f[0].r;
            sc[C_X0I] = R.f[0].i;
            sc[C_ACTIVE] = 1.f;
        }
    }
    WMX_CTA_PHASE_END
    return true;
}

// segment 2: start-up noise blend, decision-directed SNR, LRT average (per body bin)
template <int ANA, typename WarpT>
WMX_HD void w_seg2(WarpT& W, float* rec, float* sh, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    float* sv = sh + G::kShSum;
    float* sc = sh + G::kShScal;
    WMX_CTA_PHASE_BEGIN(L)
    {
        const int frame_idx = f2i(sc[S_FRAME_IDX]);
        const float avg_magn = sc[C_AVGMAGN], avg_pause = sc[C_AVGPAUSE];
        const float pnum = sc[C_PNUM], pexp = sc[C_PEXP];
        const bool use_pink = sc[C_USE_PINK] != 0.f;
        const float white = sc[S_WHITE];
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) {
            const int b = 32 * s + lane;
            float pn = 0.f;
            bin_snr<ANA>(T, sv, b, frame_idx, use_pink, white, pnum, pexp, avg_magn, avg_pause, R.mag[s], R.noise[s], R.st[A_MAGN_PREV][s],
                         R.st[A_NOISE_PREV][s], R.st[A_SMOOTH][s], R.st[A_PAUSE][s], R.st[A_LRT][s], R.prev[s], pn);
            if (frame_idx < kStartupShort) rec[G::kOffArrays + A_PARAM_NOISE * G::kBody + b] = pn;
        }
    }
    WMX_CTA_PHASE_END
}

// segment 3a: probability, noise update, Wiener gain, state back (everything that does not involve the Nyquist bin)
template <int ANA, typename WarpT>
WMX_HD void w_seg3a(WarpT& W, float* rec, uint16_t* hist, float* sh, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    float* tb = sh + G::kShTime;
    float* sv = sh + G::kShSum;
    float* sc = sh + G::kShScal;

    WMX_CTA_PHASE_BEGIN(L)
    {
        const float gain_prior = sc[C_GAIN_PRIOR];
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) {
            const float p = bin_prob<ANA>(T, R.st[A_LRT][s], gain_prior);
            R.prob[s] = p;
            sv[0 * G::kSumStride + 32 * s + lane] = p;
        }
        if (sc[C_RELEARNED] != 0.f) {
            uint32_t* h32 = reinterpret_cast<uint32_t*>(hist);
            for (int i = lane; i < 3 * kHistBins / 2; i += 32) h32[i] = 0u;
        }
    }
    WMX_CTA_PHASE_END

    WMX_CTA_PHASE_BEGIN(L)
    {
        const int frame_idx = f2i(sc[S_FRAME_IDX]);
        const bool startup = frame_idx < kStartupShort;
#pragma unroll
        for (int s = 0; s < G::kSlots; ++s) {
            const int b = 32 * s + lane;
            float init_magn = 0.f, param_noise = 0.f;
            if (startup) {
                init_magn = rec[G::kOffArrays + A_INIT_MAGN * G::kBody + b];
                param_noise = rec[G::kOffArrays + A_PARAM_NOISE * G::kBody + b];
            }
            const float mag = R.mag[s];
            const float h = bin_filter<ANA>(T, frame_idx, mag, R.prob[s], b > 0 && sv[b - 1] > 0.2f, R.st[A_NOISE_PREV][s], R.st[A_PAUSE][s],
                                            R.prev[s], init_magn, param_noise);
            if (startup) rec[G::kOffArrays + A_INIT_MAGN * G::kBody + b] = init_magn;
            R.st[A_SMOOTH][s] = h;
            R.st[A_MAGN_PREV][s] = mag;
            // filtered spectrum, kept in registers as element b of the packed sequence (ns_core.c:1296-1311)
            R.f[s].r = tb[b] * h;
            R.f[s].i = (b >= 1) ? tb[G::kBody + b] * h : 0.f;
        }
        // state arrays back to the record (whole lines)
#pragma unroll
        for (int a = 0; a < kNumRegArrays; ++a) {
            if (tracker_array(a)) continue;
#pragma unroll
            for (int s = 0; s < G::kSlots; ++s) rec[G::kOffArrays + a * G::kBody + 32 * s + lane] = R.st[a][s];
        }
    }
    WMX_CTA_PHASE_END

    cta_fetch_mirrors<ANA>(W);
}

// segment 3b: Nyquist line back, inverse split, inverse transform, scale.
// Leaves the scaled time signal in f[] (element pair lane + 32 r -> samples 2c, 2c+1) and its squares in the sum rows.
template <int ANA, typename WarpT>
WMX_HD void w_seg3b(WarpT& W, float* rec, float* sh, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    constexpr int NR = G::kNc / 32;
    float* xb = sh + G::kShX;
    float* sv = sh + G::kShSum;
    float* sc = sh + G::kShScal;
    float* nq = sh + G::kShNyq;

    // inverse real split (rdft isgn<0 head + rftbsub, fft4g.c:345-350, :1259-1283): every lane its elements lane + 32 r
    WMX_CTA_PHASE_BEGIN(L)
    {
        rec[G::kOffNyq + lane] = lane < 16 ? nq[lane] : 0.f;        // the reducer finished the Nyquist line in segment 2b
        const float nyq_fre = sc[C_NYQ_FRE];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int c = lane + 32 * r;
            float vr, vi;
            if (c == 0) {
                const float a0 = R.f[0].r, a1 = nyq_fre;
                const float h1 = 0.5f * (a0 - a1);
                vr = a0 - h1;
                vi = -h1;
            } else if (c == G::kNc / 2) {
                vr = R.f[r].r;
                vi = -R.f[r].i;
            } else {
                const bool low = c < G::kNc / 2;
                const int cj = low ? c : G::kNc - c;
                const float jr = low ? R.f[r].r : R.m[r].r, ji = low ? R.f[r].i : R.m[r].i;
                const float kr = low ? R.m[r].r : R.f[r].r, ki = low ? R.m[r].i : R.f[r].i;
                const float wkr = 0.5f - T.c[ANA / 4 - cj], wki = T.c[cj];
                const float xr = jr - kr, xi = ji + ki;
                const float yr = wkr * xr + wki * xi, yi = wkr * xi - wki * xr;
                if (low) { vr = jr - yr; vi = yi - ji; }
                else { vr = kr + yr; vi = yi - ki; }
            }
            const int p = xpos(c);
            xb[p] = vr;
            xb[p + 1] = vi;
        }
    }
    WMX_CTA_PHASE_END
    // bit-reversed gather for pass 1
    WMX_CTA_PHASE_BEGIN(L)
    if (lane < G::kBfly) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = gather_index<ANA>(lane, q);
            R.f[q].r = xb[xpos(c)];
            R.f[q].i = xb[xpos(c) + 1];
        }
    }
    WMX_CTA_PHASE_END
    cta_passes<ANA>(W, sh, T.w, true);

    // scale by 2/N (ns_core.c:941-943); the squares are staged for the reducer's in-order output energy
    WMX_CTA_PHASE_BEGIN(L)
    {
        const bool want_e = sc[C_WANT_E] != 0.f;
        F2* sv2 = reinterpret_cast<F2*>(sv);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int c = lane + 32 * r;
            const float a = R.f[r].r * (2.f / ANA), b = R.f[r].i * (2.f / ANA);
            R.f[r].r = a;
            R.f[r].i = b;
            if (want_e) {
                F2 e;
                e.x = a * a;
                e.y = b * b;
                sv2[c] = e;
            }
        }
    }
    WMX_CTA_PHASE_END
}

// segment 4: window, overlap-add, gain-map factor, saturate, emit; scalar line back to the record
template <int ANA, typename WarpT>
WMX_HD void w_seg4(WarpT& W, float* rec, int16_t* out, float* sh, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    typedef WLane<ANA> L;
    constexpr int NR = G::kNc / 32;
    float* sv = sh + G::kShSum;
    float* sc = sh + G::kShScal;
    float* syn = sh + G::kShSynth;
    WMX_CTA_PHASE_BEGIN(L)
    {
        const float factor = sc[C_FACTOR];
        const F2* win2 = reinterpret_cast<const F2*>(T.window);
        const F2* syn2 = reinterpret_cast<const F2*>(syn);
        S2* out2 = reinterpret_cast<S2*>(out);
        F2* rec2 = reinterpret_cast<F2*>(rec);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int c = lane + 32 * r;                               // samples 2c, 2c+1
            const F2 w = win2[c];
            F2 prev;
            prev.x = prev.y = 0.f;
            if (c < G::kOverlap / 2) prev = syn2[c];
            const float v0 = prev.x + factor * (w.x * R.f[r].r);
            const float v1 = prev.y + factor * (w.y * R.f[r].i);
            if (c < G::kBlock / 2) {
                const float s0 = v0 > 32767 ? 32767 : (v0 < -32768 ? -32768 : v0);
                const float s1 = v1 > 32767 ? 32767 : (v1 < -32768 ? -32768 : v1);
                S2 o;
                o.x = (int16_t)s0;
                o.y = (int16_t)s1;
                out2[c] = o;
            } else {
                F2 t;
                t.x = v0;
                t.y = v1;
                rec2[G::kOffSynth / 2 + c - G::kBlock / 2] = t;
            }
        }
        rec[G::kOffScal + lane] = sc[lane];
        // the energy staging ran over the zero padding of sum row 0 (and the rows behind it): restore what the sums rely on
        if (sc[C_WANT_E] != 0.f) {
#pragma unroll
            for (int k = 0; k < G::kNumSums; ++k)
                if (lane < G::kSumStride - G::kBins && k * G::kSumStride + G::kBins + lane < ANA) sv[k * G::kSumStride + G::kBins + lane] = 0.f;
        }
    }
    WMX_CTA_PHASE_END
}

// ---------------------------------------------------------------------------------------------------------------
// reducer segments.  Lane 4j+k serves worker j (tile base + j * tile_stride): k = 0 is the stream's scalar /
// Nyquist lane, k = 0..3 each walk one in-order sum.
// ---------------------------------------------------------------------------------------------------------------
template <int ANA, int N4>
WMX_HD float seq_sum_rows(const float* row) { return seq_sum4<N4>(row); }

// segment 1: Nyquist tracker, sums (signal energy, sum of magnitudes, flatness numerator, pause average, start-up
// regressors), counters, start-up model, flatness (ns_core.c:217-283, :1089-1162, :523-557)
template <int ANA, typename RWarpT>
WMX_HD void r_seg1(RWarpT& W, float* tiles, int tile_stride, int n_workers, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sv = sh + G::kShSum;
        float* sc = sh + G::kShScal;
        float* nq = sh + G::kShNyq;
        float* tb = sh + G::kShTime;
        R.active = (j < n_workers) && sc[C_ACTIVE] != 0.f;
        if (R.active && k == 0) {
            const TrackerCtl c = tracker_ctl(sc);
            R.frame_idx = c.frame_idx;
            const float re = sc[C_X0R] - sc[C_X0I];                    // xi = a[0] - a[1]
            R.re = re;
            tb[G::kBody] = re;
            bin_analyze<ANA>(T, sv, G::kBody, re, 0.f, c.startup, c.cf, c.rcf, c.cfm1, c.quant_from, nq[A_DENS0], nq[A_DENS1], nq[A_DENS2],
                             nq[A_LQ0], nq[A_LQ1], nq[A_LQ2], nq[A_QUANT], R.mag, R.noise);
            // counters advance (ns_core.c:265-283); the workers read the old values before barrier 1
            int cnt[3] = {c.counter[0], c.counter[1], c.counter[2]};
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (cnt[t] >= kStartupLong) cnt[t] = 0;
                cnt[t]++;
            }
            sc[S_COUNTER0] = i2f(cnt[0]);
            sc[S_COUNTER1] = i2f(cnt[1]);
            sc[S_COUNTER2] = i2f(cnt[2]);
            sc[S_UPDATES] = i2f(c.updates);
            sc[S_FRAME_IDX] = i2f(c.frame_idx);
        }
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sv = sh + G::kShSum;
        float* sc = sh + G::kShScal;
        if (R.active) {
            sc[C_COUNT_ + k] = seq_sum4<G::kSumStride / 4>(sv + k * G::kSumStride);
            const bool startup = f2i(sc[S_FRAME_IDX]) < kStartupShort;
            if (startup && k < 2) sc[C_COUNT_ + 4 + k] = seq_sum4<G::kSumStride / 4>(sv + (4 + k) * G::kSumStride);
        }
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sc = sh + G::kShScal;
        if (R.active && k == 0) {
            const int frame_idx = R.frame_idx;
            const float nb = (float)G::kBins;
            const float sig_e = sc[C_COUNT_ + 0] / nb;
            const float sum_magn = sc[C_COUNT_ + 1];
            sc[C_COUNT_ + 0] = sig_e;                                  // kept for the difference feature of segment 2
            float use_pink = 0.f, pnum = 0.f, pexp = 0.f;
            if (frame_idx < kStartupShort) {
                const float s_li = T.sum_log_i, s_li2 = T.sum_log_i_sq;
                const float s_lm = sc[C_COUNT_ + 4], s_lilm = sc[C_COUNT_ + 5];
                float white = sc[S_WHITE], pink_num = sc[S_PINK_NUM], pink_exp = sc[S_PINK_EXP];
                white += sum_magn / nb * T.overdrive;
                float f1 = s_li2 * ((float)(G::kBins - 5));
                f1 -= (s_li * s_li);
                float f2 = (s_li2 * s_lm - s_li * s_lilm);
                float f3 = f2 / f1;
                if (f3 < 0.f) f3 = 0.f;
                pink_num += f3;
                f2 = (s_li * s_lm);
                f2 -= ((float)(G::kBins - 5)) * s_lilm;
                f3 = f2 / f1;
                if (f3 < 0.f) f3 = 0.f;
                if (f3 > 1.f) f3 = 1.f;
                pink_exp += f3;
                if (pink_exp > 0.f) {
                    pnum = exp_f(pink_num / (float)(frame_idx + 1), T.dm);
                    pnum *= (float)(frame_idx + 1);
                    pexp = pink_exp / (float)(frame_idx + 1);
                }
                use_pink = (pink_exp == 0.f) ? 0.f : 1.f;
                sc[S_WHITE] = white;
                sc[S_PINK_NUM] = pink_num;
                sc[S_PINK_EXP] = pink_exp;
            }
            sc[C_PNUM] = pnum;
            sc[C_PEXP] = pexp;
            sc[C_USE_PINK] = use_pink;
            sc[C_AVGPAUSE] = sc[C_COUNT_ + 3] / nb;
            sc[C_AVGMAGN] = sum_magn / nb;
            R.mag0 = (sh + G::kShSum)[1 * G::kSumStride + 0];      // mag of bin 0 (row 1's first entry), for the flatness in 1b
        }
    }
    WMX_CTA_PHASE_END
}

// The reducer's chains are what the workers wait for, so what the workers do NOT need at a hand-over is taken off that chain.
//
// segment 1b: Nyquist SNR / LRT (ns_core.c:566-640) and the two features only the reducer's own segment 2 reads.  The kernel
// runs it AFTER releasing the workers into their segment 2 (which touches neither the scalar line nor the Nyquist line and
// writes other columns of the sum rows), i.e. concurrently with it.  Measured: 0.623 against 0.632 ms per 100 000-stream tick.
// (Deferring the Nyquist probability / filter of segment 2 the same way needs one more meeting point before the workers'
// inverse transform; that cost more than it hid: 0.672 ms with a barrier per worker, 0.700 ms with one CTA-wide barrier.)
template <int ANA, typename RWarpT>
WMX_HD void r_seg1b(RWarpT& W, float* tiles, int tile_stride, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sv = sh + G::kShSum;
        float* sc = sh + G::kShScal;
        float* nq = sh + G::kShNyq;
        if (R.active && k == 0) {
            float pn = 0.f;
            bin_snr<ANA>(T, sv, G::kBody, R.frame_idx, sc[C_USE_PINK] != 0.f, sc[S_WHITE], sc[C_PNUM], sc[C_PEXP], sc[C_AVGMAGN], sc[C_AVGPAUSE],
                         R.mag, R.noise, nq[A_MAGN_PREV], nq[A_NOISE_PREV], nq[A_SMOOTH], nq[A_PAUSE], nq[A_LRT], R.prev, pn);
            if (R.frame_idx < kStartupShort) nq[A_PARAM_NOISE] = pn;
            // the two features only the reducer's own segment 2a reads: long-term signal energy (ns_core.c:1141-1147) ...
            const int frame_idx = R.frame_idx;
            if (frame_idx < kStartupLong) {
                float f5 = sc[S_FEAT5];
                f5 *= frame_idx;
                f5 += sc[C_COUNT_ + 0];
                f5 /= (frame_idx + 1);
                sc[S_FEAT5] = f5;
            }
            // ... and spectral flatness (ns_core.c:523-557); |X|+1 >= 1 so the log(0) escape never fires
            {
                float den = sc[C_COUNT_ + 1];
                den -= R.mag0;
                float num = sc[C_COUNT_ + 2];
                den = den / G::kBins;
                num = num / G::kBins;
                const float v = exp_f(num, T.dm) / den;
                float f0 = sc[S_FEAT0];
                f0 += 0.3f * (v - f0);
                sc[S_FEAT0] = f0;
            }
        }
    }
    WMX_CTA_PHASE_END
}

// segment 2a: sums (covariance, two variances, LRT sum), difference feature, histograms and threshold re-learning,
// indicator functions, prior (ns_core.c:641-738, :293-520) — ends with everything the workers' segment 3 starts from
template <int ANA, typename RWarpT>
WMX_HD void r_seg2a(RWarpT& W, float* tiles, int tile_stride, uint16_t* const* hists, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sv = sh + G::kShSum;
        float* sc = sh + G::kShScal;
        if (R.active) sc[C_COUNT_ + 8 + k] = seq_sum4<G::kSumStride / 4>(sv + k * G::kSumStride);
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sc = sh + G::kShScal;
        if (R.active && k == 0) {
            uint16_t* hist = hists[j];
            const float nb = (float)G::kBins;
            {
                const float cov = sc[C_COUNT_ + 8] / nb, vp = sc[C_COUNT_ + 9] / nb, vm = sc[C_COUNT_ + 10] / nb;
                sc[S_FEAT6] = sc[S_FEAT6] + sc[C_COUNT_ + 0];
                float d = vm - (cov * cov) / (vp + 0.0001f);
                d = (float)(d / (sc[S_FEAT5] + 0.0001f));
                float f4 = sc[S_FEAT4];
                f4 += 0.3f * (d - f4);
                sc[S_FEAT4] = f4;
            }
            // histogram update / threshold re-learn (ns_core.c:755-790, :293-520); the LRT feature used here is still last
            // frame's (featureData[3] is refreshed further down)
            const int upd_mode = f2i(sc[S_UPD_MODE]);
            sc[C_RELEARNED] = 0.f;
            if (upd_mode >= 1) {
                int countdown = f2i(sc[S_UPD_COUNTDOWN]) - 1;
                if (countdown > 0) {
                    const float v3 = sc[S_FEAT3], v0 = sc[S_FEAT0], v4 = sc[S_FEAT4];
                    if (v3 < kHistBins * 0.1f && v3 >= 0.0) hist_inc(hist, 0 * kHistBins + (int)(v3 / 0.1f));
                    if (v0 < kHistBins * 0.05f && v0 >= 0.0) hist_inc(hist, 1 * kHistBins + (int)(v0 / 0.05f));
                    if (v4 < kHistBins * 0.1f && v4 >= 0.0) hist_inc(hist, 2 * kHistBins + (int)(v4 / 0.1f));
                }
                if (countdown == 0) {
                    const int window = 500;
                    float avg = 0.f, avg_all = 0.f, avg_sq = 0.f;
                    int n = 0;
                    for (int i = 0; i < kHistBins; ++i) {
                        const int h = hist[i];
                        if (h == 0) continue;                      // adding 0.f never changes a float sum
                        const float mid = ((float)i + 0.5f) * 0.1f;
                        if (mid <= 1.f) { avg += h * mid; n += h; }
                        avg_sq += h * mid * mid;
                        avg_all += h * mid;
                    }
                    if (n > 0) avg = avg / ((float)n);
                    avg_all = avg_all / ((float)window);
                    avg_sq = avg_sq / ((float)window);
                    const float fluct = avg_sq - avg * avg_all;
                    float pm0;
                    if (fluct < 0.05f) pm0 = 1.f;
                    else {
                        pm0 = 1.2f * avg;
                        if (pm0 < 0.2f) pm0 = 0.2f;
                        if (pm0 > 1.f) pm0 = 1.f;
                    }
                    sc[S_PM0] = pm0;
                    int use_flat = 1, use_diff = 1;
                    for (int which = 1; which <= 2; ++which) {
                        const float bin = which == 1 ? 0.05f : 0.1f;
                        const uint16_t* h = hist + which * kHistBins;
                        int m1 = 0, m2 = 0, w1 = 0, w2 = 0;
                        float p1 = 0.f, p2 = 0.f;
                        for (int i = 0; i < kHistBins; ++i) {
                            const int v = h[i];
                            const float mid = ((float)i + 0.5f) * bin;
                            if (v > m1) { m2 = m1; w2 = w1; p2 = p1; m1 = v; w1 = v; p1 = mid; }
                            else if (v > m2) { m2 = v; w2 = v; p2 = mid; }
                        }
                        if ((fabs(p2 - p1) < 2 * bin) && (w2 > 0.5f * w1)) { w1 += w2; p1 = 0.5f * (p1 + p2); }
                        const int min_weight = (int)(0.3 * (window));
                        if (which == 1) {
                            if (w1 < min_weight || p1 < 0.6f) use_flat = 0;
                            if (use_flat) {
                                float pm1 = 0.9f * p1;
                                if (pm1 < 0.1f) pm1 = 0.1f;
                                if (pm1 > 0.95f) pm1 = 0.95f;
                                sc[S_PM1] = pm1;
                            }
                        } else {
                            float pm3 = 1.2f * p1;
                            if (w1 < min_weight) use_diff = 0;
                            if (pm3 < 0.16f) pm3 = 0.16f;
                            if (pm3 > 1.f) pm3 = 1.f;
                            sc[S_PM3] = pm3;
                            if (fluct < 0.05f) use_diff = 0;
                        }
                    }
                    const float fsum = (float)(1 + use_flat + use_diff);
                    sc[S_PM4] = 1.f / fsum;
                    sc[S_PM5] = ((float)use_flat) / fsum;
                    sc[S_PM6] = ((float)use_diff) / fsum;
                    sc[C_RELEARNED] = 1.f;                         // the worker clears the histograms in segment 3
                    countdown = window;
                    if (upd_mode == 1) {
                        sc[S_UPD_MODE] = i2f(0);
                    } else {
                        float f6 = sc[S_FEAT6] / ((float)window);
                        sc[S_FEAT5] = 0.5f * (f6 + sc[S_FEAT5]);
                        sc[S_FEAT6] = 0.f;
                    }
                }
                sc[S_UPD_COUNTDOWN] = i2f(countdown);
            }
            // arguments of the three indicator functions (ns_core.c:689-730)
            {
                const float thr0 = sc[S_PM0], thr1 = sc[S_PM1], thr2 = sc[S_PM3];
                const int sgn = (int)(sc[S_PM2]);
                float ksum = sc[C_COUNT_ + 11];
                ksum = (float)ksum / (G::kBins);
                sc[S_FEAT3] = ksum;
                float width = 4.f;
                if (ksum < thr0) width = 2.f * 4.f;
                sc[C_COUNT_ + 12] = width * (ksum - thr0);
                float x = sc[S_FEAT0];
                width = 4.f;
                if (sgn == 1 && (x > thr1)) width = 2.f * 4.f;
                if (sgn == -1 && (x < thr1)) width = 2.f * 4.f;
                sc[C_COUNT_ + 13] = (float)sgn * width * (thr1 - x);
                x = sc[S_FEAT4];
                width = 4.f;
                if (x < thr2) width = 2.f * 4.f;
                sc[C_COUNT_ + 14] = width * (x - thr2);
            }
        }
    }
    WMX_CTA_PHASE_END
    // the three indicator functions side by side in lanes k = 0..2 of every stream
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sc = tiles + (size_t)j * tile_stride + G::kShScal;
        if (R.active && k < 3) sc[C_COUNT_ + 12 + k] = 0.5f * ((float)tanh((double)sc[C_COUNT_ + 12 + k]) + 1.f);
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sv = sh + G::kShSum;
        float* sc = sh + G::kShScal;
        if (R.active && k == 0) {
            // prior update (ns_core.c:731-738)
            const float ind = sc[S_PM4] * sc[C_COUNT_ + 12] + sc[S_PM5] * sc[C_COUNT_ + 13] + sc[S_PM6] * sc[C_COUNT_ + 14];
            float pp = sc[S_PRIOR_PROB];
            pp += 0.1f * (ind - pp);
            if (pp > 1.f) pp = 1.f;
            if (pp < 0.01f) pp = 0.01f;
            sc[S_PRIOR_PROB] = pp;
            const float gain_prior = fdiv(1.f - pp, pp + 0.0001f);
            sc[C_GAIN_PRIOR] = gain_prior;
            sc[C_WANT_E] = (T.gainmap == 1 && R.frame_idx > kStartupLong) ? 1.f : 0.f;
            sc[C_FACTOR] = 1.f;
            R.gain_prior = gain_prior;
            R.lrt_prev = sv[3 * G::kSumStride + G::kBody - 1];     // the workers reuse the sum rows from here on
        }
    }
    WMX_CTA_PHASE_END
}

// segment 2b: Nyquist probability, noise update, gain, filtered value (ns_core.c:741-846, :985-1010)
template <int ANA, typename RWarpT>
WMX_HD void r_seg2b(RWarpT& W, float* tiles, int tile_stride, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sc = sh + G::kShScal;
        float* nq = sh + G::kShNyq;
        if (R.active && k == 0) {
            const float gain_prior = R.gain_prior;
            // the look-back neighbour, bin kBody-1, is re-derived here from the LRT the worker staged — the same arithmetic,
            // so the same value the worker computes
            const float p_prev = bin_prob<ANA>(T, R.lrt_prev, gain_prior);
            const float ps = bin_prob<ANA>(T, nq[A_LRT], gain_prior);
            float init_magn = nq[A_INIT_MAGN];
            const float h = bin_filter<ANA>(T, R.frame_idx, R.mag, ps, p_prev > 0.2f, nq[A_NOISE_PREV], nq[A_PAUSE], R.prev, init_magn,
                                            nq[A_PARAM_NOISE]);
            if (R.frame_idx < kStartupShort) nq[A_INIT_MAGN] = init_magn;
            nq[A_SMOOTH] = h;
            nq[A_MAGN_PREV] = R.mag;
            sc[C_NYQ_FRE] = R.re * h;
        }
    }
    WMX_CTA_PHASE_END
}

// segment 3: the two energies of the gain map in sample order (lanes k = 0, 1), then the factor (ns_core.c:1314-1342)
template <int ANA, typename RWarpT>
WMX_HD void r_seg3(RWarpT& W, float* tiles, int tile_stride, const Tables<ANA>& T)
{
    typedef Geo<ANA> G;
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sh = tiles + (size_t)j * tile_stride;
        float* sc = sh + G::kShScal;
        if (R.active && k < 2 && sc[C_WANT_E] != 0.f)
            sc[C_COUNT_ + 16 + k] = seq_sum4<ANA / 4>(k == 0 ? sh + G::kShSq : sh + G::kShSum);
    }
    WMX_CTA_PHASE_END
    WMX_CTA_PHASE_BEGIN(RLane)
    {
        const int j = lane >> 2, k = lane & 3;
        float* sc = tiles + (size_t)j * tile_stride + G::kShScal;
        if (R.active && k == 0 && sc[C_WANT_E] != 0.f) {
            // (float)sqrt((double)x) == sqrtf(x): a double carries more than 2*24+2 bits, so rounding twice is innocuous
            float gain = sqrtf(sc[C_COUNT_ + 17] / (sc[C_COUNT_ + 16] + 1.f));
            float f1 = 1.f, f2 = 1.f;
            if (gain > 0.5f) {
                f1 = 1.f + 1.3f * (gain - 0.5f);
                if (gain * f1 > 1.f) f1 = 1.f / gain;
            }
            if (gain < 0.5f) {
                if (gain <= T.floor_gain) gain = T.floor_gain;
                f2 = 1.f - 0.3f * (0.5f - gain);
            }
            const float pp = sc[S_PRIOR_PROB];
            sc[C_FACTOR] = pp * f1 + (1.f - pp) * f2;
        }
    }
    WMX_CTA_PHASE_END
}

}  // namespace ns
}  // namespace wmx
