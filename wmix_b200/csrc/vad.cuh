// WebRTC VAD, one stream per thread, state in structure-of-arrays 32-bit words.
//
// Follows T:webrtc/common_audio/vad/{vad_sp.c,vad_filterbank.c,vad_gmm.c,vad_core.c} and the
// wmix wrapper R:src/webrtc.c:91-151 bit for bit (integer arithmetic, int16 narrowing wraps,
// arithmetic >> on negatives).  The IIR chains are serial in time, so all parallelism is
// across streams; consecutive threads own consecutive streams and every state word access is
// one coalesced line per warp (SoaWords).
#pragma once
#include "common.cuh"

namespace wmx {
namespace vad {

// ---- state word map (per stream); int16 fields are packed two per word ----
enum {
    W_DS = 0,        // 2 words : 16k->8k decimator all-pass states (int32)
    W_NMEAN = 2,     // 6 words : noise means,  word ch = (gaussian 0, gaussian 1)
    W_SMEAN = 8,     // 6 words : speech means
    W_NSTD = 14,     // 6 words : noise stds
    W_SSTD = 20,     // 6 words : speech stds
    W_FRAMES = 26,   // 1 word  : frame_counter
    W_HANG = 27,     // 1 word  : (over_hang, num_of_speech)
    W_AGE = 28,      // 48 words: age[ch][16]
    W_LOW = 76,      // 48 words: low_value[ch][16]
    W_MEANVAL = 124, // 3 words : mean_value[6]
    W_UPPER = 127,   // 3 words : split-filter upper states [5] (+pad)
    W_LOWER = 130,   // 3 words : split-filter lower states [5] (+pad)
    W_HP = 133,      // 2 words : 80 Hz high-pass state [4]
    W_REDUCE = 135,  // 1 word  : wmix wrapper's attenuation shift 0..4
    W_DS32 = 136,    // 2 words : 32k->16k decimator all-pass states (int32), T:.../vad/vad_core.c:632
    N_WORDS = 138
};

struct Params {            // mode-dependent thresholds for the frame length in use
    int16_t over_hang_1, over_hang_2, local_thr, global_thr;
};

// trained model constants, T:.../vad/vad_core.c:22-63.  Function-local const arrays so the
// same source serves the device pass (constant bank) and the host emulation build.
#define WMX_VAD_MODEL_TABLES                                                                      \
    const int16_t k_spec_w[6] = {6, 8, 10, 12, 14, 16};                                          \
    const int16_t k_min_diff[6] = {544, 544, 576, 576, 576, 576};                                \
    const int16_t k_max_speech[6] = {11392, 11392, 11520, 11520, 11520, 11520};                  \
    const int16_t k_max_noise[6] = {9216, 9088, 8960, 8832, 8704, 8576};                         \
    const int16_t k_noise_w[12] = {34, 62, 72, 66, 53, 25, 94, 66, 56, 62, 75, 103};             \
    const int16_t k_speech_w[12] = {48, 82, 45, 87, 50, 47, 80, 46, 83, 41, 78, 81};

// 16 kHz -> 8 kHz: two first-order all-pass branches summed (T:.../vad/vad_sp.c:27-54)
template <int LEN8>
WMX_HD void decimate(const int16_t* in, int16_t* out, int32_t& s0, int32_t& s1)
{
    for (int n = 0; n < LEN8; ++n) {
        int x0 = in[2 * n], x1 = in[2 * n + 1];
        int16_t a = (int16_t)((s0 >> 1) + ((5243 * x0) >> 14));
        s0 = x0 - ((5243 * a) >> 12);
        int16_t b = (int16_t)((s1 >> 1) + ((1392 * x1) >> 14));
        s1 = x1 - ((1392 * b) >> 12);
        out[n] = (int16_t)(a + b);
    }
}

// first-order all-pass over every second input sample (T:.../vad/vad_filterbank.c:83-108)
WMX_HD void allpass2(const int16_t* in, int n, int coef, int16_t& state, int16_t* out)
{
    int32_t s = wshl((int32_t)state, 16);
    for (int i = 0; i < n; ++i) {
        int x = in[2 * i];
        int16_t y = (int16_t)(wadd(s, coef * x) >> 16);
        out[i] = y;
        s = wshl(wsub(wshl(x, 14), coef * y), 1);
    }
    state = (int16_t)(s >> 16);
}

// half-band split: hp = up - low, lp = low + up (T:.../vad/vad_filterbank.c:121-140)
WMX_HD void split(const int16_t* in, int n, int16_t& up_st, int16_t& lo_st, int16_t* hp, int16_t* lp)
{
    int half = n >> 1;
    allpass2(in, half, 20972, up_st, hp);
    allpass2(in + 1, half, 5571, lo_st, lp);
    for (int i = 0; i < half; ++i) {
        int16_t u = hp[i], l = lp[i];
        hp[i] = (int16_t)(u - l);
        lp[i] = (int16_t)(l + u);
    }
}

// WebRtcSpl_Energy with its data-dependent pre-shift (T:.../signal_processing/energy.c:18-35,
// get_scaling_square.c:20-49); |x| is taken in int16 so -32768 never wins the max
WMX_HD int32_t energy16(const int16_t* v, int n, int& scale)
{
    int16_t peak = -1;
    for (int i = 0; i < n; ++i) {
        int16_t a = (int16_t)(v[i] > 0 ? v[i] : -v[i]);
        if (a > peak) peak = a;
    }
    int nbits = size_in_bits((uint32_t)n);
    int t = norm_w32((int32_t)peak * peak);
    int sh = (peak == 0) ? 0 : ((t > nbits) ? 0 : nbits - t);
    int32_t e = 0;
    for (int i = 0; i < n; ++i) e = wadd(e, ((int32_t)v[i] * v[i]) >> sh);
    scale = sh;
    return e;
}

// 10*log10(energy) in Q4 (T:.../vad/vad_filterbank.c:155-236)
WMX_HD int16_t log_energy(const int16_t* in, int n, int16_t offset, int16_t& total)
{
    int rsh = 0;
    uint32_t e = (uint32_t)energy16(in, n, rsh);
    if (e == 0) return offset;
    int norm = 17 - norm_u32(e);
    rsh += norm;
    e = (norm < 0) ? (e << -norm) : (e >> norm);
    int16_t log2e = (int16_t)(14336 + (int16_t)((e & 0x3FFF) >> 4));
    int16_t r = (int16_t)(((24660 * log2e) >> 19) + ((rsh * 24660) >> 9));
    if (r < 0) r = 0;
    if (total <= 10) {
        if (rsh >= 0) total = (int16_t)(total + 11);
        else total = (int16_t)(total + (int16_t)(e >> -rsh));
    }
    return (int16_t)(r + offset);
}

// five split stages -> six band log-energies (T:.../vad/vad_filterbank.c:246-333)
template <int LEN8>
WMX_HD int16_t features(const SoaWords& st, const int16_t* in, int16_t* feat)
{
    int16_t hpA[LEN8 / 2], lpA[LEN8 / 2], hpB[LEN8 / 4], lpB[LEN8 / 4];
    int16_t up[6], lo[6], hp[4];
    for (int w = 0; w < 3; ++w) {
        int32_t a = st.get(W_UPPER + w), b = st.get(W_LOWER + w);
        up[2 * w] = lo16(a); up[2 * w + 1] = hi16(a);
        lo[2 * w] = lo16(b); lo[2 * w + 1] = hi16(b);
    }
    {
        int32_t a = st.get(W_HP), b = st.get(W_HP + 1);
        hp[0] = lo16(a); hp[1] = hi16(a); hp[2] = lo16(b); hp[3] = hi16(b);
    }
    int16_t total = 0;
    int n = LEN8 / 4;
    split(in, LEN8, up[0], lo[0], hpA, lpA);            // 2-4 kHz | 0-2 kHz
    split(hpA, LEN8 / 2, up[1], lo[1], hpB, lpB);       // 3-4 | 2-3
    feat[5] = log_energy(hpB, n, 176, total);
    feat[4] = log_energy(lpB, n, 176, total);
    split(lpA, LEN8 / 2, up[2], lo[2], hpB, lpB);       // 1-2 | 0-1
    feat[3] = log_energy(hpB, n, 176, total);
    split(lpB, n, up[3], lo[3], hpA, lpA);              // 0.5-1 | 0-0.5
    n >>= 1;
    feat[2] = log_energy(hpA, n, 272, total);
    split(lpA, n, up[4], lo[4], hpB, lpB);              // 250-500 | 0-250
    n >>= 1;
    feat[1] = log_energy(hpB, n, 368, total);
    // 80 Hz high-pass on the lowest band (T:.../vad/vad_filterbank.c:41-72)
    for (int i = 0; i < n; ++i) {
        int x = lpB[i];
        int32_t acc = 6631 * x;
        acc += -13262 * hp[0];
        acc += 6631 * hp[1];
        hp[1] = hp[0];
        hp[0] = (int16_t)x;
        acc -= -7756 * hp[2];
        acc -= 5620 * hp[3];
        hp[3] = hp[2];
        hp[2] = (int16_t)(acc >> 14);
        hpA[i] = hp[2];
    }
    feat[0] = log_energy(hpA, n, 368, total);
    for (int w = 0; w < 3; ++w) {
        st.set(W_UPPER + w, pack16(up[2 * w], up[2 * w + 1]));
        st.set(W_LOWER + w, pack16(lo[2 * w], lo[2 * w + 1]));
    }
    st.set(W_HP, pack16(hp[0], hp[1]));
    st.set(W_HP + 1, pack16(hp[2], hp[3]));
    return total;
}

// T:.../vad/vad_gmm.c:30-83
WMX_HD int32_t gaussian_body(int16_t input, int16_t mean, int16_t sd, int16_t& delta)
{
    int16_t inv_std = (int16_t)div_w32_w16(131072 + (int32_t)(sd >> 1), sd);
    int16_t t16 = (int16_t)(inv_std >> 2);
    int16_t inv_std2 = (int16_t)((t16 * t16) >> 2);
    t16 = (int16_t)(input << 3);
    t16 = (int16_t)(t16 - mean);
    delta = (int16_t)((inv_std2 * t16) >> 10);
    int32_t t32 = (delta * t16) >> 9;
    int16_t expv = 0;
    if (t32 < 22005) {
        t16 = (int16_t)((5909 * t32) >> 12);
        t16 = (int16_t)-t16;
        expv = (int16_t)(0x0400 | (t16 & 0x03FF));
        t16 = (int16_t)~t16;
        t16 >>= 10;
        t16 += 1;
        expv >>= t16;
    }
    return inv_std * expv;
}

// 24 calls per frame: one out-of-line copy on the device (post_kernel is bound by instruction fetch, see common.cuh);
// probability and delta come back packed in one 64-bit value
#if defined(__CUDACC__)
static __device__ __noinline__ long long gaussian_device(int input, int mean, int sd)
{
    int16_t d;
    const int32_t p = gaussian_body((int16_t)input, (int16_t)mean, (int16_t)sd, d);
    return (long long)(((unsigned long long)(uint32_t)p << 16) | (uint16_t)d);
}
#endif
WMX_HD int32_t gaussian(int16_t input, int16_t mean, int16_t sd, int16_t& delta)
{
#if defined(__CUDA_ARCH__)
    const unsigned long long r = (unsigned long long)gaussian_device(input, mean, sd);
    delta = (int16_t)(r & 0xFFFFu);
    return (int32_t)(uint32_t)(r >> 16);
#else
    return gaussian_body(input, mean, sd, delta);
#endif
}

// 16 smallest feature values of the last 100 frames + smoothed "median" (T:.../vad/vad_sp.c:59-177).
// The two 16-entry lists stay PACKED as the record holds them — entry 2w in the low half of word w, 2w+1 in the high half —
// and every step works on the pairs: a warp carries 32 streams and some lane evicts an entry on most frames, so the
// "rare" eviction runs for the whole warp; as funnel shifts of eight words it costs a quarter of moving 2 x 15 unpacked
// entries, and nothing is unpacked on entry or re-packed on exit.  All indices are static after unrolling.
WMX_HD uint32_t pair_down(uint32_t cur, uint32_t next) { return (cur >> 16) | (next << 16); }   // (cur.hi, next.lo)
WMX_HD int16_t find_minimum(const SoaWords& st, int16_t feature, int ch, int32_t frame_counter)
{
    uint32_t A[8], Lw[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        A[w] = (uint32_t)st.get(W_AGE + ch * 8 + w);
        Lw[w] = (uint32_t)st.get(W_LOW + ch * 8 + w);
    }
    // every entry gets one frame older; an entry that has reached 100 is removed instead: the entries behind it move down
    // one place on the LIVE list and (age 101, value 10000) enters at the end.  As in the reference the scan then goes on
    // with the next POSITION, so the entry that moved into the freed place is neither aged nor examined this frame, and its
    // copy loop's read one past the list lands in the slot the sentinel overwrites (vad_sp.c:78-90).
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int w = i >> 1;
        const uint32_t age_i = (i & 1) ? A[w] >> 16 : A[w] & 0xFFFFu;
        if (age_i != 100u) {
            A[w] = (i & 1) ? A[w] + 0x10000u : (A[w] & 0xFFFF0000u) | ((A[w] + 1u) & 0xFFFFu);
        } else {
#pragma unroll
            for (int v = w; v < 8; ++v) {
                const uint32_t na = v < 7 ? A[v + 1] : 101u, nl = v < 7 ? Lw[v + 1] : 10000u;
                if (v == w && (i & 1)) {
                    A[v] = (A[v] & 0xFFFFu) | (na << 16);
                    Lw[v] = (Lw[v] & 0xFFFFu) | (nl << 16);
                } else {
                    A[v] = pair_down(A[v], na);
                    Lw[v] = pair_down(Lw[v], nl);
                }
            }
        }
    }
#define WMX_LOW(i) ((int16_t)(((i) & 1) ? Lw[(i) >> 1] >> 16 : Lw[(i) >> 1] & 0xFFFFu))
    // fixed comparison tree of the reference (vad_sp.c:93-146); 16 = nothing to insert
    int pos = 16;
    if (feature < WMX_LOW(7)) {
        if (feature < WMX_LOW(3)) {
            if (feature < WMX_LOW(1)) pos = (feature < WMX_LOW(0)) ? 0 : 1;
            else pos = (feature < WMX_LOW(2)) ? 2 : 3;
        } else if (feature < WMX_LOW(5)) pos = (feature < WMX_LOW(4)) ? 4 : 5;
        else pos = (feature < WMX_LOW(6)) ? 6 : 7;
    } else if (feature < WMX_LOW(15)) {
        if (feature < WMX_LOW(11)) {
            if (feature < WMX_LOW(9)) pos = (feature < WMX_LOW(8)) ? 8 : 9;
            else pos = (feature < WMX_LOW(10)) ? 10 : 11;
        } else if (feature < WMX_LOW(13)) pos = (feature < WMX_LOW(12)) ? 12 : 13;
        else pos = (feature < WMX_LOW(14)) ? 14 : 15;
    }
    // insert (feature, age 1) at pos: the entries from pos on move up one place, the last one drops out (vad_sp.c:150-158).
    // Downwards over the words so that word w - 1 is still the old one when word w takes its high half.
    {
        const uint32_t fv = (uint32_t)(uint16_t)feature;
#pragma unroll
        for (int w = 7; w >= 0; --w) {
            const uint32_t pa = w > 0 ? A[w - 1] : 0u, pl = w > 0 ? Lw[w - 1] : 0u;
            if (pos < 2 * w) {
                A[w] = (pa >> 16) | (A[w] << 16);
                Lw[w] = (pl >> 16) | (Lw[w] << 16);
            } else if (pos == 2 * w) {
                A[w] = 1u | (A[w] << 16);
                Lw[w] = fv | (Lw[w] << 16);
            } else if (pos == 2 * w + 1) {
                A[w] = (A[w] & 0xFFFFu) | (1u << 16);
                Lw[w] = (Lw[w] & 0xFFFFu) | (fv << 16);
            }
        }
    }
    int16_t median = 1600, alpha = 0;
    if (frame_counter > 2) median = WMX_LOW(2);
    else if (frame_counter > 0) median = WMX_LOW(0);
#undef WMX_LOW
    int w = st.get(W_MEANVAL + (ch >> 1));
    int16_t mv = (ch & 1) ? hi16(w) : lo16(w);
    if (frame_counter > 0) alpha = (median < mv) ? 6553 : 32439;
    int32_t acc = (alpha + 1) * mv;
    acc += (32767 - alpha) * median;
    acc += 16384;
    mv = (int16_t)(acc >> 15);
    st.set(W_MEANVAL + (ch >> 1), (ch & 1) ? pack16(lo16(w), mv) : pack16(mv, hi16(w)));
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        st.set(W_AGE + ch * 8 + k, (int32_t)A[k]);
        st.set(W_LOW + ch * 8 + k, (int32_t)Lw[k]);
    }
    return mv;
}

// shift both Gaussians' means of one band and return their weighted sum
// (T:.../vad/vad_core.c:110-122 WeightedAverage)
WMX_HD int32_t shift_avg(int16_t m[2], int16_t shift, int16_t w0, int16_t w1)
{
    m[0] = (int16_t)(m[0] + shift);
    m[1] = (int16_t)(m[1] + shift);
    return m[0] * w0 + m[1] * w1;
}

// GMM hypothesis test + model adaptation + hang-over (T:.../vad/vad_core.c:124-479)
WMX_HD int gmm(const SoaWords& st, const int16_t* feat, int16_t total_power, const Params& P)
{
    WMX_VAD_MODEL_TABLES
    int16_t vadflag = 0;
    int16_t dN[12], dS[12], pN[12], pS[12];
    for (int i = 0; i < 12; ++i) { pN[i] = 0; pS[i] = 0; dN[i] = 0; dS[i] = 0; }

    if (total_power > 10) {
        int32_t sum_llr = 0;
        for (int ch = 0; ch < 6; ++ch) {
            int32_t wnm = st.get(W_NMEAN + ch), wsm = st.get(W_SMEAN + ch);
            int32_t wns = st.get(W_NSTD + ch), wss = st.get(W_SSTD + ch);
            int32_t probN0 = k_noise_w[ch] * gaussian(feat[ch], lo16(wnm), lo16(wns), dN[ch]);
            int32_t probS0 = k_speech_w[ch] * gaussian(feat[ch], lo16(wsm), lo16(wss), dS[ch]);
            int32_t probN1 = k_noise_w[ch + 6] * gaussian(feat[ch], hi16(wnm), hi16(wns), dN[ch + 6]);
            int32_t probS1 = k_speech_w[ch + 6] * gaussian(feat[ch], hi16(wsm), hi16(wss), dS[ch + 6]);
            int32_t h0 = probN0 + probN1, h1 = probS0 + probS1;
            int16_t sh0 = (int16_t)(h0 ? norm_w32(h0) : 31), sh1 = (int16_t)(h1 ? norm_w32(h1) : 31);
            int16_t llr = (int16_t)(sh0 - sh1);
            sum_llr += (int32_t)(llr * k_spec_w[ch]);
            if ((llr << 2) > P.local_thr) vadflag = 1;
            int16_t q = (int16_t)(h0 >> 12);
            if (q > 0) {
                pN[ch] = (int16_t)div_w32_w16((int32_t)(((uint32_t)probN0 & 0xFFFFF000u) << 2), q);
                pN[ch + 6] = (int16_t)(16384 - pN[ch]);
            } else {
                pN[ch] = 16384;
            }
            q = (int16_t)(h1 >> 12);
            if (q > 0) {
                pS[ch] = (int16_t)div_w32_w16((int32_t)(((uint32_t)probS0 & 0xFFFFF000u) << 2), q);
                pS[ch + 6] = (int16_t)(16384 - pS[ch]);
            }
        }
        vadflag |= (sum_llr >= P.global_thr);

        int32_t frame_counter = st.get(W_FRAMES);
        int16_t maxspe = 12800;
        st.prefetch(W_AGE, 8);
        st.prefetch(W_LOW, 8);
        for (int ch = 0; ch < 6; ++ch) {
            if (ch < 5) {                                   // the next band's minimum lists arrive while this band is worked on
                st.prefetch(W_AGE + (ch + 1) * 8, 8);
                st.prefetch(W_LOW + (ch + 1) * 8, 8);
            }
            int16_t fmin = find_minimum(st, feat[ch], ch, frame_counter);
            int32_t wnm = st.get(W_NMEAN + ch), wsm = st.get(W_SMEAN + ch);
            int32_t wns = st.get(W_NSTD + ch), wss = st.get(W_SSTD + ch);
            int16_t nm[2] = {lo16(wnm), hi16(wnm)}, sm[2] = {lo16(wsm), hi16(wsm)};
            int16_t ns[2] = {lo16(wns), hi16(wns)}, ss[2] = {lo16(wss), hi16(wss)};
            const int16_t nw0 = k_noise_w[ch], nw1 = k_noise_w[ch + 6];
            const int16_t sw0 = k_speech_w[ch], sw1 = k_speech_w[ch + 6];
            int32_t ngm = shift_avg(nm, 0, nw0, nw1);
            int16_t ngm_q8 = (int16_t)(ngm >> 6);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                int g = ch + k * 6;
                int16_t nmk = nm[k], smk = sm[k], nsk = ns[k], ssk = ss[k];
                int16_t nmk2 = nmk, t16;
                if (!vadflag) {
                    int16_t d = (int16_t)((pN[g] * dN[g]) >> 11);
                    nmk2 = (int16_t)(nmk + (int16_t)((d * 655) >> 22));
                }
                int16_t nd = (int16_t)((fmin << 4) - ngm_q8);
                int16_t nmk3 = (int16_t)(nmk2 + (int16_t)((nd * 154) >> 9));
                int16_t lim = (int16_t)((k + 5) << 7);
                if (nmk3 < lim) nmk3 = lim;
                lim = (int16_t)((72 + k - ch) << 7);
                if (nmk3 > lim) nmk3 = lim;
                nm[k] = nmk3;
                if (vadflag) {
                    int16_t d = (int16_t)((pS[g] * dS[g]) >> 11);
                    t16 = (int16_t)((d * 6554) >> 21);
                    int16_t smk2 = (int16_t)(smk + ((t16 + 1) >> 1));
                    int16_t maxmu = (int16_t)(maxspe + 640);
                    int16_t minm = (k == 0) ? 640 : 768;
                    if (smk2 < minm) smk2 = minm;
                    if (smk2 > maxmu) smk2 = maxmu;
                    sm[k] = smk2;
                    t16 = (int16_t)((smk + 4) >> 3);
                    t16 = (int16_t)(feat[ch] - t16);
                    int32_t a32 = (dS[g] * t16) >> 3;
                    int32_t b32 = a32 - 4096;
                    t16 = (int16_t)(pS[g] >> 2);
                    a32 = wmul((int32_t)t16, b32);
                    b32 = a32 >> 4;
                    if (b32 > 0) t16 = (int16_t)div_w32_w16(b32, (int16_t)(ssk * 10));
                    else { t16 = (int16_t)div_w32_w16(wsub(0, b32), (int16_t)(ssk * 10)); t16 = (int16_t)-t16; }
                    t16 = (int16_t)(t16 + 128);
                    ssk = (int16_t)(ssk + (t16 >> 8));
                    if (ssk < 384) ssk = 384;
                    ss[k] = ssk;
                } else {
                    t16 = (int16_t)(feat[ch] - (nmk >> 3));
                    int32_t a32 = (dN[g] * t16) >> 3;
                    a32 -= 4096;
                    t16 = (int16_t)((pN[g] + 2) >> 2);
                    int32_t b32 = wmul((int32_t)t16, a32);
                    a32 = b32 >> 14;
                    if (a32 > 0) t16 = (int16_t)div_w32_w16(a32, nsk);
                    else { t16 = (int16_t)div_w32_w16(wsub(0, a32), nsk); t16 = (int16_t)-t16; }
                    t16 = (int16_t)(t16 + 32);
                    nsk = (int16_t)(nsk + (t16 >> 6));
                    if (nsk < 384) nsk = 384;
                    ns[k] = nsk;
                }
            }
            // keep the two models apart and inside their bounds (vad_core.c:406-457)
            ngm = shift_avg(nm, 0, nw0, nw1);
            int32_t sgm = shift_avg(sm, 0, sw0, sw1);
            int16_t diff = (int16_t)((int16_t)(sgm >> 9) - (int16_t)(ngm >> 9));
            if (diff < k_min_diff[ch]) {
                int16_t gap = (int16_t)(k_min_diff[ch] - diff);
                sgm = shift_avg(sm, (int16_t)((13 * gap) >> 2), sw0, sw1);
                ngm = shift_avg(nm, (int16_t)-(int16_t)((3 * gap) >> 2), nw0, nw1);
            }
            maxspe = k_max_speech[ch];
            int16_t t = (int16_t)(sgm >> 7);
            if (t > maxspe) {
                t = (int16_t)(t - maxspe);
                sm[0] = (int16_t)(sm[0] - t);
                sm[1] = (int16_t)(sm[1] - t);
            }
            t = (int16_t)(ngm >> 7);
            if (t > k_max_noise[ch]) {
                t = (int16_t)(t - k_max_noise[ch]);
                nm[0] = (int16_t)(nm[0] - t);
                nm[1] = (int16_t)(nm[1] - t);
            }
            st.set(W_NMEAN + ch, pack16(nm[0], nm[1]));
            st.set(W_SMEAN + ch, pack16(sm[0], sm[1]));
            st.set(W_NSTD + ch, pack16(ns[0], ns[1]));
            st.set(W_SSTD + ch, pack16(ss[0], ss[1]));
        }
        st.set(W_FRAMES, frame_counter + 1);
    }
    // hang-over smoothing (vad_core.c:462-477)
    int32_t hw = st.get(W_HANG);
    int16_t over_hang = lo16(hw), num_speech = hi16(hw);
    if (!vadflag) {
        if (over_hang > 0) { vadflag = (int16_t)(2 + over_hang); over_hang--; }
        num_speech = 0;
    } else {
        num_speech++;
        if (num_speech > 6) { num_speech = 6; over_hang = P.over_hang_2; }
        else over_hang = P.over_hang_1;
    }
    st.set(W_HANG, pack16(over_hang, num_speech));
    return vadflag;
}

// the wrapper's mute ramp (R:src/webrtc.c:138-141): every sample >> reduce, arithmetic.  A rolled loop: unrolled it was
// 570 instructions of a kernel that is bound by instruction fetch.
WMX_HD void attenuate(int16_t* x, int n, int reduce)
{
#pragma unroll 4
    for (int i = 0; i < n; ++i) x[i] = (int16_t)(x[i] >> reduce);
}

// One packet of one stream: WebRtcVad_Process (T:.../vad/webrtc_vad.c:71-105) + the wmix
// wrapper's mute ramp (R:src/webrtc.c:127-141).  `x` holds LEN8*(FS16?2:1) samples and is
// attenuated in place.  Returns the 0/1 decision.
// The packet in three stretches, so that a kernel may align the warps of a CTA between them (post_kernel):
// decimator + filterbank -> features, GMM decision, wrapper's mute ramp.
template <int LEN8, bool FS16>
WMX_HD int16_t packet_features(const SoaWords& st, const int16_t* x, int16_t* feat)
{
    st.prefetch(W_NMEAN, 24);                               // the GMM is read right after the filterbank
    st.prefetch(W_UPPER, 8);
    if (FS16) {
        int16_t nb[LEN8];
        int32_t s0 = st.get(W_DS), s1 = st.get(W_DS + 1);
        decimate<LEN8>(x, nb, s0, s1);
        st.set(W_DS, s0);
        st.set(W_DS + 1, s1);
        return features<LEN8>(st, nb, feat);
    }
    return features<LEN8>(st, x, feat);
}
template <int LEN8, bool FS16>
WMX_HD int packet_finish(const SoaWords& st, int16_t* x, int flag)
{
    int reduce = st.get(W_REDUCE);
    if (flag == 0) { if (reduce < 4) reduce++; }
    else if (reduce > 0) reduce--;
    st.set(W_REDUCE, reduce);
    const int n = LEN8 * (FS16 ? 2 : 1);
    attenuate(x, n, reduce);
    return flag > 0 ? 1 : 0;
}
template <int LEN8, bool FS16>
WMX_HD int process_packet(const SoaWords& st, int16_t* x, const Params& P)
{
    int16_t feat[6];
    const int16_t power = packet_features<LEN8, FS16>(st, x, feat);
    const int flag = gmm(st, feat, power, P);
    return packet_finish<LEN8, FS16>(st, x, flag);
}

// 32 kHz packet (T:.../vad/vad_core.c:623-643): 32k -> 16k -> 8k through the same decimator with two state pairs,
// then the 8 kHz detector; the wrapper attenuates all 4*LEN8 samples.
template <int LEN8>
WMX_HD int process_packet32(const SoaWords& st, int16_t* x, const Params& P)
{
    int16_t feat[6], wb[2 * LEN8], nb[LEN8];
    int32_t s0 = st.get(W_DS32), s1 = st.get(W_DS32 + 1);
    decimate<2 * LEN8>(x, wb, s0, s1);
    st.set(W_DS32, s0);
    st.set(W_DS32 + 1, s1);
    s0 = st.get(W_DS);
    s1 = st.get(W_DS + 1);
    decimate<LEN8>(wb, nb, s0, s1);
    st.set(W_DS, s0);
    st.set(W_DS + 1, s1);
    const int16_t power = features<LEN8>(st, nb, feat);
    int flag = gmm(st, feat, power, P);
    int reduce = st.get(W_REDUCE);
    if (flag == 0) { if (reduce < 4) reduce++; }
    else if (reduce > 0) reduce--;
    st.set(W_REDUCE, reduce);
    attenuate(x, 4 * LEN8, reduce);
    return flag > 0 ? 1 : 0;
}

}  // namespace vad
}  // namespace wmx
