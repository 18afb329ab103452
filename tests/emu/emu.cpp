// TEST INFRASTRUCTURE — lane-loop / thread-loop emulation of the kernel bodies on the CPU.
//
// Compiles the very same .cuh sources the sm_100a kernels are built from (wmix_b200/csrc/
// {ns,vad,agc,g711_mix}.cuh + host_tables.cpp) with g++, replacing "32 lanes in lock-step
// between warp barriers" by a loop over lanes per phase (WMX_NS_PHASE_* in ns.cuh) and "one
// thread per stream" by a plain call.  It exists so the kernel LOGIC can be checked against the
// reference in the GPU-less CI container; it is not part of libwmix_b200.so and nothing in the
// product can reach it.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../wmix_b200/csrc/aec.cuh"
#include "../../wmix_b200/csrc/agc.cuh"
#include "../../wmix_b200/csrc/g711_mix.cuh"
#include "../../wmix_b200/csrc/host_tables.h"
#include "../../wmix_b200/csrc/ns.cuh"
#include "../../wmix_b200/csrc/ns_cta.cuh"
#include "../../wmix_b200/csrc/nsx.cuh"
#include "../../wmix_b200/csrc/vad.cuh"

using namespace wmx;

template <int ANA>
static void fill_tables(ns::Tables<ANA>& T, int policy)
{
    memset(&T, 0, sizeof(T));
    host::dmath_tables(T.dm.log_invc, T.dm.log_logc, T.dm.exp_2jn);
    host::ns_window(ANA, ns::Geo<ANA>::kBlock, T.window);
    host::fft_w_table(ANA / 4, T.w);
    host::fft_c_table(ANA / 4, T.c);
    host::ns_log_table(ANA / 2 + 1, T.log_i, &T.sum_log_i, &T.sum_log_i_sq);
    host::ns_policy(policy, &T.overdrive, &T.floor_gain, &T.gainmap);
}

struct EmuNs {
    int ana;
    std::vector<float> rec, sh, hb;
    std::vector<uint16_t> hist;
    ns::Tables<256> t256;
    ns::Tables<128> t128;
    ns::Warp<256> w256;
    ns::Warp<128> w128;
};

// ---- CTA-cooperative NS (ns_cta.cuh): W worker warps + one reducer warp, the segments called in barrier order ----
template <int ANA>
struct EmuNsCtaT {
    typedef ns::Geo<ANA> G;
    int W;
    std::vector<float> rec, tiles;
    std::vector<uint16_t> hist;
    std::vector<uint16_t*> hptr;
    std::vector<ns::WWarp<ANA>> ww;
    ns::RWarp rw;
    ns::Tables<ANA> T;
    bool defer_first = false;
    explicit EmuNsCtaT(int w) : W(w), rec((size_t)w * G::kRecFloats, 0.f), tiles((size_t)8 * G::kShFloats, 0.f), hist((size_t)w * 3 * ns::kHistBins, 0),
                                hptr(8, nullptr), ww(w)
    {
        fill_tables(T, 2);
        memset(&rw, 0, sizeof rw);
        for (int j = 0; j < w; ++j) {
            for (int l = 0; l < 32; ++l) ns::init_record_values<ANA>(rec.data() + (size_t)j * G::kRecFloats, l, 32);
            hptr[j] = hist.data() + (size_t)j * 3 * ns::kHistBins;
        }
    }
    // in / out: [W][kBlock]; live[j] = 0 leaves worker j idle this frame (a CTA at the ragged end of the stream range)
    void frame(const int16_t* in, int16_t* out, const uint8_t* live)
    {
        bool act[8] = {};
        for (int j = 0; j < W; ++j) {
            float* sh = tiles.data() + (size_t)j * G::kShFloats;
            if (live && !live[j]) { sh[G::kShScal + ns::C_ACTIVE] = 0.f; continue; }
            act[j] = ns::w_seg1<ANA>(ww[j], rec.data() + (size_t)j * G::kRecFloats, in + (size_t)j * G::kBlock, out + (size_t)j * G::kBlock, sh, T);
        }
        ns::r_seg1<ANA>(rw, tiles.data(), G::kShFloats, W, T);
        // r_seg1b runs concurrently with the workers' segment 2 on the device; `defer_first` picks which side goes first here, and
        // the tests run both orders.  (r_seg2b would be legal behind w_seg3a too — kept in the second order as a check of that.)
        if (defer_first) ns::r_seg1b<ANA>(rw, tiles.data(), G::kShFloats, T);
        for (int j = 0; j < W; ++j)
            if (act[j]) ns::w_seg2<ANA>(ww[j], rec.data() + (size_t)j * G::kRecFloats, tiles.data() + (size_t)j * G::kShFloats, T);
        if (!defer_first) ns::r_seg1b<ANA>(rw, tiles.data(), G::kShFloats, T);
        ns::r_seg2a<ANA>(rw, tiles.data(), G::kShFloats, hptr.data(), T);
        if (defer_first) ns::r_seg2b<ANA>(rw, tiles.data(), G::kShFloats, T);
        for (int j = 0; j < W; ++j)
            if (act[j]) ns::w_seg3a<ANA>(ww[j], rec.data() + (size_t)j * G::kRecFloats, hptr[j], tiles.data() + (size_t)j * G::kShFloats, T);
        if (!defer_first) ns::r_seg2b<ANA>(rw, tiles.data(), G::kShFloats, T);
        for (int j = 0; j < W; ++j)
            if (act[j]) ns::w_seg3b<ANA>(ww[j], rec.data() + (size_t)j * G::kRecFloats, tiles.data() + (size_t)j * G::kShFloats, T);
        ns::r_seg3<ANA>(rw, tiles.data(), G::kShFloats, T);
        for (int j = 0; j < W; ++j)
            if (act[j]) ns::w_seg4<ANA>(ww[j], rec.data() + (size_t)j * G::kRecFloats, out + (size_t)j * G::kBlock, tiles.data() + (size_t)j * G::kShFloats, T);
    }
};
struct EmuNsCta {
    int ana;
    EmuNsCtaT<256>* e256 = nullptr;
    EmuNsCtaT<128>* e128 = nullptr;
};

extern "C" {

void* emu_ns_create(int freq)
{
    EmuNs* e = new EmuNs();
    e->ana = freq == 8000 ? 128 : 256;
    e->hist.assign(3 * ns::kHistBins, 0);
    if (e->ana == 256) {
        fill_tables(e->t256, 2);
        e->rec.assign(ns::Geo<256>::kRecFloats, 0.f);
        e->sh.assign(ns::Geo<256>::kShFloats + ns::Geo<256>::kBlock, 0.f);
        e->hb.assign(ns::Geo<256>::kOverlap, 0.f);
        for (int l = 0; l < 32; ++l) ns::init_record_values<256>(e->rec.data(), l, 32);
    } else {
        fill_tables(e->t128, 2);
        e->rec.assign(ns::Geo<128>::kRecFloats, 0.f);
        e->sh.assign(ns::Geo<128>::kShFloats + ns::Geo<128>::kBlock, 0.f);
        e->hb.assign(ns::Geo<128>::kOverlap, 0.f);
        for (int l = 0; l < 32; ++l) ns::init_record_values<128>(e->rec.data(), l, 32);
    }
    return e;
}
void emu_ns_frame(void* h, const int16_t* in, int16_t* out)
{
    EmuNs* e = (EmuNs*)h;
    if (e->ana == 256) ns::frame<256>(e->w256, e->rec.data(), e->hist.data(), in, out, e->sh.data(), e->t256);
    else ns::frame<128>(e->w128, e->rec.data(), e->hist.data(), in, out, e->sh.data(), e->t128);
}
// wmix's stereo NS: the right channel as a second band (ns.cuh, frame<ANA, true>)
void emu_ns_frame_hb(void* h, const int16_t* in, int16_t* out, const int16_t* in_hb, int16_t* out_hb)
{
    EmuNs* e = (EmuNs*)h;
    if (e->ana == 256) ns::frame<256, true>(e->w256, e->rec.data(), e->hist.data(), in, out, e->sh.data(), e->t256, e->hb.data(), in_hb, out_hb);
    else ns::frame<128, true>(e->w128, e->rec.data(), e->hist.data(), in, out, e->sh.data(), e->t128, e->hb.data(), in_hb, out_hb);
}
void emu_ns_destroy(void* h) { delete (EmuNs*)h; }
const float* emu_ns_record(void* h) { return ((EmuNs*)h)->rec.data(); }

struct EmuInt {
    int freq;
    std::vector<int32_t> vad, agc;
    int32_t table[32];
    vad::Params vp;
};

void* emu_int_create(int freq, int agc_gain_db, int vad_mode)
{
    EmuInt* e = new EmuInt();
    e->freq = freq;
    e->vad.assign(vad::N_WORDS, 0);
    e->agc.assign(agc::N_WORDS, 0);
    host::vad_initial_words(e->vad.data());
    host::agc_initial_words(e->agc.data());
    host::agc_gain_table(e->table, (int16_t)agc_gain_db, 0, 0, host::agc_analog_target((int16_t)agc_gain_db));
    int16_t th[4];
    host::vad_thresholds(vad_mode, 10, th);
    e->vp = {th[0], th[1], th[2], th[3]};
    return e;
}
void emu_agc_frame(void* h, int16_t* x)
{
    EmuInt* e = (EmuInt*)h;
    SoaWords st{e->agc.data(), 1};
    if (e->freq == 16000) agc::process_packet<true>(st, x, e->table);
    else agc::process_packet<false>(st, x, e->table);
}
int emu_vad_frame(void* h, int16_t* x)
{
    EmuInt* e = (EmuInt*)h;
    SoaWords st{e->vad.data(), 1};
    if (e->freq == 16000) return vad::process_packet<80, true>(st, x, e->vp);
    return vad::process_packet<80, false>(st, x, e->vp);
}
// 20 ms packets (thresholds of the 20 ms column): x holds 160 (8 kHz) / 320 (16 kHz) samples
int emu_vad_frame20(void* h, int16_t* x, int vad_mode)
{
    EmuInt* e = (EmuInt*)h;
    SoaWords st{e->vad.data(), 1};
    int16_t th[4];
    host::vad_thresholds(vad_mode, 20, th);
    const vad::Params vp{th[0], th[1], th[2], th[3]};
    if (e->freq == 16000) return vad::process_packet<160, true>(st, x, vp);
    return vad::process_packet<160, false>(st, x, vp);
}
// 32 kHz packets of 10 ms: x holds 320 samples (the emulated engine must have been created at 16 kHz)
int emu_vad_frame32(void* h, int16_t* x)
{
    EmuInt* e = (EmuInt*)h;
    SoaWords st{e->vad.data(), 1};
    return vad::process_packet32<80>(st, x, e->vp);
}
void emu_int_destroy(void* h) { delete (EmuInt*)h; }

struct EmuAec {
    int mult, depth;
    std::vector<float> rec, sh;
    aec::Tables T;
    aec::Warp W;
};
void* emu_aec_create(int freq, int depth)
{
    EmuAec* e = new EmuAec();
    e->mult = freq / 8000;
    e->depth = depth;
    memset(&e->T, 0, sizeof e->T);
    host::dmath_tables(e->T.dm.log_invc, e->T.dm.log_logc, e->T.dm.exp_2jn);
    host::aec_tables(e->T.w, e->T.c, e->T.hann, e->T.weight, e->T.over, e->T.lcg_mul, e->T.lcg_add);
    e->rec.assign(aec::rec_floats(depth), 0.f);
    e->sh.assign(aec::Geo::kShFloats, 0.f);
    for (int l = 0; l < 32; ++l) aec::init_record_values(e->rec.data(), l, 32);
    return e;
}
// far / near may be NULL (aec_setFrameFar / aec_process); n = 80 or 160 samples
void emu_aec_tick(void* h, const int16_t* far, const int16_t* near, int16_t* out, int n, int delay_ms)
{
    EmuAec* e = (EmuAec*)h;
    aec::tick(e->W, e->rec.data(), e->depth, e->mult, n, far, near, out, delay_ms, e->sh.data(), e->T);
}
int emu_aec_error(void* h) { return ns::f2i(((EmuAec*)h)->rec[aec::Geo::kOffScal + aec::S_ERROR]); }
int emu_aec_scalar(void* h, int id) { return ns::f2i(((EmuAec*)h)->rec[aec::Geo::kOffScal + id]); }
void emu_aec_destroy(void* h) { delete (EmuAec*)h; }
void emu_aec_rdft(float* a, int inverse)
{
    EmuAec* e = (EmuAec*)emu_aec_create(8000, 4);
    float* sp = e->sh.data() + aec::Geo::kShSp;
    memcpy(sp, a, 128 * sizeof(float));
    if (inverse) aec::rdft_inv(e->W, sp, (float*)nullptr, e->sh.data() + aec::Geo::kShX, e->T, 1);
    else aec::rdft_fwd(e->W, sp, (float*)nullptr, e->sh.data() + aec::Geo::kShX, e->T, 1);
    memcpy(a, sp, 128 * sizeof(float));
    delete e;
}

void emu_g711(const int16_t* pcm, int n, uint8_t* alaw, uint8_t* ulaw)
{
    for (int i = 0; i < n; ++i) { alaw[i] = linear2alaw(pcm[i]); ulaw[i] = linear2ulaw(pcm[i]); }
}
void emu_g711_dec(const uint8_t* codes, int n, int16_t* a, int16_t* u)
{
    for (int i = 0; i < n; ++i) { a[i] = alaw2linear(codes[i]); u[i] = ulaw2linear(codes[i]); }
}
int16_t emu_mix_step(int16_t bus, int16_t src, int rdce) { return mix_step(bus, src, rdce); }
void* emu_nscta_create(int freq, int workers)
{
    if (workers < 1 || workers > 8) return nullptr;
    EmuNsCta* e = new EmuNsCta();
    e->ana = freq == 8000 ? 128 : 256;
    if (e->ana == 256) e->e256 = new EmuNsCtaT<256>(workers);
    else e->e128 = new EmuNsCtaT<128>(workers);
    return e;
}
void emu_nscta_order(void* h, int defer_first)
{
    EmuNsCta* e = (EmuNsCta*)h;
    if (e->e256) e->e256->defer_first = defer_first != 0;
    else e->e128->defer_first = defer_first != 0;
}
void emu_nscta_frame(void* h, const int16_t* in, int16_t* out, const uint8_t* live)
{
    EmuNsCta* e = (EmuNsCta*)h;
    if (e->e256) e->e256->frame(in, out, live);
    else e->e128->frame(in, out, live);
}
const float* emu_nscta_record(void* h, int j)
{
    EmuNsCta* e = (EmuNsCta*)h;
    return e->e256 ? e->e256->rec.data() + (size_t)j * ns::Geo<256>::kRecFloats : e->e128->rec.data() + (size_t)j * ns::Geo<128>::kRecFloats;
}
const uint16_t* emu_nscta_hist(void* h, int j)
{
    EmuNsCta* e = (EmuNsCta*)h;
    return (e->e256 ? e->e256->hist.data() : e->e128->hist.data()) + (size_t)j * 3 * ns::kHistBins;
}
void emu_nscta_destroy(void* h)
{
    EmuNsCta* e = (EmuNsCta*)h;
    delete e->e256;
    delete e->e128;
    delete e;
}
// ---- fixed-point NS (nsx.cuh): one warp per stream, phases as lane loops, warp reductions as loops over the lanes ----
struct EmuNsx {
    int ana;
    std::vector<uint32_t> rec, tile;
    std::vector<int16_t> hist;
    nsx::Tables T;
    nsx::Warp<256> w256;
    nsx::Warp<128> w128;
};
void* emu_nsx_create(int freq, int policy)
{
    EmuNsx* e = new EmuNsx();
    int32_t thr = 0;
    if (host::nsx_tables(freq, policy, &e->T, &thr) != 0) { delete e; return nullptr; }
    e->ana = freq == 8000 ? 128 : 256;
    e->hist.assign(3 * nsx::kHistBins, 0);
    memset(&e->w256, 0, sizeof e->w256);
    memset(&e->w128, 0, sizeof e->w128);
    if (e->ana == 256) {
        e->rec.assign(nsx::Geo<256>::kRecWords, 0u);
        e->tile.assign(nsx::Geo<256>::kShWords, 0u);
        nsx::init_record<256>(e->rec.data(), e->hist.data(), 0, 1, thr);
    } else {
        e->rec.assign(nsx::Geo<128>::kRecWords, 0u);
        e->tile.assign(nsx::Geo<128>::kShWords, 0u);
        nsx::init_record<128>(e->rec.data(), e->hist.data(), 0, 1, thr);
    }
    return e;
}
void emu_nsx_frame(void* h, const int16_t* in, int16_t* out)
{
    EmuNsx* e = (EmuNsx*)h;
    if (e->ana == 256) nsx::frame<256>(e->w256, e->rec.data(), e->hist.data(), in, out, e->tile.data(), e->T);
    else nsx::frame<128>(e->w128, e->rec.data(), e->hist.data(), in, out, e->tile.data(), e->T);
}
// two bands (wmix's stereo): hb = the second band's delay line, int16 [kKeep], owned by the caller
void emu_nsx_frame_hb(void* h, const int16_t* in, int16_t* out, int16_t* hb, const int16_t* in_hb, int16_t* out_hb)
{
    EmuNsx* e = (EmuNsx*)h;
    if (e->ana == 256) {
        const int g = nsx::frame<256, true>(e->w256, e->rec.data(), e->hist.data(), in, out, e->tile.data(), e->T);
        nsx::second_band<256>(e->w256, hb, in_hb, out_hb, g);
    } else {
        const int g = nsx::frame<128, true>(e->w128, e->rec.data(), e->hist.data(), in, out, e->tile.data(), e->T);
        nsx::second_band<128>(e->w128, hb, in_hb, out_hb, g);
    }
}
void emu_nsx_destroy(void* h) { delete (EmuNsx*)h; }
const uint32_t* emu_nsx_record(void* h) { return ((EmuNsx*)h)->rec.data(); }
const int16_t* emu_nsx_hist(void* h) { return ((EmuNsx*)h)->hist.data(); }
int emu_nsx_rec_words(int freq) { return freq == 8000 ? nsx::Geo<128>::kRecWords : nsx::Geo<256>::kRecWords; }
}   // extern "C"
// the record in the canonical order of oracle/orc_nsx.c's orc_nsx_state (tests compare the two word for word)
template <int ANA>
static int nsx_canonical(const uint32_t* rec, int32_t* out)
{
    typedef nsx::Geo<ANA> G;
    const int32_t* sc = (const int32_t*)(rec + G::kOffScal);
    auto word = [&](int a, int bin) { return bin < G::kHalf ? rec[G::kOffArrays + a * G::kHalf + bin] : rec[G::kOffNyq + a]; };
    int k = 0;
    const int order[25] = {nsx::S_FRAME_IDX, nsx::S_MODEL_COUNT, nsx::S_COUNTER0, nsx::S_COUNTER1, nsx::S_COUNTER2, nsx::S_Q_NOISE, nsx::S_Q_NOISE_PREV,
                           nsx::S_Q_MAGN_PREV, nsx::S_MIN_NORM, nsx::S_PRIOR, nsx::S_FEAT_LRT, nsx::S_THR_LRT, nsx::S_FEAT_FLAT, nsx::S_THR_FLAT,
                           nsx::S_FEAT_DIFF, nsx::S_THR_DIFF, nsx::S_W_LRT, nsx::S_W_FLAT, nsx::S_W_DIFF, nsx::S_CUR_AVG_E, nsx::S_TIME_AVG_E,
                           nsx::S_TIME_AVG_ACC, nsx::S_WHITE, nsx::S_PINK_NUM, nsx::S_PINK_EXP};
    for (int i = 0; i < 25; ++i) out[k++] = sc[order[i]];
    while (k < 40) out[k++] = 0;
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)(word(nsx::A_QF, b) >> 16);
    for (int e = 0; e < 3; ++e)
        for (int b = 0; b < G::kBins; ++b) out[k++] = lo16((int32_t)word(nsx::A_LQD0 + e, b));
    for (int e = 0; e < 3; ++e)
        for (int b = 0; b < G::kBins; ++b) out[k++] = hi16((int32_t)word(nsx::A_LQD0 + e, b));
    for (int b = 0; b < G::kBins; ++b) out[k++] = lo16((int32_t)word(nsx::A_QF, b));
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)word(nsx::A_LRT, b);
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)word(nsx::A_PAUSE, b);
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)word(nsx::A_INIT, b);
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)word(nsx::A_NPREV, b);
    for (int b = 0; b < G::kBins; ++b) out[k++] = (int32_t)(word(nsx::A_MPREV, b) & 0xFFFFu);
    const int16_t* hist = (const int16_t*)(rec + G::kOffHist);
    const int16_t* syn = (const int16_t*)(rec + G::kOffSyn);
    for (int i = 0; i < G::kKeep; ++i) out[k++] = hist[i];
    for (int i = 0; i < G::kKeep; ++i) out[k++] = syn[i];
    return k;
}
extern "C" {
int emu_nsx_canonical(int freq, const uint32_t* rec, int32_t* out) { return freq == 8000 ? nsx_canonical<128>(rec, out) : nsx_canonical<256>(rec, out); }
// the parallel two-peak search against the reference's scan order: h int16 [1000] -> pos1, pos2, w1, w2
void emu_nsx_two_peaks(const int16_t* h, int32_t* out4)
{
    static nsx::Warp<256> w;
    std::vector<uint32_t> tile(nsx::Geo<256>::kShWords, 0u);
    const nsx::Peaks pk = nsx::two_peaks<256>(w, h, tile.data());
    out4[0] = (int32_t)pk.pos1; out4[1] = (int32_t)pk.pos2; out4[2] = pk.w1; out4[3] = pk.w2;
}
int emu_nsx_tables(int freq, int policy, void* out, int cap)
{
    nsx::Tables T;
    int32_t thr = 0;
    if (cap < (int)sizeof T || host::nsx_tables(freq, policy, &T, &thr) != 0) return -(int)sizeof T;
    memcpy(out, &T, sizeof T);
    return (int)sizeof T;
}
// the VAD's minimum tracker on its own (vad::find_minimum on the packed lists): a feature sequence for one channel, the
// smoothed value per step in `out`, the channel's 8 + 8 list words and the mean-value word at the end in `words_out[17]`
void emu_vad_minimum_run(const int16_t* feats, int n, int ch, int16_t* out, int32_t* words_out)
{
    std::vector<int32_t> w(vad::N_WORDS, 0);
    host::vad_initial_words(w.data());
    SoaWords st{w.data(), 1};
    for (int t = 0; t < n; ++t) out[t] = vad::find_minimum(st, feats[t], ch, t);
    for (int k = 0; k < 8; ++k) { words_out[k] = w[vad::W_AGE + ch * 8 + k]; words_out[8 + k] = w[vad::W_LOW + ch * 8 + k]; }
    words_out[16] = w[vad::W_MEANVAL + (ch >> 1)];
}
int emu_ns_rec_floats(int freq) { return freq == 8000 ? ns::Geo<128>::kRecFloats : ns::Geo<256>::kRecFloats; }
const uint16_t* emu_ns_hist(void* h) { return ((EmuNs*)h)->hist.data(); }
int emu_agc_gain_table(int32_t* t, int comp, int target, int lim, int at) { return host::agc_gain_table(t, (int16_t)comp, (int16_t)target, lim, (int16_t)at); }
int emu_agc_analog_target(int comp) { return host::agc_analog_target((int16_t)comp); }
void emu_ns_window(int ana, int block, float* w) { host::ns_window(ana, block, w); }
// ns::div_by_counter against the IEEE division: every `stride`-th float significand at three binades,
// divisors d_lo..d_hi; returns the number of disagreements
long emu_div_by_counter_mismatches(int d_lo, int d_hi, int stride)
{
    long bad = 0;
    const int exps[3] = {127 - 3, 127 + 2, 127 + 13};      // 0.125.., 4.., 8192..: the ranges the trackers produce
    for (int d = d_lo; d <= d_hi; ++d) {
        const float cf = (float)d, rc = 1.f / cf;
        for (int e = 0; e < 3; ++e)
            for (uint32_t m = 0; m < (1u << 23); m += (uint32_t)stride) {
                const float x = ns::i2f((int32_t)(((uint32_t)exps[e] << 23) | m));
                bad += (x / cf != ns::div_by_counter(x, cf, rc));
            }
    }
    return bad;
}
uint32_t emu_zoom_map(int ic, int ifr, uint32_t in_bytes, int oc, int ofr, int32_t* map) { return host::zoom_map(ic, ifr, in_bytes, oc, ofr, map); }
uint32_t emu_mix_plan(int chn, int freq, uint32_t src_bytes, int mix_freq, int32_t* map, uint16_t* ramp) { return host::mix_plan(chn, freq, src_bytes, mix_freq, map, ramp); }
int emu_play_fifo_slot(int count, int n_pkg, int delay_pkgs) { return host::play_fifo_slot(count, n_pkg, delay_pkgs); }
uint32_t emu_len_of_out(int ic, int ifr, uint32_t n, int oc, int ofr) { return host::zoom_len_of_out(ic, ifr, n, oc, ofr); }
uint32_t emu_len_of_in(int ic, int ifr, int oc, int ofr, uint32_t n) { return host::zoom_len_of_in(ic, ifr, oc, ofr, n); }
void emu_logexp(const float* x, int n, float* lg, float* ex)
{
    static ns::DMath dm;
    static bool init = false;
    if (!init) { host::dmath_tables(dm.log_invc, dm.log_logc, dm.exp_2jn); init = true; }
    for (int i = 0; i < n; ++i) { if (lg) lg[i] = ns::log_f(x[i], dm); if (ex) ex[i] = ns::exp_f(x[i], dm); }
}
}
