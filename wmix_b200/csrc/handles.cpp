// Drop-in handle layer: include/webrtc.h and include/g711codec.h on top of wmixb.
// Mirrors the marshaling of R:src/webrtc.c (channel averaging / replication, packet loop,
// early return on error) on the host; the DSP itself is a one-stream wmixb engine on the GPU.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/g711codec.h"
#include "../../include/webrtc.h"
#include "../../include/wmix_zoom.h"
#include "../../include/wmixb.h"
#include "host_tables.h"
#include "scratch.h"

namespace {

struct Handle {
    wmixb_engine* eng = nullptr;
    int chn = 1, freq = 0, pkg = 0, stage = 0;
    bool* debug = nullptr;
    bool bad_rate = false;   // 24 kHz: the reference's vad_init / agc_init hand out a handle whose every process call fails
    std::vector<int16_t> mono, res, far, hb, hb_res;   // hb: right channel of a stereo NS handle
};

bool dbg(const bool* d) { return d && *d; }

// 32 kHz handles run on a 16 kHz engine: at that rate the reference's AGC and NS see 160-sample packets and take
// exactly their 16 kHz paths, and the VAD only adds a 32k -> 16k decimator in front (wmixb_vad32_*).
Handle* make(int stage, int chn, int freq, int gain, bool* debug, const char* who, bool ns_high_band = false)
{
    wmixb_config c;
    memset(&c, 0, sizeof c);
    c.n_streams = 1;
    c.freq = (freq == 32000 || freq == 24000) ? 16000 : freq;
    c.stages = stage;
    c.ns_policy = 2;      // NS_AGGRESSIVE, R:src/webrtc.c:532
    c.agc_gain_db = gain; // compressionGaindB, R:src/webrtc.c:707
    c.vad_mode = 3;       // VAD_AGGRESSIVE, R:src/webrtc.c:16
    c.ns_high_band = ns_high_band ? 1 : 0;
    c.ns_core = (stage & WMIXB_NS) ? wmixb_default_ns_core() : 0;   // the reference's MAKE_WEBRTC_NSX switch, R:src/webrtc.c:511-523
    c.device = wmixb_default_device();
    wmixb_engine* e = nullptr;
    if (wmixb_create(&c, &e) != WMIXB_OK) {
        if (dbg(debug)) printf("%s failed !! (%s)\r\n", who, wmixb_last_error());
        return nullptr;
    }
    Handle* h = new Handle();
    h->eng = e;
    h->chn = chn;
    h->freq = freq;
    h->pkg = c.freq / 100;
    h->stage = stage;
    h->bad_rate = freq == 24000;
    h->debug = debug;
    h->mono.resize((size_t)h->pkg);
    h->res.resize((size_t)h->pkg);
    return h;
}

void drop(void* fp, const char* who)
{
    Handle* h = (Handle*)fp;
    wmixb_destroy(h->eng);
    if (dbg(h->debug)) printf("%s\r\n", who);
    delete h;
}

bool rate_ok(int freq, int max) { return !(freq > max || freq % 8000 != 0); }

}  // namespace

extern "C" {

// ---------------------------------------------------------------- VAD
void* vad_init(int chn, int freq, int intervalMs, bool* debug)
{
    if (!rate_ok(freq, 32000)) return nullptr;                      // R:src/webrtc.c:43
    Handle* h = make(WMIXB_VAD, chn, freq, 0, debug, "vad_init");
    if (!h) return nullptr;
    const int ms = (freq <= 16000 && intervalMs % 20 == 0) ? 20 : 10;   // R:src/webrtc.c:56-67
    h->pkg = freq / 1000 * ms;
    h->mono.resize((size_t)h->pkg);
    h->res.resize((size_t)h->pkg);
    if (dbg(debug)) printf("vad_init: chn/%d freq/%d intervalMs/%d pkgFrame/%d\r\n", chn, freq, ms, h->pkg);
    return h;
}

void vad_process(void* fp, int16_t* frame, int frameNum)
{
    Handle* h = (Handle*)fp;
    const int total = frameNum * h->chn;
    int mono = total;
    if (h->chn > 1) {                                               // R:src/webrtc.c:107-118
        int pos = 0;
        for (mono = 0; pos < total;) {
            int32_t s = 0;
            for (int c = 0; c < h->chn; ++c) s += frame[pos++];
            frame[mono++] = (int16_t)(s / h->chn);
        }
    }
    if (h->bad_rate) {
        // 24 kHz passes vad_init's rate test (R:src/webrtc.c:43) but WebRtcVad_Process refuses the rate: the reference returns
        // here, with the frame untouched — or, for a stereo frame, with its first half already averaged to mono and never
        // expanded again (R:src/webrtc.c:121-127)
        if (dbg(h->debug)) printf("WebRtcVad_Process failed !!, ret %d \r\n", -1);
        return;
    }
    // R:src/webrtc.c:120-141 always hands the detector the START of the buffer and attenuates
    // [cLen, pkgFrame): with more than one packet per call only the first is ever touched.
    for (int pos = 0; pos < mono; pos += h->pkg) {
        memcpy(h->mono.data(), frame, (size_t)h->pkg * 2);
        uint8_t flag = 0;
        int rc;
        if (h->freq == 32000) {                                     // 320-sample packet, CalcVad32khz
            memcpy(h->res.data(), h->mono.data(), (size_t)h->pkg * 2);
            rc = wmixb_vad32_host(h->eng, h->res.data(), &flag);
        } else if (h->pkg == h->freq / 100) {
            rc = wmixb_tick_host(h->eng, h->mono.data(), h->res.data(), &flag, WMIXB_VAD);
        } else {                                                    // 20 ms packet
            memcpy(h->res.data(), h->mono.data(), (size_t)h->pkg * 2);
            rc = wmixb_vad20_host(h->eng, h->res.data(), &flag);
        }
        if (rc != WMIXB_OK) {
            if (dbg(h->debug)) printf("WebRtcVad_Process failed !!, %s \r\n", wmixb_last_error());
            return;
        }
        for (int i = pos; i < h->pkg; ++i) frame[i] = h->res[(size_t)i];
    }
    if (h->chn > 1) {                                               // R:src/webrtc.c:144-150
        int m = mono - 1;
        for (int pos = total - 1; pos >= 0; --m)
            for (int c = 0; c < h->chn; ++c) frame[pos--] = frame[m];
    }
}

void vad_release(void* fp) { drop(fp, "vad_release"); }

// ---------------------------------------------------------------- NS
void* ns_init(int chn, int freq, bool* debug)
{
    if (!rate_ok(freq, 32000)) return nullptr;                      // R:src/webrtc.c:563
    if (chn != 1 && chn != 2) return nullptr;                        // the reference only has in[2] / out[2]
    if (freq == 24000) return nullptr;                               // passes the rate test, WebRtcNs(x)_Init then refuses it (R:src/webrtc.c:569)
    // two channels: the right one is WebRtcNs's "high band" (num_bands = chn, R:src/webrtc.c:624-636) -> wmixb_ns2_host
    Handle* h = make(WMIXB_NS, chn, freq, 0, debug, "ns_init", chn == 2);
    if (!h) return nullptr;
    if (freq == 32000) {
        // WebRtcNs at 32 kHz still works on 160-sample blocks with the 16 kHz window (T:.../ns/ns_core.c:89-98) and
        // wmix hands it the channels as bands, so of each 320-sample packet only the first 160 samples are analysed
        // and written; the rest of the reference's calloc'ed out buffers stays zero.
        h->pkg = 320;
    }
    h->mono.assign((size_t)h->pkg, 0);
    h->res.assign((size_t)h->pkg, 0);
    h->hb.assign((size_t)h->pkg, 0);
    h->hb_res.assign((size_t)h->pkg, 0);
    if (dbg(debug)) printf("ns_init: chn/%d freq/%d intervalMs/%d pkgFrame/%d x %d\r\n", chn, freq, 10, h->pkg, chn);
    return h;
}

void ns_process(void* fp, int16_t* frame, int16_t* frameOut, int frameNum)
{
    Handle* h = (Handle*)fp;
    for (int pos = 0; pos < frameNum; pos += h->pkg) {              // R:src/webrtc.c:624-643
        int rc;
        // (a 32 kHz packet: the engine's frame is the first 160 samples, res[160..319] stays zero)
        if (h->chn == 1) {
            memcpy(h->mono.data(), frame + pos, (size_t)h->pkg * 2);
            rc = wmixb_tick_host(h->eng, h->mono.data(), h->res.data(), nullptr, WMIXB_NS);
        } else {
            for (int i = 0; i < h->pkg; ++i) { h->mono[(size_t)i] = frame[2 * (pos + i)]; h->hb[(size_t)i] = frame[2 * (pos + i) + 1]; }
            rc = wmixb_ns2_host(h->eng, h->mono.data(), h->hb.data(), h->res.data(), h->hb_res.data());
        }
        if (rc != WMIXB_OK) {
            if (dbg(h->debug)) printf("ns_process failed !!, %s \r\n", wmixb_last_error());
            return;
        }
        if (h->chn == 1) memcpy(frameOut + pos, h->res.data(), (size_t)h->pkg * 2);
        else for (int i = 0; i < h->pkg; ++i) { frameOut[2 * (pos + i)] = h->res[(size_t)i]; frameOut[2 * (pos + i) + 1] = h->hb_res[(size_t)i]; }
    }
}

void ns_release(void* fp) { drop(fp, "ns_release"); }

// ---------------------------------------------------------------- AGC
void* agc_init(int chn, int freq, int intervalMs, int value, bool* debug)
{
    (void)intervalMs;
    if (!rate_ok(freq, 32000)) return nullptr;                      // R:src/webrtc.c:711
    // 32 kHz: 5 ms packets of 160 samples through the 16 kHz path (R:src/webrtc.c:724-728; WebRtcAgc_Process accepts 160
    // samples at 16 / 32 / 48 kHz alike, T:.../agc/legacy/analog_agc.c:1152-1163, and nothing else reads fs in wmix's use)
    Handle* h = make(WMIXB_AGC, chn, freq, value, debug, "agc_init");
    if (h && dbg(debug)) printf("agc_init: chn/%d freq/%d intervalMs/%d pkgFrame/%d x %d\r\n", chn, freq, freq <= 16000 ? 10 : 5, h->pkg, chn);
    return h;
}

int agc_process(void* fp, int16_t* frame, int16_t* frameOut, int frameNum)
{
    Handle* h = (Handle*)fp;
    const int total = frameNum * h->chn, step = h->pkg * h->chn;
    if (h->bad_rate && total > 0) {
        // 24 kHz: agc_init succeeds in the reference, WebRtcAgc_Process then fails on the packet length and agc_process returns
        // its code with frameOut untouched (R:src/webrtc.c:806-811)
        if (dbg(h->debug)) printf("WebRtcAgc_Process failed !!, ret %d \r\n", -1);
        return -1;
    }
    for (int pos = 0; pos < total; pos += step) {                   // R:src/webrtc.c:786-818
        for (int i = 0; i < h->pkg; ++i) {
            int32_t s = 0;
            for (int c = 0; c < h->chn; ++c) s += *frame++;
            h->mono[(size_t)i] = (int16_t)(s / h->chn);
        }
        if (wmixb_tick_host(h->eng, h->mono.data(), h->res.data(), nullptr, WMIXB_AGC) != WMIXB_OK) {
            if (dbg(h->debug)) printf("WebRtcAgc_Process failed !!, %s \r\n", wmixb_last_error());
            return -1;
        }
        for (int i = 0; i < h->pkg; ++i)
            for (int c = 0; c < h->chn; ++c) *frameOut++ = h->res[(size_t)i];
    }
    return 0;
}

void agc_addition(void* fp, uint8_t value)
{
    Handle* h = (Handle*)fp;
    if (wmixb_set_agc_gain(h->eng, value) != WMIXB_OK && dbg(h->debug))
        printf("WebRtcAgc_set_config failed !!, %s \r\n", wmixb_last_error());
}

void agc_release(void* fp) { drop(fp, "agc_release"); }

// ---------------------------------------------------------------- AEC
// R:src/webrtc.c:217-505.  The engine keeps the reference's full far-end history (250 partitions + the
// two the block being read spans), so any cadence of aec_setFrameFar / aec_process calls the reference
// accepts behaves the same here.
void* aec_init(int chn, int freq, int intervalMs, bool* debug)
{
    if (!rate_ok(freq, 16000)) return nullptr;                      // R:src/webrtc.c:220
    wmixb_config c;
    memset(&c, 0, sizeof c);
    c.n_streams = 1;
    c.freq = freq;
    c.stages = WMIXB_AEC;
    c.aec_far_depth = 252;
    wmixb_engine* e = nullptr;
    if (wmixb_create(&c, &e) != WMIXB_OK) {
        if (dbg(debug)) printf("WebRtcAecX_Create failed !! (%s)\r\n", wmixb_last_error());
        return nullptr;
    }
    Handle* h = new Handle();
    h->eng = e;
    h->chn = chn;
    h->freq = freq;
    h->stage = WMIXB_AEC;
    h->debug = debug;
    // 10 ms packets, or 20 ms at 8 kHz when the caller's interval is a multiple of 20 (R:src/webrtc.c:239-249)
    const int interval = (freq <= 8000 && intervalMs % 20 == 0) ? 20 : 10;
    h->pkg = freq / 1000 * interval;
    h->mono.resize((size_t)h->pkg);
    h->res.resize((size_t)h->pkg);
    h->far.resize((size_t)h->pkg);
    if (dbg(debug)) printf("aec_init: chn/%d freq/%d intervalMs/%d pkgFrame/%d x %d\r\n", chn, freq, interval, h->pkg, chn);
    return h;
}

// one packet through the engine; mirrors WebRtcAec_Process's return code for an out-of-range delay
// (it clamps, processes, and still reports -1: T:.../aec/echo_cancellation.c:366-376)
static int aec_packet(Handle* h, const int16_t* far, const int16_t* near, int16_t* out, int delayms)
{
    int rc_ref = 0, d = (int16_t)delayms;
    if (d < 0) { d = 0; rc_ref = -1; }
    else if (d > 500) { d = 500; rc_ref = -1; }
    if (wmixb_aec_host(h->eng, far, near, out, h->pkg, d) != WMIXB_OK) return -1;
    return rc_ref;
}

int aec_setFrameFar(void* fp, int16_t* frameFar, int frameNum)
{
    Handle* h = (Handle*)fp;
    const int total = frameNum * h->chn, step = h->pkg * h->chn;
    for (int pos = 0; pos < total; pos += step) {                   // R:src/webrtc.c:298-321
        for (int i = 0; i < h->pkg; ++i) { h->far[(size_t)i] = *frameFar; frameFar += h->chn; }
        if (wmixb_aec_host(h->eng, h->far.data(), nullptr, nullptr, h->pkg, 0) != WMIXB_OK) {
            if (dbg(h->debug)) printf("WebRtcAecX_BufferFarend failed !!, %s \r\n", wmixb_last_error());
            return -1;
        }
    }
    return 0;
}

int aec_process(void* fp, int16_t* frameNear, int16_t* frameOut, int frameNum, int delayms)
{
    Handle* h = (Handle*)fp;
    const int total = frameNum * h->chn, step = h->pkg * h->chn;
    for (int pos = 0; pos < total; pos += step) {                   // R:src/webrtc.c:349-403
        for (int i = 0; i < h->pkg; ++i) { h->mono[(size_t)i] = *frameNear; frameNear += h->chn; }
        const int ret = aec_packet(h, nullptr, h->mono.data(), h->res.data(), delayms);
        if (ret != 0) {
            if (dbg(h->debug)) printf("WebRtcAecX_Process failed !!, ret %d \r\n", ret);
            return ret;
        }
        for (int i = 0; i < h->pkg; ++i)
            for (int c = 0; c < h->chn; ++c) *frameOut++ = h->res[(size_t)i];
    }
    return 0;
}

int aec_process2(void* fp, int16_t* frameFar, int16_t* frameNear, int16_t* frameOut, int frameNum, int delayms)
{
    Handle* h = (Handle*)fp;
    const int total = frameNum * h->chn, step = h->pkg * h->chn;
    for (int pos = 0; pos < total; pos += step) {                   // R:src/webrtc.c:422-480
        for (int i = 0; i < h->pkg; ++i) {
            h->far[(size_t)i] = *frameFar;
            h->mono[(size_t)i] = *frameNear;
            frameFar += h->chn;
            frameNear += h->chn;
        }
        const int ret = aec_packet(h, h->far.data(), h->mono.data(), h->res.data(), delayms);
        if (ret != 0) {
            if (dbg(h->debug)) printf("WebRtcAecX_Process failed !!, ret %d \r\n", ret);
            return ret;
        }
        for (int i = 0; i < h->pkg; ++i)
            for (int c = 0; c < h->chn; ++c) *frameOut++ = h->res[(size_t)i];
    }
    return 0;
}

void aec_release(void* fp) { drop(fp, "aec_release"); }

// ---------------------------------------------------------------- G.711 host entry points
static int g711_host(int law, bool encode, const void* in, void* out, int n)
{
    if (n <= 0) return 0;
    // grow-only staging and a private stream per calling thread: no cudaMalloc / cudaFree / device-wide sync per call
    wmx::host::Scratch* sc = wmx::host::scratch(wmixb_default_device());
    if (!sc) return -1;
    const size_t in_b = encode ? (size_t)n * 2 : (size_t)n, out_b = encode ? (size_t)n : (size_t)n * 2;
    void *din = sc->need(0, in_b), *dout = sc->need(1, out_b);
    if (!din || !dout) return -1;
    if (cudaMemcpyAsync(din, in, in_b, cudaMemcpyHostToDevice, sc->st) != cudaSuccess) return -1;
    const int r = encode ? wmixb_g711_encode_device(law, (const int16_t*)din, (uint8_t*)dout, (size_t)n, sc->st)
                         : wmixb_g711_decode_device(law, (const uint8_t*)din, (int16_t*)dout, (size_t)n, sc->st);
    if (r != WMIXB_OK) return -1;
    if (cudaMemcpyAsync(out, dout, out_b, cudaMemcpyDeviceToHost, sc->st) != cudaSuccess) return -1;
    return cudaStreamSynchronize(sc->st) == cudaSuccess ? 0 : -1;
}

int g711a_encode(unsigned char g711_data[], const short amp[], int len) { return g711_host(0, true, amp, g711_data, len) == 0 ? len : -1; }
int g711u_encode(unsigned char g711_data[], const short amp[], int len) { return g711_host(1, true, amp, g711_data, len) == 0 ? len : -1; }
int g711a_decode(short amp[], const unsigned char g711a_data[], int n) { return g711_host(0, false, g711a_data, amp, n) == 0 ? (n > 0 ? n : 0) * 2 : -1; }
int g711u_decode(short amp[], const unsigned char g711u_data[], int n) { return g711_host(1, false, g711u_data, amp, n) == 0 ? (n > 0 ? n : 0) * 2 : -1; }

int PCM2G711a(char* in, char* out, int len, int) { if (!in && !out && len == 0) { printf("Error, empty data or transmit failed, exit !\n"); return -1; } return g711a_encode((unsigned char*)out, (const short*)in, len / 2); }
int PCM2G711u(char* in, char* out, int len, int) { if (!in && !out && len == 0) { printf("Error, empty data or transmit failed, exit !\n"); return -1; } return g711u_encode((unsigned char*)out, (const short*)in, len / 2); }
int G711a2PCM(char* in, char* out, int len, int) { if (!in && !out && len == 0) { printf("Error, empty data or transmit failed, exit !\n"); return -1; } return g711a_decode((short*)out, (const unsigned char*)in, len); }
int G711u2PCM(char* in, char* out, int len, int) { if (!in && !out && len == 0) { printf("Error, empty data or transmit failed, exit !\n"); return -1; } return g711u_decode((short*)out, (const unsigned char*)in, len); }

}  // extern "C"

// ---------------------------------------------------------------- wmix_pcm_zoom (R:src/wmix.c:49-222)
extern "C" {

uint32_t wmix_len_of_out(uint8_t inChn, uint16_t inFreq, uint32_t inLen, uint8_t outChn, uint16_t outFreq)
{
    return wmx::host::zoom_len_of_out(inChn, inFreq, inLen, outChn, outFreq);
}

uint32_t wmix_len_of_in(uint8_t inChn, uint16_t inFreq, uint8_t outChn, uint16_t outFreq, uint32_t outLen)
{
    return wmx::host::zoom_len_of_in(inChn, inFreq, outChn, outFreq, outLen);
}

// Plans (gather tables) are cached per (formats, length) in a small LRU — the reference's callers reuse one buffer size
// per task, but a short last read or a variable decode length must not grow the cache for the life of the daemon.  The
// lock covers the cache lookup only; the samples make their H2D -> kernel -> D2H round trip on the calling thread's own
// staging and stream (scratch.h), so concurrent task threads do not serialise on each other.
uint32_t wmix_pcm_zoom(uint8_t inChn, uint16_t inFreq, uint8_t* in, uint32_t inLen, uint8_t outChn, uint16_t outFreq,
                       uint8_t* out)
{
    if (inFreq == outFreq && inChn == outChn) {                      // R:src/wmix.c:153-157: a plain copy
        memcpy(out, in, inLen);
        return inLen;
    }
    typedef std::tuple<int, int, uint32_t, int, int, int> Key;
    struct Plan {
        wmixb_zoom* z = nullptr;
        ~Plan() { if (z) wmixb_zoom_destroy(z); }
    };
    constexpr size_t kMaxPlans = 32;
    static std::mutex mu;
    static std::list<std::pair<Key, std::shared_ptr<Plan>>> lru;      // most recently used first
    const int device = wmixb_default_device();
    const Key key = std::make_tuple((int)inChn, (int)inFreq, inLen, (int)outChn, (int)outFreq, device);
    std::shared_ptr<Plan> plan;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (auto it = lru.begin(); it != lru.end(); ++it)
            if (it->first == key) { plan = it->second; lru.splice(lru.begin(), lru, it); break; }
    }
    if (!plan) {
        plan = std::make_shared<Plan>();
        if (wmixb_zoom_create(inChn, inFreq, inLen, outChn, outFreq, device, &plan->z) != WMIXB_OK) return 0;
        std::lock_guard<std::mutex> lock(mu);
        lru.emplace_front(key, plan);
        while (lru.size() > kMaxPlans) lru.pop_back();               // a plan still in use elsewhere lives until that call returns
    }
    const uint32_t ob = wmixb_zoom_out_bytes(plan->z);
    if (ob == 0) return 0;
    wmx::host::Scratch* sc = wmx::host::scratch(device);
    if (!sc) return 0;
    void *d_in = sc->need(0, (size_t)inLen + 2), *d_out = sc->need(1, ob);
    if (!d_in || !d_out) return 0;
    if (cudaMemsetAsync((char*)d_in + (inLen & ~1u), 0, 2 + (inLen & 1u), sc->st) != cudaSuccess) return 0;   // the walk may look one sample past
    if (cudaMemcpyAsync(d_in, in, inLen, cudaMemcpyHostToDevice, sc->st) != cudaSuccess) return 0;
    if (wmixb_zoom_device(plan->z, (const int16_t*)d_in, (int16_t*)d_out, 1, sc->st) != WMIXB_OK) return 0;
    if (cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, sc->st) != cudaSuccess) return 0;
    if (cudaStreamSynchronize(sc->st) != cudaSuccess) return 0;
    return ob;
}

}  // extern "C"
