// Per-thread, per-device staging for the drop-in entry points (G.711, wmix_pcm_zoom, wmix_load_data).
//
// Those calls are one H2D -> kernel -> D2H round trip each.  They must not touch the device-wide
// synchronisation points (cudaMalloc / cudaFree / cudaDeviceSynchronize / the legacy default stream): a
// daemon's RTP thread calling the codec every 20 ms would otherwise stall the pipelined ticks of a batched
// engine in the same process.  So every host thread keeps, per device, a few grow-only device buffers and a
// private non-blocking stream, and waits with cudaStreamSynchronize on that stream only.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <map>

namespace wmx {
namespace host {

struct Scratch {
    static constexpr int kSlots = 4;
    int device = 0;
    cudaStream_t st = nullptr;
    void* buf[kSlots] = {};
    size_t cap[kSlots] = {};
    // device buffer `slot` of at least `bytes` (contents undefined after growth); nullptr on failure
    void* need(int slot, size_t bytes)
    {
        if (bytes < 16) bytes = 16;
        if (cap[slot] >= bytes) return buf[slot];
        // growth is the rare path: the old buffer may still be in use by this thread's stream only
        if (st) cudaStreamSynchronize(st);
        cudaFree(buf[slot]);
        buf[slot] = nullptr;
        cap[slot] = 0;
        size_t want = 4096;
        while (want < bytes) want *= 2;
        if (cudaMalloc(&buf[slot], want) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        cap[slot] = want;
        return buf[slot];
    }
};

struct ScratchSet {
    std::map<int, Scratch> per_device;
    ~ScratchSet()
    {
        for (auto& kv : per_device) {
            Scratch& s = kv.second;
            if (cudaSetDevice(s.device) != cudaSuccess) continue;   // runtime already unloading at process exit
            for (int k = 0; k < Scratch::kSlots; ++k) cudaFree(s.buf[k]);
            if (s.st) cudaStreamDestroy(s.st);
        }
        (void)cudaGetLastError();
    }
};

// the calling thread's staging on `device` (makes `device` current); nullptr if the device is unusable
inline Scratch* scratch(int device)
{
    static thread_local ScratchSet set;
    if (cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    Scratch& s = set.per_device[device];
    if (!s.st) {
        s.device = device;
        if (cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); s.st = nullptr; return nullptr; }
    }
    return &s;
}

}  // namespace host
}  // namespace wmx
