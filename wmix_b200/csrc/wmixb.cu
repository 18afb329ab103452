// wmix_b200 — sm_100a kernels and the C-ABI engine behind include/wmixb.h.
//
// Kernels (all new; the reference has no GPU code):
//   ns_kernel<ANA>      WebRTC NS, one stream-frame per warp, persistent grid  (ns.cuh)
//   post_kernel<FS16>   AGC -> VAD (+ mute ramp), one stream per thread        (agc.cuh, vad.cuh)
//   bus_sum / nminus1   conference bus: exact int32 sum, N-minus-one read-out
//   g711_*              A-law / mu-law codecs, 16-byte vectorised
//   mix_load            same-format branch of wmix_load_data on a device ring
// Build: nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -lineinfo
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <dlfcn.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/wmix.h"
#include "../../include/wmix_rtp.h"
#include "../../include/wmix_zoom.h"
#include "../../include/wmixb.h"
#include "aec.cuh"
#include "agc.cuh"
#include "g711_mix.cuh"
#include "host_tables.h"
#include "ns.cuh"
#include "ns_cta.cuh"
#include "nsx.cuh"
#include "peer_bus.cuh"
#include "scratch.h"
#include "vad.cuh"

using namespace wmx;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_default_device{0};   // device of the drop-in handle / codec / zoom entry points

static int fail_cuda(cudaError_t e, const char* what, int line)
{
    snprintf(g_err, sizeof g_err, "%s failed at wmixb.cu:%d: %s", what, line, cudaGetErrorString(e));
    return WMIXB_ECUDA;
}
#define CK(call)                                                         \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call, __LINE__);  \
    } while (0)
#define CK_LAUNCH()                                                      \
    do {                                                                 \
        g_launches.fetch_add(1, std::memory_order_relaxed);              \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, "kernel launch", __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------
// NS kernel: persistent grid, one warp per stream, K frames per stream per launch
// ------------------------------------------------------------------------------------------

// bulk L2 prefetch (bytes a multiple of 16, address 16-byte aligned): one instruction, no
// destination, no completion to wait for
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int ANA>
struct NsSmem {
    static constexpr size_t kTableFloats = (sizeof(ns::Tables<ANA>) + 15) / 16 * 4;
};
template <int ANA>
constexpr size_t ns_smem_bytes(int warps) { return (NsSmem<ANA>::kTableFloats + (size_t)warps * ns::Geo<ANA>::kShFloats) * sizeof(float); }

template <int ANA, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
ns_kernel(float* __restrict__ rec, uint16_t* __restrict__ hist, const ns::Tables<ANA>* __restrict__ tables,
          const int16_t* in, int16_t* out, int n_streams, int n_frames, int align)
{
    typedef ns::Geo<ANA> G;
    extern __shared__ __align__(16) float smem[];
    ns::Tables<ANA>* T = reinterpret_cast<ns::Tables<ANA>*>(smem);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(ns::Tables<ANA>) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    float* tile = smem + NsSmem<ANA>::kTableFloats + (size_t)warp * G::kShFloats;
    ns::Warp<ANA> W;
    W.lane_id = threadIdx.x & 31;
    for (int i = W.lane_id; i < G::kShFloats; i += 32) tile[i] = 0.f;   // sum rows rely on +0.0f padding
    __syncwarp();
    const int total_warps = gridDim.x * WARPS;
    // The loop body is ~110 KB of code against a 32 KB instruction cache: warps that drift apart each stream
    // their own copy of it from L2.  With `align` the warps of a CTA start every frame together, so they run
    // the same phases at the same time and share the fetched lines (uniform trip count: barriers inside).
    const int iters = (n_streams - blockIdx.x * WARPS + total_warps - 1) / total_warps;   // of warp 0; >= every other warp's
    for (int it = 0; it < iters; ++it) {
        const int s = blockIdx.x * WARPS + warp + it * total_warps;
        const bool live = s < n_streams;
        float* r = rec + (size_t)s * G::kRecFloats;
        uint16_t* h = hist + (size_t)s * 3 * ns::kHistBins;
        const int16_t* pi = in + (size_t)s * n_frames * G::kBlock;
        int16_t* po = out + (size_t)s * n_frames * G::kBlock;
        for (int f = 0; f < n_frames; ++f) {
            if (align && (it % align) == 0) __syncthreads();
            if (!live) continue;
            if (f == n_frames - 1 && W.lane_id == 0 && s + total_warps < n_streams) {
                // pull the next stream's record and first frame towards L2 while this one computes
                l2_prefetch(rec + (size_t)(s + total_warps) * G::kRecFloats, G::kRecFloats * sizeof(float));
                l2_prefetch(in + (size_t)(s + total_warps) * n_frames * G::kBlock, G::kBlock * sizeof(int16_t));
            }
            ns::frame<ANA>(W, r, h, pi + (size_t)f * G::kBlock, po + (size_t)f * G::kBlock, tile, *T);
        }
    }
}

// ------------------------------------------------------------------------------------------
// NS kernel, CTA-cooperative form (ns_cta.cuh): W worker warps (one stream-frame each per round) + ONE reducer
// warp that walks the in-order sums, the scalar model and the Nyquist bin of all W streams at once.  Persistent
// grid; worker j of CTA b serves streams b*W + j + round * gridDim.x*W.  Hand-over through the worker tiles and six
// named barriers per frame (ids 1..6; odd = workers arrive / reducer waits, even = reducer arrives / workers wait).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// TMA bulk copies of a whole per-stream record (1-D cp.async.bulk, completion on an mbarrier / a bulk group)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// every thread that wrote the buffer through the generic proxy runs this before the (one) thread that issues the bulk store
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// STAGED (the persistent offline mode): every worker also keeps its stream's whole record on chip
template <int ANA>
constexpr size_t ns_cta_smem_bytes(int workers, bool staged = false)
{
    return (NsSmem<ANA>::kTableFloats + (size_t)workers * ns::Geo<ANA>::kShFloats + (staged ? (size_t)workers * ns::Geo<ANA>::kRecFloats : 0)) * sizeof(float) +
           8 * sizeof(uint16_t*) + (staged ? 8 * sizeof(uint64_t) : 0);
}

// STAGED = persistent offline mode: a worker pulls its stream's whole record (8 KB at 16 kHz) into shared memory with ONE
// TMA bulk copy, runs all n_frames frames of the stream against that copy — every state access of the segments becomes a
// shared-memory access, the only DRAM traffic per frame is the PCM — and writes the record back with one bulk store.
template <int ANA, int W, int MINB, bool STAGED = false>
__global__ void __launch_bounds__((W + 1) * 32, MINB)
ns_cta_kernel(float* __restrict__ rec, uint16_t* __restrict__ hist, const ns::Tables<ANA>* __restrict__ tables,
              const int16_t* in, int16_t* out, int n_streams, int n_frames)
{
    typedef ns::Geo<ANA> G;
    static_assert(W >= 1 && W <= 8, "the reducer serves at most 8 workers (4 lanes each)");
    constexpr int kThreads = (W + 1) * 32;
    extern __shared__ __align__(16) float smem[];
    ns::Tables<ANA>* T = reinterpret_cast<ns::Tables<ANA>*>(smem);
    float* tiles = smem + NsSmem<ANA>::kTableFloats;
    float* recs = tiles + (size_t)W * G::kShFloats;                                   // [W][kRecFloats], STAGED only
    uint16_t** hptr = reinterpret_cast<uint16_t**>(recs + (STAGED ? (size_t)W * G::kRecFloats : 0));
    uint64_t* mbars = reinterpret_cast<uint64_t*>(hptr + 8);                         // [W], STAGED only
    if (STAGED && threadIdx.x < W) mbar_init(&mbars[threadIdx.x], 1);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(ns::Tables<ANA>) / 4); i += kThreads) dst[i] = src[i];
        float4* t4 = reinterpret_cast<float4*>(tiles);
        for (int i = threadIdx.x; i < W * G::kShFloats / 4; i += kThreads) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // sum rows rely on +0.0f padding
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = gridDim.x * W;
    const int first = blockIdx.x * W;
    const int rounds = first < n_streams ? (n_streams - first + total - 1) / total : 0;
    if (warp < W) {
        float* tile = tiles + (size_t)warp * G::kShFloats;
        ns::WWarp<ANA> Wk;
        Wk.lane_id = lane;
        float* staged = recs + (size_t)warp * G::kRecFloats;
        uint32_t stage_phase = 0;
        for (int it = 0; it < rounds; ++it) {
            const int s = first + warp + it * total;
            const bool live = s < n_streams;
            float* r_global = rec + (size_t)s * G::kRecFloats;
            float* r = STAGED ? staged : r_global;
            uint16_t* h = hist + (size_t)s * 3 * ns::kHistBins;
            const int16_t* pi = in + (size_t)s * n_frames * G::kBlock;
            int16_t* po = out + (size_t)s * n_frames * G::kBlock;
            if (lane == 0) hptr[warp] = h;
            if (STAGED && live) {
                if (lane == 0) {
                    bulk_store_wait_read();                                   // the previous stream's write-back has left the buffer
                    bulk_load(staged, r_global, G::kRecFloats * sizeof(float), &mbars[warp]);
                }
                mbar_wait(&mbars[warp], stage_phase);
                stage_phase ^= 1;
            }
            for (int f = 0; f < n_frames; ++f) {
                bool act = false;
                if (live) {
                    if (f == n_frames - 1 && lane == 0 && s + total < n_streams) {
                        // pull the next stream's record and first frame towards L2 while this one computes
                        l2_prefetch(rec + (size_t)(s + total) * G::kRecFloats, G::kRecFloats * sizeof(float));
                        l2_prefetch(in + (size_t)(s + total) * n_frames * G::kBlock, G::kBlock * sizeof(int16_t));
                    }
                    if (STAGED && lane == 0 && f + 1 < n_frames) l2_prefetch(pi + (size_t)(f + 1) * G::kBlock, G::kBlock * sizeof(int16_t));
                    act = ns::w_seg1<ANA>(Wk, r, pi + (size_t)f * G::kBlock, po + (size_t)f * G::kBlock, tile, *T);
                } else if (lane == 0) {
                    tile[G::kShScal + ns::C_ACTIVE] = 0.f;
                }
                named_bar_arrive(1, kThreads);
                named_bar_sync(2, kThreads);
                if (act) ns::w_seg2<ANA>(Wk, r, tile, *T);
                named_bar_arrive(3, kThreads);
                named_bar_sync(4, kThreads);
                if (act) {
                    ns::w_seg3a<ANA>(Wk, r, h, tile, *T);
                    ns::w_seg3b<ANA>(Wk, r, tile, *T);
                }
                named_bar_arrive(5, kThreads);
                named_bar_sync(6, kThreads);
                if (act) ns::w_seg4<ANA>(Wk, r, po + (size_t)f * G::kBlock, tile, *T);
            }
            if (STAGED && live) {
                fence_async_smem();
                __syncwarp();
                if (lane == 0) bulk_store(r_global, staged, G::kRecFloats * sizeof(float));
            }
        }
        if (STAGED && lane == 0) bulk_store_wait_all();
    } else {
        ns::RWarp Rd;
        Rd.lane_id = lane;
        for (int it = 0; it < rounds; ++it) {
            for (int f = 0; f < n_frames; ++f) {
                named_bar_sync(1, kThreads);
                ns::r_seg1<ANA>(Rd, tiles, G::kShFloats, W, *T);
                named_bar_arrive(2, kThreads);
                ns::r_seg1b<ANA>(Rd, tiles, G::kShFloats, *T);                 // behind the workers' segment 2, off their critical path
                named_bar_sync(3, kThreads);
                ns::r_seg2a<ANA>(Rd, tiles, G::kShFloats, hptr, *T);
                ns::r_seg2b<ANA>(Rd, tiles, G::kShFloats, *T);
                named_bar_arrive(4, kThreads);
                named_bar_sync(5, kThreads);
                ns::r_seg3<ANA>(Rd, tiles, G::kShFloats, *T);
                named_bar_arrive(6, kThreads);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// NSX kernel: the fixed-point suppressor (nsx.cuh), persistent grid, one self-contained warp per stream-frame.
// All sums are warp reductions, so there is nothing to hand to a reducer warp: the CTA is just WARPS independent warps
// sharing the constant tables; `align` starts the warps of a CTA together on every stream (instruction-cache sharing).
// ------------------------------------------------------------------------------------------
constexpr size_t kNsxTableWords = (sizeof(nsx::Tables) + 15) / 16 * 4;
template <int ANA>
constexpr size_t nsx_smem_bytes(int warps) { return (kNsxTableWords + (size_t)warps * ((nsx::Geo<ANA>::kShWords + 3) / 4 * 4)) * sizeof(uint32_t); }

template <int ANA, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
nsx_kernel(uint32_t* __restrict__ rec, int16_t* __restrict__ hist, const nsx::Tables* __restrict__ tables, const int16_t* in, int16_t* out,
           int n_streams, int n_frames, int align)
{
    typedef nsx::Geo<ANA> G;
    constexpr int kTile = (G::kShWords + 3) / 4 * 4;
    extern __shared__ __align__(16) uint32_t smem_u[];
    nsx::Tables* T = reinterpret_cast<nsx::Tables*>(smem_u);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        for (int i = threadIdx.x; i < (int)(sizeof(nsx::Tables) / 4); i += blockDim.x) smem_u[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    uint32_t* tile = smem_u + kNsxTableWords + (size_t)warp * kTile;
    nsx::Warp<ANA> W;
    W.lane_id = threadIdx.x & 31;
    const int total_warps = gridDim.x * WARPS;
    const int iters = (n_streams - blockIdx.x * WARPS + total_warps - 1) / total_warps;   // of warp 0; >= every other warp's
    for (int it = 0; it < iters; ++it) {
        const int s = blockIdx.x * WARPS + warp + it * total_warps;
        const bool live = s < n_streams;
        uint32_t* r = rec + (size_t)s * G::kRecWords;
        int16_t* h = hist + (size_t)s * 3 * nsx::kHistBins;
        const int16_t* pi = in + (size_t)s * n_frames * G::kBlock;
        int16_t* po = out + (size_t)s * n_frames * G::kBlock;
        for (int f = 0; f < n_frames; ++f) {
            // the warps of the CTA start every frame together: they then run the same ~90 KB of code at about the same time
            // and share the fetched lines (uniform trip count, so the barrier is in uniform control flow).  Meeting points
            // INSIDE the frame were measured too (profiles/r2_p_nsx_summary.md): they help small CTAs, cost the shipped
            // 32-warp one, and need warps that leave the frame early to arrive at a different bar.sync — removed.
            if (align) __syncthreads();
            if (!live) continue;
            if (f == n_frames - 1 && W.lane_id == 0 && s + total_warps < n_streams) {
                // pull the next stream's record and first frame towards L2 while this one computes
                l2_prefetch(rec + (size_t)(s + total_warps) * G::kRecWords, G::kRecWords * sizeof(uint32_t));
                l2_prefetch(in + (size_t)(s + total_warps) * n_frames * G::kBlock, G::kBlock * sizeof(int16_t));
            }
            nsx::frame<ANA>(W, r, h, pi + (size_t)f * G::kBlock, po + (size_t)f * G::kBlock, tile, *T);
        }
    }
}

// wmix's stereo case with the fixed-point core: the right channel as WebRtcNsx's second band (nsx.cuh, frame<ANA, true> +
// second_band).  One frame per launch; the drop-in handle's path, kept apart so the mono kernel is untouched.
template <int ANA>
__global__ void __launch_bounds__(256, 2)
nsx_hb_kernel(uint32_t* __restrict__ rec, int16_t* __restrict__ hist, int16_t* __restrict__ hb, const nsx::Tables* __restrict__ tables,
              const int16_t* in, int16_t* out, const int16_t* in_hb, int16_t* out_hb, int n_streams)
{
    typedef nsx::Geo<ANA> G;
    constexpr int kTile = (G::kShWords + 3) / 4 * 4;
    extern __shared__ __align__(16) uint32_t smem_u[];
    nsx::Tables* T = reinterpret_cast<nsx::Tables*>(smem_u);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        for (int i = threadIdx.x; i < (int)(sizeof(nsx::Tables) / 4); i += blockDim.x) smem_u[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    uint32_t* tile = smem_u + kNsxTableWords + (size_t)warp * kTile;
    nsx::Warp<ANA> W;
    W.lane_id = threadIdx.x & 31;
    for (int s = blockIdx.x * 8 + warp; s < n_streams; s += gridDim.x * 8) {
        const int gain = nsx::frame<ANA, true>(W, rec + (size_t)s * G::kRecWords, hist + (size_t)s * 3 * nsx::kHistBins, in + (size_t)s * G::kBlock,
                                               out + (size_t)s * G::kBlock, tile, *T);
        nsx::second_band<ANA>(W, hb + (size_t)s * G::kKeep, in_hb + (size_t)s * G::kBlock, out_hb + (size_t)s * G::kBlock, gain);
    }
}

template <int ANA>
__global__ void nsx_init_kernel(uint32_t* rec, int16_t* hist, int first, int count, int32_t thr_lrt)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;
    const int s = first + warp;
    nsx::init_record<ANA>(rec + (size_t)s * nsx::Geo<ANA>::kRecWords, hist + (size_t)s * 3 * nsx::kHistBins, lane, 32, thr_lrt);
}

// wmix's stereo case (ns_init(2, ..)): the right channel rides through WebRtcNs as a "high band" (ns.cuh, frame<ANA, true>).
// Same warp-per-stream shape as ns_kernel, one frame per launch; kept apart so the mono kernel's code and registers
// are untouched.  The drop-in handle's path, not a throughput path.
template <int ANA>
__global__ void __launch_bounds__(256, 2)
ns_hb_kernel(float* __restrict__ rec, uint16_t* __restrict__ hist, float* __restrict__ hb_hist, const ns::Tables<ANA>* __restrict__ tables,
             const int16_t* in, int16_t* out, const int16_t* in_hb, int16_t* out_hb, int n_streams)
{
    typedef ns::Geo<ANA> G;
    constexpr int kTile = G::kShFloats + G::kBlock;
    extern __shared__ __align__(16) float smem[];
    ns::Tables<ANA>* T = reinterpret_cast<ns::Tables<ANA>*>(smem);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(ns::Tables<ANA>) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    float* tile = smem + NsSmem<ANA>::kTableFloats + (size_t)warp * kTile;
    ns::Warp<ANA> W;
    W.lane_id = threadIdx.x & 31;
    for (int i = W.lane_id; i < kTile; i += 32) tile[i] = 0.f;
    __syncwarp();
    const int total_warps = gridDim.x * 8;
    for (int s = blockIdx.x * 8 + warp; s < n_streams; s += total_warps) {
        ns::frame<ANA, true>(W, rec + (size_t)s * G::kRecFloats, hist + (size_t)s * 3 * ns::kHistBins, in + (size_t)s * G::kBlock,
                             out + (size_t)s * G::kBlock, tile, *T, hb_hist + (size_t)s * G::kOverlap, in_hb + (size_t)s * G::kBlock,
                             out_hb + (size_t)s * G::kBlock);
        __syncwarp();
    }
}
template <int ANA>
constexpr size_t ns_hb_smem_bytes() { return (NsSmem<ANA>::kTableFloats + (size_t)8 * (ns::Geo<ANA>::kShFloats + ns::Geo<ANA>::kBlock)) * sizeof(float); }

template <int ANA>
__global__ void ns_init_kernel(float* rec, uint16_t* hist, int first, int count)
{
    typedef ns::Geo<ANA> G;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;
    const int s = first + warp;
    float* r = rec + (size_t)s * G::kRecFloats;
    ns::init_record<ANA>(r, hist + (size_t)s * 3 * ns::kHistBins, lane, 32);
    __syncwarp();
    ns::init_record_values<ANA>(r, lane, 32);
}

// ------------------------------------------------------------------------------------------
// AEC kernel: persistent grid, one warp per stream (aec.cuh)
// ------------------------------------------------------------------------------------------
constexpr int kAecWarps = 8;   // grid sizing unit; the shipped shape is ONE 16-warp CTA per SM (aec_warps = 16: 0.424 ms per 16 384-stream
                               // tick against 0.435 ms for two CTAs of 8, profiles/r2_s_aec_warps.txt); wmixb_set_tuning("aec_warps", 8) for the other
constexpr size_t kAecTableFloats = (sizeof(aec::Tables) + 15) / 16 * 4;
constexpr size_t aec_smem_bytes(int warps) { return (kAecTableFloats + (size_t)warps * aec::Geo::kShFloats) * sizeof(float); }
constexpr size_t kAecSmemBytes = aec_smem_bytes(kAecWarps);

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS)
aec_kernel(float* __restrict__ rec, size_t rec_floats, const aec::Tables* __restrict__ tables, const int16_t* far,
           const int16_t* near, int16_t* out, int n_streams, int n, int mult, int depth, int delay_ms, int pf_mode, int row_stride, int align)
{
    extern __shared__ __align__(16) float smem[];
    aec::Tables* T = reinterpret_cast<aec::Tables*>(smem);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables);
        uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
        for (int i = threadIdx.x; i < (int)(sizeof(aec::Tables) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    float* tile = smem + kAecTableFloats + (size_t)warp * aec::Geo::kShFloats;
    aec::Warp W;
    W.lane_id = threadIdx.x & 31;
    const int total_warps = gridDim.x * WARPS;
    // The tick is ~90 KB of code: warps that drift apart each stream their own copy of it through the instruction cache.
    // With `align` the warps of a CTA start every stream together (uniform trip count, so the barrier is legal).
    const int first_of_cta = blockIdx.x * WARPS;
    const int iters = first_of_cta < n_streams ? (n_streams - first_of_cta + total_warps - 1) / total_warps : 0;
    for (int it = 0; it < iters; ++it) {
        const int s = first_of_cta + warp + it * total_warps;
        if (align) __syncthreads();
        if (s >= n_streams) continue;
        // L2 staging of the record (WMIXB_AEC_PF): 2 (default) = this stream's fixed part as ONE bulk request when the tick
        // starts — the first loads wait for DRAM once, the rest of the tick hits L2; 1 = the next stream's record a whole
        // tick ahead (measured slower: 2368 resident warps x two 23-31 KB records outgrow the L2 and the lines are evicted
        // before use); 5 = 2 plus the next record requested mid-tick (also slower); 0 = none.
        if (W.lane_id == 0) {
            if (pf_mode == 1 && s + total_warps < n_streams)
                l2_prefetch(rec + (size_t)(s + total_warps) * rec_floats, (uint32_t)(rec_floats * sizeof(float)));
            if (pf_mode == 2 || pf_mode == 5) l2_prefetch(rec + (size_t)s * rec_floats, (uint32_t)(aec::Geo::kFixedFloats * sizeof(float)));
        }
        W.pf_next = (pf_mode == 5 && s + total_warps < n_streams) ? rec + (size_t)(s + total_warps) * rec_floats : nullptr;
        W.pf_bytes = (uint32_t)(aec::Geo::kFixedFloats * sizeof(float));
        aec::tick(W, rec + (size_t)s * rec_floats, depth, mult, n, far ? far + (size_t)s * row_stride : nullptr,
                  near ? near + (size_t)s * row_stride : nullptr, out ? out + (size_t)s * row_stride : nullptr, delay_ms, tile, *T);
        __syncwarp();
    }
}

__global__ void aec_init_kernel(float* rec, size_t rec_floats, int depth, int first, int count)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= count) return;
    float* r = rec + (size_t)(first + warp) * rec_floats;
    aec::init_record(r, depth, lane, 32);
    __syncwarp();
    aec::init_record_values(r, lane, 32);
}

__global__ void aec_status_kernel(const float* rec, size_t rec_floats, int n_streams, int* result)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    const int f = __float_as_int(rec[(size_t)s * rec_floats + aec::Geo::kOffScal + aec::S_ERROR]);
    if (f) {
        atomicOr(&result[0], f);
        atomicAdd(&result[1], 1);
    }
}

// ------------------------------------------------------------------------------------------
// post kernel: AGC -> VAD, one stream per thread; PCM tile staged through shared memory so
// global traffic is whole lines while each thread walks its own row without bank conflicts
// (row pitch = frame+2 int16 = an odd number of 32-bit words).
// ------------------------------------------------------------------------------------------
constexpr int kPostThreads = 128;

template <bool FS16, int MINB, int THREADS = kPostThreads>
__global__ void __launch_bounds__(THREADS, MINB)
post_kernel(int32_t* __restrict__ agc_words, int32_t* __restrict__ vad_words, const int32_t* __restrict__ agc_table,
            vad::Params vp, const int16_t* in, int16_t* out, uint8_t* vad_out, int n_streams, size_t stride,
            int n_frames, int stages)
{
    constexpr int L = FS16 ? 160 : 80;
    constexpr int ROWW = L / 2 + 1;                     // row pitch in 32-bit words (odd)
    extern __shared__ __align__(16) int32_t post_smem[];      // [THREADS][ROWW] PCM tile, then the 32-entry AGC gain table
    int32_t* tile = post_smem;
    int32_t* tab = post_smem + THREADS * ROWW;
    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * THREADS;
    const int s = s0 + tid;
    if (tid < 32) tab[tid] = agc_table ? agc_table[tid] : 0;
    const int rows = min(THREADS, n_streams - s0);
    const int32_t* in32 = reinterpret_cast<const int32_t*>(in);
    int32_t* out32 = reinterpret_cast<int32_t*>(out);
    SoaWords agc_st{agc_words ? agc_words + s : nullptr, stride};
    SoaWords vad_st{vad_words ? vad_words + s : nullptr, stride};
    for (int f = 0; f < n_frames; ++f) {
        __syncthreads();
        // a warp moves whole rows, a lane the words lane, lane + 32, (lane + 64) of each: no index division, and four rows'
        // loads are in flight before the first store (the flat idx = r * 40 + w loop was a tenth of the kernel's instructions
        // and its load -> store pairs waited for DRAM one after the other)
        {
            constexpr int WPR = L / 2, NW = THREADS / 32;               // words per row, warps
            const int wid = tid >> 5, ln = tid & 31;
#pragma unroll 1
            for (int r0 = wid; r0 < rows; r0 += 4 * NW) {
                int32_t v[4][(WPR + 31) / 32];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = r0 + q * NW;
                    const int32_t* src = in32 + ((size_t)(s0 + r) * n_frames + f) * WPR;
#pragma unroll
                    for (int k = 0; k < (WPR + 31) / 32; ++k)
                        if (r < rows && ln + 32 * k < WPR) v[q][k] = src[ln + 32 * k];
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = r0 + q * NW;
#pragma unroll
                    for (int k = 0; k < (WPR + 31) / 32; ++k)
                        if (r < rows && ln + 32 * k < WPR) tile[r * ROWW + ln + 32 * k] = v[q][k];
                }
            }
        }
        __syncthreads();
        int16_t* x = reinterpret_cast<int16_t*>(tile + tid * ROWW);
        const bool active = s < n_streams;
        if (THREADS <= 128) {
            if (active) {
                if (stages & WMIXB_AGC) agc::process_packet<FS16>(agc_st, x, tab);
                if (stages & WMIXB_VAD) {
                    const int flag = vad::process_packet<80, FS16>(vad_st, x, vp);
                    if (vad_out) vad_out[(size_t)s * n_frames + f] = (uint8_t)flag;
                }
            }
        } else {
            // the big-CTA shape: the kernel is bound by instruction fetch (~120 KB of straight-line code per thread, warps
            // scattered all over it).  Re-aligning the 22 warps of the SM between the stretches of the chain makes them run
            // the same ~30 KB at the same time and share the fetched lines: 0.184 -> 0.137 ms per 100 000 streams.  (More
            // alignment points INSIDE the stages were measured slower: 0.138 - 0.144 ms, the waits outgrow the sharing.)
            if (active && (stages & WMIXB_AGC)) agc::process_packet<FS16>(agc_st, x, tab);
            __syncthreads();
            int16_t feat[6];
            int16_t power = 0;
            if (active && (stages & WMIXB_VAD)) power = vad::packet_features<80, FS16>(vad_st, x, feat);
            __syncthreads();
            int flag = 0;
            if (active && (stages & WMIXB_VAD)) flag = vad::gmm(vad_st, feat, power, vp);
            __syncthreads();
            if (active && (stages & WMIXB_VAD)) {
                flag = vad::packet_finish<80, FS16>(vad_st, x, flag);
                if (vad_out) vad_out[(size_t)s * n_frames + f] = (uint8_t)flag;
            }
        }
        __syncthreads();
        {
            constexpr int WPR = L / 2, NW = THREADS / 32;
            const int wid = tid >> 5, ln = tid & 31;
#pragma unroll 2
            for (int r = wid; r < rows; r += NW) {
                int32_t* dst = out32 + ((size_t)(s0 + r) * n_frames + f) * WPR;
#pragma unroll
                for (int k = 0; k < (WPR + 31) / 32; ++k)
                    if (ln + 32 * k < WPR) dst[ln + 32 * k] = tile[r * ROWW + ln + 32 * k];
            }
        }
    }
}

// VAD on 20 ms packets (what wmix itself asks for: vad_init(.., WMIX_INTERVAL_MS = 20, ..), R:src/wmix.c:703,
// R:src/webrtc.c:56-65; the 20 ms threshold column of T:.../vad/vad_core.c:149-164).  One stream per thread,
// the packet staged in local memory: a correctness path for the drop-in handle, not a throughput path.
// VAD on whole packets that are not the tick's 10 ms frame: 20 ms packets (what wmix itself configures) and the 32 kHz
// handle packets.  One stream per thread like post_kernel, and like there the packets of a CTA's 64 streams move between
// global and shared memory as one coalesced block (row pitch odd in 32-bit words: conflict-free per-thread rows) — a thread
// walking its own 640-byte row in global memory touches a different line than each of its 31 neighbours on every access,
// and a per-thread local array is the same pattern behind the L1.
constexpr int kPktThreads = 64;
template <int L>
constexpr size_t pkt_smem_bytes() { return (size_t)kPktThreads * (L / 2 + 1) * sizeof(int32_t); }

template <int L>
__device__ __forceinline__ int16_t* pkt_stage_in(int32_t* tile, const int16_t* pcm, int s0, int rows)
{
    constexpr int ROWW = L / 2 + 1;
    const int32_t* in32 = reinterpret_cast<const int32_t*>(pcm);
    for (int idx = threadIdx.x; idx < rows * (L / 2); idx += kPktThreads) {
        const int r = idx / (L / 2), w = idx - r * (L / 2);
        tile[r * ROWW + w] = in32[(size_t)(s0 + r) * (L / 2) + w];
    }
    __syncthreads();
    return reinterpret_cast<int16_t*>(tile + threadIdx.x * ROWW);
}
template <int L>
__device__ __forceinline__ void pkt_stage_out(const int32_t* tile, int16_t* pcm, int s0, int rows)
{
    constexpr int ROWW = L / 2 + 1;
    int32_t* out32 = reinterpret_cast<int32_t*>(pcm);
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * (L / 2); idx += kPktThreads) {
        const int r = idx / (L / 2), w = idx - r * (L / 2);
        out32[(size_t)(s0 + r) * (L / 2) + w] = tile[r * ROWW + w];
    }
}

template <bool FS16>
__global__ void __launch_bounds__(kPktThreads)
vad_packet20_kernel(int32_t* __restrict__ vad_words, vad::Params vp, int16_t* pcm, uint8_t* vad_out, int n_streams, size_t stride)
{
    constexpr int L = FS16 ? 320 : 160;
    extern __shared__ __align__(16) int32_t pkt_smem[];
    const int s0 = blockIdx.x * kPktThreads, s = s0 + threadIdx.x;
    const int rows = min(kPktThreads, n_streams - s0);
    int16_t* x = pkt_stage_in<L>(pkt_smem, pcm, s0, rows);
    if (s < n_streams) {
        SoaWords st{vad_words + s, stride};
        const int flag = vad::process_packet<160, FS16>(st, x, vp);
        if (vad_out) vad_out[s] = (uint8_t)flag;
    }
    pkt_stage_out<L>(pkt_smem, pcm, s0, rows);
}

// VAD on 32 kHz packets of 10 ms (CalcVad32khz, T:.../vad/vad_core.c:623-643): the handle API's 32 kHz case and the VAD
// stage of a 32 kHz engine.
__global__ void __launch_bounds__(kPktThreads)
vad_packet32_kernel(int32_t* __restrict__ vad_words, vad::Params vp, int16_t* pcm, uint8_t* vad_out, int n_streams, size_t stride)
{
    constexpr int L = 320;
    extern __shared__ __align__(16) int32_t pkt_smem[];
    const int s0 = blockIdx.x * kPktThreads, s = s0 + threadIdx.x;
    const int rows = min(kPktThreads, n_streams - s0);
    int16_t* x = pkt_stage_in<L>(pkt_smem, pcm, s0, rows);
    if (s < n_streams) {
        SoaWords st{vad_words + s, stride};
        const int flag = vad::process_packet32<80>(st, x, vp);
        if (vad_out) vad_out[s] = (uint8_t)flag;
    }
    pkt_stage_out<L>(pkt_smem, pcm, s0, rows);
}

__global__ void words_init_kernel(int32_t* words, const int32_t* init, int n_words, size_t stride, int first, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    for (int w = 0; w < n_words; ++w) words[(size_t)w * stride + first + i] = init[w];
}

// ------------------------------------------------------------------------------------------
// conference bus
// ------------------------------------------------------------------------------------------
// grid = (n_conf, chunks): thread i owns sample i of the frame and walks a slice of the
// conference's members, so every load is a coalesced row segment; slices are combined with one
// int32 atomicAdd per sample (exact and order-independent).  LAW < 0: PCM input, else G.711.
template <int LAW>
__global__ void bus_sum_kernel(const void* __restrict__ src, int32_t* __restrict__ bus, const int32_t* __restrict__ conf_start,
                               int frame, int chunk)
{
    const int c = blockIdx.x;
    const int first = conf_start[c] + blockIdx.y * chunk;
    const int last = min(conf_start[c + 1], first + chunk);
    if (first >= last) return;
    for (int i = threadIdx.x; i < frame; i += blockDim.x) {
        int32_t acc = 0;
        for (int p = first; p < last; ++p) {
            if (LAW < 0) acc += static_cast<const int16_t*>(src)[(size_t)p * frame + i];
            else if (LAW == 0) acc += alaw2linear(static_cast<const uint8_t*>(src)[(size_t)p * frame + i]);
            else acc += ulaw2linear(static_cast<const uint8_t*>(src)[(size_t)p * frame + i]);
        }
        if (gridDim.y == 1) bus[(size_t)c * frame + i] = acc;
        else atomicAdd(&bus[(size_t)c * frame + i], acc);
    }
}

// int16 PCM legs, frame a multiple of 8: a thread owns EIGHT consecutive samples of one conference (one 16-byte load per member
// row, a row of 160 samples is 20 such threads) instead of one sample and a 2-byte load per member — a third of the load
// instructions for the same bytes.  grid.y slices big conferences exactly like bus_sum_kernel.
__global__ void __launch_bounds__(256)
bus_sum_vec_kernel(const int16_t* __restrict__ src, int32_t* __restrict__ bus, const int32_t* __restrict__ conf_start, int n_conf, int frame,
                   int chunk, int slices)
{
    const int vpr = frame >> 3;                                        // 16-byte vectors per row
    const long long total = (long long)n_conf * vpr;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx / vpr), v = (int)(idx - (long long)c * vpr);
        const int first = conf_start[c] + (int)blockIdx.y * chunk;
        const int last = min(conf_start[c + 1], first + chunk);
        if (first >= last) continue;
        int32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint4* row = reinterpret_cast<const uint4*>(src) + (size_t)first * vpr + v;
#pragma unroll 4
        for (int p = first; p < last; ++p, row += vpr) {
            const uint4 w = *row;
            const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[2 * k] += (int32_t)(int16_t)(u[k] & 0xFFFFu);
                acc[2 * k + 1] += (int32_t)u[k] >> 16;
            }
        }
        int32_t* dst = bus + (size_t)c * frame + 8 * v;
        if (slices == 1) {
            reinterpret_cast<int4*>(dst)[0] = make_int4(acc[0], acc[1], acc[2], acc[3]);
            reinterpret_cast<int4*>(dst)[1] = make_int4(acc[4], acc[5], acc[6], acc[7]);
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(dst + k, acc[k]);
        }
    }
}

template <int LAW>
__global__ void nminus1_kernel(const int32_t* __restrict__ bus, const void* __restrict__ own, void* __restrict__ out,
                               const int32_t* __restrict__ conf_of, int frame, size_t total)
{
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t s = idx / frame;
        const int i = (int)(idx - s * frame);
        const int32_t b = bus[(size_t)conf_of[s] * frame + i];
        if (LAW < 0) {
            static_cast<int16_t*>(out)[idx] = sat16(b - static_cast<const int16_t*>(own)[idx]);
        } else {
            const uint8_t code = static_cast<const uint8_t*>(own)[idx];
            const int16_t mine = LAW == 0 ? alaw2linear(code) : ulaw2linear(code);
            const int16_t v = sat16(b - mine);
            static_cast<uint8_t*>(out)[idx] = LAW == 0 ? linear2alaw(v) : linear2ulaw(v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// G.711 element-wise, 8 samples (16 B of PCM / 8 B of codes) per thread-iteration
// ------------------------------------------------------------------------------------------
template <int LAW>
__global__ void g711_encode_kernel(const int16_t* __restrict__ pcm, uint8_t* __restrict__ codes, size_t n)
{
    const size_t nvec = n / 8;
    const int4* p4 = reinterpret_cast<const int4*>(pcm);
    uint2* c2 = reinterpret_cast<uint2*>(codes);
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const int4 x = p4[v];
        const int w[4] = {x.x, x.y, x.z, x.w};
        uint32_t o[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int16_t smp = (int16_t)((k & 1) ? (w[k >> 1] >> 16) : (w[k >> 1] & 0xFFFF));
            const uint32_t code = LAW == 0 ? linear2alaw(smp) : linear2ulaw(smp);
            o[k >> 2] |= code << (8 * (k & 3));
        }
        c2[v] = make_uint2(o[0], o[1]);
    }
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        codes[i] = LAW == 0 ? linear2alaw(pcm[i]) : linear2ulaw(pcm[i]);
}

template <int LAW>
__global__ void g711_decode_kernel(const uint8_t* __restrict__ codes, int16_t* __restrict__ pcm, size_t n)
{
    const size_t nvec = n / 8;
    const uint2* c2 = reinterpret_cast<const uint2*>(codes);
    int4* p4 = reinterpret_cast<int4*>(pcm);
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
        const uint2 c = c2[v];
        const uint32_t w[2] = {c.x, c.y};
        int o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t a = (uint8_t)(w[k >> 1] >> (16 * (k & 1))), b = (uint8_t)(w[k >> 1] >> (16 * (k & 1) + 8));
            const int16_t lo = LAW == 0 ? alaw2linear(a) : ulaw2linear(a);
            const int16_t hi = LAW == 0 ? alaw2linear(b) : ulaw2linear(b);
            o[k] = pack16(lo, hi);
        }
        p4[v] = make_int4(o[0], o[1], o[2], o[3]);
    }
    for (size_t i = nvec * 8 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        pcm[i] = LAW == 0 ? alaw2linear(codes[i]) : ulaw2linear(codes[i]);
}

// same-format branch of wmix_load_data on a device ring (R:src/wmix.c:1678-1702)
__global__ void mix_load_kernel(int16_t* ring, uint32_t ring_len, uint32_t pos, const int16_t* __restrict__ src,
                                uint32_t n, int rdce)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t at = (pos + i) % ring_len;
        ring[at] = mix_step(ring[at], src[i], rdce);
    }
}

// ------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------
constexpr int kPipe = 8;   // CUDA streams of the host-buffer chunk pipeline
struct wmixb_engine {
    wmixb_config cfg;
    int frame = 0, ana = 0, sm_count = 0;
    int row = 0;                            // samples per stream and tick in the caller's buffers: frame, or 320 for a 32 kHz engine
    int16_t* pack32 = nullptr;              // [n][160] staging of a 32 kHz engine's NS stage
    uint8_t* d_codes = nullptr;             // 2 x [n][frame] G.711 codes in / out of wmixb_tick_host_g711
    size_t stride = 0;                      // SoA row pitch (streams rounded up to 32)
    float* ns_rec = nullptr;
    float* ns_hb = nullptr;                 // [n][OVERLAP] high-band history (cfg.ns_high_band: wmix's stereo NS)
    int16_t* ns_stage = nullptr;            // 4 x [n][frame] staging of wmixb_ns2_host
    uint16_t* ns_hist = nullptr;
    void* ns_tables = nullptr;
    uint32_t* nsx_rec = nullptr;            // cfg.ns_core = 1: records of the fixed-point suppressor (nsx.cuh); ns_hist holds its histograms
    void* nsx_tables = nullptr;
    int16_t* nsx_hb = nullptr;              // [n][kKeep] second-band history (cfg.ns_high_band with ns_core = 1)
    int32_t nsx_thr_lrt = 0;
    int nsx_grid = 0, nsx_cfg = 0;
    int nsx_sync = 1;                       // nsx_kernel: the warps of a CTA start every frame together
    int32_t* agc_words = nullptr;
    int32_t* vad_words = nullptr;
    int32_t* agc_table = nullptr;
    int32_t* agc_init = nullptr;
    int32_t* vad_init = nullptr;
    vad::Params vp{}, vp20{};
    int16_t *d_in = nullptr, *d_out = nullptr;   // staging for the host-buffer entry point
    int16_t* d_pkt20 = nullptr;             // staging of wmixb_vad20_host
    uint8_t* d_vad = nullptr;
    int32_t* conf_start = nullptr;          // [n_conf+1]
    int32_t* conf_of = nullptr;             // [n_streams]
    int n_conf = 0, max_conf = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[kPipe] = {};                        // chunk pipeline of the host-buffer tick
    cudaEvent_t pipe_ev[kPipe] = {};
    // pipelined submission (wmixb_tick_host_submit / _wait): two ticks in flight, d_out double-buffered
    int16_t* d_out2 = nullptr;
    cudaEvent_t tail_ev[kPipe] = {}, done_ev[2] = {nullptr, nullptr};
    unsigned long long submitted = 0, waited = 0;
    int32_t* d_bus = nullptr;               // staging of the conference bus for the host-buffer tick
    size_t d_bus_bytes = 0;
    int ns_grid = 0, ns_staged_grid = 0;
    int ns_offline_staged = 1;              // several frames per launch: records staged in shared memory (persistent offline mode)
    int ns_cfg = 0;                         // index into kNsCfgs: 2 CTAs of 8 worker warps + 1 reducer warp
    int ns_align = 1;                       // CTA barrier at the top of every frame (instruction-cache sharing)
    int post_occ = 0;                       // post_kernel shape: 0 = automatic (see run_stages), 2..5 = CTAs of 128 threads per SM, 22 = one aligned CTA of 704
    int host_chunks_sync = 8, host_chunks_pipe = 4, host_lanes = 0;   // chunk pipeline of the host-buffer tick (wmixb_set_tuning)
    // small engines (the drop-in handles: one or two streams): the blocking host tick runs the kernels straight on a pinned,
    // device-mapped staging block — no copy engine round trips, one launch + one stream synchronisation per call
    int host_zero_copy = 1;
    char* zc_host = nullptr;                 // [in | out | vad], kZcBytes each
    char* zc_dev = nullptr;
    float* aec_rec = nullptr;               // [n_streams][aec_rec_floats]
    void* aec_tables = nullptr;
    int* aec_result = nullptr;              // [2] flags OR, flagged count
    int16_t* aec_stage = nullptr;           // far / near / out staging of the host-buffer entry point
    int aec_depth = 0, aec_grid = 0, aec_grid_max = 0, aec_pf = 2, aec_align = 1, aec_warps = 16;
    size_t aec_rec_floats = 0;
};

static int nsx_rec_words(const wmixb_engine* e) { return e->ana == 256 ? nsx::Geo<256>::kRecWords : nsx::Geo<128>::kRecWords; }
static int ns_rec_floats(const wmixb_engine* e) { return e->ana == 256 ? ns::Geo<256>::kRecFloats : ns::Geo<128>::kRecFloats; }

// compiled shapes of the NS kernel; wmixb_set_tuning("ns_cfg", index) picks one (default 0).
// cta = 1: ns_cta_kernel, `warps` WORKER warps + one reducer warp per CTA; cta = 0: ns_kernel (one self-contained warp per stream)
constexpr int kNsStagedWorkers = 7;
struct NsCfg { int cta, warps, minb; };
static const NsCfg kNsCfgs[] = {{1, 8, 2}, {1, 6, 3}, {1, 4, 4}, {1, 4, 5}, {1, 7, 2}, {1, 5, 3}, {0, 10, 2}, {0, 8, 2}, {0, 20, 1}, {0, 24, 1}, {0, 16, 1}, {0, 28, 1}};
template <int ANA>
static const void* ns_fn(int cfg)
{
    switch (cfg) {
    case 1: return (const void*)ns_cta_kernel<ANA, 6, 3>;
    case 2: return (const void*)ns_cta_kernel<ANA, 4, 4>;
    case 3: return (const void*)ns_cta_kernel<ANA, 4, 5>;
    case 4: return (const void*)ns_cta_kernel<ANA, 7, 2>;
    case 5: return (const void*)ns_cta_kernel<ANA, 5, 3>;
    case 6: return (const void*)ns_kernel<ANA, 10, 2>;
    case 7: return (const void*)ns_kernel<ANA, 8, 2>;
    case 8: return (const void*)ns_kernel<ANA, 20, 1>;
    case 9: return (const void*)ns_kernel<ANA, 24, 1>;
    case 10: return (const void*)ns_kernel<ANA, 16, 1>;
    case 11: return (const void*)ns_kernel<ANA, 28, 1>;
    default: return (const void*)ns_cta_kernel<ANA, 8, 2>;
    }
}
static int ns_threads(int cfg) { return (kNsCfgs[cfg].warps + kNsCfgs[cfg].cta) * 32; }
template <int ANA>
static size_t ns_cfg_smem(int cfg) { return kNsCfgs[cfg].cta ? ns_cta_smem_bytes<ANA>(kNsCfgs[cfg].warps) : ns_smem_bytes<ANA>(kNsCfgs[cfg].warps); }

template <int ANA>
static int launch_ns(wmixb_engine* e, int grid, cudaStream_t st, const int16_t* in, int16_t* out, int first, int n, int n_frames)
{
    float* rec = e->ns_rec + (size_t)first * ns::Geo<ANA>::kRecFloats;
    uint16_t* hist = e->ns_hist + (size_t)first * 3 * ns::kHistBins;
    const ns::Tables<ANA>* T = (const ns::Tables<ANA>*)e->ns_tables;
    int align = e->ns_align;
    void* args[] = {&rec, &hist, &T, &in, &out, &n, &n_frames, &align};   // the CTA form takes the first seven
    if (n_frames > 1 && e->ns_offline_staged) {
        // persistent offline mode: state on chip for the whole run of frames (7 workers + reducer, 2 CTAs per SM: 225 KB of smem)
        const int need = (n + kNsStagedWorkers - 1) / kNsStagedWorkers;
        const int g = need < e->ns_staged_grid ? need : e->ns_staged_grid;
        CK(cudaLaunchKernel((const void*)ns_cta_kernel<ANA, kNsStagedWorkers, 2, true>, dim3(g), dim3((kNsStagedWorkers + 1) * 32), args,
                            ns_cta_smem_bytes<ANA>(kNsStagedWorkers, true), st));
        return WMIXB_OK;
    }
    CK(cudaLaunchKernel(ns_fn<ANA>(e->ns_cfg), dim3(grid), dim3(ns_threads(e->ns_cfg)), args, ns_cfg_smem<ANA>(e->ns_cfg), st));
    return WMIXB_OK;
}

// launch shape of the NS kernel for e->ns_cfg: opt into its dynamic shared memory, size the persistent grid
template <int ANA>
static int ns_configure(wmixb_engine* e)
{
    const void* fn = ns_fn<ANA>(e->ns_cfg);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ns_cfg_smem<ANA>(e->ns_cfg)));
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, ns_threads(e->ns_cfg), ns_cfg_smem<ANA>(e->ns_cfg)));
    if (per_sm < 1) per_sm = 1;
    e->ns_grid = e->sm_count * per_sm;
    // the staged (persistent offline) shape
    const void* sfn = (const void*)ns_cta_kernel<ANA, kNsStagedWorkers, 2, true>;
    const size_t ssm = ns_cta_smem_bytes<ANA>(kNsStagedWorkers, true);
    CK(cudaFuncSetAttribute(sfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
    CK(cudaFuncSetAttribute(sfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sfn, (kNsStagedWorkers + 1) * 32, ssm));
    if (per_sm < 1) per_sm = 1;
    e->ns_staged_grid = e->sm_count * per_sm;
    return WMIXB_OK;
}

template <int ANA>
static int upload_ns_tables(wmixb_engine* e)
{
    ns::Tables<ANA> T;
    memset(&T, 0, sizeof T);
    host::dmath_tables(T.dm.log_invc, T.dm.log_logc, T.dm.exp_2jn);
    host::ns_window(ANA, ns::Geo<ANA>::kBlock, T.window);
    host::fft_w_table(ANA / 4, T.w);
    host::fft_c_table(ANA / 4, T.c);
    host::ns_log_table(ANA / 2 + 1, T.log_i, &T.sum_log_i, &T.sum_log_i_sq);
    if (host::ns_policy(e->cfg.ns_policy, &T.overdrive, &T.floor_gain, &T.gainmap) != 0) {
        snprintf(g_err, sizeof g_err, "ns_policy %d out of range 0..3", e->cfg.ns_policy);
        return WMIXB_EINVAL;
    }
    CK(cudaMalloc(&e->ns_tables, sizeof T));
    CK(cudaMemcpy(e->ns_tables, &T, sizeof T, cudaMemcpyHostToDevice));
    // measured (profiles/r2_*): at 16 kHz the CTA-cooperative kernel (8 workers + reducer, 2 CTAs per SM) is the faster one,
    // at 8 kHz — half the bins per stream, so half as much lane-sparse work to hand over — the self-contained warp kernel is
    e->ns_cfg = ANA == 256 ? 0 : 6;
    return ns_configure<ANA>(e);
}

// compiled shapes of the NSX kernel; wmixb_set_tuning("nsx_cfg", index) picks one
struct NsxCfg { int warps, minb; };
static const NsxCfg kNsxCfgs[] = {{8, 2}, {8, 3}, {8, 4}, {4, 6}, {16, 1}, {8, 1}, {24, 1}, {32, 1}, {16, 2}, {12, 2}};
template <int ANA>
static const void* nsx_fn(int cfg)
{
    switch (cfg) {
    case 1: return (const void*)nsx_kernel<ANA, 8, 3>;
    case 2: return (const void*)nsx_kernel<ANA, 8, 4>;
    case 3: return (const void*)nsx_kernel<ANA, 4, 6>;
    case 4: return (const void*)nsx_kernel<ANA, 16, 1>;
    case 5: return (const void*)nsx_kernel<ANA, 8, 1>;
    case 6: return (const void*)nsx_kernel<ANA, 24, 1>;
    case 7: return (const void*)nsx_kernel<ANA, 32, 1>;
    case 8: return (const void*)nsx_kernel<ANA, 16, 2>;
    case 9: return (const void*)nsx_kernel<ANA, 12, 2>;
    default: return (const void*)nsx_kernel<ANA, 8, 2>;
    }
}
template <int ANA>
static int nsx_configure(wmixb_engine* e)
{
    const void* fn = nsx_fn<ANA>(e->nsx_cfg);
    const int threads = kNsxCfgs[e->nsx_cfg].warps * 32;
    const size_t smem = nsx_smem_bytes<ANA>(kNsxCfgs[e->nsx_cfg].warps);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, smem));
    if (per_sm < 1) per_sm = 1;
    e->nsx_grid = e->sm_count * per_sm;
    return WMIXB_OK;
}
template <int ANA>
static int launch_nsx(wmixb_engine* e, cudaStream_t st, const int16_t* in, int16_t* out, int first, int n, int n_frames)
{
    uint32_t* rec = e->nsx_rec + (size_t)first * nsx::Geo<ANA>::kRecWords;
    int16_t* hist = reinterpret_cast<int16_t*>(e->ns_hist) + (size_t)first * 3 * nsx::kHistBins;
    const nsx::Tables* T = (const nsx::Tables*)e->nsx_tables;
    int align = e->nsx_sync;
    const int nw = kNsxCfgs[e->nsx_cfg].warps;
    const int need = (n + nw - 1) / nw;
    const int grid = need < e->nsx_grid ? need : e->nsx_grid;
    void* args[] = {&rec, &hist, &T, &in, &out, &n, &n_frames, &align};
    CK(cudaLaunchKernel(nsx_fn<ANA>(e->nsx_cfg), dim3(grid), dim3(nw * 32), args, nsx_smem_bytes<ANA>(nw), st));
    return WMIXB_OK;
}
static int upload_nsx_tables(wmixb_engine* e)
{
    nsx::Tables T;
    if (host::nsx_tables(e->cfg.freq, e->cfg.ns_policy, &T, &e->nsx_thr_lrt) != 0) {
        snprintf(g_err, sizeof g_err, "ns_policy %d out of range 0..3", e->cfg.ns_policy);
        return WMIXB_EINVAL;
    }
    CK(cudaMalloc(&e->nsx_tables, sizeof T));
    CK(cudaMemcpy(e->nsx_tables, &T, sizeof T, cudaMemcpyHostToDevice));
    // measured (profiles/r2_p_nsx_sweep.jsonl): ONE CTA of 32 warps per SM (64 registers), its warps starting every frame
    // together, is the fastest shape at both rates — 0.750 ms per 100 000-stream tick against 1.008 ms for 2 CTAs of 8 warps at
    // 128 registers: the frame is ~90 KB of straight-line code and what the warps of an SM share of it decides the speed
    e->nsx_cfg = 7;
    return e->ana == 256 ? nsx_configure<256>(e) : nsx_configure<128>(e);
}

static int upload_agc_table(wmixb_engine* e, int gain_db)
{
    int32_t tab[32];
    if (host::agc_gain_table(tab, (int16_t)gain_db, 0, 0, host::agc_analog_target((int16_t)gain_db)) != 0) {
        snprintf(g_err, sizeof g_err, "agc gain %d dB out of range", gain_db);
        return WMIXB_EINVAL;
    }
    CK(cudaMemcpyAsync(e->agc_table, tab, sizeof tab, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_reset(wmixb_engine* e, int first, int count)
{
    if (!e || first < 0 || count < 0 || first + count > e->cfg.n_streams) return WMIXB_EINVAL;
    if (count == 0) return WMIXB_OK;
    CK(cudaSetDevice(e->cfg.device));
    if (e->ns_rec) {
        const int blocks = (count * 32 + 255) / 256;
        if (e->ana == 256) ns_init_kernel<256><<<blocks, 256, 0, e->stream>>>(e->ns_rec, e->ns_hist, first, count);
        else ns_init_kernel<128><<<blocks, 256, 0, e->stream>>>(e->ns_rec, e->ns_hist, first, count);
        CK_LAUNCH();
        if (e->ns_hb) {
            const size_t ov = e->ana == 256 ? ns::Geo<256>::kOverlap : ns::Geo<128>::kOverlap;
            CK(cudaMemsetAsync(e->ns_hb + (size_t)first * ov, 0, (size_t)count * ov * sizeof(float), e->stream));
        }
    }
    if (e->nsx_rec) {
        const int blocks = (count * 32 + 255) / 256;
        int16_t* hist = reinterpret_cast<int16_t*>(e->ns_hist);
        if (e->ana == 256) nsx_init_kernel<256><<<blocks, 256, 0, e->stream>>>(e->nsx_rec, hist, first, count, e->nsx_thr_lrt);
        else nsx_init_kernel<128><<<blocks, 256, 0, e->stream>>>(e->nsx_rec, hist, first, count, e->nsx_thr_lrt);
        CK_LAUNCH();
        if (e->nsx_hb) CK(cudaMemsetAsync(e->nsx_hb + (size_t)first * (e->ana - e->frame), 0, (size_t)count * (e->ana - e->frame) * sizeof(int16_t), e->stream));
    }
    if (e->aec_rec) {
        aec_init_kernel<<<(count * 32 + 255) / 256, 256, 0, e->stream>>>(e->aec_rec, e->aec_rec_floats, e->aec_depth, first, count);
        CK_LAUNCH();
    }
    if (e->agc_words) {
        words_init_kernel<<<(count + 255) / 256, 256, 0, e->stream>>>(e->agc_words, e->agc_init, agc::N_WORDS, e->stride, first, count);
        CK_LAUNCH();
    }
    if (e->vad_words) {
        words_init_kernel<<<(count + 255) / 256, 256, 0, e->stream>>>(e->vad_words, e->vad_init, vad::N_WORDS, e->stride, first, count);
        CK_LAUNCH();
    }
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" void wmixb_destroy(wmixb_engine* e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->ns_rec); cudaFree(e->nsx_rec); cudaFree(e->nsx_tables); cudaFree(e->nsx_hb); cudaFree(e->ns_hist); cudaFree(e->ns_tables); cudaFree(e->ns_hb); cudaFree(e->ns_stage);
    cudaFree(e->agc_words); cudaFree(e->vad_words); cudaFree(e->agc_table);
    cudaFree(e->agc_init); cudaFree(e->vad_init);
    cudaFree(e->d_in); cudaFree(e->d_out); cudaFree(e->pack32); cudaFree(e->d_codes); cudaFree(e->d_out2); cudaFree(e->d_vad); cudaFree(e->d_pkt20);
    for (int k = 0; k < kPipe; ++k) if (e->tail_ev[k]) cudaEventDestroy(e->tail_ev[k]);
    for (int k = 0; k < 2; ++k) if (e->done_ev[k]) cudaEventDestroy(e->done_ev[k]);
    cudaFree(e->conf_start); cudaFree(e->conf_of);
    cudaFree(e->aec_rec); cudaFree(e->aec_tables); cudaFree(e->aec_result); cudaFree(e->aec_stage);
    for (int k = 0; k < kPipe; ++k) {
        if (e->pipe[k]) cudaStreamDestroy(e->pipe[k]);
        if (e->pipe_ev[k]) cudaEventDestroy(e->pipe_ev[k]);
    }
    cudaFree(e->d_bus);
    if (e->zc_host) cudaFreeHost(e->zc_host);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

static int create_impl(const wmixb_config* cfg, wmixb_engine* e)
{
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        snprintf(g_err, sizeof g_err, "no CUDA device (%s) — wmix_b200 has no CPU path", cudaGetErrorString(ce));
        return WMIXB_ENODEV;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { snprintf(g_err, sizeof g_err, "device %d of %d", cfg->device, ndev); return WMIXB_EINVAL; }
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    e->sm_count = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    const size_t n = (size_t)cfg->n_streams;
    e->stride = (n + 31) / 32 * 32;
    if ((cfg->stages & WMIXB_NS) && cfg->ns_core == 1) {
        // the fixed-point suppressor (R:src/webrtc.c:512, MAKE_WEBRTC_NSX)
        const size_t words = e->ana == 256 ? nsx::Geo<256>::kRecWords : nsx::Geo<128>::kRecWords;
        if (cfg->ns_high_band) {
            CK(cudaMalloc(&e->nsx_hb, n * (size_t)(e->ana - e->frame) * sizeof(int16_t)));
            if (e->ana == 256) CK(cudaFuncSetAttribute(nsx_hb_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsx_smem_bytes<256>(8)));
            else CK(cudaFuncSetAttribute(nsx_hb_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nsx_smem_bytes<128>(8)));
        }
        CK(cudaMalloc(&e->nsx_rec, n * words * sizeof(uint32_t)));
        CK(cudaMalloc(&e->ns_hist, n * 3 * nsx::kHistBins * sizeof(uint16_t)));
        int rc = upload_nsx_tables(e);
        if (rc) return rc;
    } else if (cfg->stages & WMIXB_NS) {
        CK(cudaMalloc(&e->ns_rec, n * ns_rec_floats(e) * sizeof(float)));
        CK(cudaMalloc(&e->ns_hist, n * 3 * ns::kHistBins * sizeof(uint16_t)));
        int rc = e->ana == 256 ? upload_ns_tables<256>(e) : upload_ns_tables<128>(e);
        if (rc) return rc;
        if (cfg->ns_high_band) {
            const size_t ov = e->ana == 256 ? ns::Geo<256>::kOverlap : ns::Geo<128>::kOverlap;
            CK(cudaMalloc(&e->ns_hb, n * ov * sizeof(float)));
            if (e->ana == 256) CK(cudaFuncSetAttribute(ns_hb_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ns_hb_smem_bytes<256>()));
            else CK(cudaFuncSetAttribute(ns_hb_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ns_hb_smem_bytes<128>()));
        }
    }
    if (cfg->stages & WMIXB_AEC) {
        e->aec_depth = cfg->aec_far_depth > 0 ? cfg->aec_far_depth : 32;
        if (e->aec_depth < 8 || e->aec_depth > 1024) { snprintf(g_err, sizeof g_err, "aec_far_depth %d outside 8..1024", e->aec_depth); return WMIXB_EINVAL; }
        e->aec_rec_floats = (size_t)aec::rec_floats(e->aec_depth);
        aec::Tables* T = new aec::Tables();
        memset(T, 0, sizeof *T);
        host::dmath_tables(T->dm.log_invc, T->dm.log_logc, T->dm.exp_2jn);
        host::aec_tables(T->w, T->c, T->hann, T->weight, T->over, T->lcg_mul, T->lcg_add);
        cudaError_t ce1 = cudaMalloc(&e->aec_tables, sizeof *T);
        if (ce1 == cudaSuccess) ce1 = cudaMemcpy(e->aec_tables, T, sizeof *T, cudaMemcpyHostToDevice);
        delete T;
        CK(ce1);
        CK(cudaMalloc(&e->aec_rec, n * e->aec_rec_floats * sizeof(float)));
        CK(cudaMalloc(&e->aec_result, 2 * sizeof(int)));
        CK(cudaFuncSetAttribute(aec_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)aec_smem_bytes(8)));
        CK(cudaFuncSetAttribute(aec_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)aec_smem_bytes(16)));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, aec_kernel<8>, kAecWarps * 32, kAecSmemBytes));
        if (per_sm < 1) per_sm = 1;
        e->aec_grid = e->sm_count * per_sm;
        e->aec_grid_max = e->aec_grid;
    }
    if (cfg->stages & WMIXB_AGC) {
        int32_t init[agc::N_WORDS];
        host::agc_initial_words(init);
        CK(cudaMalloc(&e->agc_words, e->stride * agc::N_WORDS * sizeof(int32_t)));
        CK(cudaMalloc(&e->agc_init, sizeof init));
        CK(cudaMemcpy(e->agc_init, init, sizeof init, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&e->agc_table, 32 * sizeof(int32_t)));
        int rc = upload_agc_table(e, cfg->agc_gain_db);
        if (rc) return rc;
    }
    if (cfg->stages & WMIXB_VAD) {
        int32_t init[vad::N_WORDS];
        host::vad_initial_words(init);
        CK(cudaMalloc(&e->vad_words, e->stride * vad::N_WORDS * sizeof(int32_t)));
        CK(cudaMalloc(&e->vad_init, sizeof init));
        CK(cudaMemcpy(e->vad_init, init, sizeof init, cudaMemcpyHostToDevice));
        int16_t th[4];
        if (host::vad_thresholds(cfg->vad_mode, 10, th) != 0) { snprintf(g_err, sizeof g_err, "vad_mode %d out of range 0..3", cfg->vad_mode); return WMIXB_EINVAL; }
        e->vp = vad::Params{th[0], th[1], th[2], th[3]};
        host::vad_thresholds(cfg->vad_mode, 20, th);
        e->vp20 = vad::Params{th[0], th[1], th[2], th[3]};
    }
    if (cfg->stages & (WMIXB_AGC | WMIXB_VAD)) {
        const int big = (int)(((size_t)704 * (e->frame / 2 + 1) + 32) * sizeof(int32_t));
        if (e->frame == 160) CK(cudaFuncSetAttribute(post_kernel<true, 1, 704>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
        else CK(cudaFuncSetAttribute(post_kernel<false, 1, 704>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    }
    CK(cudaMalloc(&e->d_in, n * e->row * sizeof(int16_t)));
    CK(cudaMalloc(&e->d_out, n * e->row * sizeof(int16_t)));
    if (e->row != e->frame && (cfg->stages & WMIXB_NS)) CK(cudaMalloc(&e->pack32, n * e->frame * sizeof(int16_t)));
    CK(cudaMalloc(&e->d_vad, n));
    CK(cudaMalloc(&e->conf_of, n * sizeof(int32_t)));
    return wmixb_reset(e, 0, cfg->n_streams);
}

extern "C" int wmixb_create(const wmixb_config* cfg, wmixb_engine** out)
{
    if (!cfg || !out) return WMIXB_EINVAL;
    *out = nullptr;
    if (cfg->n_streams < 1) { snprintf(g_err, sizeof g_err, "n_streams must be >= 1"); return WMIXB_EINVAL; }
    // the reference accepts freq <= 32000 && freq % 8000 == 0 (R:src/webrtc.c:43, :563, :711);
    // the batched engine covers the two rates BASELINE.json's configs use
    // 32000 runs on the 16 kHz cores exactly as the reference's handles do at that rate (see run_stages_32k); 24000 passes the
    // reference's rate test too but every WebRTC call then fails there, so it is refused here
    if (cfg->freq != 8000 && cfg->freq != 16000 && cfg->freq != 32000) { snprintf(g_err, sizeof g_err, "freq %d: batched engine supports 8000, 16000 and 32000", cfg->freq); return WMIXB_EINVAL; }
    if (cfg->freq == 32000 && ((cfg->stages & WMIXB_AEC) || cfg->ns_high_band)) { snprintf(g_err, sizeof g_err, "32 kHz engine: NS, AGC and VAD only (the reference's AEC stops at 16 kHz, R:src/webrtc.c:233)"); return WMIXB_EINVAL; }
    if ((cfg->stages & ~(WMIXB_NS | WMIXB_AGC | WMIXB_VAD | WMIXB_AEC)) != 0) { snprintf(g_err, sizeof g_err, "unknown stage bits"); return WMIXB_EINVAL; }
    if (cfg->ns_core != 0 && cfg->ns_core != 1) { snprintf(g_err, sizeof g_err, "ns_core %d: 0 = float core, 1 = fixed-point core", cfg->ns_core); return WMIXB_EINVAL; }
    wmixb_engine* e = new (std::nothrow) wmixb_engine();
    if (!e) return WMIXB_ENOMEM;
    e->cfg = *cfg;
    e->row = cfg->freq / 100;
    e->frame = cfg->freq == 32000 ? 160 : e->row;          // the cores' frame: a 32 kHz engine is the 16 kHz machinery on 320-sample rows
    e->ana = cfg->freq == 8000 ? 128 : 256;
    const int rc = create_impl(cfg, e);
    if (rc != WMIXB_OK) { wmixb_destroy(e); return rc; }
    *out = e;
    return WMIXB_OK;
}

extern "C" int wmixb_set_agc_gain(wmixb_engine* e, int gain_db)
{
    if (!e || !e->agc_table) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    return upload_agc_table(e, gain_db);
}

// row_stride: samples between consecutive streams' rows (0 = packed rows of `samples`)
static int launch_aec(wmixb_engine* e, const int16_t* d_far, const int16_t* d_near, int16_t* d_out, int samples, int delay_ms,
                      cudaStream_t st, int row_stride = 0)
{
    if (row_stride == 0) row_stride = samples;
    const int n = e->cfg.n_streams;
    const int aw = e->aec_warps;
    const int need = (n + aw - 1) / aw;
    const int grid = need < e->aec_grid ? need : e->aec_grid;
    if (aw == 16) aec_kernel<16><<<(grid + 1) / 2, 512, aec_smem_bytes(16), st>>>(e->aec_rec, e->aec_rec_floats, (const aec::Tables*)e->aec_tables, d_far,
                                                             d_near, d_out, n, samples, e->cfg.freq / 8000, e->aec_depth, delay_ms, e->aec_pf, row_stride, e->aec_align);
    else aec_kernel<8><<<grid, 256, aec_smem_bytes(8), st>>>(e->aec_rec, e->aec_rec_floats, (const aec::Tables*)e->aec_tables, d_far,
                                                             d_near, d_out, n, samples, e->cfg.freq / 8000, e->aec_depth, delay_ms, e->aec_pf, row_stride, e->aec_align);
    CK_LAUNCH();
    return WMIXB_OK;
}

// Streams [first, first+n) of the engine; d_in / d_out / d_vad / d_far point at stream `first`.
// `first` must be a multiple of 32 (SoA rows stay line-aligned).
static int run_stages(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int n_frames, int stages,
                      cudaStream_t st, const int16_t* d_far = nullptr, int delay_ms = 0, int first = 0, int n = -1);

// One 10 ms tick of a 32 kHz engine: rows of 320 samples through the 16 kHz cores, the way the reference's own handles treat
// that rate.  NS (R:src/webrtc.c:563-644): WebRtcNs(x) keeps its 160-sample block at every rate above 8 kHz and wmix hands it
// channels, not bands, so the first 160 samples of the row are suppressed and the rest of the reference's calloc'ed output
// stays zero.  AGC (R:src/webrtc.c:724-728, :786-818): 5 ms packets of 160 samples through the 16 kHz path — the row is two
// consecutive frames.  VAD (R:src/webrtc.c:56-67; CalcVad32khz, T:.../vad/vad_core.c:623-643): one 320-sample packet,
// decimated 32k -> 16k -> 8k.  Whole engine per call.
static int run_stages_32k(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int stages, cudaStream_t st)
{
    const int n = e->cfg.n_streams;
    const size_t row_b = (size_t)e->row * 2, core_b = (size_t)e->frame * 2;
    const int16_t* cur = d_in;
    if (stages & WMIXB_NS) {
        CK(cudaMemcpy2DAsync(e->pack32, core_b, cur, row_b, core_b, (size_t)n, cudaMemcpyDeviceToDevice, st));
        const int rc = run_stages(e, e->pack32, e->pack32, nullptr, 1, WMIXB_NS, st, nullptr, 0, 0, -2);
        if (rc) return rc;
        CK(cudaMemcpy2DAsync(d_out, row_b, e->pack32, core_b, core_b, (size_t)n, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemset2DAsync(d_out + e->frame, row_b, 0, row_b - core_b, (size_t)n, st));
        cur = d_out;
    }
    if (stages & WMIXB_AGC) {
        const int rc = run_stages(e, cur, d_out, nullptr, 2, WMIXB_AGC, st, nullptr, 0, 0, -2);   // [n][2][160]
        if (rc) return rc;
        cur = d_out;
    }
    if (cur == d_in && d_in != d_out) CK(cudaMemcpyAsync(d_out, d_in, (size_t)n * row_b, cudaMemcpyDeviceToDevice, st));
    if (stages & WMIXB_VAD) {
        vad_packet32_kernel<<<(n + kPktThreads - 1) / kPktThreads, kPktThreads, pkt_smem_bytes<320>(), st>>>(e->vad_words, e->vp, d_out, d_vad, n, e->stride);
        CK_LAUNCH();
    }
    return WMIXB_OK;
}

static int run_stages(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int n_frames, int stages,
                      cudaStream_t st, const int16_t* d_far, int delay_ms, int first, int n)
{
    if (stages == 0) stages = e->cfg.stages;
    if (e->row != e->frame && n != -2) {
        if (stages & ~e->cfg.stages) { snprintf(g_err, sizeof g_err, "stage mask 0x%x not configured (engine has 0x%x)", stages, e->cfg.stages); return WMIXB_EINVAL; }
        if (n_frames != 1 || first != 0 || (n >= 0 && n != e->cfg.n_streams)) { snprintf(g_err, sizeof g_err, "32 kHz engine: one tick of the whole engine per call"); return WMIXB_EINVAL; }
        return run_stages_32k(e, d_in, d_out, d_vad, stages, st);
    }
    if (n == -2) n = -1;                                     // a core stage of the 32 kHz path
    if (stages & ~e->cfg.stages) { snprintf(g_err, sizeof g_err, "stage mask 0x%x not configured (engine has 0x%x)", stages, e->cfg.stages); return WMIXB_EINVAL; }
    if (n < 0) n = e->cfg.n_streams - first;
    const int16_t* cur = d_in;
    if ((stages & WMIXB_NS) && e->nsx_rec) {
        const int rc = e->ana == 256 ? launch_nsx<256>(e, st, cur, d_out, first, n, n_frames) : launch_nsx<128>(e, st, cur, d_out, first, n, n_frames);
        if (rc) return rc;
        CK_LAUNCH();
        cur = d_out;
    } else if (stages & WMIXB_NS) {
        const int nw = kNsCfgs[e->ns_cfg].warps;
        const int need = (n + nw - 1) / nw;
        const int grid = need < e->ns_grid ? need : e->ns_grid;
        const int rc = e->ana == 256 ? launch_ns<256>(e, grid, st, cur, d_out, first, n, n_frames) : launch_ns<128>(e, grid, st, cur, d_out, first, n, n_frames);
        if (rc) return rc;
        CK_LAUNCH();
        cur = d_out;
    }
    if (stages & WMIXB_AEC) {
        if (n_frames != 1 || !d_far || first != 0 || n != e->cfg.n_streams) { snprintf(g_err, sizeof g_err, "the AEC stage needs a far-end buffer and runs one tick of the whole engine per call"); return WMIXB_EINVAL; }
        const int rc = launch_aec(e, d_far, cur, d_out, e->frame, delay_ms, st);
        if (rc) return rc;
        cur = d_out;
    }
    if (stages & (WMIXB_AGC | WMIXB_VAD)) {
        // shapes: post_occ = 2..5 CTAs of 128 threads per SM (231 / 164 / 128 / 96 registers); 22 = ONE CTA of 704 threads per SM
        // (22 warps, 80 registers): the smallest register budget at which 100 000 streams on 148 SMs are a single wave
        // (148 x 704 = 104 192 resident threads) instead of 1.76 waves of 12-warp SMs
        // default (post_occ = 0): the aligned big-CTA shape once the batch fills a few of them, the small shape below that
        const int occ = e->post_occ ? e->post_occ : (n >= 4 * 704 ? 22 : 3);
        const int pthreads = occ == 22 ? 704 : kPostThreads;
        const int grid = (n + pthreads - 1) / pthreads;
        const size_t psmem = ((size_t)pthreads * (e->frame / 2 + 1) + 32) * sizeof(int32_t);
        int32_t* aw = e->agc_words ? e->agc_words + first : nullptr;
        int32_t* vw = e->vad_words ? e->vad_words + first : nullptr;
#define WMX_POST(FS, MB, TH) post_kernel<FS, MB, TH><<<grid, TH, psmem, st>>>(aw, vw, e->agc_table, e->vp, cur, d_out, d_vad, n, e->stride, n_frames, stages)
        if (e->frame == 160) { if (occ == 2) WMX_POST(true, 2, 128); else if (occ == 4) WMX_POST(true, 4, 128); else if (occ == 5) WMX_POST(true, 5, 128); else if (occ == 22) WMX_POST(true, 1, 704); else WMX_POST(true, 3, 128); }
        else { if (occ == 2) WMX_POST(false, 2, 128); else if (occ == 4) WMX_POST(false, 4, 128); else if (occ == 5) WMX_POST(false, 5, 128); else if (occ == 22) WMX_POST(false, 1, 704); else WMX_POST(false, 3, 128); }
#undef WMX_POST
        CK_LAUNCH();
        cur = d_out;
    }
    if (cur == d_in && d_in != d_out) CK(cudaMemcpyAsync(d_out, d_in, (size_t)n * n_frames * e->frame * 2, cudaMemcpyDeviceToDevice, st));
    return WMIXB_OK;
}

extern "C" int wmixb_tick_device(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int stages, void* stream)
{
    if (!e || !d_in || !d_out) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    return run_stages(e, d_in, d_out, d_vad, 1, stages, (cudaStream_t)stream);
}

extern "C" int wmixb_vad20_device(wmixb_engine* e, int16_t* d_pcm, uint8_t* d_vad, void* stream)
{
    if (!e || !d_pcm) return WMIXB_EINVAL;
    if (!e->vad_words) { snprintf(g_err, sizeof g_err, "vad20: the engine was created without WMIXB_VAD"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams, grid = (n + 63) / 64;
    if (e->frame == 160) vad_packet20_kernel<true><<<grid, kPktThreads, pkt_smem_bytes<320>(), (cudaStream_t)stream>>>(e->vad_words, e->vp20, d_pcm, d_vad, n, e->stride);
    else vad_packet20_kernel<false><<<grid, kPktThreads, pkt_smem_bytes<160>(), (cudaStream_t)stream>>>(e->vad_words, e->vp20, d_pcm, d_vad, n, e->stride);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_vad20_host(wmixb_engine* e, int16_t* h_pcm, uint8_t* h_vad)
{
    if (!e || !h_pcm) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    if (!e->d_pkt20) CK(cudaMalloc(&e->d_pkt20, (size_t)e->cfg.n_streams * 2 * e->frame * sizeof(int16_t)));
    const size_t bytes = (size_t)e->cfg.n_streams * 2 * e->frame * sizeof(int16_t);
    CK(cudaMemcpyAsync(e->d_pkt20, h_pcm, bytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = wmixb_vad20_device(e, e->d_pkt20, e->d_vad, e->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_pcm, e->d_pkt20, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (h_vad) CK(cudaMemcpyAsync(h_vad, e->d_vad, (size_t)e->cfg.n_streams, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_ns2_device(wmixb_engine* e, const int16_t* d_in, const int16_t* d_in_hb, int16_t* d_out, int16_t* d_out_hb, void* stream)
{
    if (!e || !d_in || !d_in_hb || !d_out || !d_out_hb) return WMIXB_EINVAL;
    if (!e->ns_hb && !e->nsx_hb) { snprintf(g_err, sizeof g_err, "ns2: the engine was created without WMIXB_NS + ns_high_band"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    const int need = (n + 7) / 8, cap = e->sm_count * 2, grid = need < cap ? need : cap;
    if (e->nsx_hb) {
        int16_t* hist = reinterpret_cast<int16_t*>(e->ns_hist);
        const nsx::Tables* T = (const nsx::Tables*)e->nsx_tables;
        if (e->ana == 256) nsx_hb_kernel<256><<<grid, 256, nsx_smem_bytes<256>(8), (cudaStream_t)stream>>>(e->nsx_rec, hist, e->nsx_hb, T, d_in, d_out, d_in_hb, d_out_hb, n);
        else nsx_hb_kernel<128><<<grid, 256, nsx_smem_bytes<128>(8), (cudaStream_t)stream>>>(e->nsx_rec, hist, e->nsx_hb, T, d_in, d_out, d_in_hb, d_out_hb, n);
        CK_LAUNCH();
        return WMIXB_OK;
    }
    if (e->ana == 256)
        ns_hb_kernel<256><<<grid, 256, ns_hb_smem_bytes<256>(), (cudaStream_t)stream>>>(e->ns_rec, e->ns_hist, e->ns_hb, (const ns::Tables<256>*)e->ns_tables,
                                                                                      d_in, d_out, d_in_hb, d_out_hb, n);
    else
        ns_hb_kernel<128><<<grid, 256, ns_hb_smem_bytes<128>(), (cudaStream_t)stream>>>(e->ns_rec, e->ns_hist, e->ns_hb, (const ns::Tables<128>*)e->ns_tables,
                                                                                      d_in, d_out, d_in_hb, d_out_hb, n);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_ns2_host(wmixb_engine* e, const int16_t* h_in, const int16_t* h_in_hb, int16_t* h_out, int16_t* h_out_hb)
{
    if (!e || !h_in || !h_in_hb || !h_out || !h_out_hb) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const size_t cnt = (size_t)e->cfg.n_streams * e->frame, bytes = cnt * sizeof(int16_t);
    if (!e->ns_stage) CK(cudaMalloc(&e->ns_stage, 4 * bytes));
    int16_t *a = e->ns_stage, *b = a + cnt, *c = b + cnt, *d = c + cnt;
    CK(cudaMemcpyAsync(a, h_in, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(b, h_in_hb, bytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = wmixb_ns2_device(e, a, b, c, d, e->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_out, c, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaMemcpyAsync(h_out_hb, d, bytes, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_vad32_device(wmixb_engine* e, int16_t* d_pcm, uint8_t* d_vad, void* stream)
{
    if (!e || !d_pcm) return WMIXB_EINVAL;
    if (!e->vad_words || e->frame != 160) { snprintf(g_err, sizeof g_err, "vad32: needs a 16 kHz engine created with WMIXB_VAD"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams, grid = (n + 63) / 64;
    vad_packet32_kernel<<<grid, kPktThreads, pkt_smem_bytes<320>(), (cudaStream_t)stream>>>(e->vad_words, e->vp, d_pcm, d_vad, n, e->stride);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_vad32_host(wmixb_engine* e, int16_t* h_pcm, uint8_t* h_vad)
{
    if (!e || !h_pcm) return WMIXB_EINVAL;
    if (!e->vad_words || e->frame != 160) { snprintf(g_err, sizeof g_err, "vad32: needs a 16 kHz engine created with WMIXB_VAD"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const size_t bytes = (size_t)e->cfg.n_streams * 320 * sizeof(int16_t);      // same size as a 20 ms packet at 16 kHz
    if (!e->d_pkt20) CK(cudaMalloc(&e->d_pkt20, bytes));
    CK(cudaMemcpyAsync(e->d_pkt20, h_pcm, bytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = wmixb_vad32_device(e, e->d_pkt20, e->d_vad, e->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_pcm, e->d_pkt20, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (h_vad) CK(cudaMemcpyAsync(h_vad, e->d_vad, (size_t)e->cfg.n_streams, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_tick_chain_device(wmixb_engine* e, const int16_t* d_far, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad,
                                       int stages, int delay_ms, void* stream)
{
    if (!e || !d_in || !d_out) return WMIXB_EINVAL;
    if (delay_ms < 0 || delay_ms > 500) { snprintf(g_err, sizeof g_err, "delay_ms %d outside 0..500 (the reference returns -1 for it)", delay_ms); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    return run_stages(e, d_in, d_out, d_vad, 1, stages, (cudaStream_t)stream, d_far, delay_ms);
}

extern "C" int wmixb_aec_device(wmixb_engine* e, const int16_t* d_far, const int16_t* d_near, int16_t* d_out, int samples,
                                int delay_ms, void* stream)
{
    if (!e || !e->aec_rec || (!d_far && !d_near) || (d_near && !d_out)) return WMIXB_EINVAL;
    // WebRtcAec_BufferFarend / _Process accept 80 or 160 samples (T:.../aec/echo_cancellation.c:297, :362)
    if (samples != 80 && samples != 160) { snprintf(g_err, sizeof g_err, "aec: %d samples per call (80 or 160)", samples); return WMIXB_EINVAL; }
    if (samples % (80 * (e->cfg.freq / 8000)) != 0) { snprintf(g_err, sizeof g_err, "aec: %d samples is not a whole 10 ms at %d Hz", samples, e->cfg.freq); return WMIXB_EINVAL; }
    if (delay_ms < 0 || delay_ms > 500) { snprintf(g_err, sizeof g_err, "delay_ms %d outside 0..500 (the reference returns -1 for it)", delay_ms); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    return launch_aec(e, d_far, d_near, d_out, samples, delay_ms, (cudaStream_t)stream);
}

extern "C" int wmixb_aec_host(wmixb_engine* e, const int16_t* h_far, const int16_t* h_near, int16_t* h_out, int samples, int delay_ms)
{
    if (!e || !e->aec_rec || (!h_far && !h_near) || (h_near && !h_out)) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const size_t bytes = (size_t)e->cfg.n_streams * 160 * sizeof(int16_t);
    if (!e->aec_stage) CK(cudaMalloc(&e->aec_stage, 3 * bytes));
    int16_t* d_far = e->aec_stage;
    int16_t* d_near = e->aec_stage + (size_t)e->cfg.n_streams * 160;
    int16_t* d_out = d_near + (size_t)e->cfg.n_streams * 160;
    const size_t used = (size_t)e->cfg.n_streams * samples * sizeof(int16_t);
    if (h_far) CK(cudaMemcpyAsync(d_far, h_far, used, cudaMemcpyHostToDevice, e->stream));
    if (h_near) CK(cudaMemcpyAsync(d_near, h_near, used, cudaMemcpyHostToDevice, e->stream));
    const int rc = wmixb_aec_device(e, h_far ? d_far : nullptr, h_near ? d_near : nullptr, d_out, samples, delay_ms, e->stream);
    if (rc) return rc;
    if (h_near) CK(cudaMemcpyAsync(h_out, d_out, used, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_aec_status(wmixb_engine* e, int* h_flags, int* h_flagged)
{
    if (!e || !e->aec_rec) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemsetAsync(e->aec_result, 0, 2 * sizeof(int), e->stream));
    aec_status_kernel<<<(e->cfg.n_streams + 255) / 256, 256, 0, e->stream>>>(e->aec_rec, e->aec_rec_floats, e->cfg.n_streams, e->aec_result);
    CK_LAUNCH();
    int r[2] = {0, 0};
    CK(cudaMemcpyAsync(r, e->aec_result, sizeof r, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (h_flags) *h_flags = r[0];
    if (h_flagged) *h_flagged = r[1];
    return WMIXB_OK;
}

// ---- the daemon's record tick at its own 20 ms cadence (R:src/wmix.c:528-760), for every stream ----
struct wmixb_record {
    wmixb_engine* e = nullptr;
    int pkg = 0, n_pkg = 0, count = 0, delay_pkgs = 0;
    int16_t* fifo = nullptr;                // [n_pkg][n_streams][pkg]: slot-major, so a slot is a ready [n_streams][pkg] far-end batch
    int16_t* stage = nullptr;               // play + mic staging of wmixb_record_tick_host
};

extern "C" int wmixb_record_create(wmixb_engine* e, int aec_interval_ms, wmixb_record** out)
{
    if (!e || !out) return WMIXB_EINVAL;
    *out = nullptr;
    if (e->row != e->frame) { snprintf(g_err, sizeof g_err, "record: the daemon's record tick runs at 8 or 16 kHz (R:src/wmixConf.h WMIX_FREQ)"); return WMIXB_EINVAL; }
    const int interval = 20;                                         // WMIX_INTERVAL_MS, R:src/wmixConf.h:112
    if (aec_interval_ms < 0 || aec_interval_ms % interval != 0 || aec_interval_ms > 2000) {
        snprintf(g_err, sizeof g_err, "record: AEC_INTERVALMS %d must be a whole number of 20 ms packages (with a remainder the reference reads before its ring row)", aec_interval_ms);
        return WMIXB_EINVAL;
    }
    wmixb_record* r = new (std::nothrow) wmixb_record();
    if (!r) return WMIXB_ENOMEM;
    r->e = e;
    r->pkg = 2 * e->frame;
    r->delay_pkgs = aec_interval_ms / interval;
    r->n_pkg = r->delay_pkgs + 2;                                    // AEC_FIFO_PKG_NUM, R:src/wmixConf.h:141
    const size_t bytes = (size_t)r->n_pkg * e->cfg.n_streams * r->pkg * sizeof(int16_t);
    cudaError_t ce = cudaSetDevice(e->cfg.device);
    if (ce == cudaSuccess) ce = cudaMalloc(&r->fifo, bytes);
    if (ce == cudaSuccess) ce = cudaMemset(r->fifo, 0, bytes);       // the reference's ring is a zero-initialised static
    if (ce != cudaSuccess) { cudaFree(r->fifo); delete r; return fail_cuda(ce, "record_create", __LINE__); }
    *out = r;
    return WMIXB_OK;
}

extern "C" void wmixb_record_destroy(wmixb_record* r)
{
    if (!r) return;
    cudaSetDevice(r->e->cfg.device);
    cudaFree(r->fifo);
    cudaFree(r->stage);
    delete r;
}

extern "C" int wmixb_record_far_slot(const wmixb_record* r) { return r ? host::play_fifo_slot(r->count, r->n_pkg, r->delay_pkgs) : -1; }

extern "C" int wmixb_record_tick_device(wmixb_record* r, const int16_t* d_play, const int16_t* d_mic, int16_t* d_out, uint8_t* d_vad,
                                        int16_t* d_far_used, int stages, void* stream)
{
    if (!r || !d_play || !d_mic || !d_out) return WMIXB_EINVAL;
    wmixb_engine* e = r->e;
    cudaStream_t st = (cudaStream_t)stream;
    if (stages == 0) stages = e->cfg.stages;
    if (stages & ~e->cfg.stages) { snprintf(g_err, sizeof g_err, "record: stages 0x%x not all configured (0x%x)", stages, e->cfg.stages); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const size_t slot_elems = (size_t)e->cfg.n_streams * r->pkg, slot_bytes = slot_elems * sizeof(int16_t);
    // playPkgBuff_add(playBuff) — R:src/wmix.c:1419, right before the record tick under WMIX_RECORD_PLAY_SYNC (:1438)
    CK(cudaMemcpyAsync(r->fifo + (size_t)r->count * slot_elems, d_play, slot_bytes, cudaMemcpyDeviceToDevice, st));
    r->count = (r->count + 1) % r->n_pkg;
    // playPkgBuff_get(playPkgBuff, AEC_INTERVALMS) — R:src/wmix.c:653
    const int16_t* far = r->fifo + (size_t)host::play_fifo_slot(r->count, r->n_pkg, r->delay_pkgs) * slot_elems;
    if (d_far_used) CK(cudaMemcpyAsync(d_far_used, far, slot_bytes, cudaMemcpyDeviceToDevice, st));
    const int16_t* cur = d_mic;
    int rc;
    if (stages & WMIXB_NS) {                                         // ns_process(buffSrc, WMIX_FRAME_NUM): two 10 ms packets
        if ((rc = run_stages(e, cur, d_out, nullptr, 2, WMIXB_NS, st)) != WMIXB_OK) return rc;
        cur = d_out;
    }
    if (stages & WMIXB_AEC) {                                        // aec_process2(far, buffSrc, buffSrc, WMIX_FRAME_NUM, 0)
        if (e->frame == 80) {                                        // aec_init(.., 20, ..) at 8 kHz: ONE 160-sample packet
            if ((rc = launch_aec(e, far, cur, d_out, 160, 0, st)) != WMIXB_OK) return rc;
        } else {                                                     // 16 kHz: 10 ms packets, two per package
            for (int k = 0; k < 2; ++k)
                if ((rc = launch_aec(e, far + 160 * k, cur + 160 * k, d_out + 160 * k, 160, 0, st, 320)) != WMIXB_OK) return rc;
        }
        cur = d_out;
    }
    if (stages & WMIXB_AGC) {                                        // agc_process: 10 ms packets
        if ((rc = run_stages(e, cur, d_out, nullptr, 2, WMIXB_AGC, st)) != WMIXB_OK) return rc;
        cur = d_out;
    }
    if (cur != d_out) CK(cudaMemcpyAsync(d_out, cur, slot_bytes, cudaMemcpyDeviceToDevice, st));
    if (stages & WMIXB_VAD) return wmixb_vad20_device(e, d_out, d_vad, st);     // vad_init(.., WMIX_INTERVAL_MS = 20, ..)
    return WMIXB_OK;
}

extern "C" int wmixb_record_tick_host(wmixb_record* r, const int16_t* h_play, const int16_t* h_mic, int16_t* h_out, uint8_t* h_vad, int stages)
{
    if (!r || !h_play || !h_mic || !h_out) return WMIXB_EINVAL;
    wmixb_engine* e = r->e;
    CK(cudaSetDevice(e->cfg.device));
    const size_t elems = (size_t)e->cfg.n_streams * r->pkg, bytes = elems * sizeof(int16_t);
    if (!r->stage) CK(cudaMalloc(&r->stage, 2 * bytes));
    int16_t *d_play = r->stage, *d_mic = r->stage + elems;
    CK(cudaMemcpyAsync(d_play, h_play, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(d_mic, h_mic, bytes, cudaMemcpyHostToDevice, e->stream));
    const int rc = wmixb_record_tick_device(r, d_play, d_mic, d_mic, h_vad ? e->d_vad : nullptr, nullptr, stages, e->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h_out, d_mic, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (h_vad) CK(cudaMemcpyAsync(h_vad, e->d_vad, (size_t)e->cfg.n_streams, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

extern "C" int wmixb_offline_device(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int n_frames,
                                    int stages, void* stream)
{
    if (!e || !d_in || !d_out || n_frames < 1) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    return run_stages(e, d_in, d_out, d_vad, n_frames, stages, (cudaStream_t)stream);
}

// Host-buffer tick.  The streams are cut into chunks that flow through three CUDA streams, so the
// H2D copy of chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 overlap (the two copy
// engines run full duplex); with h_bus the conference bus is summed on the device-resident result
// and copied out last.
// pipelined = true: nothing is waited for; the tick's completion is the event done_ev[slot] (see wmixb_tick_host_submit).
constexpr size_t kZcBytes = 8192;            // zero-copy staging per direction: up to 25 streams at 16 kHz, 51 at 8 kHz
static int tick_host_impl(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int32_t* h_bus, int stages,
                          bool pipelined = false, int slot = 0)
{
    if (!e || !h_in || (!h_out && !h_bus && !h_vad)) return WMIXB_EINVAL;
    if (h_bus && e->n_conf < 1) { snprintf(g_err, sizeof g_err, "tick_host_bus: call wmixb_set_conferences first"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    if (e->row != e->frame) {
        // 32 kHz engine: one blocking round trip (the chunk pipeline and the conference bus are built for the 8 / 16 kHz rows)
        if (h_bus || pipelined) { snprintf(g_err, sizeof g_err, "32 kHz engine: wmixb_tick_device / wmixb_tick_host only"); return WMIXB_EINVAL; }
        const size_t bytes = (size_t)n * e->row * sizeof(int16_t);
        CK(cudaMemcpyAsync(e->d_in, h_in, bytes, cudaMemcpyHostToDevice, e->stream));
        const int rc = run_stages(e, e->d_in, e->d_out, e->d_vad, 1, stages, e->stream);
        if (rc) return rc;
        if (h_out) CK(cudaMemcpyAsync(h_out, e->d_out, bytes, cudaMemcpyDeviceToHost, e->stream));
        if (h_vad) CK(cudaMemcpyAsync(h_vad, e->d_vad, (size_t)n, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        return WMIXB_OK;
    }
    {
        const int eff = stages ? stages : e->cfg.stages;
        const size_t bytes = (size_t)n * e->frame * sizeof(int16_t);
        if (e->host_zero_copy && !pipelined && !h_bus && bytes <= kZcBytes && (eff & (WMIXB_NS | WMIXB_AGC | WMIXB_VAD)) && !(eff & WMIXB_AEC)) {
            if (!e->zc_host) {
                void* hp = nullptr;
                void* dp = nullptr;
                CK(cudaHostAlloc(&hp, 3 * kZcBytes, cudaHostAllocMapped));
                if (cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess) { cudaFreeHost(hp); (void)cudaGetLastError(); e->host_zero_copy = 0; }
                else { e->zc_host = (char*)hp; e->zc_dev = (char*)dp; }
            }
            if (e->zc_host) {
                memcpy(e->zc_host, h_in, bytes);
                const int rc = run_stages(e, (const int16_t*)e->zc_dev, (int16_t*)(e->zc_dev + kZcBytes), (uint8_t*)(e->zc_dev + 2 * kZcBytes), 1, stages, e->stream);
                if (rc) return rc;
                CK(cudaStreamSynchronize(e->stream));
                if (h_out) memcpy(h_out, e->zc_host + kZcBytes, bytes);
                if (h_vad) memcpy(h_vad, e->zc_host + 2 * kZcBytes, (size_t)n);
                return WMIXB_OK;
            }
        }
    }
    // One chunk per pipeline stream (a stream that gets two chunks serialises them and unbalances the pipeline).  Measured per
    // 100 k-stream tick, chunks = streams: blocking call 1.50 / 1.39 / 1.35 / 1.32 / 1.27 ms for 3 / 4 / 5 / 6 / 8; pipelined
    // ticks 1.07 / 0.96 / 0.96 / 0.97 / 0.99 ms.  wmixb_set_tuning("host_chunks" / "host_lanes") overrides (experiments).
    int chunks = pipelined ? e->host_chunks_pipe : e->host_chunks_sync;
    if (n < 8192 && !pipelined) chunks = 1;
    if (pipelined && chunks < 3) chunks = 3;
    int per = ((n + chunks - 1) / chunks + 127) / 128 * 128;     // whole post_kernel CTAs, line-aligned SoA rows
    // (blocking call only: measured 1.40 against 1.44 ms; with pipelined ticks the unequal chunks unbalance the three
    // streams, 1.08 against 1.06 ms)
    if (!pipelined && chunks > 1 && (stages == 0 ? e->cfg.stages : stages) & WMIXB_NS) {
        // the NS kernel is a persistent grid of ns_grid x warps warps, one stream each per round: a chunk that is a whole
        // number of rounds wastes no tail (33 408 streams = 11.3 rounds ran as 12; 35 456 = 11.98 rounds does not)
        const long long lanes = (long long)e->ns_grid * kNsCfgs[e->ns_cfg].warps;
        const long long rounds = ((n + chunks - 1) / chunks + lanes - 1) / lanes;
        const long long cand = rounds * lanes / 128 * 128;
        if (cand >= 128 && cand * (chunks - 1) < n && cand * chunks >= n) per = (int)cand;
    }
    const bool multi = chunks > 1;
    int lanes_used = chunks < kPipe ? chunks : kPipe;            // pipeline streams this tick deals its chunks to
    if (e->host_lanes >= 1 && e->host_lanes < lanes_used) lanes_used = e->host_lanes;
    if (multi && !e->pipe[0]) {
        for (int k = 0; k < kPipe; ++k) {
            CK(cudaStreamCreateWithFlags(&e->pipe[k], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&e->pipe_ev[k], cudaEventDisableTiming));
        }
    }
    int16_t* d_out = (pipelined && slot) ? e->d_out2 : e->d_out;
    int used = 0;
    for (int first = 0, c = 0; first < n; first += per, ++c) {
        const int cnt = n - first < per ? n - first : per;
        cudaStream_t st = multi ? e->pipe[c % lanes_used] : e->stream;
        const size_t off = (size_t)first * e->frame, bytes = (size_t)cnt * e->frame * sizeof(int16_t);
        CK(cudaMemcpyAsync(e->d_in + off, h_in + off, bytes, cudaMemcpyHostToDevice, st));
        const int rc = run_stages(e, e->d_in + off, d_out + off, e->d_vad + first, 1, stages, st, nullptr, 0, first, cnt);
        if (rc) return rc;
        if (multi && h_bus) CK(cudaEventRecord(e->pipe_ev[c % lanes_used], st));   // last record per stream covers its chunks
        if (h_out) CK(cudaMemcpyAsync(h_out + off, d_out + off, bytes, cudaMemcpyDeviceToHost, st));
        if (h_vad) CK(cudaMemcpyAsync(h_vad + first, e->d_vad + first, (size_t)cnt, cudaMemcpyDeviceToHost, st));
        used = c + 1 < lanes_used ? c + 1 : lanes_used;
    }
    if (h_bus) {
        if (multi)
            for (int k = 0; k < used; ++k) CK(cudaStreamWaitEvent(e->stream, e->pipe_ev[k], 0));
        const size_t bus_bytes = (size_t)e->n_conf * e->frame * sizeof(int32_t);
        if (e->d_bus_bytes < bus_bytes) {
            CK(cudaStreamSynchronize(e->stream));
            cudaFree(e->d_bus);
            e->d_bus = nullptr;
            e->d_bus_bytes = 0;
            CK(cudaMalloc(&e->d_bus, bus_bytes));
            e->d_bus_bytes = bus_bytes;
        }
        const int rc = wmixb_bus_sum_device(e, d_out, e->d_bus, e->stream);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h_bus, e->d_bus, bus_bytes, cudaMemcpyDeviceToHost, e->stream));
    }
    if (pipelined) {
        // completion of this tick = the tail of every chunk stream + the bus read-out, gathered on the engine stream
        for (int k = 0; k < used; ++k) {
            CK(cudaEventRecord(e->tail_ev[k], e->pipe[k]));
            CK(cudaStreamWaitEvent(e->stream, e->tail_ev[k], 0));
        }
        CK(cudaEventRecord(e->done_ev[slot], e->stream));
        return WMIXB_OK;
    }
    if (multi)
        for (int k = 0; k < used; ++k) CK(cudaStreamSynchronize(e->pipe[k]));
    CK(cudaStreamSynchronize(e->stream));
    return WMIXB_OK;
}

// Pipelined form of wmixb_tick_host[_bus] for a host that feeds ticks back to back: submit returns once the tick's
// copies and kernels are queued, wait returns when the OLDEST submitted tick's h_out / h_vad / h_bus are complete.
// Up to two ticks are in flight, so tick t+1's H2D copy and first kernels overlap tick t's last kernels and D2H copy
// instead of the pipeline draining at every tick boundary.  Per-stream state order is preserved: chunk c of every tick
// runs on the same CUDA stream.  The caller keeps h_in / h_out / h_vad / h_bus of a tick alive (and pinned, for
// overlap) until its wait returns.
extern "C" int wmixb_tick_host_submit(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int32_t* h_bus, int stages)
{
    if (!e) return WMIXB_EINVAL;
    if (e->row != e->frame) { snprintf(g_err, sizeof g_err, "32 kHz engine: wmixb_tick_device / wmixb_tick_host only"); return WMIXB_EINVAL; }
    if (e->submitted - e->waited >= 2) { snprintf(g_err, sizeof g_err, "tick_host_submit: two ticks already in flight — call wmixb_tick_host_wait"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    if (!e->d_out2) {
        CK(cudaMalloc(&e->d_out2, (size_t)e->cfg.n_streams * e->frame * sizeof(int16_t)));
        for (int k = 0; k < kPipe; ++k) CK(cudaEventCreateWithFlags(&e->tail_ev[k], cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) CK(cudaEventCreateWithFlags(&e->done_ev[k], cudaEventDisableTiming));
    }
    const int rc = tick_host_impl(e, h_in, h_out, h_vad, h_bus, stages, true, (int)(e->submitted & 1));
    if (rc == WMIXB_OK) e->submitted++;
    else {
        // part of the tick may be queued with no completion event recorded: drain it before the caller frees its buffers
        for (int k = 0; k < kPipe; ++k) if (e->pipe[k]) cudaStreamSynchronize(e->pipe[k]);
        cudaStreamSynchronize(e->stream);
        (void)cudaGetLastError();
    }
    return rc;
}

extern "C" int wmixb_tick_host_wait(wmixb_engine* e)
{
    if (!e) return WMIXB_EINVAL;
    if (e->submitted == e->waited) return WMIXB_OK;
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaEventSynchronize(e->done_ev[e->waited & 1]));
    e->waited++;
    return WMIXB_OK;
}

extern "C" int wmixb_tick_host(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int stages)
{
    return tick_host_impl(e, h_in, h_out, h_vad, nullptr, stages);
}

extern "C" int wmixb_tick_host_bus(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int32_t* h_bus, int stages)
{
    if (!h_bus) return WMIXB_EINVAL;
    return tick_host_impl(e, h_in, h_out, h_vad, h_bus, stages);
}

// G.711 legs in and out (include/wmixb.h): one byte per sample over the host bus, the codecs on the device around the stages
extern "C" int wmixb_tick_host_g711(wmixb_engine* e, int law, const uint8_t* h_codes_in, uint8_t* h_codes_out, uint8_t* h_vad,
                                    int32_t* h_bus, int nminus1, int stages)
{
    if (!e || !h_codes_in || (law != 0 && law != 1) || (!h_codes_out && !h_vad && !h_bus)) return WMIXB_EINVAL;
    if (e->row != e->frame) { snprintf(g_err, sizeof g_err, "tick_host_g711: 8 / 16 kHz engines"); return WMIXB_EINVAL; }
    if ((h_bus || nminus1) && e->n_conf < 1) { snprintf(g_err, sizeof g_err, "tick_host_g711: the bus and the N-minus-one read-out need wmixb_set_conferences"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const int n = e->cfg.n_streams;
    const size_t cnt = (size_t)n * e->frame;
    if (!e->d_codes) CK(cudaMalloc(&e->d_codes, 2 * cnt));
    uint8_t *c_in = e->d_codes, *c_out = e->d_codes + cnt;
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(c_in, h_codes_in, cnt, cudaMemcpyHostToDevice, st));
    int rc = wmixb_g711_decode_device(law, c_in, e->d_in, cnt, st);
    if (rc) return rc;
    rc = run_stages(e, e->d_in, e->d_out, e->d_vad, 1, stages, st);
    if (rc) return rc;
    const int16_t* leg = e->d_out;
    if (h_bus || nminus1) {
        const size_t bus_bytes = (size_t)e->n_conf * e->frame * sizeof(int32_t);
        if (e->d_bus_bytes < bus_bytes) {
            CK(cudaStreamSynchronize(st));
            cudaFree(e->d_bus);
            e->d_bus = nullptr;
            e->d_bus_bytes = 0;
            CK(cudaMalloc(&e->d_bus, bus_bytes));
            e->d_bus_bytes = bus_bytes;
        }
        rc = wmixb_bus_sum_device(e, e->d_out, e->d_bus, st);
        if (rc) return rc;
        if (h_bus) CK(cudaMemcpyAsync(h_bus, e->d_bus, bus_bytes, cudaMemcpyDeviceToHost, st));
        if (nminus1 && h_codes_out) {
            rc = wmixb_bus_nminus1_device(e, e->d_bus, e->d_out, e->d_in, st);      // d_in is free again: the read-out lands there
            if (rc) return rc;
            leg = e->d_in;
        }
    }
    if (h_codes_out) {
        rc = wmixb_g711_encode_device(law, leg, c_out, cnt, st);
        if (rc) return rc;
        CK(cudaMemcpyAsync(h_codes_out, c_out, cnt, cudaMemcpyDeviceToHost, st));
    }
    if (h_vad) CK(cudaMemcpyAsync(h_vad, e->d_vad, (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return WMIXB_OK;
}

// ---- conference bus ----
extern "C" int wmixb_set_conferences(wmixb_engine* e, const int32_t* h_conf_start, int n_conf)
{
    if (!e || !h_conf_start || n_conf < 1) return WMIXB_EINVAL;
    if (e->row != e->frame) { snprintf(g_err, sizeof g_err, "32 kHz engine: the conference bus is built for 8 / 16 kHz rows"); return WMIXB_EINVAL; }
    if (h_conf_start[0] != 0 || h_conf_start[n_conf] != e->cfg.n_streams) { snprintf(g_err, sizeof g_err, "conference ranges must cover [0, n_streams)"); return WMIXB_EINVAL; }
    std::vector<int32_t> of((size_t)e->cfg.n_streams);
    int biggest = 0;
    for (int c = 0; c < n_conf; ++c) {
        if (h_conf_start[c + 1] < h_conf_start[c]) { snprintf(g_err, sizeof g_err, "conference ranges must be non-decreasing"); return WMIXB_EINVAL; }
        for (int s = h_conf_start[c]; s < h_conf_start[c + 1]; ++s) of[s] = c;
        if (h_conf_start[c + 1] - h_conf_start[c] > biggest) biggest = h_conf_start[c + 1] - h_conf_start[c];
    }
    CK(cudaSetDevice(e->cfg.device));
    cudaFree(e->conf_start);
    e->conf_start = nullptr;
    CK(cudaMalloc(&e->conf_start, (size_t)(n_conf + 1) * sizeof(int32_t)));
    CK(cudaMemcpy(e->conf_start, h_conf_start, (size_t)(n_conf + 1) * sizeof(int32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->conf_of, of.data(), of.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    e->n_conf = n_conf;
    e->max_conf = biggest;
    return WMIXB_OK;
}

template <int LAW>
static int bus_sum_impl(wmixb_engine* e, const void* src, int32_t* d_bus, cudaStream_t st)
{
    if (!e || !src || !d_bus || e->n_conf < 1) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    // enough CTAs to fill the chip: split big conferences into member slices
    int chunks = 1;
    const int want = 4 * e->sm_count;
    if (e->n_conf < want) chunks = (want + e->n_conf - 1) / e->n_conf;
    int chunk = (e->max_conf + chunks - 1) / chunks;
    if (chunk < 8) chunk = 8;
    chunks = (e->max_conf + chunk - 1) / chunk;
    if (chunks < 1) chunks = 1;
    if (chunks > 1) CK(cudaMemsetAsync(d_bus, 0, (size_t)e->n_conf * e->frame * sizeof(int32_t), st));
    if (LAW < 0 && (e->frame & 7) == 0 && ((uintptr_t)src & 15) == 0 && ((uintptr_t)d_bus & 15) == 0) {
        const long long threads = (long long)e->n_conf * (e->frame >> 3);
        long long blocks = (threads + 255) / 256;
        const long long cap = 8LL * e->sm_count;
        if (blocks > cap) blocks = cap;
        bus_sum_vec_kernel<<<dim3((unsigned)blocks, (unsigned)chunks), 256, 0, st>>>(static_cast<const int16_t*>(src), d_bus, e->conf_start, e->n_conf,
                                                                                 e->frame, chunk, chunks);
        CK_LAUNCH();
        return WMIXB_OK;
    }
    dim3 grid((unsigned)e->n_conf, (unsigned)chunks);
    bus_sum_kernel<LAW><<<grid, e->frame <= 96 ? 96 : 160, 0, st>>>(src, d_bus, e->conf_start, e->frame, chunk);
    CK_LAUNCH();
    return WMIXB_OK;
}

// int16 PCM legs, frame a multiple of 8: eight samples per thread (one 16-byte load of the leg, two of its bus row, one 16-byte
// store) instead of one sample with an index division each
__global__ void __launch_bounds__(256)
nminus1_vec_kernel(const int32_t* __restrict__ bus, const int16_t* __restrict__ own, int16_t* __restrict__ out, const int32_t* __restrict__ conf_of,
                   int frame, size_t total_vec)
{
    const int vpr = frame >> 3;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t s = idx / vpr;
        const int v = (int)(idx - s * vpr);
        const int4* b4 = reinterpret_cast<const int4*>(bus + (size_t)conf_of[s] * frame + 8 * v);
        const int4 b0 = b4[0], b1 = b4[1];
        const uint4 w = reinterpret_cast<const uint4*>(own)[idx];
        const int32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint32_t u[4] = {w.x, w.y, w.z, w.w};
        uint32_t r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int32_t lo = sat16(b[2 * k] - (int32_t)(int16_t)(u[k] & 0xFFFFu)), hi = sat16(b[2 * k + 1] - ((int32_t)u[k] >> 16));
            r[k] = ((uint32_t)hi << 16) | ((uint32_t)lo & 0xFFFFu);
        }
        reinterpret_cast<uint4*>(out)[idx] = make_uint4(r[0], r[1], r[2], r[3]);
    }
}
template <int LAW>
static int nminus1_impl(wmixb_engine* e, const int32_t* d_bus, const void* own, void* out, cudaStream_t st)
{
    if (!e || !d_bus || !own || !out || e->n_conf < 1) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const size_t total = (size_t)e->cfg.n_streams * e->frame;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)e->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (LAW < 0 && (e->frame & 7) == 0 && (((uintptr_t)d_bus | (uintptr_t)own | (uintptr_t)out) & 15) == 0) {
        const size_t total_vec = total >> 3;
        size_t vb = (total_vec + 255) / 256;
        if (vb > cap) vb = cap;
        nminus1_vec_kernel<<<(unsigned)vb, 256, 0, st>>>(d_bus, static_cast<const int16_t*>(own), static_cast<int16_t*>(out), e->conf_of, e->frame, total_vec);
        CK_LAUNCH();
        return WMIXB_OK;
    }
    nminus1_kernel<LAW><<<(unsigned)blocks, 256, 0, st>>>(d_bus, own, out, e->conf_of, e->frame, total);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_bus_sum_device(wmixb_engine* e, const int16_t* d_pcm, int32_t* d_bus, void* stream)
{
    return bus_sum_impl<-1>(e, d_pcm, d_bus, (cudaStream_t)stream);
}
extern "C" int wmixb_bus_nminus1_device(wmixb_engine* e, const int32_t* d_bus, const int16_t* d_pcm, int16_t* d_out, void* stream)
{
    return nminus1_impl<-1>(e, d_bus, d_pcm, d_out, (cudaStream_t)stream);
}
extern "C" int wmixb_g711_bus_sum_device(wmixb_engine* e, int law, const uint8_t* d_codes, int32_t* d_bus, void* stream)
{
    if (law == 0) return bus_sum_impl<0>(e, d_codes, d_bus, (cudaStream_t)stream);
    if (law == 1) return bus_sum_impl<1>(e, d_codes, d_bus, (cudaStream_t)stream);
    return WMIXB_EINVAL;
}
extern "C" int wmixb_g711_nminus1_device(wmixb_engine* e, int law, const int32_t* d_bus, const uint8_t* d_codes,
                                         uint8_t* d_out_codes, void* stream)
{
    if (law == 0) return nminus1_impl<0>(e, d_bus, d_codes, d_out_codes, (cudaStream_t)stream);
    if (law == 1) return nminus1_impl<1>(e, d_bus, d_codes, d_out_codes, (cudaStream_t)stream);
    return WMIXB_EINVAL;
}

// ---- cross-GPU conference bus over peer memory (peer_bus.cuh) ----
struct wmixb_peer_bus {
    wmixb_engine* e = nullptr;
    int rank = 0, world = 1, n_conf = 0, frame = 0, grid = 0, threads = 0, slices = 0, tile = 16;
    size_t smem = 0;
    void* mailbox = nullptr;            // [slots | flags], cudaMalloc'ed on e's device
    size_t slot_bytes = 0, flag_bytes = 0, result_bytes = 0, rflag_bytes = 0;
    bool rs = false;                    // reduce-scatter / all-gather exchange (peer::reduce_scatter_for)
    void* mapped[peer::kMaxWorld] = {}; // IPC mappings to close
    peer::Ring ring{};
    bool connected = false;
    uint32_t seq = 0;
    int* d_error = nullptr;
    unsigned long long timeout_ns = 2000000000ull;
    long long timeout_cycles = 4000000000ll;                 // timeout_ns on the SM cycle counter (set at create)
};

struct PeerMeta { int32_t magic, world, n_conf, frame; };   // rides behind the 64-byte IPC handle
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
static_assert(sizeof(cudaIpcMemHandle_t) + sizeof(PeerMeta) == WMIXB_PEER_HANDLE_BYTES, "handle blob size");

static void peer_set_ring(wmixb_peer_bus* pb, int r, void* base)
{
    pb->ring.slots[r] = (int32_t*)base;
    pb->ring.flags[r] = (uint32_t*)((char*)base + pb->slot_bytes);
    pb->ring.result[r] = (int32_t*)((char*)base + pb->slot_bytes + pb->flag_bytes);
    pb->ring.rflags[r] = (uint32_t*)((char*)base + pb->slot_bytes + pb->flag_bytes + pb->result_bytes);
}

extern "C" int wmixb_peer_bus_create(wmixb_engine* e, int rank, int world, wmixb_peer_bus** out)
{
    return wmixb_peer_bus_create_ex(e, rank, world, nullptr, out);
}

extern "C" int wmixb_peer_bus_create_ex(wmixb_engine* e, int rank, int world, const wmixb_peer_opts* opts, wmixb_peer_bus** out)
{
    wmixb_peer_opts o;
    memset(&o, 0, sizeof o);
    o.reduce_scatter = -1;
    if (opts) o = *opts;
    if (o.tile != 0 && o.tile != 16 && !(e && o.tile == e->frame)) { snprintf(g_err, sizeof g_err, "peer_bus: tile must be 0, 16 or the frame length"); return WMIXB_EINVAL; }
    if (o.ranks_per_device < 0 || o.ranks_per_device > 8 || o.timeout_ms < 0) return WMIXB_EINVAL;
    if (!e || !out || world < 1 || world > peer::kMaxWorld || rank < 0 || rank >= world) return WMIXB_EINVAL;
    *out = nullptr;
    if (e->n_conf < 1) { snprintf(g_err, sizeof g_err, "peer_bus: call wmixb_set_conferences first"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    wmixb_peer_bus* pb = new (std::nothrow) wmixb_peer_bus();
    if (!pb) return WMIXB_ENOMEM;
    pb->e = e; pb->rank = rank; pb->world = world; pb->n_conf = e->n_conf; pb->frame = e->frame;
    pb->slot_bytes = (size_t)2 * world * e->n_conf * e->frame * sizeof(int32_t);
    pb->flag_bytes = (size_t)2 * world * e->n_conf * (e->frame / peer::kTile) * sizeof(uint32_t);
    pb->result_bytes = (size_t)2 * e->n_conf * e->frame * sizeof(int32_t);
    pb->rflag_bytes = ((size_t)2 * e->n_conf * sizeof(uint32_t) + 15) / 16 * 16;
    // tile = 16 samples or the whole bus row (the same on every rank: a function of n_conf only, peer::tile_for)
    pb->tile = peer::tile_for(e->n_conf, e->frame);
    if (o.tile) pb->tile = o.tile;                                                       // experiments / tests: the same on ALL ranks
    pb->rs = peer::reduce_scatter_for(e->n_conf, e->frame, world) && pb->tile == e->frame;
    if (o.reduce_scatter >= 0) pb->rs = o.reduce_scatter != 0 && pb->tile == e->frame;    // experiments: the same on ALL ranks
    // member slices per tile: a quarter of the largest local conference, power of two, 2..32 (1.. for row tiles),
    // as many as fit one CTA
    int slices = pb->tile == peer::kTile ? 2 : 1;
    while (slices < 32 && slices * 4 < e->max_conf && slices * 2 * pb->tile <= peer::kThreads) slices *= 2;
    pb->slices = slices;
    {
        const int gthreads = slices * pb->tile;
        int groups = (gthreads >= 256 ? gthreads : 256) / gthreads;
        if (pb->tile != peer::kTile) groups = peer::kThreads / gthreads;          // row tiles: fill the CTA
        if (groups < 1) groups = 1;
        pb->threads = groups * gthreads;
        if (pb->threads < 256 && pb->tile == peer::kTile) pb->threads = 256;
    }
    const int groups = pb->threads / (slices * pb->tile);
    pb->smem = (size_t)groups * (slices + 1) * pb->tile * sizeof(int32_t);
    if (o.timeout_ms > 0) pb->timeout_ns = (unsigned long long)o.timeout_ms * 1000000ull;
    {
        int khz = 0;
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->cfg.device) != cudaSuccess || khz < 1) khz = 2000000;
        pb->timeout_cycles = (long long)((double)pb->timeout_ns * (double)khz / 1e6);
    }
    const size_t mailbox_bytes = pb->slot_bytes + pb->flag_bytes + pb->result_bytes + pb->rflag_bytes;
    cudaError_t ce = cudaMalloc(&pb->mailbox, mailbox_bytes);
    if (ce == cudaSuccess) ce = cudaMemset(pb->mailbox, 0, mailbox_bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&pb->d_error, sizeof(int));
    if (ce == cudaSuccess) ce = cudaMemset(pb->d_error, 0, sizeof(int));
    int per_sm = 0;
    if (ce == cudaSuccess) {
        if (pb->rs && pb->tile == 80) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_bus_rs_kernel<0, 80>, pb->threads, pb->smem);
        else if (pb->rs) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_bus_rs_kernel<0, 160>, pb->threads, pb->smem);
        else if (pb->tile == 16) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_bus_kernel<0, 16>, pb->threads, pb->smem);
        else if (pb->tile == 80) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_bus_kernel<0, 80>, pb->threads, pb->smem);
        else ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer::peer_bus_kernel<0, 160>, pb->threads, pb->smem);
    }
    if (ce != cudaSuccess) { cudaFree(pb->mailbox); cudaFree(pb->d_error); delete pb; return fail_cuda(ce, "peer_bus_create", __LINE__); }
    // every CTA must be resident (phase 2 waits on other ranks): never more than one wave, shared between the ranks whose
    // peer kernels may run on this GPU at the same time (opts->ranks_per_device; default 2, so that a second rank living
    // on the same device — the one-process tests — still fits beside this one).  If the share is less than one CTA per
    // SM the guarantee cannot be given: refuse instead of risking a spin until the timeout.
    const int share = o.ranks_per_device > 0 ? o.ranks_per_device : 2;
    if (per_sm < 1 || e->sm_count * per_sm / share < 1) {
        cudaFree(pb->mailbox); cudaFree(pb->d_error); delete pb;
        snprintf(g_err, sizeof g_err, "peer_bus: %d ranks per device cannot all be resident (%d CTAs per SM)", share, per_sm);
        return WMIXB_EINVAL;
    }
    int cap = e->sm_count * per_sm / share;
    if (cap < 1) cap = 1;
    const int n_steps = (e->n_conf * (e->frame / pb->tile) + groups - 1) / groups;
    pb->grid = n_steps < cap ? n_steps : cap;
    peer_set_ring(pb, rank, pb->mailbox);
    pb->connected = world == 1;
    *out = pb;
    return WMIXB_OK;
}

extern "C" void wmixb_peer_bus_destroy(wmixb_peer_bus* pb)
{
    if (!pb) return;
    cudaSetDevice(pb->e->cfg.device);
    cudaDeviceSynchronize();
    for (int r = 0; r < pb->world; ++r)
        if (pb->mapped[r]) cudaIpcCloseMemHandle(pb->mapped[r]);
    cudaFree(pb->mailbox);
    cudaFree(pb->d_error);
    delete pb;
}

extern "C" int wmixb_peer_bus_handle(const wmixb_peer_bus* pb, void* handle_out)
{
    if (!pb || !handle_out) return WMIXB_EINVAL;
    CK(cudaSetDevice(pb->e->cfg.device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, pb->mailbox));
    const PeerMeta m{0x77425042, pb->world, pb->n_conf, pb->frame};
    memcpy(handle_out, &h, sizeof h);
    memcpy((char*)handle_out + sizeof h, &m, sizeof m);
    return WMIXB_OK;
}

extern "C" int wmixb_peer_bus_connect(wmixb_peer_bus* pb, const void* handles)
{
    if (!pb || !handles) return WMIXB_EINVAL;
    CK(cudaSetDevice(pb->e->cfg.device));
    for (int r = 0; r < pb->world; ++r) {
        const char* blob = (const char*)handles + (size_t)r * WMIXB_PEER_HANDLE_BYTES;
        PeerMeta m;
        memcpy(&m, blob + sizeof(cudaIpcMemHandle_t), sizeof m);
        if (m.magic != 0x77425042 || m.world != pb->world || m.n_conf != pb->n_conf || m.frame != pb->frame) {
            snprintf(g_err, sizeof g_err, "peer_bus_connect: rank %d announces world/n_conf/frame %d/%d/%d, this rank has %d/%d/%d",
                     r, m.world, m.n_conf, m.frame, pb->world, pb->n_conf, pb->frame);
            return WMIXB_EINVAL;
        }
        if (r == pb->rank || pb->mapped[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, blob, sizeof h);
        void* base = nullptr;
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        pb->mapped[r] = base;
        peer_set_ring(pb, r, base);
    }
    pb->connected = true;
    return WMIXB_OK;
}

extern "C" int wmixb_peer_bus_connect_local(wmixb_peer_bus* pb, wmixb_peer_bus* const* peers)
{
    if (!pb || !peers) return WMIXB_EINVAL;
    CK(cudaSetDevice(pb->e->cfg.device));
    for (int r = 0; r < pb->world; ++r) {
        const wmixb_peer_bus* q = peers[r];
        if (!q || q->rank != r || q->world != pb->world || q->n_conf != pb->n_conf || q->frame != pb->frame) {
            snprintf(g_err, sizeof g_err, "peer_bus_connect_local: peers[%d] does not match (rank/world/n_conf/frame)", r);
            return WMIXB_EINVAL;
        }
        if (r == pb->rank) continue;
        if (q->e->cfg.device != pb->e->cfg.device) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, pb->e->cfg.device, q->e->cfg.device));
            if (!can) { snprintf(g_err, sizeof g_err, "peer_bus: device %d cannot access device %d", pb->e->cfg.device, q->e->cfg.device); return WMIXB_ECUDA; }
            const cudaError_t ce = cudaDeviceEnablePeerAccess(q->e->cfg.device, 0);
            if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) return fail_cuda(ce, "cudaDeviceEnablePeerAccess", __LINE__);
            (void)cudaGetLastError();
        }
        peer_set_ring(pb, r, q->mailbox);
    }
    pb->connected = true;
    return WMIXB_OK;
}

extern "C" int wmixb_peer_bus_tick_device(wmixb_peer_bus* pb, int law, const void* d_in, void* d_out, int32_t* d_bus, void* stream)
{
    if (!pb || !d_in || law < -1 || law > 1) return WMIXB_EINVAL;
    if (!pb->connected) { snprintf(g_err, sizeof g_err, "peer_bus: not connected to its %d peers yet", pb->world - 1); return WMIXB_EINVAL; }
    wmixb_engine* e = pb->e;
    if (e->n_conf != pb->n_conf) { snprintf(g_err, sizeof g_err, "peer_bus: the conference table changed after create"); return WMIXB_EINVAL; }
    CK(cudaSetDevice(e->cfg.device));
    const uint32_t seq_before = pb->seq;
    if (++pb->seq == 0) pb->seq = 2;     // 0 is the "never written" flag value; keep the parity sequence alternating
    cudaStream_t st = (cudaStream_t)stream;
#define WMX_PEER(LAW, TILE) peer::peer_bus_kernel<LAW, TILE><<<pb->grid, pb->threads, pb->smem, st>>>(pb->ring, pb->rank, pb->world, pb->seq, d_in, d_out, d_bus, e->conf_start, pb->n_conf, e->frame, pb->slices, pb->timeout_cycles, pb->d_error)
#define WMX_PEER_T(TILE) do { if (law < 0) WMX_PEER(-1, TILE); else if (law == 0) WMX_PEER(0, TILE); else WMX_PEER(1, TILE); } while (0)
#define WMX_PEER_RS(LAW, TILE) peer::peer_bus_rs_kernel<LAW, TILE><<<pb->grid, pb->threads, pb->smem, st>>>(pb->ring, pb->rank, pb->world, pb->seq, d_in, d_out, d_bus, e->conf_start, pb->n_conf, pb->slices, pb->timeout_cycles, pb->d_error)
#define WMX_PEER_RS_T(TILE) do { if (law < 0) WMX_PEER_RS(-1, TILE); else if (law == 0) WMX_PEER_RS(0, TILE); else WMX_PEER_RS(1, TILE); } while (0)
    if (pb->rs) { if (pb->tile == 80) WMX_PEER_RS_T(80); else WMX_PEER_RS_T(160); }
    else if (pb->tile == 16) WMX_PEER_T(16); else if (pb->tile == 80) WMX_PEER_T(80); else WMX_PEER_T(160);
#undef WMX_PEER_RS_T
#undef WMX_PEER_RS
#undef WMX_PEER_T
#undef WMX_PEER
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) {
        pb->seq = seq_before;            // nothing ran: this rank's tick parity must stay in step with its peers
        return fail_cuda(le, "peer_bus kernel launch", __LINE__);
    }
    return WMIXB_OK;
}

// ---- the conference-bus exchange over NCCL, reachable from C (include/wmixb.h).  libnccl is opened at run time. ----
namespace {
struct NcclId { char internal[WMIXB_NCCL_ID_BYTES]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
constexpr int kNcclInt32 = 2, kNcclSum = 0;              // ncclDataType_t / ncclRedOp_t values (nccl.h, stable across 2.x)

int nccl_open(const char* path)
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return WMIXB_OK;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib || n == path) break;
    }
    if (!lib) { snprintf(g_err, sizeof g_err, "NCCL is not available: %s", dlerror()); return WMIXB_ENODEV; }
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (int (*)(NcclId*))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    a.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(lib, "ncclAllReduce");
    a.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllReduce || !a.CommDestroy) {
        snprintf(g_err, sizeof g_err, "NCCL library lacks a required symbol");
        dlclose(lib);
        return WMIXB_ENODEV;
    }
    g_nccl = a;
    return WMIXB_OK;
}
int nccl_fail(int rc, const char* what)
{
    snprintf(g_err, sizeof g_err, "%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    return WMIXB_ECUDA;
}
}  // namespace

struct wmixb_nccl_bus {
    wmixb_engine* e = nullptr;
    void* comm = nullptr;
    int rank = 0, world = 1;
};

extern "C" int wmixb_nccl_load(const char* libnccl_path) { return nccl_open(libnccl_path); }

extern "C" int wmixb_nccl_unique_id(void* id_out)
{
    if (!id_out) return WMIXB_EINVAL;
    const int rc = nccl_open(nullptr);
    if (rc) return rc;
    NcclId id;
    const int n = g_nccl.GetUniqueId(&id);
    if (n) return nccl_fail(n, "ncclGetUniqueId");
    memcpy(id_out, &id, sizeof id);
    return WMIXB_OK;
}

extern "C" int wmixb_nccl_bus_create(wmixb_engine* e, int rank, int world, const void* id, wmixb_nccl_bus** out)
{
    if (!e || !id || !out || world < 1 || rank < 0 || rank >= world) return WMIXB_EINVAL;
    *out = nullptr;
    if (e->n_conf < 1) { snprintf(g_err, sizeof g_err, "nccl_bus: call wmixb_set_conferences first"); return WMIXB_EINVAL; }
    void* comm = nullptr;
    if (world > 1) {                         // a single rank exchanges nothing: no communicator, NCCL need not even be installed
        const int rc = nccl_open(nullptr);
        if (rc) return rc;
        CK(cudaSetDevice(e->cfg.device));
        NcclId nid;
        memcpy(&nid, id, sizeof nid);
        const int n = g_nccl.CommInitRank(&comm, world, nid, rank);
        if (n) return nccl_fail(n, "ncclCommInitRank");
    }
    wmixb_nccl_bus* nb = new (std::nothrow) wmixb_nccl_bus();
    if (!nb) { g_nccl.CommDestroy(comm); return WMIXB_ENOMEM; }
    nb->e = e; nb->comm = comm; nb->rank = rank; nb->world = world;
    *out = nb;
    return WMIXB_OK;
}

extern "C" void wmixb_nccl_bus_destroy(wmixb_nccl_bus* nb)
{
    if (!nb) return;
    if (nb->comm && g_nccl.CommDestroy) {
        cudaSetDevice(nb->e->cfg.device);
        g_nccl.CommDestroy(nb->comm);
    }
    delete nb;
}

extern "C" int wmixb_nccl_bus_tick_device(wmixb_nccl_bus* nb, int law, const void* d_in, void* d_out, int32_t* d_bus, void* stream)
{
    if (!nb || !d_in || !d_bus || law < -1 || law > 1) return WMIXB_EINVAL;
    wmixb_engine* e = nb->e;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = law < 0 ? bus_sum_impl<-1>(e, d_in, d_bus, st) : (law == 0 ? bus_sum_impl<0>(e, d_in, d_bus, st) : bus_sum_impl<1>(e, d_in, d_bus, st));
    if (rc) return rc;
    if (nb->world > 1) {
        const int n = g_nccl.AllReduce(d_bus, d_bus, (size_t)e->n_conf * e->frame, kNcclInt32, kNcclSum, nb->comm, st);
        if (n) return nccl_fail(n, "ncclAllReduce");
    }
    if (d_out) rc = law < 0 ? nminus1_impl<-1>(e, d_bus, d_in, d_out, st) : (law == 0 ? nminus1_impl<0>(e, d_bus, d_in, d_out, st) : nminus1_impl<1>(e, d_bus, d_in, d_out, st));
    return rc;
}

extern "C" int wmixb_peer_bus_status(wmixb_peer_bus* pb, int* h_error)
{
    if (!pb || !h_error) return WMIXB_EINVAL;
    CK(cudaSetDevice(pb->e->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_error, pb->d_error, sizeof(int), cudaMemcpyDeviceToHost));
    return WMIXB_OK;
}

// ---- G.711 / mix ring ----
static unsigned ew_blocks(size_t work_items)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    size_t b = (work_items + 255) / 256;
    const size_t cap = (size_t)sms * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

extern "C" int wmixb_g711_encode_device(int law, const int16_t* d_pcm, uint8_t* d_codes, size_t n, void* stream)
{
    if ((law != 0 && law != 1) || (n && (!d_pcm || !d_codes))) return WMIXB_EINVAL;
    if (n == 0) return WMIXB_OK;
    if (((uintptr_t)d_pcm & 15) || ((uintptr_t)d_codes & 7)) { snprintf(g_err, sizeof g_err, "g711: buffers must be 16-byte (pcm) / 8-byte (codes) aligned"); return WMIXB_EINVAL; }
    if (law == 0) g711_encode_kernel<0><<<ew_blocks(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(d_pcm, d_codes, n);
    else g711_encode_kernel<1><<<ew_blocks(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(d_pcm, d_codes, n);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_g711_decode_device(int law, const uint8_t* d_codes, int16_t* d_pcm, size_t n, void* stream)
{
    if ((law != 0 && law != 1) || (n && (!d_pcm || !d_codes))) return WMIXB_EINVAL;
    if (n == 0) return WMIXB_OK;
    if (((uintptr_t)d_pcm & 15) || ((uintptr_t)d_codes & 7)) { snprintf(g_err, sizeof g_err, "g711: buffers must be 16-byte (pcm) / 8-byte (codes) aligned"); return WMIXB_EINVAL; }
    if (law == 0) g711_decode_kernel<0><<<ew_blocks(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(d_codes, d_pcm, n);
    else g711_decode_kernel<1><<<ew_blocks(n / 8 + 1), 256, 0, (cudaStream_t)stream>>>(d_codes, d_pcm, n);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_mix_load_device(int16_t* d_ring, uint32_t ring_len, uint32_t pos, const int16_t* d_src, uint32_t n,
                                     int rdce, uint32_t* new_pos, void* stream)
{
    if (!d_ring || ring_len == 0 || pos >= ring_len || (n && !d_src) || rdce < 1 || n > ring_len) return WMIXB_EINVAL;
    if (n) {
        mix_load_kernel<<<ew_blocks(n), 256, 0, (cudaStream_t)stream>>>(d_ring, ring_len, pos, d_src, n, rdce);
        CK_LAUNCH();
    }
    if (new_pos) *new_pos = (uint32_t)(((uint64_t)pos + n) % ring_len);
    return WMIXB_OK;
}

// ---- resampling branches of wmix_load_data on a device ring (R:src/wmix.c:1704-1939, mono bus) ----
// Thread i owns bus sample i of the chunk: it evaluates, for every source IN ORDER, the copied or ramped sample
// the host plan names and chains the reference's saturating add, so n_src producers that the reference would
// mix one after another land in one launch with the same bits.
__global__ void mix_plan_kernel(int16_t* ring, uint32_t ring_len, uint32_t pos, const int16_t* __restrict__ src, uint32_t src_stride,
                                int src_chn, const int32_t* __restrict__ map, const uint16_t* __restrict__ ramp, uint32_t n_out,
                                int n_src, const uint8_t* __restrict__ rdce)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(((uint64_t)pos + i) % ring_len);
        const int32_t m = map[i];
        const uint32_t code = ramp[i];
        const int n = (int)(code >> 8), k = (int)(code & 0xffu);
        int16_t bus = ring[p];
        for (int s = 0; s < n_src; ++s) {
            const int16_t* row = src + (size_t)s * src_stride;
            int16_t v = row[m];
            if (code) {
                // repairStep = (float)(next - last) / n; the k-th ramp value adds k accumulated steps (R:src/wmix.c:1911-1920)
                const float step = __fdiv_rn((float)((int)row[m + src_chn] - (int)v), (float)n);
                float run = step;
                for (int j = 1; j < k; ++j) run = __fadd_rn(run, step);
                v = (int16_t)(int)__fadd_rn((float)v, run);
            }
            int d = rdce ? rdce[s] : 1;
            if (d == 0) d = 1;
            bus = wmx::mix_step(bus, v, d);
        }
        ring[p] = bus;
    }
}

struct wmixb_mixplan {
    int device = 0, src_chn = 1;
    uint32_t src_bytes = 0, out_samples = 0;
    int32_t* d_map = nullptr;
    uint16_t* d_ramp = nullptr;
    std::vector<int32_t> h_map;
    std::vector<uint16_t> h_ramp;
};

extern "C" int wmixb_mixplan_create(int src_chn, int src_freq, uint32_t src_bytes, int mix_freq, int device, wmixb_mixplan** out)
{
    if (!out) return WMIXB_EINVAL;
    *out = nullptr;
    if (src_chn < 1 || src_chn > 2 || src_freq < 1 || src_freq > 65535 || mix_freq < 1 || mix_freq > 65535 ||
        src_bytes % (2u * (uint32_t)src_chn) != 0 || (src_freq == mix_freq && src_chn == 1)) {
        snprintf(g_err, sizeof g_err, "mixplan: 1 or 2 channels of whole 16-bit frames, rates 1..65535 Hz, and a format that differs from the bus");
        return WMIXB_EINVAL;
    }
    const uint32_t n = host::mix_plan(src_chn, src_freq, src_bytes, mix_freq, nullptr, nullptr);
    if (n == 0xFFFFFFFFu) {
        snprintf(g_err, sizeof g_err, "mixplan: %d Hz into %d Hz needs a ramp longer than the reference's 64-entry buffer", src_freq, mix_freq);
        return WMIXB_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(g_err, sizeof g_err, "no CUDA device — wmix_b200 has no CPU path"); return WMIXB_ENODEV; }
    if (device < 0 || device >= ndev) return WMIXB_EINVAL;
    wmixb_mixplan* m = new (std::nothrow) wmixb_mixplan();
    if (!m) return WMIXB_ENOMEM;
    m->device = device;
    m->src_chn = src_chn;
    m->src_bytes = src_bytes;
    m->out_samples = n;
    m->h_map.resize(n ? n : 1);
    m->h_ramp.resize(n ? n : 1);
    host::mix_plan(src_chn, src_freq, src_bytes, mix_freq, m->h_map.data(), m->h_ramp.data());
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaMalloc(&m->d_map, m->h_map.size() * sizeof(int32_t));
    if (ce == cudaSuccess) ce = cudaMalloc(&m->d_ramp, m->h_ramp.size() * sizeof(uint16_t));
    if (ce == cudaSuccess) ce = cudaMemcpy(m->d_map, m->h_map.data(), m->h_map.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(m->d_ramp, m->h_ramp.data(), m->h_ramp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(m->d_map); cudaFree(m->d_ramp); delete m; return fail_cuda(ce, "mixplan_create", __LINE__); }
    *out = m;
    return WMIXB_OK;
}

extern "C" void wmixb_mixplan_destroy(wmixb_mixplan* m)
{
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_map);
    cudaFree(m->d_ramp);
    delete m;
}

extern "C" uint32_t wmixb_mixplan_out_samples(const wmixb_mixplan* m) { return m ? m->out_samples : 0u; }

extern "C" int wmixb_mixplan_tables(const wmixb_mixplan* m, int32_t* h_map, uint16_t* h_ramp)
{
    if (!m) return WMIXB_EINVAL;
    if (h_map) memcpy(h_map, m->h_map.data(), (size_t)m->out_samples * sizeof(int32_t));
    if (h_ramp) memcpy(h_ramp, m->h_ramp.data(), (size_t)m->out_samples * sizeof(uint16_t));
    return WMIXB_OK;
}

extern "C" int wmixb_mix_load_plan_device(const wmixb_mixplan* m, int16_t* d_ring, uint32_t ring_len, uint32_t pos, const int16_t* d_src,
                                          int n_src, const uint8_t* d_rdce, uint32_t* new_pos, void* stream)
{
    if (!m || !d_ring || ring_len == 0 || pos >= ring_len || n_src < 0 || (n_src && m->out_samples && !d_src) || m->out_samples > ring_len)
        return WMIXB_EINVAL;
    if (n_src && m->out_samples) {
        CK(cudaSetDevice(m->device));
        mix_plan_kernel<<<ew_blocks(m->out_samples), 256, 0, (cudaStream_t)stream>>>(d_ring, ring_len, pos, d_src, m->src_bytes / 2, m->src_chn,
                                                                                     m->d_map, m->d_ramp, m->out_samples, n_src, d_rdce);
        CK_LAUNCH();
    }
    if (new_pos) *new_pos = (uint32_t)(((uint64_t)pos + m->out_samples) % ring_len);
    return WMIXB_OK;
}

// ---- wmix_load_data itself, on the daemon's HOST ring (R:src/wmix.h:40-49, R:src/wmix.c:1639-1956) ----
// The bookkeeping (head restart, background-reduce choice, tick) is the reference's, on the host; the adds run on the
// GPU: the span of the ring the call touches goes up, the same kernels as wmixb_mix_load*_device add into it, and it
// comes back.  A drop-in for the daemon's call sites, priced like the handle API: one round trip per call.
extern "C" uint8_t* wmixb_load_data_host(const wmixb_mix_view* w, const uint8_t* src, uint32_t src_bytes, uint16_t freq,
                                         uint8_t channels, uint8_t sample, uint8_t* head, uint8_t reduce, uint32_t* tick)
{
    if (!w || !w->run || !src || src_bytes < 1 || !tick || !w->ring_start || w->ring_bytes < 2) return head;   // R:src/wmix.c:1664
    const uint32_t ring_len = w->ring_bytes / 2;
    if (!head || *tick < w->tick) {                                                                          // :1667-1674
        head = w->ring_start + w->head_off + w->play_correct;
        *tick = w->tick + w->play_correct;
        if (head >= w->ring_start + w->ring_bytes) head = w->ring_start;
    }
    const int rdce = (reduce == w->reduce_mode) ? 1 : (w->reduce_mode ? w->reduce_mode : 1);                 // :1676-1677
    const uint32_t pos = (uint32_t)(head - w->ring_start) / 2;
    const bool same = freq == w->mix_freq && channels == 1 && sample == 16;
    if (!same && !(sample == 16 && (channels == 1 || channels == 2))) return head;                           // the empty 8 / 32-bit cases
    // plans are immutable once built and keyed by device; the lock covers the lookup only, the round trip runs on the calling
    // thread's own staging and stream (scratch.h)
    static std::mutex mu;
    static std::map<std::tuple<int, int, uint32_t, int, int>, wmixb_mixplan*> plans;
    // The reference's loop `for (count = 0; count < srcU8Len; count += 2)` (R:src/wmix.c:1681) consumes ceil(n / 2) samples
    // of an odd-length source, reading one byte past it; here the last, half-present sample is completed with a zero byte.
    uint32_t n_out = (src_bytes + 1) / 2;
    wmixb_mixplan* plan = nullptr;
    if (!same) {
        std::lock_guard<std::mutex> lock(mu);
        wmixb_mixplan*& slot = plans[std::make_tuple((int)channels, (int)freq, src_bytes, (int)w->mix_freq, w->device)];
        if (!slot && wmixb_mixplan_create(channels, freq, src_bytes, w->mix_freq, w->device, &slot) != WMIXB_OK) { slot = nullptr; return head; }
        plan = slot;
        n_out = wmixb_mixplan_out_samples(plan);
    }
    if (n_out == 0) return head;
    if (n_out > ring_len) { snprintf(g_err, sizeof g_err, "load_data: the chunk (%u samples) is longer than the ring", n_out); return head; }
    host::Scratch* sc = host::scratch(w->device);
    if (!sc) return head;
    const size_t src_elems = (src_bytes + 1) / 2 + 2;
    int16_t* d_span = (int16_t*)sc->need(0, (size_t)n_out * 2);
    int16_t* d_src = (int16_t*)sc->need(1, src_elems * 2);
    uint8_t* d_rd = (uint8_t*)sc->need(2, 16);
    if (!d_span || !d_src || !d_rd) return head;
    cudaStream_t st = sc->st;
    // the touched span [pos, pos + n_out) of the ring, unwrapped, becomes a private device ring of exactly n_out samples
    int16_t* ring16 = reinterpret_cast<int16_t*>(w->ring_start);
    const uint32_t first = n_out < ring_len - pos ? n_out : ring_len - pos;
    bool ok = cudaMemcpyAsync(d_span, ring16 + pos, (size_t)first * 2, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok && first < n_out) ok = cudaMemcpyAsync(d_span + first, ring16, (size_t)(n_out - first) * 2, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok && (src_bytes & 1)) ok = cudaMemsetAsync(d_src + src_bytes / 2, 0, 2, st) == cudaSuccess;
    if (ok) ok = cudaMemcpyAsync(d_src, src, src_bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
    if (ok) {
        if (same) ok = wmixb_mix_load_device(d_span, n_out, 0, d_src, n_out, rdce, nullptr, st) == WMIXB_OK;
        else {
            const uint8_t rd8 = (uint8_t)rdce;
            ok = cudaMemcpyAsync(d_rd, &rd8, 1, cudaMemcpyHostToDevice, st) == cudaSuccess &&
                 wmixb_mix_load_plan_device(plan, d_span, n_out, 0, d_src, 1, d_rd, nullptr, st) == WMIXB_OK;
        }
    }
    if (ok) ok = cudaMemcpyAsync(ring16 + pos, d_span, (size_t)first * 2, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (ok && first < n_out) ok = cudaMemcpyAsync(ring16, d_span + first, (size_t)(n_out - first) * 2, cudaMemcpyDeviceToHost, st) == cudaSuccess;
    if (ok) ok = cudaStreamSynchronize(st) == cudaSuccess;
    if (!ok) { snprintf(g_err, sizeof g_err, "load_data: CUDA call failed: %s", cudaGetErrorString(cudaGetLastError())); return head; }
    *tick += n_out * 2;                                                                                      // :1942-1953
    return w->ring_start + (size_t)((pos + n_out) % ring_len) * 2;
}

// wmix_load_data under the reference's own prototype (include/wmix.h): the view is read off the daemon's struct
static std::atomic<uint32_t> g_mix_freq{8000}, g_play_correct{3200};   // R:platform/alsa/plat.h:17-21
extern "C" void wmix_load_data_config(uint16_t mix_freq, uint32_t play_correct_bytes)
{
    g_mix_freq.store(mix_freq);
    g_play_correct.store(play_correct_bytes);
}
extern "C" WMix_Point wmix_load_data(WMix_Struct* wmix, WMix_Point src, uint32_t srcU8Len, uint16_t freq, uint8_t channels,
                                     uint8_t sample, WMix_Point head, uint8_t reduce, uint32_t* tick)
{
    if (!wmix) return head;                                                                                  // R:src/wmix.c:1664
    wmixb_mix_view v;
    memset(&v, 0, sizeof v);
    v.ring_start = wmix->start.U8;
    v.ring_bytes = (uint32_t)(wmix->end.U8 - wmix->start.U8);
    v.head_off = (uint32_t)(wmix->head.U8 - wmix->start.U8);
    v.tick = wmix->tick;
    v.play_correct = g_play_correct.load();
    v.mix_freq = (uint16_t)g_mix_freq.load();
    v.reduce_mode = wmix->reduceMode;
    v.run = wmix->run ? 1 : 0;
    v.device = g_default_device.load();
    WMix_Point out;
    out.U8 = wmixb_load_data_host(&v, src.U8, srcU8Len, freq, channels, sample, head.U8, reduce, tick);
    return out;
}

// ---- state snapshot ----
extern "C" size_t wmixb_stream_state_bytes(const wmixb_engine* e)
{
    if (!e) return 0;
    size_t b = 0;
    if (e->ns_rec) b += (size_t)ns_rec_floats(e) * 4 + 3 * ns::kHistBins * 2;
    if (e->nsx_rec) b += (size_t)nsx_rec_words(e) * 4 + 3 * nsx::kHistBins * 2;
    if (e->nsx_hb) b += (size_t)(e->ana - e->frame) * 2;
    if (e->ns_hb) b += (size_t)(e->ana - e->frame) * 4;
    if (e->agc_words) b += agc::N_WORDS * 4;
    if (e->vad_words) b += vad::N_WORDS * 4;
    if (e->aec_rec) b += e->aec_rec_floats * 4;
    return b;
}
extern "C" size_t wmixb_state_bytes_per_stream(const wmixb_engine* e) { return wmixb_stream_state_bytes(e); }

static int state_xfer(wmixb_engine* e, int s, void* buf, bool get)
{
    if (!e || !buf || s < 0 || s >= e->cfg.n_streams) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaStreamSynchronize(e->stream));
    char* p = (char*)buf;
    auto xfer = [&](void* dev, size_t bytes) -> cudaError_t {
        cudaError_t r = get ? cudaMemcpy(p, dev, bytes, cudaMemcpyDeviceToHost) : cudaMemcpy(dev, p, bytes, cudaMemcpyHostToDevice);
        p += bytes;
        return r;
    };
    auto xfer2d = [&](int32_t* dev, int words) -> cudaError_t {
        cudaError_t r = get ? cudaMemcpy2D(p, 4, dev + s, e->stride * 4, 4, words, cudaMemcpyDeviceToHost)
                            : cudaMemcpy2D(dev + s, e->stride * 4, p, 4, 4, words, cudaMemcpyHostToDevice);
        p += (size_t)words * 4;
        return r;
    };
    if (e->ns_rec) {
        CK(xfer(e->ns_rec + (size_t)s * ns_rec_floats(e), (size_t)ns_rec_floats(e) * 4));
        CK(xfer(e->ns_hist + (size_t)s * 3 * ns::kHistBins, 3 * ns::kHistBins * 2));
        if (e->ns_hb) CK(xfer(e->ns_hb + (size_t)s * (e->ana - e->frame), (size_t)(e->ana - e->frame) * 4));
    }
    if (e->nsx_rec) {
        CK(xfer(e->nsx_rec + (size_t)s * nsx_rec_words(e), (size_t)nsx_rec_words(e) * 4));
        CK(xfer(e->ns_hist + (size_t)s * 3 * nsx::kHistBins, 3 * nsx::kHistBins * 2));
        if (e->nsx_hb) CK(xfer(e->nsx_hb + (size_t)s * (e->ana - e->frame), (size_t)(e->ana - e->frame) * 2));
    }
    if (e->agc_words) CK(xfer2d(e->agc_words, agc::N_WORDS));
    if (e->vad_words) CK(xfer2d(e->vad_words, vad::N_WORDS));
    if (e->aec_rec) CK(xfer(e->aec_rec + (size_t)s * e->aec_rec_floats, e->aec_rec_floats * 4));
    return WMIXB_OK;
}
extern "C" int wmixb_get_stream_state(wmixb_engine* e, int s, void* h_buf) { return state_xfer(e, s, h_buf, true); }
extern "C" int wmixb_set_stream_state(wmixb_engine* e, int s, const void* h_buf) { return state_xfer(e, s, (void*)h_buf, false); }

// ---- RTP / G.711 wire framing for batches of legs (include/wmix_rtp.h) ----
static_assert(sizeof(wmixb_rtp_meta) == 16 && sizeof(wmixb_rtp_state) == 12, "wire-framing structs are part of the ABI");
constexpr int kRtpWords = WMIXB_RTP_PCMA_PAYLOAD / 4;       // 40 payload words per packet
__host__ __device__ inline uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }

extern "C" void wmixb_rtp_write_header(uint8_t out[12], uint8_t vpxcc, uint8_t marker, uint8_t pt, uint16_t seq, uint32_t timestamp,
                                       uint32_t ssrc)
{
    out[0] = vpxcc;
    out[1] = (uint8_t)((pt & 127) | ((marker & 1) << 7));
    out[2] = (uint8_t)(seq >> 8);
    out[3] = (uint8_t)seq;
    for (int k = 0; k < 4; ++k) { out[4 + k] = (uint8_t)(timestamp >> (24 - 8 * k)); out[8 + k] = (uint8_t)(ssrc >> (24 - 8 * k)); }
}

extern "C" void wmixb_rtp_read_header(const uint8_t in[12], wmixb_rtp_meta* m)
{
    memset(m, 0, sizeof *m);
    m->vpxcc = in[0];
    m->pt = in[1] & 127;
    m->marker = in[1] >> 7;
    m->seq = (uint16_t)((in[2] << 8) | in[3]);
    m->timestamp = ((uint32_t)in[4] << 24) | ((uint32_t)in[5] << 16) | ((uint32_t)in[6] << 8) | in[7];
    m->ssrc = ((uint32_t)in[8] << 24) | ((uint32_t)in[9] << 16) | ((uint32_t)in[10] << 8) | in[11];
    m->ok = (uint8_t)((in[0] >> 6) == 2 && (m->pt == WMIXB_RTP_PT_PCMA || m->pt == WMIXB_RTP_PT_PCMU));
}

// thread per payload word: word w of leg l moves from slab[l*stride + 12 + 4w] to codes[l*160 + 4w]
__global__ void rtp_unpack_kernel(const uint32_t* __restrict__ slab, const int32_t* __restrict__ sizes, int n, int stride_words,
                                  uint32_t fill, uint32_t* __restrict__ codes, wmixb_rtp_meta* __restrict__ meta)
{
    const size_t total = (size_t)n * kRtpWords;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int leg = (int)(idx / kRtpWords), w = (int)(idx - (size_t)leg * kRtpWords);
        const uint32_t* pkt = slab + (size_t)leg * stride_words;
        const uint32_t h0 = pkt[0];                                   // bytes 0..3 (little-endian load)
        const uint32_t b0 = h0 & 0xFF, pt = (h0 >> 8) & 127;
        bool ok = (b0 >> 6) == 2 && (pt == WMIXB_RTP_PT_PCMA || pt == WMIXB_RTP_PT_PCMU);
        if (sizes) ok = ok && sizes[leg] >= WMIXB_RTP_HEADER + WMIXB_RTP_PCMA_PAYLOAD;
        codes[idx] = ok ? pkt[WMIXB_RTP_HEADER / 4 + w] : fill;
        if (w == 0 && meta) {
            wmixb_rtp_meta m;
            m.timestamp = bswap32(pkt[1]);
            m.ssrc = bswap32(pkt[2]);
            m.seq = (uint16_t)((((h0 >> 16) & 0xFF) << 8) | (h0 >> 24));
            m.pt = (uint8_t)pt;
            m.marker = (uint8_t)((h0 >> 15) & 1);
            m.vpxcc = (uint8_t)b0;
            m.ok = (uint8_t)ok;
            m.reserved = 0;
            meta[leg] = m;
        }
    }
}

__global__ void rtp_pack_kernel(const uint32_t* __restrict__ codes, int n, int chn, wmixb_rtp_state* __restrict__ state,
                                uint32_t* __restrict__ slab, int stride_words)
{
    const size_t total = (size_t)n * kRtpWords;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int leg = (int)(idx / kRtpWords), w = (int)(idx - (size_t)leg * kRtpWords);
        uint32_t* pkt = slab + (size_t)leg * stride_words;
        pkt[WMIXB_RTP_HEADER / 4 + w] = codes[idx];
        if (w == 0) {
            wmixb_rtp_state st = state[leg];
            st.timestamp += (uint32_t)(WMIXB_RTP_PCMA_PAYLOAD / chn);            // R:src/wmixTask.c:1141
            // V=2, P=X=CC=0 (R:src/wmixTask.c:1058); seq big-endian in bytes 2..3
            pkt[0] = 0x80u | ((uint32_t)((st.pt & 127) | ((st.marker & 1) << 7)) << 8) | ((uint32_t)(st.seq >> 8) << 16) | ((uint32_t)(st.seq & 0xFF) << 24);
            pkt[1] = bswap32(st.timestamp);
            pkt[2] = bswap32(st.ssrc);
            st.seq = (uint16_t)(st.seq + 1);                                      // R:src/rtp.c:68
            state[leg] = st;
        }
    }
}

extern "C" int wmixb_rtp_unpack_device(const uint8_t* d_slab, const int32_t* d_sizes, int n, int stride, int law_fill, uint8_t* d_codes,
                                       wmixb_rtp_meta* d_meta, void* stream)
{
    if (n < 0 || (n && (!d_slab || !d_codes)) || (law_fill != 0 && law_fill != 1)) return WMIXB_EINVAL;
    if (stride < WMIXB_RTP_HEADER + WMIXB_RTP_PCMA_PAYLOAD || (stride & 3) || ((uintptr_t)d_slab & 3) || ((uintptr_t)d_codes & 3)) {
        snprintf(g_err, sizeof g_err, "rtp_unpack: stride must be >= 172 and a multiple of 4, buffers 4-byte aligned");
        return WMIXB_EINVAL;
    }
    if (n == 0) return WMIXB_OK;
    const uint32_t fill = law_fill == 0 ? 0xD5D5D5D5u : 0xFFFFFFFFu;      // linear 0 in A-law / mu-law
    rtp_unpack_kernel<<<ew_blocks((size_t)n * kRtpWords), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)d_slab, d_sizes, n, stride / 4, fill,
                                                                                           (uint32_t*)d_codes, d_meta);
    CK_LAUNCH();
    return WMIXB_OK;
}

extern "C" int wmixb_rtp_pack_device(const uint8_t* d_codes, int n, int chn, wmixb_rtp_state* d_state, uint8_t* d_slab, int stride,
                                     void* stream)
{
    if (n < 0 || (n && (!d_slab || !d_codes || !d_state)) || chn < 1 || chn > 2) return WMIXB_EINVAL;
    if (stride < WMIXB_RTP_HEADER + WMIXB_RTP_PCMA_PAYLOAD || (stride & 3) || ((uintptr_t)d_slab & 3) || ((uintptr_t)d_codes & 3)) {
        snprintf(g_err, sizeof g_err, "rtp_pack: stride must be >= 172 and a multiple of 4, buffers 4-byte aligned");
        return WMIXB_EINVAL;
    }
    if (n == 0) return WMIXB_OK;
    rtp_pack_kernel<<<ew_blocks((size_t)n * kRtpWords), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)d_codes, n, chn, d_state, (uint32_t*)d_slab, stride / 4);
    CK_LAUNCH();
    return WMIXB_OK;
}

// ---- wmix_pcm_zoom as a gather (include/wmix_zoom.h) ----
// out[s][k] = in[s][map[k]]: thread per output sample pair where possible; the table is shared by all streams
__global__ void zoom_gather_kernel(const int16_t* __restrict__ in, int16_t* __restrict__ out, const int32_t* __restrict__ map,
                                   uint32_t in_samples, uint32_t out_samples, int n_streams)
{
    const size_t total = (size_t)n_streams * out_samples;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t s = idx / out_samples;
        const uint32_t k = (uint32_t)(idx - s * out_samples);
        out[idx] = in[s * in_samples + map[k]];
    }
}

struct wmixb_zoom {
    int device = 0;
    uint32_t in_bytes = 0, in_samples = 0, out_samples = 0;
    int32_t* d_map = nullptr;
    std::vector<int32_t> h_map;
};

extern "C" int wmixb_zoom_create(int in_chn, int in_freq, uint32_t in_bytes, int out_chn, int out_freq, int device, wmixb_zoom** out)
{
    if (!out) return WMIXB_EINVAL;
    *out = nullptr;
    if (in_chn < 1 || in_chn > 2 || out_chn < 1 || out_chn > 2 || in_freq < 1 || out_freq < 1 || in_freq > 65535 || out_freq > 65535) {
        snprintf(g_err, sizeof g_err, "zoom: channels must be 1 or 2 and rates 1..65535 Hz");
        return WMIXB_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { snprintf(g_err, sizeof g_err, "no CUDA device — wmix_b200 has no CPU path"); return WMIXB_ENODEV; }
    if (device < 0 || device >= ndev) return WMIXB_EINVAL;
    wmixb_zoom* z = new (std::nothrow) wmixb_zoom();
    if (!z) return WMIXB_ENOMEM;
    z->device = device;
    z->in_bytes = in_bytes;
    z->in_samples = (in_bytes + 1) / 2;
    z->out_samples = host::zoom_map(in_chn, in_freq, in_bytes, out_chn, out_freq, nullptr);
    z->h_map.resize(z->out_samples ? z->out_samples : 1);
    host::zoom_map(in_chn, in_freq, in_bytes, out_chn, out_freq, z->h_map.data());
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaMalloc(&z->d_map, z->h_map.size() * sizeof(int32_t));
    if (ce == cudaSuccess) ce = cudaMemcpy(z->d_map, z->h_map.data(), z->h_map.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(z->d_map); delete z; return fail_cuda(ce, "zoom_create", __LINE__); }
    *out = z;
    return WMIXB_OK;
}

extern "C" void wmixb_zoom_destroy(wmixb_zoom* z)
{
    if (!z) return;
    cudaSetDevice(z->device);
    cudaFree(z->d_map);
    delete z;
}

extern "C" uint32_t wmixb_zoom_out_bytes(const wmixb_zoom* z) { return z ? z->out_samples * 2u : 0u; }

extern "C" int wmixb_zoom_map(const wmixb_zoom* z, int32_t* h_map)
{
    if (!z || !h_map) return WMIXB_EINVAL;
    memcpy(h_map, z->h_map.data(), (size_t)z->out_samples * sizeof(int32_t));
    return WMIXB_OK;
}

extern "C" int wmixb_zoom_device(const wmixb_zoom* z, const int16_t* d_in, int16_t* d_out, int n_streams, void* stream)
{
    if (!z || n_streams < 0 || (n_streams && z->out_samples && (!d_in || !d_out))) return WMIXB_EINVAL;
    if (n_streams == 0 || z->out_samples == 0) return WMIXB_OK;
    CK(cudaSetDevice(z->device));
    zoom_gather_kernel<<<ew_blocks((size_t)n_streams * z->out_samples), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, z->d_map, z->in_samples, z->out_samples, n_streams);
    CK_LAUNCH();
    return WMIXB_OK;
}

// ---- device self-test of ns::fdiv against the IEEE division ----
__global__ void fdiv_selftest_kernel(unsigned long long n, uint32_t seed, float a_lo_log2, float a_hi_log2, float b_lo_log2,
                                     float b_hi_log2, unsigned long long* mismatches)
{
    unsigned long long bad = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        // counter-based hash -> two floats with log-uniform magnitude and full random mantissas
        uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const uint32_t ma = (uint32_t)z & 0x7FFFFF, mb = (uint32_t)(z >> 23) & 0x7FFFFF;
        const float ua = (float)((z >> 46) & 0x1FF) / 512.f, ub = (float)((z >> 55) & 0x1FF) / 512.f;
        const int ea = (int)floorf(a_lo_log2 + ua * (a_hi_log2 - a_lo_log2)), eb = (int)floorf(b_lo_log2 + ub * (b_hi_log2 - b_lo_log2));
        float a = __int_as_float(((ea + 127) << 23) | ma), b = __int_as_float(((eb + 127) << 23) | mb);
        if (z & (1ull << 63)) a = -a;
        if ((i & 1023) == 0) a = 0.f;
        bad += (__float_as_int(ns::fdiv(a, b)) != __float_as_int(__fdiv_rn(a, b)));
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" int wmixb_selftest_fdiv(unsigned long long n, unsigned seed, float a_lo_log2, float a_hi_log2, float b_lo_log2,
                                   float b_hi_log2, unsigned long long* h_mismatches)
{
    if (!h_mismatches) return WMIXB_EINVAL;
    unsigned long long* d = nullptr;
    CK(cudaMalloc(&d, sizeof *d));
    CK(cudaMemset(d, 0, sizeof *d));
    fdiv_selftest_kernel<<<148 * 8, 256>>>(n, seed, a_lo_log2, a_hi_log2, b_lo_log2, b_hi_log2, d);
    CK_LAUNCH();
    CK(cudaMemcpy(h_mismatches, d, sizeof *d, cudaMemcpyDeviceToHost));
    CK(cudaFree(d));
    return WMIXB_OK;
}

// ---- bookkeeping ----
// ---- experiment / test knobs (none of them changes results) ----
extern "C" int wmixb_set_tuning(wmixb_engine* e, const char* key, int value)
{
    if (!e || !key) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    if (!strcmp(key, "ns_cfg")) {
        if (value < 0 || value >= (int)(sizeof kNsCfgs / sizeof kNsCfgs[0]) || !e->ns_rec) return WMIXB_EINVAL;
        e->ns_cfg = value;
        return e->ana == 256 ? ns_configure<256>(e) : ns_configure<128>(e);
    }
    if (!strcmp(key, "nsx_cfg")) {
        if (value < 0 || value >= (int)(sizeof kNsxCfgs / sizeof kNsxCfgs[0]) || !e->nsx_rec) return WMIXB_EINVAL;
        e->nsx_cfg = value;
        return e->ana == 256 ? nsx_configure<256>(e) : nsx_configure<128>(e);
    }
    if (!strcmp(key, "nsx_sync")) { e->nsx_sync = value != 0; return WMIXB_OK; }
    if (!strcmp(key, "ns_offline_staged")) { e->ns_offline_staged = value != 0; return WMIXB_OK; }
    if (!strcmp(key, "ns_align")) { if (value < 0 || value > 64) return WMIXB_EINVAL; e->ns_align = value; return WMIXB_OK; }
    if (!strcmp(key, "post_occ")) { if ((value < 2 || value > 5) && value != 22 && value != 0) return WMIXB_EINVAL; e->post_occ = value; return WMIXB_OK; }
    if (!strcmp(key, "aec_pf")) { e->aec_pf = value; return WMIXB_OK; }
    if (!strcmp(key, "aec_warps")) { if (value != 8 && value != 16) return WMIXB_EINVAL; e->aec_warps = value; return WMIXB_OK; }
    if (!strcmp(key, "aec_align")) { e->aec_align = value != 0; return WMIXB_OK; }
    if (!strcmp(key, "aec_grid")) { if (value < 1 || value > e->aec_grid_max) return WMIXB_EINVAL; e->aec_grid = value; return WMIXB_OK; }
    if (!strcmp(key, "host_zero_copy")) { e->host_zero_copy = value != 0; return WMIXB_OK; }
    if (!strcmp(key, "host_chunks")) { if (value < 1 || value > 64) return WMIXB_EINVAL; e->host_chunks_sync = e->host_chunks_pipe = value; return WMIXB_OK; }
    if (!strcmp(key, "host_lanes")) { if (value < 0 || value > kPipe) return WMIXB_EINVAL; e->host_lanes = value; return WMIXB_OK; }
    snprintf(g_err, sizeof g_err, "set_tuning: unknown key '%s'", key);
    return WMIXB_EINVAL;
}

extern "C" int wmixb_set_default_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { (void)cudaGetLastError(); return WMIXB_EINVAL; }
    g_default_device.store(device);
    return WMIXB_OK;
}
extern "C" int wmixb_default_device(void) { return g_default_device.load(); }

// The reference picks its suppressor at compile time (R:src/webrtc.c:511-523, `#define MAKE_WEBRTC_NSX`); the drop-in
// ns_init picks it from this process-wide switch (a build with -DMAKE_WEBRTC_NSX starts with the fixed-point core).
#ifdef MAKE_WEBRTC_NSX
static std::atomic<int> g_default_ns_core{1};
#else
static std::atomic<int> g_default_ns_core{0};
#endif
extern "C" int wmixb_set_default_ns_core(int ns_core)
{
    if (ns_core != 0 && ns_core != 1) return WMIXB_EINVAL;
    g_default_ns_core.store(ns_core);
    return WMIXB_OK;
}
extern "C" int wmixb_default_ns_core(void) { return g_default_ns_core.load(); }

// ---- pinned host buffers placed for a device ----
// Pinned pages are allocated by the calling thread inside cudaHostAlloc, so the NUMA node they land on is the one the
// thread runs on at that moment: the thread is moved onto the CPUs of the device's node (sysfs: the PCI function's
// numa_node and that node's cpulist) for the duration of the allocation and the first touch, then put back.
static bool numa_cpus_of_device(int device, cpu_set_t* set)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return false;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return false;
    char list[4096] = "";
    const bool have = fgets(list, sizeof list, f) != nullptr;
    fclose(f);
    if (!have) return false;
    CPU_ZERO(set);
    int n = 0;
    for (char* p = list; *p && *p != '\n';) {
        char* end;
        long a = strtol(p, &end, 10), b = a;
        if (end == p) break;
        if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, set); ++n; }
        p = *end == ',' ? end + 1 : end;
    }
    return n > 0;
}

extern "C" void* wmixb_host_alloc(size_t bytes, int device, int flags)
{
    if (bytes == 0 || cudaSetDevice(device) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    cpu_set_t old_set, node_set;
    const bool have_old = sched_getaffinity(0, sizeof old_set, &old_set) == 0;
    bool moved = false;
    if (have_old && numa_cpus_of_device(device, &node_set)) {
        cpu_set_t both;
        CPU_AND(&both, &node_set, &old_set);                       // never leave the cpuset the process was given
        if (CPU_COUNT(&both) > 0) moved = sched_setaffinity(0, sizeof both, &both) == 0;
    }
    void* p = nullptr;
    unsigned f = cudaHostAllocPortable;
    if (flags & WMIXB_HOST_WRITE_COMBINED) f |= cudaHostAllocWriteCombined;
    const cudaError_t ce = cudaHostAlloc(&p, bytes, f);
    if (ce == cudaSuccess && !(flags & WMIXB_HOST_WRITE_COMBINED)) memset(p, 0, bytes);   // first touch on the node
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
    if (ce != cudaSuccess) { fail_cuda(ce, "cudaHostAlloc", __LINE__); return nullptr; }
    return p;
}

extern "C" void wmixb_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

// What the copy engines alone sustain for one tick's traffic: h2d_bytes host -> device and d2h_bytes device -> host as
// bare cudaMemcpyAsync calls on two streams (full duplex), `reps` times back to back; *ms_per_rep = wall time per rep.
// The ceiling the host-buffer tick is measured against (no kernels, no library logic).
extern "C" int wmixb_host_copy_ceiling(int device, const void* h_src, void* h_dst, size_t h2d_bytes, size_t d2h_bytes, int reps,
                                       double* ms_per_rep)
{
    if (!ms_per_rep || reps < 1 || (h2d_bytes && !h_src) || (d2h_bytes && !h_dst)) return WMIXB_EINVAL;
    CK(cudaSetDevice(device));
    void *d_a = nullptr, *d_b = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    cudaError_t ce = cudaSuccess;
    float ms = 0.f;
#define WMX_TRY(x) do { if (ce == cudaSuccess) ce = (x); } while (0)
    WMX_TRY(cudaMalloc(&d_a, h2d_bytes ? h2d_bytes : 16));
    WMX_TRY(cudaMalloc(&d_b, d2h_bytes ? d2h_bytes : 16));
    WMX_TRY(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    WMX_TRY(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    WMX_TRY(cudaEventCreate(&e0));
    WMX_TRY(cudaEventCreate(&e1));
    WMX_TRY(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
    for (int pass = 0; pass < 2 && ce == cudaSuccess; ++pass) {       // pass 0 warms the mappings up
        WMX_TRY(cudaEventRecord(e0, s_in));
        WMX_TRY(cudaStreamWaitEvent(s_out, e0, 0));
        for (int r = 0; r < (pass ? reps : 2) && ce == cudaSuccess; ++r) {
            if (h2d_bytes) WMX_TRY(cudaMemcpyAsync(d_a, h_src, h2d_bytes, cudaMemcpyHostToDevice, s_in));
            if (d2h_bytes) WMX_TRY(cudaMemcpyAsync(h_dst, d_b, d2h_bytes, cudaMemcpyDeviceToHost, s_out));
        }
        WMX_TRY(cudaEventRecord(e2, s_out));
        WMX_TRY(cudaStreamWaitEvent(s_in, e2, 0));
        WMX_TRY(cudaEventRecord(e1, s_in));
        WMX_TRY(cudaEventSynchronize(e1));
    }
    WMX_TRY(cudaEventElapsedTime(&ms, e0, e1));
#undef WMX_TRY
    cudaFree(d_a); cudaFree(d_b);
    if (s_in) cudaStreamDestroy(s_in);
    if (s_out) cudaStreamDestroy(s_out);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (ce != cudaSuccess) return fail_cuda(ce, "host_copy_ceiling", __LINE__);
    *ms_per_rep = (double)ms / reps;
    return WMIXB_OK;
}

extern "C" int wmixb_sync(wmixb_engine* e)
{
    if (!e) return WMIXB_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    CK(cudaDeviceSynchronize());
    return WMIXB_OK;
}
extern "C" const char* wmixb_last_error(void) { return g_err; }
extern "C" long long wmixb_kernel_launches(void) { return g_launches.load(); }
extern "C" int wmixb_frame_len(const wmixb_engine* e) { return e ? e->row : 0; }
extern "C" void wmixb_ns_window(int ana, int block, float* out) { host::ns_window(ana, block, out); }
extern "C" int wmixb_agc_gain_table(int32_t table[32], int comp_db, int target_dbfs, int limiter, int analog_target)
{
    return host::agc_gain_table(table, (int16_t)comp_db, (int16_t)target_dbfs, limiter, (int16_t)analog_target);
}
extern "C" int wmixb_agc_analog_target(int comp_db) { return host::agc_analog_target((int16_t)comp_db); }
