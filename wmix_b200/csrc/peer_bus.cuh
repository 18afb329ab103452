// Cross-GPU conference bus in ONE persistent kernel: local partial sums, push over NVLink into every
// peer's mailbox, flag, wait for the peers' rows, reduce, N-minus-one read-out (+ G.711 encode).
//
// The reference has a single mix bus inside one process (R:src/wmix.c:1639-1702: every producer adds
// into the ring); here a conference's participants may live on several GPUs of one node.  Each rank
// (one process per GPU) owns a MAILBOX that all peers can write through CUDA IPC peer mappings:
//
//     slots [2][world][n_conf * frame]  int32    partial bus rows, double-buffered by tick parity
//     flags [2][world][n_conf * frame/16]   uint32   tick sequence number the 16-sample tile belongs to
//
// Phase 1 (per 16-sample tile of a conference row this CTA owns): sum the local members (decoding G.711
//   in registers) and store the tile into slot [parity][my_rank] of EVERY rank's mailbox (16-byte stores;
//   peer stores ride NVLink); then the CTA meets, one warp fences at system scope and sets
//   flags[parity][my_rank][tile] = seq on every rank for each of the CTA's tiles (WMX_PEER_PUBLISH).
// Phase 2 (same tiles): wait until all `world` flags of the tile carry `seq`, add the `world` partial
//   tiles in rank order (int32: exact, so any order gives the same bits), and emit
//   out = clamp16(bus - own) for the local members (re-encoded for the G.711 variants).
// A CTA finishes phase 1 for all of its tiles before it waits for anything, and the grid never exceeds
// what is co-resident, so no rank can block another: tiles stream across NVLink while later tiles are
// still being summed.  Slot reuse is safe because a rank can only start tick t+2 after its tick t+1
// kernel has seen every peer's t+1 rows, i.e. after every peer's tick-t kernel has retired.
// A wait that exceeds `timeout_ns` (a peer that never ticks) sets *error and gives up instead of hanging.
#pragma once
#include "g711_mix.cuh"

namespace wmx {
namespace peer {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 512;

struct Ring {
    int32_t* slots[kMaxWorld];     // mailbox of rank r (local pointer for r == rank, IPC mapping otherwise)
    uint32_t* flags[kMaxWorld];
    int32_t* result[kMaxWorld];    // reduce-scatter mode: [2][n_conf * frame] finished bus rows pushed by their owners
    uint32_t* rflags[kMaxWorld];   //                      [2][n_conf] tick sequence number of a finished row
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Spin until *f == seq.  The time-out runs on the SM's cycle counter (%globaltimer is slow to read and coarse); `timeout`
// is in cycles.  Returns false when it gave up.
__device__ __forceinline__ bool wait_flag(const uint32_t* f, uint32_t seq, long long timeout)
{
    if (ld_acquire_sys(f) == seq) return true;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) != seq) {
        if (clock64() - t0 > timeout) return false;
        __nanosleep(20);
    }
    return true;
}
// Publish: every thread of the CTA has issued its (local and peer) stores; after the CTA barrier ONE warp fences at system
// scope — cumulativity makes the whole CTA's stores visible before anything that warp writes afterwards — and then sets
// the flags with plain system-scope stores.  (A fence in every thread and a release per flag cost a system-scope round
// trip each: fifteen warps' worth instead of one.)
#define WMX_PEER_PUBLISH(COUNT, FLAG_PTR_EXPR)                                    \
    __syncthreads();                                                              \
    if (threadIdx.x < 32) {                                                       \
        __threadfence_system();                                                   \
        for (int j = threadIdx.x; j < (COUNT); j += 32) {                         \
            uint32_t* fp_ = (FLAG_PTR_EXPR);                                      \
            if (fp_) st_relaxed_sys(fp_, seq);                                    \
        }                                                                         \
    }

// sum of the members first + slice, first + slice + S, ... of one sample column: four loads in flight per trip
template <int LAW>
__device__ __forceinline__ int32_t sum_members(const void* src, int first, int last, int slice, int S, int frame, int col);

template <int LAW>
__device__ __forceinline__ int load_sample(const void* src, size_t idx)
{
    if (LAW < 0) return static_cast<const int16_t*>(src)[idx];
    const uint8_t c = static_cast<const uint8_t*>(src)[idx];
    return LAW == 0 ? alaw2linear(c) : ulaw2linear(c);
}

template <int LAW>
__device__ __forceinline__ int32_t sum_members(const void* src, int first, int last, int slice, int S, int frame, int col)
{
    int32_t acc = 0;
    int p = first + slice;
    for (; p + 3 * S < last; p += 4 * S) {
        const int a0 = load_sample<LAW>(src, (size_t)p * frame + col), a1 = load_sample<LAW>(src, (size_t)(p + S) * frame + col);
        const int a2 = load_sample<LAW>(src, (size_t)(p + 2 * S) * frame + col), a3 = load_sample<LAW>(src, (size_t)(p + 3 * S) * frame + col);
        acc += (a0 + a1) + (a2 + a3);               // int32: exact in any order
    }
    for (; p < last; p += S) acc += load_sample<LAW>(src, (size_t)p * frame + col);
    return acc;
}
// N-minus-one read-out of the members first + slice, ... of one sample column against the finished bus value b
template <int LAW>
__device__ __forceinline__ void emit_members(const void* src, void* out, int32_t b, int first, int last, int slice, int S, int frame, int col)
{
    auto put = [&](size_t idx, int own) {
        const int16_t v = sat16(b - own);
        if (LAW < 0) static_cast<int16_t*>(out)[idx] = v;
        else static_cast<uint8_t*>(out)[idx] = LAW == 0 ? linear2alaw(v) : linear2ulaw(v);
    };
    int p = first + slice;
    for (; p + 3 * S < last; p += 4 * S) {
        const size_t i0 = (size_t)p * frame + col, i1 = (size_t)(p + S) * frame + col, i2 = (size_t)(p + 2 * S) * frame + col,
                     i3 = (size_t)(p + 3 * S) * frame + col;
        const int a0 = load_sample<LAW>(src, i0), a1 = load_sample<LAW>(src, i1), a2 = load_sample<LAW>(src, i2), a3 = load_sample<LAW>(src, i3);
        put(i0, a0);
        put(i1, a1);
        put(i2, a2);
        put(i3, a3);
    }
    for (; p < last; p += S) {
        const size_t idx = (size_t)p * frame + col;
        put(idx, load_sample<LAW>(src, idx));
    }
}

// LAW < 0: int16 PCM in/out; 0 / 1: A-law / mu-law codes in/out.
// Work unit = one TILE: kTile consecutive samples of one conference row, summed by a GROUP of kTile x S
// threads (S member slices, a power of two chosen by the host from the largest local conference).  A CTA
// holds G = blockDim / (kTile * S) groups working on G tiles at a time: big conferences get 32 slices
// and one tile per CTA step, thousands of small ones run eight one-warp tiles per CTA step.
//
// TILE is 16 samples, or the whole frame (80 / 160) when there are many conferences: a tile costs `world` release
// stores for its flags whatever its size, so thousands of small conferences are better served by one tile per bus row
// (5 / 10 times fewer flags, and 320 / 640-byte bursts per peer instead of 64-byte ones).  The choice must be the same on
// every rank (it fixes the flag indexing) and is therefore made from n_conf alone (tile_for()).
constexpr int kTile = 16;
__host__ __device__ inline int tile_for(int n_conf, int frame) { return n_conf >= 256 ? frame : kTile; }

template <int LAW, int TILE>
__global__ void __launch_bounds__(kThreads)
peer_bus_kernel(Ring ring, int rank, int world, uint32_t seq, const void* __restrict__ src, void* __restrict__ out,
                int32_t* __restrict__ bus_out, const int32_t* __restrict__ conf_start, int n_conf, int frame, int S,
                long long timeout, int* __restrict__ error)
{
    extern __shared__ int32_t sh_all[];             // per group: [S][TILE] partials + [TILE] finished tile
    const int gthreads = TILE * S;
    const int G = blockDim.x / gthreads;
    const int g = threadIdx.x / gthreads, gt = threadIdx.x - g * gthreads;
    const int tx = gt % TILE, slice = gt / TILE;
    int32_t* sh = sh_all + g * (S + 1) * TILE;
    int32_t* done = sh + S * TILE;
    const int tiles_per_row = frame / TILE;
    const int n_tiles = n_conf * tiles_per_row;
    const int par = (int)(seq & 1u);
    const size_t row_words = (size_t)n_conf * frame;
    const int steps = (n_tiles + gridDim.x * G - 1) / (gridDim.x * G);     // uniform trip count (CTA barriers inside)

    // ---- phase 1: partial tiles out to every mailbox ----
    for (int it = 0; it < steps; ++it) {
        const int t = (it * gridDim.x + blockIdx.x) * G + g;
        const bool valid = t < n_tiles;
        const int c = valid ? t / tiles_per_row : 0, x0 = valid ? (t - c * tiles_per_row) * TILE : 0;
        int32_t acc = 0;
        if (valid) acc = sum_members<LAW>(src, conf_start[c], conf_start[c + 1], slice, S, frame, x0 + tx);
        __syncthreads();                            // the previous step's readers are done with sh
        sh[slice * TILE + tx] = acc;
        __syncthreads();
        if (gt < TILE) {
            int32_t v = 0;
            for (int k = 0; k < S; ++k) v += sh[k * TILE + gt];
            done[gt] = v;
        }
        __syncthreads();
        // world x TILE/4 16-byte stores per tile
        if (valid)
            for (int j = gt; j < world * (TILE / 4); j += gthreads) {
                const int r = j / (TILE / 4), v = j % (TILE / 4);
                int4* dst = reinterpret_cast<int4*>(ring.slots[r] + ((size_t)par * world + rank) * row_words + (size_t)c * frame + x0);
                dst[v] = reinterpret_cast<const int4*>(done)[v];
            }
    }
    // one system-scope fence per CTA, then the flags of all of its tiles (a fence per tile would cost an
    // NVLink round trip each); flag j: rank j % world, tile k = j / world = it * G + g
    WMX_PEER_PUBLISH(steps * G * world,
                     ((((j / world) / G) * gridDim.x + blockIdx.x) * G + ((j / world) % G)) < n_tiles
                         ? ring.flags[j % world] + ((size_t)par * world + rank) * n_tiles + ((((j / world) / G) * gridDim.x + blockIdx.x) * G + ((j / world) % G))
                         : nullptr)

    // ---- phase 2: gather, reduce, N-minus-one ----
    const int32_t* my_slots = ring.slots[rank] + (size_t)par * world * row_words;
    const uint32_t* my_flags = ring.flags[rank] + (size_t)par * world * n_tiles;
    for (int it = 0; it < steps; ++it) {
        const int t = (it * gridDim.x + blockIdx.x) * G + g;
        const bool valid = t < n_tiles;
        const int c = valid ? t / tiles_per_row : 0, x0 = valid ? (t - c * tiles_per_row) * TILE : 0;
        if (valid && gt < world && !wait_flag(my_flags + (size_t)gt * n_tiles + t, seq, timeout)) atomicExch(error, 1 + gt);
        __syncthreads();
        if (valid && gt < TILE) {
            int32_t v = 0;
            for (int r = 0; r < world; ++r) v += __ldcg(my_slots + (size_t)r * row_words + (size_t)c * frame + x0 + gt);
            done[gt] = v;
            if (bus_out) bus_out[(size_t)c * frame + x0 + gt] = v;
        }
        __syncthreads();
        if (valid && out) emit_members<LAW>(src, out, done[tx], conf_start[c], conf_start[c + 1], slice, S, frame, x0 + tx);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Reduce-scatter / all-gather form, for MANY conferences on MANY ranks.  Pushing every partial row to every rank costs
// world^2 row transfers and world^2 flags per row across the node; here row c has an OWNER (rank c % world):
//   phase 1  every rank pushes its partial row to the owner only (+ one flag),
//   phase 2a the owner waits for the `world` partials of its rows, adds them (int32: exact) and pushes the finished row
//            to every rank's result area (+ one flag per rank),
//   phase 2b every rank waits for the finished rows and writes the N-minus-one read-out of its own members.
// 2 * world row transfers per row instead of world^2, two NVLink hops of latency instead of one.  TILE == frame (one
// tile per row).  No CTA waits before it has finished phase 1 for all of its rows, phase 2a only waits for phase-1
// flags and phase 2b only for phase-2a flags, and the grid is one resident wave: ranks cannot block each other.
// Slot reuse by tick parity is safe for the same reason as above: a rank that has seen all finished rows of tick t+1
// knows every rank has started tick t+1, i.e. retired tick t.
template <int LAW, int TILE>
__global__ void __launch_bounds__(kThreads)
peer_bus_rs_kernel(Ring ring, int rank, int world, uint32_t seq, const void* __restrict__ src, void* __restrict__ out,
                   int32_t* __restrict__ bus_out, const int32_t* __restrict__ conf_start, int n_conf, int S,
                   long long timeout, int* __restrict__ error)
{
    extern __shared__ int32_t sh_all[];             // per group: [S][TILE] partials + [TILE] finished row
    const int gthreads = TILE * S;
    const int G = blockDim.x / gthreads;
    const int g = threadIdx.x / gthreads, gt = threadIdx.x - g * gthreads;
    const int tx = gt % TILE, slice = gt / TILE;
    int32_t* sh = sh_all + g * (S + 1) * TILE;
    int32_t* done = sh + S * TILE;
    const int par = (int)(seq & 1u);
    const size_t row_words = (size_t)n_conf * TILE;
    const int steps = (n_conf + gridDim.x * G - 1) / (gridDim.x * G);
    const int n_owned = rank < n_conf ? (n_conf - rank + world - 1) / world : 0;      // rows c = k * world + rank
    const int osteps = (n_owned + gridDim.x * G - 1) / (gridDim.x * G);

    // ---- phase 1: partial rows to their owners ----
    for (int it = 0; it < steps; ++it) {
        const int c = (it * gridDim.x + blockIdx.x) * G + g;
        const bool valid = c < n_conf;
        int32_t acc = 0;
        if (valid) acc = sum_members<LAW>(src, conf_start[c], conf_start[c + 1], slice, S, TILE, tx);
        __syncthreads();
        sh[slice * TILE + tx] = acc;
        __syncthreads();
        if (gt < TILE) {
            int32_t v = 0;
            for (int k = 0; k < S; ++k) v += sh[k * TILE + gt];
            done[gt] = v;
        }
        __syncthreads();
        if (valid) {
            int4* dst = reinterpret_cast<int4*>(ring.slots[c % world] + ((size_t)par * world + rank) * row_words + (size_t)c * TILE);
            for (int j = gt; j < TILE / 4; j += gthreads) dst[j] = reinterpret_cast<const int4*>(done)[j];
        }
    }
    WMX_PEER_PUBLISH(steps * G,
                     (((j / G) * gridDim.x + blockIdx.x) * G + (j % G)) < n_conf
                         ? ring.flags[(((j / G) * gridDim.x + blockIdx.x) * G + (j % G)) % world] + ((size_t)par * world + rank) * n_conf +
                               (((j / G) * gridDim.x + blockIdx.x) * G + (j % G))
                         : nullptr)

    // ---- phase 2a: the rows this rank owns: wait for the partials, reduce, push the finished row to everybody ----
    const int32_t* my_slots = ring.slots[rank] + (size_t)par * world * row_words;
    const uint32_t* my_flags = ring.flags[rank] + (size_t)par * world * n_conf;
    for (int it = 0; it < osteps; ++it) {
        const int k = (it * gridDim.x + blockIdx.x) * G + g;
        const bool valid = k < n_owned;
        const int c = valid ? k * world + rank : 0;
        if (valid && gt < world && !wait_flag(my_flags + (size_t)gt * n_conf + c, seq, timeout)) atomicExch(error, 1 + gt);
        __syncthreads();
        if (valid && gt < TILE) {
            int32_t v = 0;
            for (int r = 0; r < world; ++r) v += __ldcg(my_slots + (size_t)r * row_words + (size_t)c * TILE + gt);
            done[gt] = v;
        }
        __syncthreads();
        if (valid)
            for (int j = gt; j < world * (TILE / 4); j += gthreads) {
                const int r = j / (TILE / 4), v = j % (TILE / 4);
                int4* dst = reinterpret_cast<int4*>(ring.result[r] + (size_t)par * row_words + (size_t)c * TILE);
                dst[v] = reinterpret_cast<const int4*>(done)[v];
            }
        __syncthreads();                            // done is rewritten by the next step
    }
    WMX_PEER_PUBLISH(osteps * G * world,
                     ((((j / world) / G) * gridDim.x + blockIdx.x) * G + ((j / world) % G)) < n_owned
                         ? ring.rflags[j % world] + (size_t)par * n_conf + (size_t)((((j / world) / G) * gridDim.x + blockIdx.x) * G + ((j / world) % G)) * world + rank
                         : nullptr)

    // ---- phase 2b: finished rows in, N-minus-one out ----
    const int32_t* my_result = ring.result[rank] + (size_t)par * row_words;
    const uint32_t* my_rflags = ring.rflags[rank] + (size_t)par * n_conf;
    for (int it = 0; it < steps; ++it) {
        const int c = (it * gridDim.x + blockIdx.x) * G + g;
        const bool valid = c < n_conf;
        if (valid && gt == 0 && !wait_flag(my_rflags + c, seq, timeout)) atomicExch(error, 1 + c % world);
        __syncthreads();
        if (valid && gt < TILE) {
            const int32_t v = __ldcg(my_result + (size_t)c * TILE + gt);
            done[gt] = v;
            if (bus_out) bus_out[(size_t)c * TILE + gt] = v;
        }
        __syncthreads();
        if (valid && out) emit_members<LAW>(src, out, done[tx], conf_start[c], conf_start[c + 1], slice, S, TILE, tx);
    }
}

// the exchange pattern, like the tile, must be the same on every rank: a function of (n_conf, frame, world) only
__host__ __device__ inline bool reduce_scatter_for(int n_conf, int frame, int world) { return tile_for(n_conf, frame) == frame && world >= 4; }

}  // namespace peer
}  // namespace wmx
