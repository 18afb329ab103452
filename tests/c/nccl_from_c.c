/* Conference bus across two GPUs from plain C: one host thread per GPU, the exchange is the library's own NCCL
 * communicator (wmixb_nccl_bus_*, include/wmixb.h).  The check is internal to the product: the two-GPU result must be
 * the single-GPU result of the same conferences (whose parity with the oracle tests/test_multi_gpu.py establishes).
 * exit 0 = identical, 77 = fewer than two GPUs / no NCCL here, anything else = failure. */
#include "wmixb.h"
#include <cuda_runtime_api.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

enum { N_CONF = 96, PER_RANK = 3, WORLD = 2, FRAME = 80, TICKS = 6 };
#define N_LOCAL (N_CONF * PER_RANK)
#define N_ALL (N_LOCAL * WORLD)

static uint8_t g_legs[TICKS][N_ALL][FRAME];      /* conference-major; inside a conference rank 0's members first */
static uint8_t g_out[WORLD][TICKS][N_LOCAL][FRAME];
static int32_t g_bus[WORLD][TICKS][N_CONF][FRAME];
static unsigned char g_id[WMIXB_NCCL_ID_BYTES];
static int g_rc[WORLD];

#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, wmixb_last_error()); return rc_; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e_)); return 100; } } while (0)

static int run_rank(int rank)
{
    wmixb_config cfg = { .n_streams = N_LOCAL, .freq = 8000, .stages = 0, .device = rank };
    wmixb_engine *e = 0;
    wmixb_nccl_bus *nb = 0;
    int32_t start[N_CONF + 1];
    uint8_t *d_in = 0, *d_out = 0;
    int32_t *d_bus = 0;
    cudaStream_t st;
    static uint8_t mine[WORLD][N_LOCAL][FRAME];
    for (int c = 0; c <= N_CONF; ++c) start[c] = c * PER_RANK;
    CU(cudaSetDevice(rank));
    CHECK(wmixb_create(&cfg, &e));
    CHECK(wmixb_set_conferences(e, start, N_CONF));
    CHECK(wmixb_nccl_bus_create(e, rank, WORLD, g_id, &nb));
    CU(cudaStreamCreate(&st));
    CU(cudaMalloc((void **)&d_in, sizeof mine[0]));
    CU(cudaMalloc((void **)&d_out, sizeof mine[0]));
    CU(cudaMalloc((void **)&d_bus, sizeof g_bus[0][0]));
    for (int t = 0; t < TICKS; ++t) {
        for (int c = 0; c < N_CONF; ++c)
            memcpy(mine[rank][c * PER_RANK], g_legs[t][c * PER_RANK * WORLD + rank * PER_RANK], (size_t)PER_RANK * FRAME);
        CU(cudaMemcpyAsync(d_in, mine[rank], sizeof mine[0], cudaMemcpyHostToDevice, st));
        CHECK(wmixb_nccl_bus_tick_device(nb, 0, d_in, d_out, d_bus, st));
        CU(cudaMemcpyAsync(g_out[rank][t], d_out, sizeof mine[0], cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(g_bus[rank][t], d_bus, sizeof g_bus[0][0], cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    wmixb_nccl_bus_destroy(nb);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_bus);
    cudaStreamDestroy(st);
    wmixb_destroy(e);
    return 0;
}

static void *thread_main(void *arg)
{
    const int rank = (int)(size_t)arg;
    g_rc[rank] = run_rank(rank);
    return 0;
}

/* the same conferences on one GPU: wmixb_g711_bus_sum_device -> wmixb_g711_nminus1_device, no exchange */
static int single_gpu(int t, uint8_t (*out)[FRAME], int32_t (*bus)[FRAME])
{
    wmixb_config cfg = { .n_streams = N_ALL, .freq = 8000, .stages = 0, .device = 0 };
    wmixb_engine *e = 0;
    int32_t start[N_CONF + 1];
    uint8_t *d_in = 0, *d_out = 0;
    int32_t *d_bus = 0;
    for (int c = 0; c <= N_CONF; ++c) start[c] = c * PER_RANK * WORLD;
    CU(cudaSetDevice(0));
    CHECK(wmixb_create(&cfg, &e));
    CHECK(wmixb_set_conferences(e, start, N_CONF));
    CU(cudaMalloc((void **)&d_in, sizeof g_legs[0]));
    CU(cudaMalloc((void **)&d_out, sizeof g_legs[0]));
    CU(cudaMalloc((void **)&d_bus, sizeof g_bus[0][0]));
    CU(cudaMemcpy(d_in, g_legs[t], sizeof g_legs[0], cudaMemcpyHostToDevice));
    CHECK(wmixb_g711_bus_sum_device(e, 0, d_in, d_bus, 0));
    CHECK(wmixb_g711_nminus1_device(e, 0, d_bus, d_in, d_out, 0));
    CU(cudaMemcpy(out, d_out, sizeof g_legs[0], cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(bus, d_bus, sizeof g_bus[0][0], cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_bus);
    wmixb_destroy(e);
    return 0;
}

int main(void)
{
    int n_dev = 0;
    pthread_t th[WORLD];
    unsigned seed = 12345u;
    static uint8_t want_out[N_ALL][FRAME];
    static int32_t want_bus[N_CONF][FRAME];
    long bad = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < WORLD) { printf("skip: %d GPU(s)\n", n_dev); return 77; }
    if (wmixb_nccl_unique_id(g_id) != WMIXB_OK) { printf("skip: %s\n", wmixb_last_error()); return 77; }
    for (size_t i = 0; i < sizeof g_legs; ++i) { seed = seed * 1664525u + 1013904223u; ((uint8_t *)g_legs)[i] = (uint8_t)(seed >> 24); }
    for (int r = 0; r < WORLD; ++r) pthread_create(&th[r], 0, thread_main, (void *)(size_t)r);
    for (int r = 0; r < WORLD; ++r) pthread_join(th[r], 0);
    for (int r = 0; r < WORLD; ++r) if (g_rc[r]) return 2;
    for (int t = 0; t < TICKS; ++t) {
        if (single_gpu(t, want_out, want_bus)) return 3;
        for (int r = 0; r < WORLD; ++r) {
            bad += memcmp(g_bus[r][t], want_bus, sizeof want_bus) != 0;
            for (int c = 0; c < N_CONF; ++c)
                bad += memcmp(g_out[r][t][c * PER_RANK], want_out[c * PER_RANK * WORLD + r * PER_RANK], (size_t)PER_RANK * FRAME) != 0;
        }
    }
    printf("nccl bus from C: %d conferences x %d legs over %d GPUs, %d ticks, mismatching blocks = %ld, kernels launched = %lld\n",
           N_CONF, PER_RANK * WORLD, WORLD, TICKS, bad, wmixb_kernel_launches());
    return bad ? 1 : 0;
}
