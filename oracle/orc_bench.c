/* ORACLE (test/bench infrastructure) — multi-threaded CPU driver for the reference-arm and
 * cpu_baseline legs of bench.py.  Runs wmix's record chain ns_process -> agc_process ->
 * vad_process (R:src/wmix.c:613-710, AEC off) per stream per 10 ms tick, then the conference
 * bus sum, over contiguous stream ranges on `n_threads` pthreads.
 *
 * kind 0 ("port")      : the C restatement in this directory (orc_*).
 * kind 1 ("reference") : the unmodified reference, dlopen'ed from oracle/_ref/libwmix_ref.so
 *                        (its own ns_init/ns_process/... symbols, R:src/webrtc.h:32-61).
 * Never linked into or called by the product library. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "oracle.h"

typedef struct {
    void *(*ns_init)(int, int, bool *);
    void (*ns_process)(void *, int16_t *, int16_t *, int);
    void (*ns_release)(void *);
    void *(*agc_init)(int, int, int, int, bool *);
    int (*agc_process)(void *, int16_t *, int16_t *, int);
    void (*agc_release)(void *);
    void *(*vad_init)(int, int, int, bool *);
    void (*vad_process)(void *, int16_t *, int);
    void (*vad_release)(void *);
} ref_api;

typedef struct {
    int kind, freq, frame, first, count, n_streams, n_ticks, conf_size, n_prime;
    const int16_t *pcm; /* [n_ticks][n_streams][frame] */
    int16_t *out;       /* same shape, nullable */
    int32_t *bus;       /* [n_ticks][n_streams/conf_size][frame], nullable */
    const ref_api *api;
    pthread_barrier_t *bar;
    double seconds;
} job;

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void *worker(void *arg)
{
    job *j = (job *)arg;
    void **ns = calloc((size_t)j->count, sizeof(void *));
    void **agc = calloc((size_t)j->count, sizeof(void *));
    void **vad = calloc((size_t)j->count, sizeof(void *));
    int16_t buf[160];
    int s, t, i;
    for (s = 0; s < j->count; ++s) {
        if (j->kind == 1) {
            ns[s] = j->api->ns_init(1, j->freq, NULL);
            agc[s] = j->api->agc_init(1, j->freq, 10, 5, NULL);
            vad[s] = j->api->vad_init(1, j->freq, 10, NULL);
        } else {
            ns[s] = orc_ns_init(1, j->freq);
            agc[s] = orc_agc_init(1, j->freq, 10, 5);
            vad[s] = orc_vad_init(1, j->freq, 10);
        }
    }
    /* untimed warm-up on the SAME handles (ticks of the pool, cyclically): takes the suppressor past its start-up model
     * (50 frames) and the gain-map switch (200), the regime the GPU arm is timed in after its own warm-up ticks */
    for (t = 0; t < j->n_prime; ++t) {
        for (s = 0; s < j->count; ++s) {
            const size_t at = ((size_t)(t % j->n_ticks) * j->n_streams + j->first + s) * j->frame;
            memcpy(buf, j->pcm + at, sizeof(int16_t) * (size_t)j->frame);
            if (j->kind == 1) {
                j->api->ns_process(ns[s], buf, buf, j->frame);
                j->api->agc_process(agc[s], buf, buf, j->frame);
                j->api->vad_process(vad[s], buf, j->frame);
            } else {
                orc_ns_process((orc_ns *)ns[s], buf, buf, j->frame);
                orc_agc_process((orc_agc *)agc[s], buf, buf, j->frame);
                orc_vad_process((orc_vad *)vad[s], buf, j->frame);
            }
        }
    }
    pthread_barrier_wait(j->bar);
    {
        const double t0 = now_s();
        for (t = 0; t < j->n_ticks; ++t) {
            for (s = 0; s < j->count; ++s) {
                const size_t at = ((size_t)t * j->n_streams + j->first + s) * j->frame;
                memcpy(buf, j->pcm + at, sizeof(int16_t) * (size_t)j->frame);
                if (j->kind == 1) {
                    j->api->ns_process(ns[s], buf, buf, j->frame);
                    j->api->agc_process(agc[s], buf, buf, j->frame);
                    j->api->vad_process(vad[s], buf, j->frame);
                } else {
                    orc_ns_process((orc_ns *)ns[s], buf, buf, j->frame);
                    orc_agc_process((orc_agc *)agc[s], buf, buf, j->frame);
                    orc_vad_process((orc_vad *)vad[s], buf, j->frame);
                }
                if (j->out)
                    memcpy(j->out + at, buf, sizeof(int16_t) * (size_t)j->frame);
                if (j->bus) {
                    /* thread ranges are multiples of conf_size, so a bus row has one writer */
                    int32_t *b = j->bus + ((size_t)t * (j->n_streams / j->conf_size) + (j->first + s) / j->conf_size) * j->frame;
                    for (i = 0; i < j->frame; ++i)
                        b[i] += buf[i];
                }
            }
        }
        j->seconds = now_s() - t0;
    }
    pthread_barrier_wait(j->bar);
    for (s = 0; s < j->count; ++s) {
        if (j->kind == 1) {
            j->api->ns_release(ns[s]);
            j->api->agc_release(agc[s]);
            j->api->vad_release(vad[s]);
        } else {
            orc_ns_release((orc_ns *)ns[s]);
            orc_agc_release((orc_agc *)agc[s]);
            orc_vad_release((orc_vad *)vad[s]);
        }
    }
    free(ns);
    free(agc);
    free(vad);
    return NULL;
}

/* Returns wall seconds of the slowest thread for n_ticks ticks over n_streams streams, or a
 * negative value on error.  n_streams must be a multiple of conf_size. */
double orc_bench_chain_primed(const char *ref_so, int freq, int n_streams, int n_prime, int n_ticks, int conf_size, int n_threads,
                              const int16_t *pcm, int16_t *out, int32_t *bus);
double orc_bench_chain(const char *ref_so, int freq, int n_streams, int n_ticks, int conf_size, int n_threads,
                       const int16_t *pcm, int16_t *out, int32_t *bus)
{
    return orc_bench_chain_primed(ref_so, freq, n_streams, 0, n_ticks, conf_size, n_threads, pcm, out, bus);
}

/* same, after n_prime untimed ticks on the same handles */
double orc_bench_chain_primed(const char *ref_so, int freq, int n_streams, int n_prime, int n_ticks, int conf_size, int n_threads,
                              const int16_t *pcm, int16_t *out, int32_t *bus)
{
    ref_api api;
    job *jobs;
    pthread_t *th;
    pthread_barrier_t bar;
    int kind = ref_so ? 1 : 0, k, groups, per, extra, first = 0;
    double worst = 0;
    if (freq != 8000 && freq != 16000)
        return -1;
    if (conf_size < 1 || n_streams % conf_size)
        return -2;
    if (kind == 1) {
        void *h = dlopen(ref_so, RTLD_NOW | RTLD_LOCAL);
        if (!h)
            return -3;
#define SYM(n) *(void **)(&api.n) = dlsym(h, #n); if (!api.n) return -4;
        SYM(ns_init) SYM(ns_process) SYM(ns_release) SYM(agc_init) SYM(agc_process) SYM(agc_release)
        SYM(vad_init) SYM(vad_process) SYM(vad_release)
#undef SYM
    }
    groups = n_streams / conf_size;
    if (n_threads > groups)
        n_threads = groups;
    if (n_threads < 1)
        n_threads = 1;
    jobs = calloc((size_t)n_threads, sizeof(job));
    th = calloc((size_t)n_threads, sizeof(pthread_t));
    pthread_barrier_init(&bar, NULL, (unsigned)n_threads);
    if (bus)
        memset(bus, 0, sizeof(int32_t) * (size_t)n_ticks * groups * (freq / 100));
    per = groups / n_threads;
    extra = groups % n_threads;
    for (k = 0; k < n_threads; ++k) {
        int g = per + (k < extra ? 1 : 0);
        jobs[k].kind = kind;
        jobs[k].freq = freq;
        jobs[k].frame = freq / 100;
        jobs[k].first = first;
        jobs[k].count = g * conf_size;
        jobs[k].n_streams = n_streams;
        jobs[k].n_ticks = n_ticks;
        jobs[k].n_prime = n_prime < 0 ? 0 : n_prime;
        jobs[k].conf_size = conf_size;
        jobs[k].pcm = pcm;
        jobs[k].out = out;
        jobs[k].bus = bus;
        jobs[k].api = &api;
        jobs[k].bar = &bar;
        first += g * conf_size;
        pthread_create(&th[k], NULL, worker, &jobs[k]);
    }
    for (k = 0; k < n_threads; ++k) {
        pthread_join(th[k], NULL);
        if (jobs[k].seconds > worst)
            worst = jobs[k].seconds;
    }
    pthread_barrier_destroy(&bar);
    free(jobs);
    free(th);
    return worst;
}
