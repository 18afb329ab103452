/* wmixb — batched, stream-parallel C entry points of wmix_b200 (additive API).
 *
 * The reference (wexiangis/wmix) runs exactly ONE record chain per process:
 *   wmix_shmem_write_circle: ns_process -> [aec_process2] -> agc_process -> vad_process
 *   (R:src/wmix.c:613-710), producers mixing through wmix_load_data (R:src/wmix.c:1639).
 * This header drives N independent streams of that same chain per 10 ms tick on one B200.
 * Per-stream results are bit-compatible with calling the reference handle API
 * (include/webrtc.h == R:src/webrtc.h:32-61) once per stream.
 *
 * C ABI only: plain pointers and sizes.  `stream` arguments are a cudaStream_t passed as void*.
 * Pointers named d_* are device memory, h_* host memory (pinned for full copy overlap).
 * Every function returns 0 on success, a negative WMIXB_E* code otherwise; there is no CPU
 * fallback: without a usable CUDA device wmixb_create fails with WMIXB_ENODEV.
 */
#ifndef WMIXB_H
#define WMIXB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    WMIXB_OK = 0,
    WMIXB_EINVAL = -1,   /* bad argument (same cases for which the reference *_init returns NULL) */
    WMIXB_ENODEV = -2,   /* no CUDA device / driver */
    WMIXB_ECUDA = -3,    /* a CUDA call failed; see wmixb_last_error() */
    WMIXB_ENOMEM = -4
};

/* stage bits, executed in the reference's order NS -> AEC -> AGC -> VAD (R:src/wmix.c:613-710) */
enum { WMIXB_NS = 1, WMIXB_AGC = 2, WMIXB_VAD = 4, WMIXB_AEC = 8 };

typedef struct wmixb_config {
    int n_streams;      /* independent mono streams on this GPU                                  */
    int freq;           /* 8000 or 16000 (10 ms tick = 80 / 160 samples per stream), or 32000: rows of 320 samples through
                           the 16 kHz cores the way the reference's handles treat that rate (NS on the first 160 samples,
                           the rest zero; AGC on two 5 ms packets; VAD on the 320-sample packet) — NS / AGC / VAD with
                           wmixb_tick_device and wmixb_tick_host only                                                  */
    int stages;         /* OR of WMIXB_NS / WMIXB_AGC / WMIXB_VAD: state is allocated for these   */
    int ns_policy;      /* 0..3; wmix uses 2   (R:src/webrtc.c:532 NS_AGGRESSIVE)                */
    int agc_gain_db;    /* compressionGaindB = wmix's `value` (R:src/webrtc.c:707), e.g. 5        */
    int vad_mode;       /* 0..3; wmix uses 3   (R:src/webrtc.c:16 VAD_AGGRESSIVE)                */
    int device;         /* CUDA device ordinal                                                   */
    int aec_far_depth;  /* far-end history kept per stream, in 64-sample partitions; 0 = 32.  The
                           reference keeps 250 (T:.../aec/aec_core.c:37); the handle API uses 252.   */
    int ns_high_band;   /* 1: allocate the high-band history wmixb_ns2_* needs (wmix's stereo NS, see below)      */
    int ns_core;        /* 0: WebRtcNs_* float core (what wmix ships); 1: WebRtcNsx_* fixed-point core — the reference's
                           own switch, `#define MAKE_WEBRTC_NSX` (R:src/webrtc.c:511-523), as a run-time choice          */
    int reserved[6];
} wmixb_config;

typedef struct wmixb_engine wmixb_engine;

int wmixb_create(const wmixb_config* cfg, wmixb_engine** out);
void wmixb_destroy(wmixb_engine* e);
/* re-initialise streams [first, first+count) — what tearing a wmix handle down and re-creating
 * it does (R:src/wmix.c:565-600) */
int wmixb_reset(wmixb_engine* e, int first, int count);
/* change the AGC compression gain for the whole engine (agc_addition, R:src/webrtc.c:824-838) */
int wmixb_set_agc_gain(wmixb_engine* e, int gain_db);

/* One 10 ms tick for all streams, PCM resident on the device.
 * d_in / d_out: int16 [n_streams][frame] (frame = freq/100), may alias.  d_vad (nullable):
 * uint8 [n_streams] speech flags of this tick.  `stages` selects a subset of the configured
 * stages (0 = all configured).  Asynchronous on `stream`. */
int wmixb_tick_device(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int stages,
                      void* stream);
/* Same, host buffers (pinned for overlap): the streams are cut into chunks whose H2D copy, kernels and
 * D2H copy are pipelined over the engine's own CUDA streams; returns when h_out / h_vad are complete. */
int wmixb_tick_host(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int stages);
/* wmixb_tick_host plus the conference bus of the processed tick (wmixb_set_conferences first):
 * h_bus int32 [n_conf][frame] = what every producer adding its 10 ms into the mix ring through
 * wmix_load_data yields while no partial sum clips (R:src/wmix.c:1678-1702). */
int wmixb_tick_host_bus(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int32_t* h_bus, int stages);
/* Output selection: in the host-buffer ticks h_out, h_vad and h_bus are each nullable — a NULL output is computed on the
 * device but not copied back (a mixer that only needs the conference bus and the speech flags moves 4 MB instead of
 * 36 MB device -> host per 100 000-stream tick).  At least one of them must be given. */

/* Pipelined form of the two calls above for a host that feeds ticks back to back (a media server's steady state):
 * _submit queues the tick's H2D copies, kernels, bus sum and D2H copies and returns; _wait returns when the OLDEST
 * submitted tick's h_out / h_vad / h_bus (nullable) are complete.  At most two ticks are in flight, so tick t+1's
 * input copy and first kernels overlap tick t's last kernels and output copy instead of the pipeline draining at every
 * tick boundary; results are identical to wmixb_tick_host_bus.  The buffers of a tick must stay valid (and should be
 * pinned) until its _wait returns.  Do not mix with the synchronous calls while a tick is in flight. */
int wmixb_tick_host_submit(wmixb_engine* e, const int16_t* h_in, int16_t* h_out, uint8_t* h_vad, int32_t* h_bus, int stages);
int wmixb_tick_host_wait(wmixb_engine* e);

/* Host tick for G.711 legs — wmix's RTP PCMA path decodes every 20 ms payload with G711a2PCM before it reaches the mixer and
 * encodes what goes back with PCM2G711a (R:src/wmixTask.c:1139-1143, :1278-1282; R:src/g711codec.c): here the tick's input
 * arrives and its output leaves as A-law / mu-law codes, ONE byte per sample over the bus instead of two, and the codecs run
 * on the device around the configured stages:
 *     codes -> G711x2PCM -> [NS -> AGC -> VAD] -> bus (wmixb_set_conferences) -> read-out -> PCM2G711x -> codes
 * h_codes_in / h_codes_out: uint8 [n_streams][frame] (h_codes_out nullable); law: 0 = A-law, 1 = mu-law; h_vad (nullable): uint8
 * [n_streams]; h_bus (nullable): int32 [n_conf][frame], the exact conference sum of the processed legs.  nminus1 != 0: the leg
 * handed back is what its conference sounds like WITHOUT it, clamp16(bus - own) (needs conferences); 0: the leg's own processed
 * PCM.  stages = 0: all configured.  Blocking; buffers should be pinned. */
int wmixb_tick_host_g711(wmixb_engine* e, int law, const uint8_t* h_codes_in, uint8_t* h_codes_out, uint8_t* h_vad, int32_t* h_bus,
                         int nminus1, int stages);

/* VAD on 20 ms packets, in place: what wmix itself configures (vad_init(.., WMIX_INTERVAL_MS = 20, ..),
 * R:src/wmix.c:703; 20 ms thresholds of T:.../vad/vad_core.c:149-164).  d_pcm: int16 [n_streams][2*frame],
 * attenuated by the wrapper's mute ramp (R:src/webrtc.c:127-141); d_vad (nullable): uint8 [n_streams].
 * Shares the per-stream VAD state with the 10 ms stage — use one packet size per engine. */
int wmixb_vad20_device(wmixb_engine* e, int16_t* d_pcm, uint8_t* d_vad, void* stream);
int wmixb_vad20_host(wmixb_engine* e, int16_t* h_pcm, uint8_t* h_vad);

/* wmix's stereo noise suppression: ns_process hands WebRtcNs the RIGHT channel as a second band (num_bands = chn,
 * R:src/webrtc.c:624-636), which ProcessCore delays by its analysis overlap and scales with one time-domain gain per
 * frame taken from the left channel's speech probability and filter (T:.../ns/ns_core.c:1214-1261, :1361-1414).
 * d_in / d_out = left (the normal NS path), d_in_hb / d_out_hb = right; all int16 [n_streams][frame], in and out may
 * alias pairwise.  Needs WMIXB_NS and ns_high_band = 1 at creation. */
int wmixb_ns2_device(wmixb_engine* e, const int16_t* d_in, const int16_t* d_in_hb, int16_t* d_out, int16_t* d_out_hb, void* stream);
int wmixb_ns2_host(wmixb_engine* e, const int16_t* h_in, const int16_t* h_in_hb, int16_t* h_out, int16_t* h_out_hb);

/* VAD on 32 kHz packets of 10 ms (320 samples), in place: the handle API's 32 kHz case (R:src/webrtc.c:43, :66-67;
 * CalcVad32khz, T:.../vad/vad_core.c:623-643: 32k -> 16k -> 8k, then the 8 kHz detector with the 10 ms thresholds).
 * Runs on an engine created with freq = 16000 and WMIXB_VAD; d_pcm: int16 [n_streams][320].  Use one packet kind
 * per engine. */
int wmixb_vad32_device(wmixb_engine* e, int16_t* d_pcm, uint8_t* d_vad, void* stream);
int wmixb_vad32_host(wmixb_engine* e, int16_t* h_pcm, uint8_t* h_vad);

/* Echo canceller (stage WMIXB_AEC), the arithmetic of aec_process2 (R:src/webrtc.c:410-483) for every
 * stream: BufferFarend(d_far) then Process(d_near) -> d_out.  d_far == NULL is aec_process (near only),
 * d_near == NULL is aec_setFrameFar (far only; d_out unused).  Buffers are int16 [n_streams][samples],
 * samples = 80 or 160 per call (10 ms at the engine's rate, or wmix's 20 ms packets at 8 kHz);
 * d_out may alias d_near.  delay_ms (0..500) is the reported sound-card delay, 0 in wmix
 * (R:src/wmix.c:651-657). */
int wmixb_aec_device(wmixb_engine* e, const int16_t* d_far, const int16_t* d_near, int16_t* d_out, int samples,
                     int delay_ms, void* stream);
/* Same with host buffers (pinned for full overlap): H2D, kernel, D2H on the engine's stream, then waits. */
int wmixb_aec_host(wmixb_engine* e, const int16_t* h_far, const int16_t* h_near, int16_t* h_out, int samples, int delay_ms);
/* The whole record chain NS -> AEC -> AGC -> VAD of one 10 ms tick (wmix's own order); d_far feeds the AEC. */
int wmixb_tick_chain_device(wmixb_engine* e, const int16_t* d_far, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad,
                            int stages, int delay_ms, void* stream);
/* Sticky per-stream AEC flags OR-ed over all streams (0 = healthy): bit 0 = a far-end rewind/backlog needed
 * more history than aec_far_depth holds, bit 1 = far-end underrun (the reference asserts it cannot happen).
 * *h_flagged (nullable) receives the number of streams with any flag.  Synchronises the engine stream. */
int wmixb_aec_status(wmixb_engine* e, int* h_flags, int* h_flagged);

/* The daemon's record tick at ITS cadence, for every stream (R:src/wmix.c:528-760; WMIX_INTERVAL_MS = 20): per 20 ms
 * package  playPkgBuff_add(play)  (R:src/wmix.c:1419)  then  ns_process -> aec_process2(playPkgBuff_get(AEC_INTERVALMS),
 * mic, mic, ..., 0) -> agc_process -> vad_process  with the handle geometries wmix itself creates: NS and AGC on two
 * 10 ms packets, the AEC on one 160-sample packet at 8 kHz / two at 16 kHz (aec_init(.., 20, ..)), the VAD on ONE
 * 20 ms packet (vad_init(.., 20, ..)).  The play FIFO (AEC_INTERVALMS / 20 + 2 packages, R:src/wmixConf.h:141) lives on
 * the device, slot-major, and the reference's slot arithmetic is reproduced as written (R:src/wmix.c:496-509) — it
 * reads the oldest package, not the one AEC_INTERVALMS back, on most ticks.  aec_interval_ms must be a whole number of
 * packages.  Buffers: int16 [n_streams][freq/50]; d_out may alias d_mic; d_vad and d_far_used (the far end the AEC
 * was given, for inspection) are nullable.  One tick = 1 copy + 4..6 launches. */
typedef struct wmixb_record wmixb_record;
int wmixb_record_create(wmixb_engine* e, int aec_interval_ms, wmixb_record** out);
void wmixb_record_destroy(wmixb_record* r);
int wmixb_record_far_slot(const wmixb_record* r);   /* slot the NEXT get would read (tests) */
int wmixb_record_tick_device(wmixb_record* r, const int16_t* d_play, const int16_t* d_mic, int16_t* d_out, uint8_t* d_vad,
                             int16_t* d_far_used, int stages, void* stream);

/* same with host buffers (pinned for overlap): H2D of play + mic, the tick, D2H of the result and the flags, then waits */
int wmixb_record_tick_host(wmixb_record* r, const int16_t* h_play, const int16_t* h_mic, int16_t* h_out, uint8_t* h_vad, int stages);

/* Persistent offline mode: every stream runs n_frames consecutive frames inside one launch per
 * stage.  d_in / d_out: int16 [n_streams][n_frames][frame].  d_vad (nullable): [n_streams][n_frames]. */
int wmixb_offline_device(wmixb_engine* e, const int16_t* d_in, int16_t* d_out, uint8_t* d_vad, int n_frames,
                         int stages, void* stream);

/* Conference bus.  Streams are grouped into conferences by contiguous ranges:
 * conference c = streams [h_conf_start[c], h_conf_start[c+1]).
 *   bus_sum     : d_bus[c][i] = sum over members of d_pcm[s][i]           (exact int32)
 *   bus_nminus1 : d_out[s][i] = clamp16(d_bus[conf(s)][i] - d_pcm[s][i])  (N-minus-one read-out)
 * d_bus: int32 [n_conf][frame] — the buffer an NCCL int32 sum all-reduce runs on between the
 * two calls when a conference spans GPUs. */
int wmixb_set_conferences(wmixb_engine* e, const int32_t* h_conf_start, int n_conf);
int wmixb_bus_sum_device(wmixb_engine* e, const int16_t* d_pcm, int32_t* d_bus, void* stream);
int wmixb_bus_nminus1_device(wmixb_engine* e, const int32_t* d_bus, const int16_t* d_pcm, int16_t* d_out,
                             void* stream);

/* Conference bus across the GPUs of one node, fused with its exchange (one process or host thread per GPU).
 * Every rank holds some members of each of the SAME n_conf conferences (its own wmixb_set_conferences table,
 * empty ranges allowed).  One call = one kernel per rank: local partial sums (G.711 decoded in registers) are
 * stored straight into every peer's mailbox over NVLink, row by row, then each rank adds the `world` partial
 * rows (exact int32) and writes the N-minus-one read-out of its own members.  Equivalent to
 * wmixb_*bus_sum_device -> int32 sum all-reduce (e.g. NCCL) -> wmixb_*bus_nminus1_device, bit for bit.
 *   law: -1 = int16 PCM in/out, 0 = A-law, 1 = mu-law codes in/out.  d_out and d_bus are nullable.
 * Ranks must call tick in lock-step (same number of calls), each on one stream of its own; a peer that never
 * arrives trips a 2 s timeout (WMIXB_PEER_TIMEOUT_MS) reported by wmixb_peer_bus_status, not a hang.
 * Wiring: create on every rank, then either exchange the WMIXB_PEER_HANDLE_BYTES blobs between processes
 * (any transport) and call _connect with all `world` blobs in rank order, or, inside one process, call
 * _connect_local with the `world` peer objects. */
typedef struct wmixb_peer_bus wmixb_peer_bus;
#define WMIXB_PEER_HANDLE_BYTES 80
int wmixb_peer_bus_create(wmixb_engine* e, int rank, int world, wmixb_peer_bus** out);
/* same with explicit options (NULL = defaults); tile / reduce_scatter must be identical on ALL ranks */
typedef struct wmixb_peer_opts {
    int tile;              /* 0 = automatic; 16, or the frame length (one tile per bus row)                          */
    int reduce_scatter;    /* -1 = automatic; 0 / 1 = all-to-all / reduce-scatter + all-gather (row tiles only)       */
    int timeout_ms;        /* 0 = 2000: how long a rank waits for a peer before it raises its error flag             */
    int ranks_per_device;  /* 0 = 2: ranks whose peer kernels may be resident on this GPU at once (sizes the grid)    */
    int reserved[4];
} wmixb_peer_opts;
int wmixb_peer_bus_create_ex(wmixb_engine* e, int rank, int world, const wmixb_peer_opts* opts, wmixb_peer_bus** out);
void wmixb_peer_bus_destroy(wmixb_peer_bus* pb);
int wmixb_peer_bus_handle(const wmixb_peer_bus* pb, void* handle_out);
int wmixb_peer_bus_connect(wmixb_peer_bus* pb, const void* handles);
int wmixb_peer_bus_connect_local(wmixb_peer_bus* pb, wmixb_peer_bus* const* peers);
int wmixb_peer_bus_tick_device(wmixb_peer_bus* pb, int law, const void* d_in, void* d_out, int32_t* d_bus, void* stream);
int wmixb_peer_bus_status(wmixb_peer_bus* pb, int* h_error);

/* The same exchange over NCCL, from C: wmixb_*bus_sum_device -> ncclAllReduce(int32, sum) over the node's NVLink / NVSwitch ->
 * wmixb_*bus_nminus1_device behind one call, for a host that has no communicator of its own.  libnccl is NOT a link-time
 * dependency of this library: it is opened on first use (dlopen "libnccl.so.2", or the path given to wmixb_nccl_load), and
 * every entry point below reports WMIXB_ENODEV with a reason when it is not there.  Wiring: rank 0 calls wmixb_nccl_unique_id
 * and ships the WMIXB_NCCL_ID_BYTES blob to the other ranks over any transport; every rank then calls wmixb_nccl_bus_create
 * (collective).  Ticks are collective too: same number of calls on every rank, each on a stream of its own device.  Same
 * arguments and the same bits as wmixb_peer_bus_tick_device (d_bus must be given: the all-reduce runs on it). */
typedef struct wmixb_nccl_bus wmixb_nccl_bus;
#define WMIXB_NCCL_ID_BYTES 128
int wmixb_nccl_load(const char* libnccl_path);          /* optional; NULL = default search */
int wmixb_nccl_unique_id(void* id_out);
int wmixb_nccl_bus_create(wmixb_engine* e, int rank, int world, const void* id, wmixb_nccl_bus** out);
void wmixb_nccl_bus_destroy(wmixb_nccl_bus* nb);
int wmixb_nccl_bus_tick_device(wmixb_nccl_bus* nb, int law, const void* d_in, void* d_out, int32_t* d_bus, void* stream);

/* G.711 on device buffers (R:src/g711codec.c).  law: 0 = A-law, 1 = mu-law.  n = samples. */
int wmixb_g711_encode_device(int law, const int16_t* d_pcm, uint8_t* d_codes, size_t n, void* stream);
int wmixb_g711_decode_device(int law, const uint8_t* d_codes, int16_t* d_pcm, size_t n, void* stream);
/* fused G.711 leg of config 5: decode -> (caller all-reduces d_bus) -> N-minus-one -> encode */
int wmixb_g711_bus_sum_device(wmixb_engine* e, int law, const uint8_t* d_codes, int32_t* d_bus, void* stream);
int wmixb_g711_nminus1_device(wmixb_engine* e, int law, const int32_t* d_bus, const uint8_t* d_codes,
                              uint8_t* d_out_codes, void* stream);

/* Same-format branch of wmix_load_data on a device-resident ring (R:src/wmix.c:1678-1702):
 * ring[(pos+i) % ring_len] = volumeAdd(ring[..], src[i] / rdce); returns the new position in
 * *new_pos.  Calls issued in order on one stream reproduce the reference's chained adds. */
int wmixb_mix_load_device(int16_t* d_ring, uint32_t ring_len, uint32_t pos, const int16_t* d_src, uint32_t n,
                          int rdce, uint32_t* new_pos, void* stream);

/* Different-format branches of wmix_load_data on a device-resident MONO ring (R:src/wmix.c:1704-1939): a 16-bit
 * source whose rate or channel count differs from the bus is dropped / linearly filled into it under the
 * reference's float phase accumulator (faster source: frames skipped; slower source: n-step float ramps between
 * neighbouring frames; stereo: left sample only).  The accumulator never depends on the audio, so a plan per
 * (format, chunk length) is walked once on the host and the kernel evaluates it for n_src sources at once:
 * d_src int16 [n_src][src_bytes/2] are added IN ORDER, source s divided by d_rdce[s] (device, nullable = 1) —
 * the same bits as n_src consecutive wmix_load_data calls starting at the same head.  src_bytes must be whole
 * frames; a rate ratio that would overrun the reference's 64-entry ramp buffer is refused (WMIXB_EINVAL), as is a
 * chunk longer than the ring.  *new_pos = position after the chunk. */
typedef struct wmixb_mixplan wmixb_mixplan;
int wmixb_mixplan_create(int src_chn, int src_freq, uint32_t src_bytes, int mix_freq, int device, wmixb_mixplan** out);
void wmixb_mixplan_destroy(wmixb_mixplan* m);
uint32_t wmixb_mixplan_out_samples(const wmixb_mixplan* m);     /* bus samples one chunk adds (tickAdd / 2) */
int wmixb_mixplan_tables(const wmixb_mixplan* m, int32_t* h_map, uint16_t* h_ramp);   /* host copies, for tests */
int wmixb_mix_load_plan_device(const wmixb_mixplan* m, int16_t* d_ring, uint32_t ring_len, uint32_t pos, const int16_t* d_src,
                               int n_src, const uint8_t* d_rdce, uint32_t* new_pos, void* stream);

/* wmix_load_data itself (R:src/wmix.h:40-49, R:src/wmix.c:1639-1956) for the daemon's HOST ring, mono 16-bit bus: the
 * same arguments, with the WMix_Struct fields the reference reads — run, start / end, head, tick, reduceMode — and the
 * two build constants WMIX_FREQ and VIEW_PLAY_CORRECT passed as a view.  Bookkeeping as in the reference (a producer
 * without a head, or whose tick fell behind, restarts play_correct bytes ahead of the play pointer; its samples are
 * divided by reduce_mode unless its own `reduce` equals it; *tick advances by the bytes written; 8- and 32-bit sources
 * write nothing); the adds run on the GPU (same-format or the resampling branches), one round trip per call.  Returns
 * the new head; on a CUDA failure the ring is untouched and the old head comes back (wmixb_last_error()). */
typedef struct wmixb_mix_view {
    uint8_t* ring_start;      /* wmix->start.U8                                     */
    uint32_t ring_bytes;      /* wmix->end.U8 - wmix->start.U8 (WMIX_BUFF_SIZE)      */
    uint32_t head_off;        /* wmix->head.U8 - wmix->start.U8: the play pointer    */
    uint32_t tick;            /* wmix->tick                                         */
    uint32_t play_correct;    /* VIEW_PLAY_CORRECT (bytes)                           */
    uint16_t mix_freq;        /* WMIX_FREQ                                          */
    uint8_t reduce_mode;      /* wmix->reduceMode                                   */
    uint8_t run;              /* wmix->run                                          */
    int device;               /* CUDA device ordinal                                */
} wmixb_mix_view;
uint8_t* wmixb_load_data_host(const wmixb_mix_view* w, const uint8_t* src, uint32_t src_bytes, uint16_t freq, uint8_t channels,
                              uint8_t sample, uint8_t* head, uint8_t reduce, uint32_t* tick);

/* state snapshot / restore of one stream (checkpointing; byte layout is engine-internal) */
size_t wmixb_stream_state_bytes(const wmixb_engine* e);
int wmixb_get_stream_state(wmixb_engine* e, int stream_index, void* h_buf);
int wmixb_set_stream_state(wmixb_engine* e, int stream_index, const void* h_buf);

/* Pinned host buffers for the host-buffer ticks, placed on the NUMA node of `device` (the calling thread is moved onto
 * that node's CPUs while the pages are allocated and first touched).  flags: WMIXB_HOST_WRITE_COMBINED for buffers the
 * host only writes (tick inputs).  NULL on failure (wmixb_last_error()). */
enum { WMIXB_HOST_WRITE_COMBINED = 1 };
void* wmixb_host_alloc(size_t bytes, int device, int flags);
void wmixb_host_free(void* p);
/* What the copy engines alone sustain for one tick's traffic: h2d_bytes up and d2h_bytes down as bare cudaMemcpyAsync
 * calls on two streams, `reps` times back to back.  The ceiling a host-buffer tick can be compared with. */
int wmixb_host_copy_ceiling(int device, const void* h_src, void* h_dst, size_t h2d_bytes, size_t d2h_bytes, int reps,
                            double* ms_per_rep);

/* device used by the drop-in entry points that have no device argument (include/webrtc.h handles, g711codec.h,
 * wmix_pcm_zoom, wmix_load_data); default 0 */
int wmixb_set_default_device(int device);
int wmixb_default_device(void);
/* suppressor behind the drop-in ns_init / ns_process: 0 = WebRtcNs_* float core (default, what wmix ships), 1 = WebRtcNsx_*
 * fixed-point core.  The reference makes this choice at compile time (`#define MAKE_WEBRTC_NSX`, R:src/webrtc.c:511-523);
 * building this library with -DMAKE_WEBRTC_NSX makes 1 the default.  Applies to handles created afterwards. */
int wmixb_set_default_ns_core(int ns_core);
int wmixb_default_ns_core(void);

/* experiment / test knobs; none changes results.  keys: "ns_cfg", "nsx_cfg", "ns_align", "post_occ", "aec_pf", "aec_grid",
 * "host_chunks", "host_lanes", "host_zero_copy" (small engines: kernels run on a pinned device-mapped staging block) */
int wmixb_set_tuning(wmixb_engine* e, const char* key, int value);

/* bookkeeping */
int wmixb_sync(wmixb_engine* e);
const char* wmixb_last_error(void);
long long wmixb_kernel_launches(void);          /* kernels this library has launched so far */
size_t wmixb_state_bytes_per_stream(const wmixb_engine* e);
int wmixb_frame_len(const wmixb_engine* e);
/* device self-test: the NS kernel's range-restricted float division against the IEEE one on n random operand
 * pairs, |a| in [2^a_lo, 2^a_hi) (and exact zeros), b in [2^b_lo, 2^b_hi); *h_mismatches must come back 0 */
int wmixb_selftest_fdiv(unsigned long long n, unsigned seed, float a_lo_log2, float a_hi_log2, float b_lo_log2,
                        float b_hi_log2, unsigned long long* h_mismatches);
/* init-time tables, exported so tests can pin them against the reference's literals */
void wmixb_ns_window(int ana, int block, float* out);
int wmixb_agc_gain_table(int32_t table[32], int comp_db, int target_dbfs, int limiter, int analog_target);
int wmixb_agc_analog_target(int comp_db);

#ifdef __cplusplus
}
#endif
#endif
