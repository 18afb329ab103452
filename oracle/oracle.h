/* ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, scalar, one-stream-at-a-time restatement of the reference hot path
 * (wexiangis/wmix: src/webrtc.c handle layer + the WebRTC C cores in pkg/webrtc_cut.tar.gz +
 * src/g711codec.c + the same-format branch of wmix_load_data).  Written from the algorithm,
 * not copied; every function cites the reference file:line it follows
 * (R: = /root/reference, T: = inside R:pkg/webrtc_cut.tar.gz).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Nothing under wmix_b200/ includes, links or dlopens it.
 *
 * Parity pin: tests/test_oracle_pin.py checks every function here against the unmodified
 * reference compiled by oracle/build_ref.sh (oracle/_ref/libwmix_ref.so), against the
 * known-answer vectors transcribed from the reference's own unit tests, and against the
 * committed fixtures under tests/golden/.
 */
#ifndef WMIX_ORACLE_H
#define WMIX_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- G.711 (R:src/g711codec.c) ---- */
uint8_t orc_linear2alaw(int pcm);
uint8_t orc_linear2ulaw(int pcm);
int16_t orc_alaw2linear(uint8_t code);
int16_t orc_ulaw2linear(uint8_t code);
int orc_PCM2G711a(const char *in, char *out, int in_bytes);   /* returns samples */
int orc_PCM2G711u(const char *in, char *out, int in_bytes);
int orc_G711a2PCM(const char *in, char *out, int n_codes);    /* returns bytes   */
int orc_G711u2PCM(const char *in, char *out, int n_codes);

/* ---- mix bus (R:src/wmix.c:1617-1702) ---- */
int16_t orc_volume_add(int16_t a, int16_t b);
/* bus[i] = volumeAdd(bus[i], src[i] / rdce) over a ring that wraps at ring_len samples.
 * Returns the new write position (in samples). */
uint32_t orc_mix_same_format(int16_t *ring, uint32_t ring_len, uint32_t pos,
                             const int16_t *src, uint32_t n, uint8_t rdce);
/* exact int32 conference bus and N-minus-one read-out (not in the reference) */
void orc_bus_sum(int32_t *bus, const int16_t *pcm, int n_part, int frame);
void orc_bus_nminus1(int16_t *out, const int32_t *bus, const int16_t *own, int frame);

/* different-format branches of wmix_load_data for a mono bus (R:src/wmix.c:1704-1939): drop / linear-fill
 * resampling driven by a float phase accumulator; returns the new position, *written = samples added */
uint32_t orc_mix_resample(int16_t *ring, uint32_t ring_len, uint32_t pos, const int16_t *src, uint32_t src_bytes,
                          uint16_t freq, uint8_t channels, uint16_t mix_freq, uint8_t rdce, uint32_t *written);

/* the whole wmix_load_data (R:src/wmix.c:1639-1956) over a view of the WMix_Struct fields it reads; offsets in bytes */
typedef struct {
    uint32_t ring_bytes, head_off, tick, play_correct;
    uint16_t mix_freq;
    uint8_t reduce_mode, run;
} orc_mix_view;
int32_t orc_wmix_load_data(const orc_mix_view *w, uint8_t *ring, const uint8_t *src, uint32_t src_bytes, uint16_t freq,
                           uint8_t channels, uint8_t sample, int32_t head_off, uint8_t reduce, uint32_t *tick);

/* play-package FIFO feeding the AEC its far end (R:src/wmix.c:482-526), whole-package delays */
#define ORC_FIFO_MAX_PKG 64
#define ORC_FIFO_MAX_BYTES 1280
typedef struct {
    int n_pkg, pkg_bytes, count;
    uint8_t buf[ORC_FIFO_MAX_PKG][ORC_FIFO_MAX_BYTES];
} orc_play_fifo;
void orc_play_fifo_init(orc_play_fifo *f, int n_pkg, int pkg_bytes);
void orc_play_fifo_add(orc_play_fifo *f, const uint8_t *pkg);
int orc_play_fifo_slot(int count, int n_pkg, int delay_pkgs);
void orc_play_fifo_get(const orc_play_fifo *f, uint8_t *out, int delay_pkgs);

/* nearest-sample rate / channel conversion (R:src/wmix.c:49-222) */
uint32_t orc_len_of_out(uint8_t in_chn, uint16_t in_freq, uint32_t in_len, uint8_t out_chn, uint16_t out_freq);
uint32_t orc_len_of_in(uint8_t in_chn, uint16_t in_freq, uint8_t out_chn, uint16_t out_freq, uint32_t out_len);
uint32_t orc_pcm_zoom(uint8_t in_chn, uint16_t in_freq, const uint8_t *in, uint32_t in_len, uint8_t out_chn,
                      uint16_t out_freq, uint8_t *out);

/* ---- RTP framing of the G.711 legs (R:src/rtp.h:51-70, R:src/rtp.c:20-99, R:src/wmixTask.c:1139-1143) ---- */
void orc_rtp_header_bytes(uint8_t out[12], uint8_t cc, uint8_t x, uint8_t p, uint8_t v, uint8_t pt, uint8_t m,
                          uint16_t seq, uint32_t timestamp, uint32_t ssrc);
void orc_rtp_send_step(uint32_t *timestamp, uint32_t ssrc, uint16_t *seq, uint8_t pt, uint8_t marker, int chn,
                       const uint8_t *codes, int n_codes, uint8_t *packet);
int orc_rtp_parse(const uint8_t *packet, uint16_t *seq, uint32_t *timestamp, uint32_t *ssrc, uint8_t *pt, uint8_t *marker);

/* ---- signal-processing primitives (T:webrtc/common_audio/signal_processing) ---- */
int16_t orc_norm_w32(int32_t a);
int16_t orc_norm_u32(uint32_t a);
int16_t orc_size_in_bits(uint32_t n);
int16_t orc_sat16(int32_t v);
int32_t orc_div_w32_w16(int32_t num, int16_t den);
int32_t orc_energy(const int16_t *v, int n, int *scale);
int32_t orc_sqrt(int32_t value);
void orc_downsample_by2(const int16_t *in, int len, int16_t *out, int32_t st[8]);

/* ---- VAD (T:webrtc/common_audio/vad) ---- */
typedef struct {
    int32_t ds_state[4];
    int16_t noise_means[12], speech_means[12], noise_stds[12], speech_stds[12];
    int32_t frame_counter;
    int16_t over_hang, num_of_speech;
    int16_t age[96], low_value[96];
    int16_t mean_value[6];
    int16_t upper_state[5], lower_state[5], hp_state[4];
    int16_t over_hang_max_1[3], over_hang_max_2[3], individual[3], total[3];
    int vad;
} orc_vad_core;
void orc_vad_core_init(orc_vad_core *v, int mode);
int16_t orc_vad_features(orc_vad_core *v, const int16_t *in, int len, int16_t feat[6]);
int32_t orc_vad_gaussian(int16_t input, int16_t mean, int16_t std, int16_t *delta);
void orc_vad_downsample(const int16_t *in, int16_t *out, int32_t st[2], int in_len);
int16_t orc_vad_find_minimum(orc_vad_core *v, int16_t feature, int channel);
int orc_vad_core_process(orc_vad_core *v, int fs, const int16_t *frame, int len); /* -1/0/1 */

/* handle layer: R:src/webrtc.c:40-167 (mono; chn>1 averaged exactly as the reference) */
typedef struct {
    orc_vad_core core;
    int chn, freq, interval_ms, pkg, reduce;
} orc_vad;
orc_vad *orc_vad_init(int chn, int freq, int interval_ms);
void orc_vad_process(orc_vad *h, int16_t *frame, int frame_num);
void orc_vad_release(orc_vad *h);

/* ---- AGC (T:webrtc/modules/audio_processing/agc/legacy) ---- */
typedef struct {
    int32_t down[8];
    int16_t hp, counter, log_ratio, mean_long;
    int32_t var_long;
    int16_t std_long, mean_short;
    int32_t var_short;
    int16_t std_short;
} orc_agc_vad;
typedef struct {
    int32_t cap_slow, cap_fast, gain;
    int32_t table[32];
    int16_t gate_prev;
    orc_agc_vad near_vad;
    int fs;
    int16_t comp_db, target_dbfs, analog_target;
    uint8_t limiter;
} orc_agc_core;
int orc_agc_gain_table(int32_t table[32], int16_t comp_db, int16_t target_dbfs,
                       uint8_t limiter, int16_t analog_target);
int16_t orc_agc_analog_target(int16_t comp_db);
void orc_agc_core_init(orc_agc_core *a, int fs, int comp_db);
int orc_agc_core_set_gain(orc_agc_core *a, int comp_db);
int16_t orc_agc_process_vad(orc_agc_vad *s, const int16_t *in, int n);
int orc_agc_core_process(orc_agc_core *a, const int16_t *in, int16_t *out, int n);

typedef struct {
    orc_agc_core core;
    int chn, freq, interval_ms, pkg;
} orc_agc;
orc_agc *orc_agc_init(int chn, int freq, int interval_ms, int value);
int orc_agc_process(orc_agc *h, int16_t *in, int16_t *out, int frame_num);
void orc_agc_addition(orc_agc *h, uint8_t value);
void orc_agc_release(orc_agc *h);

/* ---- Ooura real FFT as used by NS (T:webrtc/common_audio/fft4g.c) ---- */
void orc_rdft(int n, int isgn, float *a, int *ip, float *w);

/* ---- NS float core (T:webrtc/modules/audio_processing/ns/ns_core.c) ---- */
typedef struct orc_ns_core orc_ns_core;
typedef struct {
    orc_ns_core *core;
    int chn, freq, pkg;
} orc_ns;
orc_ns *orc_ns_init(int chn, int freq);
void orc_ns_process(orc_ns *h, const int16_t *in, int16_t *out, int frame_num);
void orc_ns_release(orc_ns *h);
/* introspection for tests */
int orc_ns_block_index(const orc_ns *h);
const float *orc_ns_prior_model(const orc_ns *h);

/* ---- NS fixed-point core, "nsx" (T:webrtc/modules/audio_processing/ns/nsx_core.c, nsx_core_c.c; the handle layer
 * of R:src/webrtc.c:563-660 built with MAKE_WEBRTC_NSX, R:src/webrtc.c:512) ---- */
typedef struct orc_nsx_core orc_nsx_core;
typedef struct {
    orc_nsx_core *core;
    int chn, freq, pkg;
} orc_nsx;
orc_nsx *orc_nsx_init(int chn, int freq);                       /* policy 2 = wmix's NS_AGGRESSIVE */
orc_nsx *orc_nsx_init_policy(int chn, int freq, int policy);
void orc_nsx_process(orc_nsx *h, const int16_t *in, int16_t *out, int frame_num);
void orc_nsx_release(orc_nsx *h);
int orc_nsx_block_index(const orc_nsx *h);
int orc_nsx_state(const orc_nsx *h, int32_t *out, int cap);     /* canonical state dump, returns the word count */
int orc_nsx_tables(int16_t *out, int cap);                      /* derived constant tables, returns the entry count */

/* ---- AEC float core as wmix drives it (T:webrtc/modules/audio_processing/aec; R:src/webrtc.c:217-500) ---- */
typedef struct orc_aec orc_aec;
orc_aec *orc_aec_init(int chn, int freq, int interval_ms);
int orc_aec_set_frame_far(orc_aec *h, const int16_t *far, int frame_num);
int orc_aec_process(orc_aec *h, const int16_t *near, int16_t *out, int frame_num, int delay_ms);
int orc_aec_process2(orc_aec *h, const int16_t *far, const int16_t *near, int16_t *out, int frame_num, int delay_ms);
void orc_aec_release(orc_aec *h);
void orc_aec_rdft(float *a128, int inverse);                 /* aec_rdft_forward_128 / inverse_128 */
void orc_aec_tables(float *w64, float *hann65, float *weight65, float *over65);
int orc_aec_far_available(const orc_aec *h);
int orc_aec_system_delay(const orc_aec *h);
int orc_aec_startup(const orc_aec *h);

#ifdef __cplusplus
}
#endif
#endif
