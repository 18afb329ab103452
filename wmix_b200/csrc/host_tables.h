// Host-side init-time tables (see host_tables.cpp).
#pragma once
#include <stdint.h>

namespace wmx {
namespace nsx {
struct Tables;
}
namespace host {
// fixed-point suppressor: every constant table of T:.../ns/nsx_core.c, nsx_core_c.c and the SPL twiddles, for one rate
// (8000, or 16000 / 32000 which share the 256-point geometry) and policy 0..3; *thr_lrt = the initial LRT threshold
int nsx_tables(int freq, int policy, nsx::Tables* T, int32_t* thr_lrt);
void ns_window(int ana, int block, float* w);
void fft_w_table(int nw, float* w);
void fft_c_table(int nc, float* c);
void ns_log_table(int bins, float* log_i, float* sum, float* sum_sq);
void dmath_tables(double* invc, double* logc, double* exp2jn);
int ns_policy(int mode, float* overdrive, float* floor_gain, int* gainmap);
int vad_thresholds(int mode, int frame_ms, int16_t out[4]);
void vad_initial_words(int32_t* words);
int16_t agc_analog_target(int16_t comp_db);
int agc_gain_table(int32_t table[32], int16_t comp_db, int16_t target_dbfs, int limiter, int16_t analog_target);
void agc_initial_words(int32_t* words);
// AEC: rdft-128 twiddles (w[32], c[32]), sqrt-Hann window, NLP weight / overdrive curves (65 each) and
// the k-step jump constants of the comfort-noise generator (65 each)
void aec_tables(float* w, float* c, float* hann, float* weight, float* over, uint32_t* lcg_mul, uint32_t* lcg_add);
// wmix_pcm_zoom's sample routing (R:src/wmix.c:139-222): map[k] = input sample feeding output sample k.
// Returns the output length in samples; map may be nullptr to only count.
uint32_t zoom_map(int in_chn, int in_freq, uint32_t in_bytes, int out_chn, int out_freq, int32_t* map);
uint32_t zoom_len_of_out(int in_chn, int in_freq, uint32_t in_len, int out_chn, int out_freq);
uint32_t zoom_len_of_in(int in_chn, int in_freq, int out_chn, int out_freq, uint32_t out_len);
// resampling branches of wmix_load_data, mono bus (R:src/wmix.c:1704-1939): per bus sample the source index and the
// ramp code (0 = copied frame, else (n << 8) | (k + 1)).  Returns the bus samples written (map/ramp may be nullptr to
// only count) or UINT32_MAX if the reference's 64-entry ramp buffer would overflow.
uint32_t mix_plan(int src_chn, int src_freq, uint32_t src_bytes, int mix_freq, int32_t* map, uint16_t* ramp);
// slot playPkgBuff_get(delay) reads (R:src/wmix.c:496-509), delay in whole packages, count = next write index
int play_fifo_slot(int count, int n_pkg, int delay_pkgs);
}  // namespace host
}  // namespace wmx
