// Host-side init-time tables (see host_tables.cpp).
#pragma once
#include <stdint.h>

namespace wmx {
namespace host {
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
}  // namespace host
}  // namespace wmx
