/*
 * ssimu2_b200_debug.h -- test / measurement hooks of libssimu2_b200.so.  NOT part of the drop-in boundary
 * (include/ssimu2_b200.h): the parity tests read intermediate planes and run the device arithmetic helpers through
 * these, bench.py reads the per-kernel CUDA-event times.
 */
#ifndef SSIMU2_B200_DEBUG_H
#define SSIMU2_B200_DEBUG_H

#include "ssimu2_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Copy an intermediate plane set of the batch slot that served `ticket` to host memory.
 * what = 0: XYB planes of `scale`:                float[2][3][h][w]  (ref X,Y,B then dis X,Y,B)
 * what = 1: H-pass output of `scale`:            float[15][h][w]    (s11,s22,s12,mu1,mu2) x 3 channels
 *           (SSIMU2_PIPELINE_SPLIT only: the default pipeline never materialises these planes)
 * Only valid until that slot is reused (i.e. right after ssimu2_get_score). */
int ssimu2_debug_read(ssimu2_t *h, uint64_t ticket, int what, int scale, float *out, size_t out_floats);
/* Run the device arithmetic helpers over an array (host pointers) so the tests can compare them bit for
 * bit with libm / IEEE division:  op 0: out[i] = cbrtf(in[i]);  op 1: powf(in[i], y);  op 2: in[i] / y (f32);
 * op 3: `in` holds n (num, den) pairs of DOUBLES, `out` n doubles: num / den;
 * op 4: `in` holds n (num, den) pairs of floats: the V-pass quotient;
 * op 5 / 6: the unchecked hot-path forms of op 0 / 1 (positive normal arguments only). */
int ssimu2_debug_math(int op, const float *in, float y, float *out, size_t n);
/* Device time (ms) of the last completed batch per kernel: front-end, hpass (fused: k_hv), vpass, finalize. */
int ssimu2_last_batch_ms(ssimu2_t *h, float ms[4]);
/* Device time (ms, CUDA events on the batch's own stream) summed per kernel over every batch completed so far:
 * front-end, hpass (fused: k_hv), vpass, finalize; optional counters; reset != 0 clears them. */
int ssimu2_kernel_ms(ssimu2_t *h, double ms_total[4], uint64_t *batches, uint64_t *pairs, int reset);

#ifdef __cplusplus
}
#endif
#endif /* SSIMU2_B200_DEBUG_H */
