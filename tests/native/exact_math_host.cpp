// Host build of turbo_metrics_b200/csrc/exact_math.cuh for tests/test_exact_math.py (test-only).
// Evaluates the restated glibc cbrtf / powf over arrays so Python can compare them bit for bit
// with the libm of this machine.
#include "../../turbo_metrics_b200/csrc/exact_math.cuh"
#include <cstddef>

static const exact_math::PowfTables kTables = {{EM_POWF_LOG2_TAB}, {EM_EXP2F_TAB}};
static const exact_math::Consts kK = EM_CONSTS_INIT;

extern "C" {
void em_cbrtf_array(const float* in, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = exact_math::cbrtf_glibc(in[i], kK);
}
void em_powf_array(const float* in, float y, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = exact_math::powf_glibc(in[i], y, kK, kTables);
}
void libm_cbrtf_array(const float* in, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = cbrtf(in[i]);
}
void libm_powf_array(const float* in, float y, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) out[i] = powf(in[i], y);
}
}
