/* TEST INFRASTRUCTURE ONLY: array wrappers over the platform libm so that the numpy restatement
 * (oracle/euler.py) evaluates pow/sqrt/exp/cos with exactly the same routines as the compiled
 * reference (euler.cpp:86-94,122-123,144,157,211 call ::pow/::sqrt/::exp through tensor.h:118-140).
 * numpy's own SIMD pow/exp may differ from glibc in the last ulp. */
#include <math.h>
#include <stddef.h>

void nsem_or_pow(const double* x, double e, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = pow(x[i], e);
}
void nsem_or_sqrt(const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = sqrt(x[i]);
}
void nsem_or_exp(const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = exp(x[i]);
}
void nsem_or_cos(const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = cos(x[i]);
}
void nsem_or_sin(const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = sin(x[i]);
}
void nsem_or_acos(const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = acos(x[i]);
}
void nsem_or_atan2(const double* y, const double* x, double* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = atan2(y[i], x[i]);
}
