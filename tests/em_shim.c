/* test shim: the HOST build of csrc/exact_math.h next to the system libm, vectorised for tests/test_exact_math.py */
#include <math.h>
#include "../radiosonde_auto_rx_b200/csrc/exact_math.h"
void em_atan2f_vec(const float *y, const float *x, float *o, long n) { for (long i = 0; i < n; ++i) o[i] = em_atan2f(y[i], x[i]); }
void libm_atan2f_vec(const float *y, const float *x, float *o, long n) { for (long i = 0; i < n; ++i) o[i] = atan2f(y[i], x[i]); }
