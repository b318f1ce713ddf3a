// Host build of the device accumulator (csrc/length_acc.cuh) so the CPU suite
// can check it sample by sample against the oracle.  Test-only.
#include "length_acc.cuh"
extern "C" void sigk_host_length_acc(const uint32_t *x, uint64_t n, uint32_t *mean, uint32_t *median, uint32_t *var,
                                     double *median_f64, double *var_f64) {
    sigk::LengthAcc a;
    for (uint64_t i = 0; i < n; ++i) a.push(x[i]);
    *mean = n ? a.S / a.n : 0;
    *median = sigk::u16_from_double(a.q2);
    *var = sigk::u16_from_double(a.var);
    *median_f64 = a.q2;
    *var_f64 = a.var;
}
extern "C" uint32_t sigk_host_u16(double d) { return sigk::u16_from_double(d); }
extern "C" int sigk_host_symbol(unsigned c) { return sigk_symbol(c); }
