// complex128 arithmetic on double2 (re = x, im = y) - device helpers.
#pragma once
#include <cuda_runtime.h>

namespace fdfd {

__device__ __forceinline__ double2 c_make(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ double2 c_zero() { return make_double2(0.0, 0.0); }
__device__ __forceinline__ double2 c_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 c_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 c_neg(double2 a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ double2 c_mul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + a*b
__device__ __forceinline__ double2 c_fma(double2 a, double2 b, double2 acc) {
    double re = fma(a.x, b.x, acc.x);
    re = fma(-a.y, b.y, re);
    double im = fma(a.x, b.y, acc.y);
    im = fma(a.y, b.x, im);
    return make_double2(re, im);
}
// acc - a*b
__device__ __forceinline__ double2 c_fms(double2 a, double2 b, double2 acc) {
    double re = fma(-a.x, b.x, acc.x);
    re = fma(a.y, b.y, re);
    double im = fma(-a.x, b.y, acc.y);
    im = fma(-a.y, b.x, im);
    return make_double2(re, im);
}
__device__ __forceinline__ double2 c_conj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double2 c_scale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
// a / b (Smith-free straightforward form; operands are O(1) scaled Krylov scalars)
__device__ __forceinline__ double2 c_div(double2 a, double2 b) {
    const double d = b.x * b.x + b.y * b.y;
    return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
}
__device__ __forceinline__ double2 ldg2(const double2 *p) { return __ldg(p); }

}  // namespace fdfd
