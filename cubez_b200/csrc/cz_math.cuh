// cz_math.cuh — device-side value math of libcubezcuda (sm_100a).
//
// Only the operations the per-step pipeline reaches (SURVEY §8a "L0 — math ops").  Every
// n-term expression is written in the reference's order and the library is compiled with
// -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so each operation rounds once
// exactly as Go's float64/float32 arithmetic does on amd64 (no FMA contraction).
// Citations are file:line in the reference (tbogdala/cubez).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/cubezcuda.h"

typedef cz_real real;
#ifdef CUBEZ_REAL_FLOAT
typedef float2 real2;
#define CZ_REAL_MAX 3.402823466e+38f
__host__ __device__ __forceinline__ real2 make_real2(real a, real b) { return make_float2(a, b); }
#else
typedef double2 real2;
#define CZ_REAL_MAX 1.7976931348623157e+308
__host__ __device__ __forceinline__ real2 make_real2(real a, real b) { return make_double2(a, b); }
#endif

// CZD functions are also compiled for the host so that tests/hostemu can run the very same
// arithmetic on the CPU (test infrastructure only; the product library never calls them on
// the host).
#define CZD __host__ __device__ __forceinline__
// Out-of-line on the device: every inlined IEEE division / square root is ~20 SASS instructions
// plus a slow-path stub, and the fused world kernel is instruction-cache bound (see profiles/).
#ifdef __CUDA_ARCH__
#define CZD_OUTLINE __device__ __noinline__
#else
#define CZD_OUTLINE __host__ __device__ __forceinline__
#endif
#define R_(x) ((real)(x))

namespace czm {

// math/math.go:27-33
#define CZ_EPSILON R_(1e-7)
#define CZ_MIN_NORMAL R_(1.1754943508222875e-38)

// math/math.go:91-98: RealAbs / RealSqrt go through float64 and round back; for IEEE types
// that is the same value as the native op (sqrt: 53 >= 2*24+2 bits).
CZD real rabs(real a) { return (real)fabs((double)a); }
CZD_OUTLINE real rsqrt_(real a) {
#ifdef __CUDA_ARCH__
#ifdef CUBEZ_REAL_FLOAT
    return __fsqrt_rn(a);
#else
    return __dsqrt_rn(a);
#endif
#else
    return (real)sqrt((double)a);
#endif
}
CZD_OUTLINE real rdiv(real a, real b) {
#ifdef __CUDA_ARCH__
#ifdef CUBEZ_REAL_FLOAT
    return __fdiv_rn(a, b);
#else
    return __ddiv_rn(a, b);
#endif
#else
    return a / b;
#endif
}
CZD real real_inf() {
#ifdef CUBEZ_REAL_FLOAT
    return (real)INFINITY;
#else
    return (real)INFINITY;
#endif
}

// math/math.go:64-78
CZD_OUTLINE bool real_equal(real a, real b) {
    if (a == b) return true;
    real diff = rabs(a - b);
    if (a * b == R_(0) || diff < CZ_MIN_NORMAL) {
        const real e = CZ_EPSILON;
        return diff < e * e;
    }
    return rdiv(diff, (real)(fabs((double)a) + fabs((double)b))) < CZ_EPSILON;
}

struct V3 { real c[3]; };
struct Q4 { real c[4]; };
struct M3 { real c[9]; };
struct M34 { real c[12]; };

CZD V3 mk3(real x, real y, real z) { V3 v; v.c[0] = x; v.c[1] = y; v.c[2] = z; return v; }
CZD V3 zero3() { return mk3(R_(0), R_(0), R_(0)); }

// math/vector.go
CZD void v_add(V3 &v, const V3 &o) { v.c[0] += o.c[0]; v.c[1] += o.c[1]; v.c[2] += o.c[2]; }                          // :7
CZD void v_add_scaled(V3 &v, const V3 &o, real s) { v.c[0] += o.c[0] * s; v.c[1] += o.c[1] * s; v.c[2] += o.c[2] * s; } // :14
CZD void v_component_product(V3 &v, const V3 &o) { v.c[0] *= o.c[0]; v.c[1] *= o.c[1]; v.c[2] *= o.c[2]; }             // :26
CZD V3 v_cross(const V3 &v, const V3 &o) {                                                                               // :33
    return mk3(v.c[1] * o.c[2] - v.c[2] * o.c[1], v.c[2] * o.c[0] - v.c[0] * o.c[2], v.c[0] * o.c[1] - v.c[1] * o.c[0]);
}
CZD real v_dot(const V3 &v, const V3 &o) { return v.c[0] * o.c[0] + v.c[1] * o.c[1] + v.c[2] * o.c[2]; }                // :42
CZD real v_sqmag(const V3 &v) { return v.c[0] * v.c[0] + v.c[1] * v.c[1] + v.c[2] * v.c[2]; }                            // :52
CZD real v_mag(const V3 &v) { return rsqrt_(v.c[0] * v.c[0] + v.c[1] * v.c[1] + v.c[2] * v.c[2]); }                      // :47
CZD void v_mul(V3 &v, real r) { v.c[0] *= r; v.c[1] *= r; v.c[2] *= r; }                                                 // :57
CZD void v_normalize(V3 &v) {                                                                                            // :64
    real m = v_mag(v);
    if (!real_equal(m, R_(0))) {
        real l = rdiv(R_(1), m);
        v.c[0] *= l; v.c[1] *= l; v.c[2] *= l;
    }
}
CZD void v_sub(V3 &v, const V3 &o) { v.c[0] -= o.c[0]; v.c[1] -= o.c[1]; v.c[2] -= o.c[2]; }                            // :82

// math/quaternion.go
CZD void q_mul(Q4 &q, const Q4 &p) {                                                                                     // :46
    real w = q.c[0] * p.c[0] - q.c[1] * p.c[1] - q.c[2] * p.c[2] - q.c[3] * p.c[3];
    real x = q.c[0] * p.c[1] + q.c[1] * p.c[0] + q.c[2] * p.c[3] - q.c[3] * p.c[2];
    real y = q.c[0] * p.c[2] + q.c[2] * p.c[0] + q.c[3] * p.c[1] - q.c[1] * p.c[3];
    real z = q.c[0] * p.c[3] + q.c[3] * p.c[0] + q.c[1] * p.c[2] - q.c[2] * p.c[1];
    q.c[0] = w; q.c[1] = x; q.c[2] = y; q.c[3] = z;
}
CZD void q_add_scaled_vector(Q4 &q, const V3 &v, real scale) {                                                           // :20
    Q4 t;
    t.c[0] = R_(0); t.c[1] = v.c[0] * scale; t.c[2] = v.c[1] * scale; t.c[3] = v.c[2] * scale;
    q_mul(t, q);
    q.c[0] += t.c[0] * R_(0.5); q.c[1] += t.c[1] * R_(0.5); q.c[2] += t.c[2] * R_(0.5); q.c[3] += t.c[3] * R_(0.5);
}
CZD real q_len(const Q4 &q) { return rsqrt_(q.c[0] * q.c[0] + q.c[1] * q.c[1] + q.c[2] * q.c[2] + q.c[3] * q.c[3]); }    // :41
CZD void q_normalize(Q4 &q) {                                                                                            // :78
    real length = q_len(q);
    if (real_equal(R_(1), length)) return;
    if (length == R_(0)) { q.c[0] = R_(1); q.c[1] = R_(0); q.c[2] = R_(0); q.c[3] = R_(0); return; }
    if (length == real_inf()) length = CZ_REAL_MAX;
    real inv = rdiv(R_(1), length);
    q.c[0] *= inv; q.c[1] *= inv; q.c[2] *= inv; q.c[3] *= inv;
}
CZD V3 q_rotate(const Q4 &q, const V3 &v) {                                                                              // :56 (tests only)
    V3 qv = mk3(q.c[1], q.c[2], q.c[3]);
    V3 cr = v_cross(qv, v);
    V3 res = v;
    v_mul(qv, R_(2));
    V3 c2 = v_cross(qv, cr);
    v_add(res, c2);
    v_mul(cr, R_(2) * q.c[0]);
    v_add(res, cr);
    return res;
}

// math/matrix.go (column-major)
CZD V3 m3_mul_v(const M3 &m, const V3 &v) {                                                                              // :80
    return mk3(m.c[0] * v.c[0] + m.c[3] * v.c[1] + m.c[6] * v.c[2], m.c[1] * v.c[0] + m.c[4] * v.c[1] + m.c[7] * v.c[2],
               m.c[2] * v.c[0] + m.c[5] * v.c[1] + m.c[8] * v.c[2]);
}
CZD M3 m3_mul_m(const M3 &a, const M3 &b) {                                                                              // :89
    M3 r;
#pragma unroll
    for (int col = 0; col < 3; col++) {
#pragma unroll
        for (int row = 0; row < 3; row++)
            r.c[col * 3 + row] = a.c[row] * b.c[col * 3] + a.c[3 + row] * b.c[col * 3 + 1] + a.c[6 + row] * b.c[col * 3 + 2];
    }
    return r;
}
CZD M3 m3_transpose(const M3 &m) {                                                                                       // :117
    M3 r;
    r.c[0] = m.c[0]; r.c[1] = m.c[3]; r.c[2] = m.c[6]; r.c[3] = m.c[1]; r.c[4] = m.c[4]; r.c[5] = m.c[7];
    r.c[6] = m.c[2]; r.c[7] = m.c[5]; r.c[8] = m.c[8];
    return r;
}
CZD real m3_det(const M3 &m) {                                                                                           // :127
    return m.c[0] * m.c[4] * m.c[8] + m.c[3] * m.c[7] * m.c[2] + m.c[6] * m.c[1] * m.c[5] - m.c[6] * m.c[4] * m.c[2] -
           m.c[3] * m.c[1] * m.c[8] - m.c[0] * m.c[7] * m.c[5];
}
CZD M3 m3_invert(const M3 &m) {                                                                                          // :133
    M3 r;
    real det = m3_det(m);
    if (real_equal(det, R_(0))) {
#pragma unroll
        for (int i = 0; i < 9; i++) r.c[i] = R_(0);
        return r;
    }
    r.c[0] = m.c[4] * m.c[8] - m.c[5] * m.c[7];
    r.c[1] = m.c[2] * m.c[7] - m.c[1] * m.c[8];
    r.c[2] = m.c[1] * m.c[5] - m.c[2] * m.c[4];
    r.c[3] = m.c[5] * m.c[6] - m.c[3] * m.c[8];
    r.c[4] = m.c[0] * m.c[8] - m.c[2] * m.c[6];
    r.c[5] = m.c[2] * m.c[3] - m.c[0] * m.c[5];
    r.c[6] = m.c[3] * m.c[7] - m.c[4] * m.c[6];
    r.c[7] = m.c[1] * m.c[6] - m.c[0] * m.c[7];
    r.c[8] = m.c[0] * m.c[4] - m.c[1] * m.c[3];
    real s = rdiv(R_(1), det);
#pragma unroll
    for (int i = 0; i < 9; i++) r.c[i] *= s;
    return r;
}
CZD V3 m3_transform_transpose(const M3 &m, const V3 &v) {                                                                // :157
    return mk3(v.c[0] * m.c[0] + v.c[1] * m.c[1] + v.c[2] * m.c[2], v.c[0] * m.c[3] + v.c[1] * m.c[4] + v.c[2] * m.c[5],
               v.c[0] * m.c[6] + v.c[1] * m.c[7] + v.c[2] * m.c[8]);
}
CZD void m34_set_as_transform(M34 &m, const V3 &pos, const Q4 &rot) {                                                    // :167
    real w = rot.c[0], x = rot.c[1], y = rot.c[2], z = rot.c[3];
    m.c[0] = R_(1) - R_(2) * y * y - R_(2) * z * z;
    m.c[1] = R_(2) * x * y + R_(2) * w * z;
    m.c[2] = R_(2) * x * z - R_(2) * w * y;
    m.c[3] = R_(2) * x * y - R_(2) * w * z;
    m.c[4] = R_(1) - R_(2) * x * x - R_(2) * z * z;
    m.c[5] = R_(2) * y * z + R_(2) * w * x;
    m.c[6] = R_(2) * x * z + R_(2) * w * y;
    m.c[7] = R_(2) * y * z - R_(2) * w * x;
    m.c[8] = R_(1) - R_(2) * x * x - R_(2) * y * y;
    m.c[9] = pos.c[0]; m.c[10] = pos.c[1]; m.c[11] = pos.c[2];
}
CZD V3 m34_mul_v(const M34 &m, const V3 &v) {                                                                            // :188
    return mk3(v.c[0] * m.c[0] + v.c[1] * m.c[3] + v.c[2] * m.c[6] + m.c[9], v.c[0] * m.c[1] + v.c[1] * m.c[4] + v.c[2] * m.c[7] + m.c[10],
               v.c[0] * m.c[2] + v.c[1] * m.c[5] + v.c[2] * m.c[8] + m.c[11]);
}
CZD M34 m34_mul_m34(const M34 &m, const M34 &o) {                                                                        // :198
    M34 r;
#pragma unroll
    for (int col = 0; col < 3; col++) {
#pragma unroll
        for (int row = 0; row < 3; row++)
            r.c[col * 3 + row] = m.c[row] * o.c[col * 3] + m.c[3 + row] * o.c[col * 3 + 1] + m.c[6 + row] * o.c[col * 3 + 2];
    }
#pragma unroll
    for (int row = 0; row < 3; row++) r.c[9 + row] = m.c[row] * o.c[9] + m.c[3 + row] * o.c[10] + m.c[6 + row] * o.c[11] + m.c[9 + row];
    return r;
}
CZD V3 m34_transform_inverse(const M34 &m, const V3 &v) {                                                                // :222
    real t0 = v.c[0] - m.c[9], t1 = v.c[1] - m.c[10], t2 = v.c[2] - m.c[11];
    return mk3(t0 * m.c[0] + t1 * m.c[1] + t2 * m.c[2], t0 * m.c[3] + t1 * m.c[4] + t2 * m.c[5], t0 * m.c[6] + t1 * m.c[7] + t2 * m.c[8]);
}
CZD V3 m34_axis(const M34 &m, int i) { return mk3(m.c[i * 3 + 0], m.c[i * 3 + 1], m.c[i * 3 + 2]); }                     // :235

// rigidbody.go:275-299
CZD void transform_inertia_tensor(M3 &w, const M3 &b, const M34 &r) {
    real t4 = r.c[0] * b.c[0] + r.c[3] * b.c[1] + r.c[6] * b.c[2];
    real t9 = r.c[0] * b.c[3] + r.c[3] * b.c[4] + r.c[6] * b.c[5];
    real t14 = r.c[0] * b.c[6] + r.c[3] * b.c[7] + r.c[6] * b.c[8];
    real t28 = r.c[1] * b.c[0] + r.c[4] * b.c[1] + r.c[7] * b.c[2];
    real t33 = r.c[1] * b.c[3] + r.c[4] * b.c[4] + r.c[7] * b.c[5];
    real t38 = r.c[1] * b.c[6] + r.c[4] * b.c[7] + r.c[7] * b.c[8];
    real t52 = r.c[2] * b.c[0] + r.c[5] * b.c[1] + r.c[8] * b.c[2];
    real t57 = r.c[2] * b.c[3] + r.c[5] * b.c[4] + r.c[8] * b.c[5];
    real t62 = r.c[2] * b.c[6] + r.c[5] * b.c[7] + r.c[8] * b.c[8];
    w.c[0] = t4 * r.c[0] + t9 * r.c[3] + t14 * r.c[6];
    w.c[3] = t4 * r.c[1] + t9 * r.c[4] + t14 * r.c[7];
    w.c[6] = t4 * r.c[2] + t9 * r.c[5] + t14 * r.c[8];
    w.c[1] = t28 * r.c[0] + t33 * r.c[3] + t38 * r.c[6];
    w.c[4] = t28 * r.c[1] + t33 * r.c[4] + t38 * r.c[7];
    w.c[7] = t28 * r.c[2] + t33 * r.c[5] + t38 * r.c[8];
    w.c[2] = t52 * r.c[0] + t57 * r.c[3] + t62 * r.c[6];
    w.c[5] = t52 * r.c[1] + t57 * r.c[4] + t62 * r.c[7];
    w.c[8] = t52 * r.c[2] + t57 * r.c[5] + t62 * r.c[8];
}

// contact.go:612-616
CZD void set_skew(M3 &m, const V3 &v) {
    m.c[0] = R_(0); m.c[3] = -v.c[2]; m.c[6] = v.c[1];
    m.c[1] = v.c[2]; m.c[4] = R_(0); m.c[7] = -v.c[0];
    m.c[2] = -v.c[1]; m.c[5] = v.c[0]; m.c[8] = R_(0);
}

}  // namespace czm
