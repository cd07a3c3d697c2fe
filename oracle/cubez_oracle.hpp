// cubez_oracle.hpp — CPU restatement of the cubez per-step rigid-body pipeline.
//
// *** TEST INFRASTRUCTURE ONLY. ***  Nothing under cubez_b200/ (the product) may include,
// link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
//
// *** PARITY STATUS: PINNED AGAINST THE REFERENCE'S OWN SOURCE, RUN HERE ***
// The reference (tbogdala/cubez) is Go and no Go toolchain exists in this image.  Its sources are therefore run
// through oracle/go2cpp.py, a syntax-directed Go -> C++ translator (it knows Go, not physics; see its header), together
// with the headless harness mains of go/harness/ — oracle/Makefile target `ref`, outputs in oracle/_ref/ only.
//   * math layer: the reference's own 22 known-answer tests (math/vector_test.go, quaternion_test.go, matrix_test.go)
//     pass on the translation, and are ported as KATs for this restatement and for the device math
//     (tests/test_oracle_math_kat.py, tests/test_gpu_math_kat.py);
//   * Integrate, every narrowphase routine, the resolver, the frame loops: the reference holds no test or vector for
//     them, so the pins are dumps PRINTED BY THE TRANSLATED REFERENCE (tests/golden/ref/*.txt, made by
//     oracle/make_ref_golden.py): this restatement reproduces them bit for bit — per frame the contact count, the
//     (body, body) sequence, the as-generated contact geometry and the raw bits of every body's state — on
//     cubedrop (600 frames), ballistic-64 (600), piles of 27 / 216 bodies, 256 batched perturbed worlds (600),
//     65 536 free bodies and five runs of a fuzz scene (random shapes, sizes, masses, collider Offsets, several planes,
//     host-painted surface materials) (tests/test_oracle_vs_reference_dump.py); the 4 096-body pile (80 frames) is
//     checked from the GPU side;
//   * float32: the reference's own switch is the one-line edit `type Real float64` -> float32 (math/math.go:23);
//     go2cpp.py --real=float32 applies it to the text it reads and gives untyped constants the type of the operand they
//     meet (rounded once from the exact value, as Go does).  The float32 instantiation of this restatement reproduces
//     the seven float32 dumps bit for bit as well.
// What is not covered: the Go compiler itself (the harnesses are ready for it: tools/compare_go_dump.py) and math.Pow,
// which is C pow() here and in the translation (<= 1 ulp from Go's; the library takes the Pow factors as host inputs).
// The restatement keeps the Go source's expression order (left to right, one IEEE rounding per operation, no FMA
// contraction: build with -ffp-contract=off).
//
// Every function cites the reference file:line (paths relative to /root/reference) it
// restates.  The structures keep the reference's AoS layout and the per-contact heap
// allocation on purpose, so that timing this code is a fair stand-in for the Go loops
// (BASELINE.md "cpu-restatement").
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace czo {

// ---------------------------------------------------------------------------------------
// math/math.go
// ---------------------------------------------------------------------------------------
template <class R> struct Lim {
    // math/math.go:27-33.  MaxValue is math.MaxFloat64 in the reference; the float32
    // build needs MaxFloat32 there.
    static constexpr R epsilon = (R)1e-7;
    static constexpr R min_normal = (R)1.1754943508222875e-38;
    static R max_value() { return std::numeric_limits<R>::max(); }
    static R inf_pos() { return std::numeric_limits<R>::infinity(); }
};

// math/math.go:91-98 — the wrappers go through float64 and round back to Real.
template <class R> static inline R rabs(R a) { return (R)std::fabs((double)a); }
template <class R> static inline R rsqrt_(R a) { return (R)std::sqrt((double)a); }

// math/math.go:64-78
template <class R> static inline bool real_equal(R a, R b) {
    if (a == b) return true;
    R diff = (R)std::fabs((double)(R)(a - b));
    if (a * b == 0 || diff < Lim<R>::min_normal) {
        const R e = Lim<R>::epsilon;
        return diff < e * e;
    }
    return diff / (R)(std::fabs((double)a) + std::fabs((double)b)) < Lim<R>::epsilon;
}

template <class R> struct V3 {
    R c[3];
    R &operator[](int i) { return c[i]; }
    const R &operator[](int i) const { return c[i]; }
};
template <class R> struct Q4 { R c[4]; R &operator[](int i) { return c[i]; } const R &operator[](int i) const { return c[i]; } };
template <class R> struct M3 { R c[9]; R &operator[](int i) { return c[i]; } const R &operator[](int i) const { return c[i]; } };
template <class R> struct M34 { R c[12]; R &operator[](int i) { return c[i]; } const R &operator[](int i) const { return c[i]; } };

// ---------------------------------------------------------------------------------------
// math/vector.go
// ---------------------------------------------------------------------------------------
template <class R> static inline void v_add(V3<R> &v, const V3<R> &o) { v[0] += o[0]; v[1] += o[1]; v[2] += o[2]; }            // :7
template <class R> static inline void v_add_scaled(V3<R> &v, const V3<R> &o, R s) { v[0] += o[0] * s; v[1] += o[1] * s; v[2] += o[2] * s; } // :14
template <class R> static inline void v_clear(V3<R> &v) { v[0] = 0; v[1] = 0; v[2] = 0; }                                        // :21
template <class R> static inline void v_component_product(V3<R> &v, const V3<R> &o) { v[0] *= o[0]; v[1] *= o[1]; v[2] *= o[2]; } // :26
template <class R> static inline V3<R> v_cross(const V3<R> &v, const V3<R> &o) {                                                  // :33
    return V3<R>{{v[1] * o[2] - v[2] * o[1], v[2] * o[0] - v[0] * o[2], v[0] * o[1] - v[1] * o[0]}};
}
template <class R> static inline R v_dot(const V3<R> &v, const V3<R> &o) { return v[0] * o[0] + v[1] * o[1] + v[2] * o[2]; }      // :42
template <class R> static inline R v_sqmag(const V3<R> &v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }                     // :52
template <class R> static inline R v_mag(const V3<R> &v) { return rsqrt_<R>(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }            // :47
template <class R> static inline void v_mul(V3<R> &v, R r) { v[0] *= r; v[1] *= r; v[2] *= r; }                                    // :57
template <class R> static inline void v_normalize(V3<R> &v) {                                                                      // :64
    R m = v_mag(v);
    if (!real_equal<R>(m, (R)0)) {
        R l = (R)1 / m;
        v[0] *= l; v[1] *= l; v[2] *= l;
    }
}
template <class R> static inline void v_sub(V3<R> &v, const V3<R> &o) { v[0] -= o[0]; v[1] -= o[1]; v[2] -= o[2]; }               // :82

// ---------------------------------------------------------------------------------------
// math/quaternion.go
// ---------------------------------------------------------------------------------------
template <class R> static inline void q_identity(Q4<R> &q) { q[0] = 1; q[1] = 0; q[2] = 0; q[3] = 0; }                             // :36
template <class R> static inline R q_len(const Q4<R> &q) { return rsqrt_<R>(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]); } // :41
template <class R> static inline void q_mul(Q4<R> &q, const Q4<R> &p) {                                                            // :46
    R w = q[0] * p[0] - q[1] * p[1] - q[2] * p[2] - q[3] * p[3];
    R x = q[0] * p[1] + q[1] * p[0] + q[2] * p[3] - q[3] * p[2];
    R y = q[0] * p[2] + q[2] * p[0] + q[3] * p[1] - q[1] * p[3];
    R z = q[0] * p[3] + q[3] * p[0] + q[1] * p[2] - q[2] * p[1];
    q[0] = w; q[1] = x; q[2] = y; q[3] = z;
}
template <class R> static inline void q_add_scaled_vector(Q4<R> &q, const V3<R> &v, R scale) {                                     // :20
    Q4<R> t;
    t[0] = 0; t[1] = v[0] * scale; t[2] = v[1] * scale; t[3] = v[2] * scale;
    q_mul(t, q);
    q[0] += t[0] * (R)0.5; q[1] += t[1] * (R)0.5; q[2] += t[2] * (R)0.5; q[3] += t[3] * (R)0.5;
}
template <class R> static inline void q_normalize(Q4<R> &q) {                                                                      // :78
    R length = q_len(q);
    if (real_equal<R>((R)1, length)) return;
    if (length == 0) { q_identity(q); return; }
    if (length == Lim<R>::inf_pos()) length = Lim<R>::max_value();
    R inv = (R)1 / length;
    q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
// math/quaternion.go:56-70 (Rotate) — not on the hot path; kept for the reference's own
// known-answer tests (math/quaternion_test.go:86-181).
template <class R> static inline V3<R> q_rotate(const Q4<R> &q, const V3<R> &v) {
    V3<R> qv{{q[1], q[2], q[3]}};
    V3<R> cr = v_cross(qv, v);
    V3<R> res = v;
    v_mul(qv, (R)2);
    V3<R> c2 = v_cross(qv, cr);
    v_add(res, c2);
    v_mul(cr, (R)2 * q[0]);
    v_add(res, cr);
    return res;
}

// ---------------------------------------------------------------------------------------
// math/matrix.go  (column-major; chart at :6-12)
// ---------------------------------------------------------------------------------------
template <class R> static inline void m3_identity(M3<R> &m) { m[0] = 1; m[3] = 0; m[6] = 0; m[1] = 0; m[4] = 1; m[7] = 0; m[2] = 0; m[5] = 0; m[8] = 1; } // :15
template <class R> static inline void m34_identity(M34<R> &m) {                                                                    // :22
    m[0] = 1; m[3] = 0; m[6] = 0; m[9] = 0; m[1] = 0; m[4] = 1; m[7] = 0; m[10] = 0; m[2] = 0; m[5] = 0; m[8] = 1; m[11] = 0;
}
template <class R> static inline void m3_add(M3<R> &m, const M3<R> &o) { for (int i = 0; i < 9; i++) m[i] += o[i]; }                // :37
template <class R> static inline void m3_set_components(M3<R> &m, const V3<R> &a, const V3<R> &b, const V3<R> &c) {                 // :52
    m[0] = a[0]; m[3] = b[0]; m[6] = c[0]; m[1] = a[1]; m[4] = b[1]; m[7] = c[1]; m[2] = a[2]; m[5] = b[2]; m[8] = c[2];
}
template <class R> static inline void m3_set_inertia_coeffs(M3<R> &m, R ix, R iy, R iz, R ixy, R ixz, R iyz) {                      // :59
    m[0] = ix; m[3] = -ixy; m[6] = -ixz; m[1] = -ixy; m[4] = iy; m[7] = -iyz; m[2] = -ixz; m[5] = -iyz; m[8] = iz;
}
template <class R> static inline void m3_set_block_inertia(M3<R> &m, const V3<R> &half, R mass) {                                  // :68
    V3<R> sq = half;
    v_component_product(sq, half);
    m3_set_inertia_coeffs<R>(m, (R)0.3 * mass * (sq[1] + sq[2]), (R)0.3 * mass * (sq[0] + sq[2]), (R)0.3 * mass * (sq[0] + sq[1]), 0, 0, 0);
}
template <class R> static inline V3<R> m3_mul_v(const M3<R> &m, const V3<R> &v) {                                                   // :80
    return V3<R>{{m[0] * v[0] + m[3] * v[1] + m[6] * v[2], m[1] * v[0] + m[4] * v[1] + m[7] * v[2], m[2] * v[0] + m[5] * v[1] + m[8] * v[2]}};
}
template <class R> static inline M3<R> m3_mul_m(const M3<R> &a, const M3<R> &b) {                                                   // :89
    return M3<R>{{a[0] * b[0] + a[3] * b[1] + a[6] * b[2], a[1] * b[0] + a[4] * b[1] + a[7] * b[2], a[2] * b[0] + a[5] * b[1] + a[8] * b[2],
                  a[0] * b[3] + a[3] * b[4] + a[6] * b[5], a[1] * b[3] + a[4] * b[4] + a[7] * b[5], a[2] * b[3] + a[5] * b[4] + a[8] * b[5],
                  a[0] * b[6] + a[3] * b[7] + a[6] * b[8], a[1] * b[6] + a[4] * b[7] + a[7] * b[8], a[2] * b[6] + a[5] * b[7] + a[8] * b[8]}};
}
template <class R> static inline void m3_mul_s(M3<R> &m, R s) { for (int i = 0; i < 9; i++) m[i] *= s; }                            // :104
template <class R> static inline M3<R> m3_transpose(const M3<R> &m) { return M3<R>{{m[0], m[3], m[6], m[1], m[4], m[7], m[2], m[5], m[8]}}; } // :117
template <class R> static inline R m3_det(const M3<R> &m) {                                                                         // :127
    return m[0] * m[4] * m[8] + m[3] * m[7] * m[2] + m[6] * m[1] * m[5] - m[6] * m[4] * m[2] - m[3] * m[1] * m[8] - m[0] * m[7] * m[5];
}
template <class R> static inline M3<R> m3_invert(const M3<R> &m) {                                                                  // :133
    R det = m3_det(m);
    if (real_equal<R>(det, (R)0)) return M3<R>{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
    M3<R> r{{m[4] * m[8] - m[5] * m[7], m[2] * m[7] - m[1] * m[8], m[1] * m[5] - m[2] * m[4],
             m[5] * m[6] - m[3] * m[8], m[0] * m[8] - m[2] * m[6], m[2] * m[3] - m[0] * m[5],
             m[3] * m[7] - m[4] * m[6], m[1] * m[6] - m[0] * m[7], m[0] * m[4] - m[1] * m[3]}};
    m3_mul_s<R>(r, (R)1 / det);
    return r;
}
template <class R> static inline V3<R> m3_transform_transpose(const M3<R> &m, const V3<R> &v) {                                     // :157
    return V3<R>{{v[0] * m[0] + v[1] * m[1] + v[2] * m[2], v[0] * m[3] + v[1] * m[4] + v[2] * m[5], v[0] * m[6] + v[1] * m[7] + v[2] * m[8]}};
}
template <class R> static inline void m34_set_as_transform(M34<R> &m, const V3<R> &pos, const Q4<R> &rot) {                         // :167
    R w = rot[0], x = rot[1], y = rot[2], z = rot[3];
    m[0] = 1 - 2 * y * y - 2 * z * z;
    m[1] = 2 * x * y + 2 * w * z;
    m[2] = 2 * x * z - 2 * w * y;
    m[3] = 2 * x * y - 2 * w * z;
    m[4] = 1 - 2 * x * x - 2 * z * z;
    m[5] = 2 * y * z + 2 * w * x;
    m[6] = 2 * x * z + 2 * w * y;
    m[7] = 2 * y * z - 2 * w * x;
    m[8] = 1 - 2 * x * x - 2 * y * y;
    m[9] = pos[0]; m[10] = pos[1]; m[11] = pos[2];
}
template <class R> static inline V3<R> m34_mul_v(const M34<R> &m, const V3<R> &v) {                                                 // :188
    return V3<R>{{v[0] * m[0] + v[1] * m[3] + v[2] * m[6] + m[9], v[0] * m[1] + v[1] * m[4] + v[2] * m[7] + m[10], v[0] * m[2] + v[1] * m[5] + v[2] * m[8] + m[11]}};
}
template <class R> static inline M34<R> m34_mul_m34(const M34<R> &m, const M34<R> &o) {                                             // :198
    return M34<R>{{m[0] * o[0] + m[3] * o[1] + m[6] * o[2], m[1] * o[0] + m[4] * o[1] + m[7] * o[2], m[2] * o[0] + m[5] * o[1] + m[8] * o[2],
                   m[0] * o[3] + m[3] * o[4] + m[6] * o[5], m[1] * o[3] + m[4] * o[4] + m[7] * o[5], m[2] * o[3] + m[5] * o[4] + m[8] * o[5],
                   m[0] * o[6] + m[3] * o[7] + m[6] * o[8], m[1] * o[6] + m[4] * o[7] + m[7] * o[8], m[2] * o[6] + m[5] * o[7] + m[8] * o[8],
                   m[0] * o[9] + m[3] * o[10] + m[6] * o[11] + m[9], m[1] * o[9] + m[4] * o[10] + m[7] * o[11] + m[10], m[2] * o[9] + m[5] * o[10] + m[8] * o[11] + m[11]}};
}
template <class R> static inline V3<R> m34_transform_inverse(const M34<R> &m, const V3<R> &v) {                                     // :222
    V3<R> t = v;
    t[0] -= m[9]; t[1] -= m[10]; t[2] -= m[11];
    return V3<R>{{t[0] * m[0] + t[1] * m[1] + t[2] * m[2], t[0] * m[3] + t[1] * m[4] + t[2] * m[5], t[0] * m[6] + t[1] * m[7] + t[2] * m[8]}};
}
template <class R> static inline V3<R> m34_axis(const M34<R> &m, int col) {                                                         // :235
    int i = col > 3 ? 3 : col;
    return V3<R>{{m[i * 3 + 0], m[i * 3 + 1], m[i * 3 + 2]}};
}

// ---------------------------------------------------------------------------------------
// rigidbody.go
// ---------------------------------------------------------------------------------------
template <class R> struct Body {   // rigidbody.go:23-101, same member order
    R linearDamping, angularDamping;
    V3<R> position;
    Q4<R> orientation;
    V3<R> velocity, acceleration, rotation;
    M3<R> inverseInertiaTensor;
    bool isAwake, canSleep;
    M3<R> iitWorld;
    R inverseMass, mass;
    M34<R> transform;
    V3<R> forceAccum, torqueAccum, lastFrameAcc;
    R motion;
};

template <class R> static inline void body_set_awake(Body<R> &b, bool awake) {   // rigidbody.go:182-192
    if (awake) {
        b.isAwake = true;
        b.motion = (R)0.6;   // sleepEpsilon*2.0 is folded by Go as the exact constant 0.6
    } else {
        b.isAwake = false;
        v_clear(b.velocity);
        v_clear(b.rotation);
    }
}
template <class R> static inline void body_init(Body<R> &b) {   // NewRigidBody, rigidbody.go:104-114
    std::memset(&b, 0, sizeof(b));
    q_identity(b.orientation);
    b.linearDamping = (R)0.95;
    b.angularDamping = (R)0.95;   // :108 assigns defaultLinearDamping
    b.acceleration = V3<R>{{(R)0.0, (R)-9.78, (R)0.0}};
    m3_identity(b.iitWorld);
    b.canSleep = true;
    body_set_awake(b, true);
}
template <class R> static inline void body_set_mass(Body<R> &b, R mass) { b.mass = mass; b.inverseMass = (R)1.0 / mass; }   // :124
template <class R> static inline void body_set_infinite_mass(Body<R> &b) { b.mass = 0; b.inverseMass = 0; }                 // :131
template <class R> static inline void body_set_inertia_tensor(Body<R> &b, const M3<R> &t) { b.inverseInertiaTensor = m3_invert(t); } // :176

// rigidbody.go:275-299
template <class R> static inline void transform_inertia_tensor(M3<R> &w, const M3<R> &b, const M34<R> &r) {
    R t4 = r[0] * b[0] + r[3] * b[1] + r[6] * b[2];
    R t9 = r[0] * b[3] + r[3] * b[4] + r[6] * b[5];
    R t14 = r[0] * b[6] + r[3] * b[7] + r[6] * b[8];
    R t28 = r[1] * b[0] + r[4] * b[1] + r[7] * b[2];
    R t33 = r[1] * b[3] + r[4] * b[4] + r[7] * b[5];
    R t38 = r[1] * b[6] + r[4] * b[7] + r[7] * b[8];
    R t52 = r[2] * b[0] + r[5] * b[1] + r[8] * b[2];
    R t57 = r[2] * b[3] + r[5] * b[4] + r[8] * b[5];
    R t62 = r[2] * b[6] + r[5] * b[7] + r[8] * b[8];
    w[0] = t4 * r[0] + t9 * r[3] + t14 * r[6];
    w[3] = t4 * r[1] + t9 * r[4] + t14 * r[7];
    w[6] = t4 * r[2] + t9 * r[5] + t14 * r[8];
    w[1] = t28 * r[0] + t33 * r[3] + t38 * r[6];
    w[4] = t28 * r[1] + t33 * r[4] + t38 * r[7];
    w[7] = t28 * r[2] + t33 * r[5] + t38 * r[8];
    w[2] = t52 * r[0] + t57 * r[3] + t62 * r[6];
    w[5] = t52 * r[1] + t57 * r[4] + t62 * r[7];
    w[8] = t52 * r[2] + t57 * r[5] + t62 * r[8];
}
template <class R> static inline void body_calculate_derived(Body<R> &b) {   // rigidbody.go:268-272
    q_normalize(b.orientation);
    m34_set_as_transform(b.transform, b.position, b.orientation);
    transform_inertia_tensor(b.iitWorld, b.inverseInertiaTensor, b.transform);
}
// rigidbody.go:213-259.  The three math.Pow results are taken as arguments (computed by the
// caller in float64 and rounded to Real exactly as :233,:234,:250 do) — see pow_factor().
template <class R> static inline R pow_factor(R base, R dt) { return (R)std::pow((double)base, (double)dt); }
template <class R> static inline void body_integrate_pows(Body<R> &b, R dt, R linPow, R angPow, R bias) {
    if (b.isAwake == false) return;
    b.lastFrameAcc = b.acceleration;
    v_add_scaled(b.lastFrameAcc, b.forceAccum, b.inverseMass);
    V3<R> angAcc = m3_mul_v(b.iitWorld, b.torqueAccum);
    v_add_scaled(b.velocity, b.lastFrameAcc, dt);
    v_add_scaled(b.rotation, angAcc, dt);
    v_mul(b.velocity, linPow);
    v_mul(b.rotation, angPow);
    v_add_scaled(b.position, b.velocity, dt);
    q_add_scaled_vector(b.orientation, b.rotation, dt);
    body_calculate_derived(b);
    v_clear(b.forceAccum); v_clear(b.torqueAccum);   // ClearAccumulators :206
    if (b.canSleep) {
        R cur = v_dot(b.velocity, b.velocity) + v_dot(b.rotation, b.rotation);
        b.motion = bias * b.motion + ((R)1.0 - bias) * cur;
        if (b.motion < (R)0.3) body_set_awake(b, false);
        else if (b.motion > (R)3.0) b.motion = (R)3.0;   // 10*sleepEpsilon is the exact constant 3.0
    }
}
template <class R> static inline void body_integrate(Body<R> &b, R dt) {
    body_integrate_pows<R>(b, dt, pow_factor<R>(b.linearDamping, dt), pow_factor<R>(b.angularDamping, dt), pow_factor<R>((R)0.5, dt));
}

// ---------------------------------------------------------------------------------------
// contact.go:17-57 — Contact
// ---------------------------------------------------------------------------------------
template <class R> struct Contact {
    Body<R> *bodies[2];
    R friction, restitution;
    V3<R> contactPoint, contactNormal;
    R penetration;
    M3<R> contactToWorld;
    V3<R> relPos[2];
    V3<R> contactVelocity;
    R desiredDeltaVelocity;
};
template <class R> static inline Contact<R> *new_contact() { Contact<R> *c = new Contact<R>; std::memset(c, 0, sizeof(*c)); return c; }

// ---------------------------------------------------------------------------------------
// colliders.go
// ---------------------------------------------------------------------------------------
enum Shape { SHAPE_NONE = 0, SHAPE_CUBE = 1, SHAPE_SPHERE = 2, SHAPE_PLANE = 3 };

template <class R> struct Plane { V3<R> normal; R offset; };   // colliders.go:29-35
template <class R> struct Collider {                           // colliders.go:39-71 (cube and sphere share this record)
    int shape;
    Body<R> *body;
    M34<R> offset, transform;
    V3<R> halfSize;
    R radius;
};
template <class R> using Contacts = std::vector<Contact<R> *>;

template <class R> static inline void collider_derive(Collider<R> &c) { c.transform = m34_mul_m34(c.body->transform, c.offset); } // :173-176, :302-304

template <class R> static inline R transform_to_axis(const Collider<R> &cube, const V3<R> &axis) {   // colliders.go:762-770
    V3<R> ax = m34_axis(cube.transform, 0), ay = m34_axis(cube.transform, 1), az = m34_axis(cube.transform, 2);
    return cube.halfSize[0] * rabs<R>(v_dot(axis, ax)) + cube.halfSize[1] * rabs<R>(v_dot(axis, ay)) + cube.halfSize[2] * rabs<R>(v_dot(axis, az));
}

// colliders.go:180-207
template <class R> static bool sphere_vs_halfspace(const Collider<R> &s, const Plane<R> &plane, Contacts<R> &out) {
    V3<R> pos = m34_axis(s.transform, 3);
    R distance = v_dot(plane.normal, pos) - s.radius;
    if ((distance <= plane.offset) == false) return false;
    Contact<R> *c = new_contact<R>();
    c->contactPoint = plane.normal;
    v_mul(c->contactPoint, distance + s.radius * (R)-1.0);
    v_add(c->contactPoint, pos);
    c->contactNormal = plane.normal;
    c->penetration = -distance;
    c->bodies[0] = s.body; c->bodies[1] = nullptr;
    c->friction = (R)0.9; c->restitution = (R)0.1;
    out.push_back(c);
    return true;
}
// colliders.go:216-254
template <class R> static bool sphere_vs_sphere(const Collider<R> &s, const Collider<R> &o, Contacts<R> &out) {
    V3<R> p1 = m34_axis(s.transform, 3), p2 = m34_axis(o.transform, 3);
    V3<R> mid = p1;
    v_sub(mid, p2);
    R size = v_mag(mid);
    if (size <= (R)0.0 || size >= s.radius + o.radius) return false;
    Contact<R> *c = new_contact<R>();
    c->contactPoint = mid;
    v_mul(c->contactPoint, (R)0.5);
    v_add(c->contactPoint, p1);
    c->contactNormal = mid;
    v_mul(c->contactNormal, (R)1.0 / size);
    c->penetration = s.radius + o.radius - size;
    c->bodies[0] = s.body; c->bodies[1] = o.body;
    c->friction = (R)0.9; c->restitution = (R)0.1;
    out.push_back(c);
    return true;
}
// colliders.go:750-760
template <class R> static inline bool intersect_cube_halfspace(const Collider<R> &cube, const Plane<R> &plane) {
    R pr = transform_to_axis(cube, plane.normal);
    V3<R> axis = m34_axis(cube.transform, 3);
    R d = v_dot(plane.normal, axis) - pr;
    return d <= plane.offset;
}
// colliders.go:308-366
template <class R> static bool cube_vs_halfspace(const Collider<R> &cube, const Plane<R> &plane, Contacts<R> &out) {
    if (!intersect_cube_halfspace(cube, plane)) return false;
    static const R mults[8][3] = {{1, 1, 1}, {-1, 1, 1}, {1, -1, 1}, {-1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {1, -1, -1}, {-1, -1, -1}};
    bool found = false;
    for (int k = 0; k < 8; k++) {
        V3<R> v{{mults[k][0], mults[k][1], mults[k][2]}};
        v_component_product(v, cube.halfSize);
        V3<R> vp = m34_mul_v(cube.transform, v);
        R vd = v_dot(vp, plane.normal);
        if (vd <= plane.offset) {
            Contact<R> *c = new_contact<R>();
            c->contactPoint = plane.normal;
            v_mul(c->contactPoint, vd - plane.offset);
            v_add(c->contactPoint, vp);
            c->contactNormal = plane.normal;
            c->penetration = plane.offset - vd;
            c->bodies[0] = cube.body; c->bodies[1] = nullptr;
            out.push_back(c);
            found = true;
            c->friction = (R)0.9; c->restitution = (R)0.1;
        }
    }
    return found;
}
// colliders.go:369-441
template <class R> static bool cube_vs_sphere(const Collider<R> &cube, const Collider<R> &sphere, Contacts<R> &out) {
    V3<R> position = m34_axis(sphere.transform, 3);
    V3<R> rel = m34_transform_inverse(cube.transform, position);
    if (rabs<R>(rel[0]) - sphere.radius > cube.halfSize[0] || rabs<R>(rel[1]) - sphere.radius > cube.halfSize[1] ||
        rabs<R>(rel[2]) - sphere.radius > cube.halfSize[2])
        return false;
    V3<R> closest;
    for (int i = 0; i < 3; i++) {
        R dist = rel[i];
        if (dist > cube.halfSize[i]) dist = cube.halfSize[i];
        else if (dist < -cube.halfSize[i]) dist = -cube.halfSize[i];
        closest[i] = dist;
    }
    V3<R> dc = closest;
    v_sub(dc, rel);
    R dist = v_sqmag(dc);
    if (dist > sphere.radius * sphere.radius) return false;
    V3<R> cw = m34_mul_v(cube.transform, closest);
    Contact<R> *c = new_contact<R>();
    c->contactPoint = cw;
    c->contactNormal = cw;
    v_sub(c->contactNormal, position);
    if (real_equal<R>(v_mag(c->contactNormal), (R)0.0)) c->contactNormal = sphere.body->velocity;
    v_normalize(c->contactNormal);
    c->penetration = sphere.radius;
    if (!real_equal<R>(dist, (R)0.0)) c->penetration -= rsqrt_<R>(dist);
    else c->penetration = (R)0.0;
    c->bodies[0] = cube.body; c->bodies[1] = sphere.body;
    out.push_back(c);
    c->friction = (R)0.9; c->restitution = (R)0.1;
    return true;
}
// colliders.go:445-456
template <class R> static inline R penetration_on_axis(const Collider<R> &one, const Collider<R> &two, const V3<R> &axis, const V3<R> &toCenter) {
    R p1 = transform_to_axis(one, axis), p2 = transform_to_axis(two, axis);
    R distance = rabs<R>(v_dot(toCenter, axis));
    return p1 + p2 - distance;
}
// colliders.go:458-477
template <class R> static inline bool try_axis(const Collider<R> &one, const Collider<R> &two, V3<R> axis, const V3<R> &toCenter, int index, R &smallest, int &smallestCase) {
    if (v_sqmag(axis) < Lim<R>::epsilon) return true;
    v_normalize(axis);
    R pen = penetration_on_axis(one, two, axis, toCenter);
    if (pen < 0) return false;
    if (pen < smallest) { smallest = pen; smallestCase = index; }
    return true;
}
// colliders.go:481-517
template <class R> static void fill_point_face(const Collider<R> &one, const Collider<R> &two, const V3<R> &toCenter, int best, R pen, Contacts<R> &out) {
    V3<R> normal = m34_axis(one.transform, best);
    if (v_dot(normal, toCenter) > 0) v_mul(normal, (R)-1.0);
    V3<R> v = two.halfSize;
    if (v_dot(m34_axis(two.transform, 0), normal) < 0) v[0] = -v[0];
    if (v_dot(m34_axis(two.transform, 1), normal) < 0) v[1] = -v[1];
    if (v_dot(m34_axis(two.transform, 2), normal) < 0) v[2] = -v[2];
    Contact<R> *c = new_contact<R>();
    c->contactNormal = normal;
    c->penetration = pen;
    c->contactPoint = m34_mul_v(two.transform, v);
    c->bodies[0] = one.body; c->bodies[1] = two.body;
    c->friction = (R)0.9; c->restitution = (R)0.1;
    out.push_back(c);
}
// colliders.go:519-573
template <class R> static V3<R> contact_point(const V3<R> &pOne, const V3<R> &dOne, R oneSize, const V3<R> &pTwo, const V3<R> &dTwo, R twoSize, bool useOne) {
    R smOne = v_sqmag(dOne), smTwo = v_sqmag(dTwo);
    R dpOneTwo = v_dot(dTwo, dOne);
    V3<R> toSt = pOne;
    v_sub(toSt, pTwo);
    R dpStaOne = v_dot(dOne, toSt), dpStaTwo = v_dot(dTwo, toSt);
    R denom = smOne * smTwo - dpOneTwo * dpOneTwo;
    if (rabs<R>(denom) < Lim<R>::epsilon) return useOne ? pOne : pTwo;
    R mua = (dpOneTwo * dpStaTwo - smTwo * dpStaOne) / denom;
    R mub = (smOne * dpStaTwo - dpOneTwo * dpStaOne) / denom;
    if (mua > oneSize || mua < -oneSize || mub > twoSize || mub < -twoSize) return useOne ? pOne : pTwo;
    V3<R> cOne = dOne; v_mul(cOne, mua); v_add(cOne, pOne);
    V3<R> cTwo = dTwo; v_mul(cTwo, mub); v_add(cTwo, pTwo);
    v_mul(cOne, (R)0.5); v_mul(cTwo, (R)0.5);
    v_add(cOne, cTwo);
    return cOne;
}
// colliders.go:576-710
template <class R> static bool cube_vs_cube(const Collider<R> &one, const Collider<R> &two, Contacts<R> &out) {
    V3<R> toCenter = m34_axis(two.transform, 3);
    V3<R> oneAxis3 = m34_axis(one.transform, 3);
    v_sub(toCenter, oneAxis3);
    R pen = Lim<R>::max_value();
    int best = 0xffffff;
    for (int i = 0; i <= 2; i++) if (!try_axis(one, two, m34_axis(one.transform, i), toCenter, i, pen, best)) return false;
    for (int i = 0; i <= 2; i++) if (!try_axis(one, two, m34_axis(two.transform, i), toCenter, i + 3, pen, best)) return false;
    int bestSingleAxis = best;
    for (int i = 0; i <= 2; i++) {
        for (int j = 0; j <= 2; j++) {
            V3<R> a1 = m34_axis(one.transform, i), a2 = m34_axis(two.transform, j);
            if (!try_axis(one, two, v_cross(a1, a2), toCenter, i * 3 + 6 + j, pen, best)) return false;
        }
    }
    if (best < 3) {
        fill_point_face(one, two, toCenter, best, pen, out);
        return true;
    } else if (best < 6) {
        V3<R> nc = toCenter;
        v_mul(nc, (R)-1.0);
        fill_point_face(two, one, nc, best - 3, pen, out);
        return true;
    }
    best -= 6;
    int oneIdx = best / 3, twoIdx = best % 3;
    V3<R> oneAxis = m34_axis(one.transform, oneIdx), twoAxis = m34_axis(two.transform, twoIdx);
    V3<R> axis = v_cross(oneAxis, twoAxis);
    v_normalize(axis);
    if (v_dot(axis, toCenter) > 0) v_mul(axis, (R)-1.0);
    V3<R> ptOne = one.halfSize, ptTwo = two.halfSize;
    for (int i = 0; i < 3; i++) {
        if (i == oneIdx) ptOne[i] = 0;
        else if (v_dot(m34_axis(one.transform, i), axis) > 0) ptOne[i] = -ptOne[i];
        if (i == twoIdx) ptTwo[i] = 0;
        else if (v_dot(m34_axis(two.transform, i), axis) < 0) ptTwo[i] = -ptTwo[i];
    }
    ptOne = m34_mul_v(one.transform, ptOne);
    ptTwo = m34_mul_v(two.transform, ptTwo);
    bool useOne = bestSingleAxis > 2;
    V3<R> vtx = contact_point(ptOne, oneAxis, one.halfSize[oneIdx], ptTwo, twoAxis, two.halfSize[twoIdx], useOne);
    Contact<R> *c = new_contact<R>();
    c->contactNormal = axis;
    c->penetration = pen;
    c->contactPoint = vtx;
    c->bodies[0] = one.body; c->bodies[1] = two.body;
    c->friction = (R)0.9; c->restitution = (R)0.1;
    out.push_back(c);
    return true;
}
// colliders.go:720-747 plus the per-type forwarding methods (:111-125, :210-213).
// `two_plane`/`one_plane` carry the plane when that operand is a CollisionPlane.
template <class R> static bool check_for_collisions(const Collider<R> *one, const Plane<R> *onePlane, const Collider<R> *two, const Plane<R> *twoPlane, Contacts<R> &out) {
    int s1 = onePlane ? SHAPE_PLANE : one->shape, s2 = twoPlane ? SHAPE_PLANE : two->shape;
    switch (s2) {
    case SHAPE_SPHERE:
        if (s1 == SHAPE_PLANE) return sphere_vs_halfspace(*two, *onePlane, out);   // :116-119
        if (s1 == SHAPE_SPHERE) return sphere_vs_sphere(*one, *two, out);
        if (s1 == SHAPE_CUBE) return cube_vs_sphere(*one, *two, out);
        return false;
    case SHAPE_CUBE:
        if (s1 == SHAPE_PLANE) return cube_vs_halfspace(*two, *onePlane, out);     // :122-125
        if (s1 == SHAPE_SPHERE) return cube_vs_sphere(*two, *one, out);            // :210-213
        if (s1 == SHAPE_CUBE) return cube_vs_cube(*one, *two, out);
        return false;
    case SHAPE_PLANE:
        if (s1 == SHAPE_PLANE) return false;                                      // :111-113
        if (s1 == SHAPE_SPHERE) return sphere_vs_halfspace(*one, *twoPlane, out);
        if (s1 == SHAPE_CUBE) return cube_vs_halfspace(*one, *twoPlane, out);
        return false;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// contact.go — resolver
// ---------------------------------------------------------------------------------------
// contact.go:118-156
template <class R> static void contact_basis(Contact<R> &c) {
    V3<R> ty, tz;
    const V3<R> &n = c.contactNormal;
    if (rabs<R>(n[0]) > rabs<R>(n[1])) {
        R s = (R)1.0 / rsqrt_<R>(n[2] * n[2] + n[0] * n[0]);
        ty[0] = n[2] * s; ty[1] = 0; ty[2] = n[0] * -s;
        tz[0] = n[1] * ty[0];
        tz[1] = n[2] * ty[0] - n[0] * ty[2];
        tz[2] = -n[1] * ty[0];
    } else {
        R s = (R)1.0 / rsqrt_<R>(n[2] * n[2] + n[1] * n[1]);
        ty[0] = 0; ty[1] = -n[2] * s; ty[2] = n[1] * s;
        tz[0] = n[1] * ty[2] - n[2] * ty[1];
        tz[1] = -n[0] * ty[2];
        tz[2] = n[0] * ty[1];
    }
    m3_set_components(c.contactToWorld, c.contactNormal, ty, tz);
}
// contact.go:159-182
template <class R> static V3<R> contact_local_velocity(Contact<R> &c, int bi, R dt) {
    Body<R> *body = c.bodies[bi];
    V3<R> vel = v_cross(body->rotation, c.relPos[bi]);
    v_add(vel, body->velocity);
    V3<R> cv = m3_transform_transpose(c.contactToWorld, vel);
    V3<R> acc = body->lastFrameAcc;
    v_mul(acc, dt);
    acc = m3_transform_transpose(c.contactToWorld, acc);
    acc[0] = 0;
    v_add(cv, acc);
    return cv;
}
// contact.go:87-112
template <class R> static void contact_desired_delta_velocity(Contact<R> &c, R dt) {
    const R velocityLimit = (R)0.25;
    R vfa = 0;
    if (c.bodies[0]->isAwake) {
        V3<R> t = c.bodies[0]->lastFrameAcc;
        v_mul(t, dt);
        vfa += v_dot(t, c.contactNormal);
    }
    if (c.bodies[1] != nullptr && c.bodies[1]->isAwake) {
        V3<R> t = c.bodies[1]->lastFrameAcc;
        v_mul(t, dt);
        vfa -= v_dot(t, c.contactNormal);
    }
    R rest = c.restitution;
    if (rabs<R>(c.contactVelocity[0]) < velocityLimit) rest = 0;
    c.desiredDeltaVelocity = -c.contactVelocity[0] - rest * (c.contactVelocity[0] - vfa);
}
// contact.go:59-85
template <class R> static void contact_internals(Contact<R> &c, R dt) {
    if (c.bodies[0] == nullptr) {
        v_mul(c.contactNormal, (R)-1.0);
        c.bodies[0] = c.bodies[1];
        c.bodies[1] = nullptr;
    }
    contact_basis(c);
    c.relPos[0] = c.contactPoint;
    v_sub(c.relPos[0], c.bodies[0]->position);
    c.contactVelocity = contact_local_velocity(c, 0, dt);
    if (c.bodies[1] != nullptr) {
        c.relPos[1] = c.contactPoint;
        v_sub(c.relPos[1], c.bodies[1]->position);
        V3<R> cv1 = contact_local_velocity(c, 1, dt);
        v_sub(c.contactVelocity, cv1);
    }
    contact_desired_delta_velocity(c, dt);
}
// contact.go:185-202
template <class R> static void contact_match_awake(Contact<R> &c) {
    if (c.bodies[1] == nullptr) return;
    bool a0 = c.bodies[0]->isAwake, a1 = c.bodies[1]->isAwake;
    if ((a0 || a1) && !(a0 && a1)) {
        if (a0) body_set_awake(*c.bodies[1], true);
        else body_set_awake(*c.bodies[0], true);
    }
}
// contact.go:286-386
template <class R> static void contact_apply_position_change(Contact<R> &c, R penetration, V3<R> linearChange[2], V3<R> angularChange[2]) {
    const R angularLimit = (R)0.2;
    R angularInertia[2] = {0, 0}, linearInertia[2] = {0, 0}, angularMove[2] = {0, 0}, linearMove[2] = {0, 0};
    R totalInertia = 0;
    for (int i = 0; i < 2; i++) { v_clear(linearChange[i]); v_clear(angularChange[i]); }
    for (int i = 0; i < 2; i++) {
        Body<R> *body = c.bodies[i];
        if (body == nullptr) continue;
        M3<R> iit = body->iitWorld;
        V3<R> aiw = v_cross(c.relPos[i], c.contactNormal);
        aiw = m3_mul_v(iit, aiw);
        aiw = v_cross(aiw, c.relPos[i]);
        angularInertia[i] = v_dot(aiw, c.contactNormal);
        linearInertia[i] = body->inverseMass;
        totalInertia += linearInertia[i] + angularInertia[i];
    }
    for (int i = 0; i < 2; i++) {
        Body<R> *body = c.bodies[i];
        if (body == nullptr) continue;
        R sign = (R)1.0;
        if (i != 0) sign = (R)-1.0;
        angularMove[i] = sign * penetration * (angularInertia[i] / totalInertia);
        linearMove[i] = sign * penetration * (linearInertia[i] / totalInertia);
        V3<R> proj = c.relPos[i];
        v_add_scaled(proj, c.contactNormal, -v_dot(c.relPos[i], c.contactNormal));
        R maxMag = angularLimit * v_mag(proj);
        if (angularMove[i] < -maxMag) {
            R total = angularMove[i] + linearMove[i];
            angularMove[i] = -maxMag;
            linearMove[i] = total - angularMove[i];
        } else if (angularMove[i] > maxMag) {
            R total = angularMove[i] + linearMove[i];
            angularMove[i] = maxMag;
            linearMove[i] = total - angularMove[i];
        }
        if (angularMove[i] == (R)0.0) {
            v_clear(angularChange[i]);
        } else {
            V3<R> target = v_cross(c.relPos[i], c.contactNormal);
            M3<R> iit = body->iitWorld;
            angularChange[i] = m3_mul_v(iit, target);
            v_mul(angularChange[i], angularMove[i] / angularInertia[i]);
        }
        linearChange[i] = c.contactNormal;
        v_mul(linearChange[i], linearMove[i]);
        v_add_scaled(body->position, c.contactNormal, linearMove[i]);
        q_add_scaled_vector(body->orientation, angularChange[i], (R)1.0);
        q_normalize(body->orientation);
        if (body->isAwake == false) body_calculate_derived(*body);
    }
}
// contact.go:233-283.  Returns the number of iterations used.
template <class R> static int adjust_positions(int maxIterations, Contacts<R> &contacts, R dt) {
    (void)dt;
    int used = 0;
    const size_t n = contacts.size();
    while (used < maxIterations) {
        R max = (R)0.01;
        size_t index = n;
        for (size_t i = 0; i < n; i++) {
            if (contacts[i]->penetration > max) { max = contacts[i]->penetration; index = i; }
        }
        if (index == n) break;
        Contact<R> *contact = contacts[index];
        contact_match_awake(*contact);
        V3<R> lin[2], ang[2];
        contact_apply_position_change(*contact, max, lin, ang);
        for (size_t i = 0; i < n; i++) {
            Contact<R> *c = contacts[i];
            for (int b = 0; b < 2; b++) {
                if (c->bodies[b] != nullptr) {
                    for (int d = 0; d < 2; d++) {
                        if (c->bodies[b] == contact->bodies[d]) {
                            V3<R> dp = v_cross(ang[d], c->relPos[b]);
                            v_add(dp, lin[d]);
                            R sign = (R)1.0;
                            if (b == 0) sign = (R)-1.0;
                            c->penetration += v_dot(dp, c->contactNormal) * sign;
                        }
                    }
                }
            }
        }
        used++;
    }
    return used;
}
// contact.go:612-616
template <class R> static inline void set_skew(M3<R> &m, const V3<R> &v) {
    m[0] = 0; m[3] = -v[2]; m[6] = v[1];
    m[1] = v[2]; m[4] = 0; m[7] = -v[0];
    m[2] = -v[1]; m[5] = v[0]; m[8] = 0;
}
// contact.go:498-531.  status: set to 1 when the reference would nil-dereference (:512-523).
template <class R> static V3<R> frictionless_impulse(Contact<R> &c, const M3<R> iit[2], int *status) {
    V3<R> dvw = v_cross(c.relPos[0], c.contactNormal);
    dvw = m3_mul_v(iit[0], dvw);
    dvw = v_cross(dvw, c.relPos[0]);
    R dv = v_dot(dvw, c.contactNormal);
    dv += c.bodies[0]->inverseMass;
    if (c.bodies[1] == nullptr) {
        // The reference dereferences Bodies[1] here (a Go panic).  Report it.
        if (status) *status = 1;
    }
    V3<R> imp;
    imp[0] = c.desiredDeltaVelocity / dv; imp[1] = 0; imp[2] = 0;
    return imp;
}
// contact.go:535-606
template <class R> static V3<R> friction_impulse(Contact<R> &c, const M3<R> iit[2]) {
    R inverseMass = c.bodies[0]->inverseMass;
    M3<R> itt;
    set_skew(itt, c.relPos[0]);
    M3<R> dvw = m3_mul_m(itt, iit[0]);
    dvw = m3_mul_m(dvw, itt);
    m3_mul_s(dvw, (R)-1.0);
    if (c.bodies[1] != nullptr) {
        set_skew(itt, c.relPos[1]);
        M3<R> dvw2 = m3_mul_m(itt, iit[1]);
        dvw2 = m3_mul_m(dvw2, itt);
        m3_mul_s(dvw2, (R)-1.0);
        m3_add(dvw, dvw2);
        inverseMass += c.bodies[1]->inverseMass;
    }
    M3<R> dv = m3_transpose(c.contactToWorld);
    dv = m3_mul_m(dv, dvw);
    dv = m3_mul_m(dv, c.contactToWorld);
    dv[0] += inverseMass; dv[4] += inverseMass; dv[8] += inverseMass;
    M3<R> im = m3_invert(dv);
    V3<R> velKill{{c.desiredDeltaVelocity, -c.contactVelocity[1], -c.contactVelocity[2]}};
    V3<R> imp = m3_mul_v(im, velKill);
    R planar = rsqrt_<R>(imp[1] * imp[1] + imp[2] * imp[2]);
    if (planar > imp[0] * c.friction) {
        imp[1] /= planar;
        imp[2] /= planar;
        imp[0] = dv[0] + dv[3] * c.friction * imp[1] + dv[6] * c.friction * imp[2];
        imp[0] = c.desiredDeltaVelocity / imp[0];
        imp[1] *= c.friction * imp[0];
        imp[2] *= c.friction * imp[0];
    }
    return imp;
}
// contact.go:448-494
template <class R> static void contact_apply_velocity_change(Contact<R> &c, V3<R> velocityChange[2], V3<R> rotationChange[2], int *status) {
    M3<R> iit[2];
    std::memset(iit, 0, sizeof(iit));
    for (int i = 0; i < 2; i++) { v_clear(velocityChange[i]); v_clear(rotationChange[i]); }
    iit[0] = c.bodies[0]->iitWorld;
    if (c.bodies[1] != nullptr) iit[1] = c.bodies[1]->iitWorld;
    V3<R> ic;
    if (c.friction == (R)0.0) ic = frictionless_impulse(c, iit, status);
    else ic = friction_impulse(c, iit);
    V3<R> impulse = m3_mul_v(c.contactToWorld, ic);
    V3<R> torque = v_cross(c.relPos[0], impulse);
    rotationChange[0] = m3_mul_v(iit[0], torque);
    v_clear(velocityChange[0]);
    v_add_scaled(velocityChange[0], impulse, c.bodies[0]->inverseMass);
    v_add(c.bodies[0]->velocity, velocityChange[0]);
    v_add(c.bodies[0]->rotation, rotationChange[0]);
    if (c.bodies[1] != nullptr) {
        torque = v_cross(impulse, c.relPos[1]);
        rotationChange[1] = m3_mul_v(iit[1], torque);
        v_clear(velocityChange[1]);
        v_add_scaled(velocityChange[1], impulse, -c.bodies[1]->inverseMass);
        v_add(c.bodies[1]->velocity, velocityChange[1]);
        v_add(c.bodies[1]->rotation, rotationChange[1]);
    }
}
// contact.go:390-445
template <class R> static int adjust_velocities(int maxIterations, Contacts<R> &contacts, R dt, int *status) {
    int used = 0;
    const size_t n = contacts.size();
    while (used < maxIterations) {
        R max = (R)0.01;
        size_t index = n;
        for (size_t i = 0; i < n; i++) {
            if (contacts[i]->desiredDeltaVelocity > max) { max = contacts[i]->desiredDeltaVelocity; index = i; }
        }
        if (index == n) break;
        Contact<R> *contact = contacts[index];
        contact_match_awake(*contact);
        V3<R> velc[2], rotc[2];
        contact_apply_velocity_change(*contact, velc, rotc, status);
        for (size_t i = 0; i < n; i++) {
            Contact<R> *c2 = contacts[i];
            for (int b = 0; b < 2; b++) {
                if (c2->bodies[b] == nullptr) continue;
                for (int d = 0; d < 2; d++) {
                    if (c2->bodies[b] == contact->bodies[d]) {
                        V3<R> dv = v_cross(rotc[d], c2->relPos[b]);
                        v_add(dv, velc[d]);
                        R sign = (R)1.0;
                        if (b == 1) sign = (R)-1.0;
                        V3<R> t = m3_transform_transpose(c2->contactToWorld, dv);
                        v_mul(t, sign);
                        v_add(c2->contactVelocity, t);
                        contact_desired_delta_velocity(*c2, dt);
                    }
                }
            }
        }
        used++;
    }
    return used;
}
// contact.go:208-222.  iters[0]/iters[1] receive the iterations used by each phase.
template <class R> static void resolve_contacts(int maxIterations, Contacts<R> &contacts, R dt, int iters[2], int *status) {
    iters[0] = iters[1] = 0;
    if (dt <= (R)0.0 || contacts.empty()) return;
    for (Contact<R> *c : contacts) contact_internals(*c, dt);   // prepareContacts :225
    iters[0] = adjust_positions(maxIterations, contacts, dt);
    iters[1] = adjust_velocities(maxIterations, contacts, dt, status);
}

// ---------------------------------------------------------------------------------------
// World step — examples/cubedrop.go:29-75 and examples/ballistic.go:27-105 re-expressed
// with an explicit pair schedule.
// ---------------------------------------------------------------------------------------
enum Schedule { SCHED_ALL_PAIRS_ORDERED = 0, SCHED_EXPLICIT = 1 };

template <class R> struct ContactRecord {   // a contact as generated, before ResolveContacts
    int32_t body[2];
    R point[3], normal[3], penetration;
    R friction, restitution;
};

template <class R> struct World {
    std::vector<Body<R>> bodies;
    std::vector<Collider<R>> colliders;      // collider i belongs to body i
    std::vector<Plane<R>> planes;
    std::vector<int32_t> activeFrom;         // body takes part from this step index on
    std::vector<uint8_t> integrateFlag;      // 0: never integrated (ballistic backboard)
    int schedule = SCHED_ALL_PAIRS_ORDERED;
    std::vector<int32_t> checkOne, checkTwo; // explicit schedule; >=0 collider index, <0 plane -(p+1)
    int64_t stepIndex = 0;
    // RL-style episodes (no reference counterpart): restore `episodeStart` whenever the
    // world's phase wraps to 0
    int episodeLength = 0, episodePhase0 = 0;
    int64_t episodeStep0 = 0;
    std::vector<Body<R>> episodeBodies;
    std::vector<Collider<R>> episodeColliders;
    // outputs of the last step
    std::vector<ContactRecord<R>> lastContacts;
    int posIters = 0, velIters = 0, status = 0;
    int64_t totalContacts = 0, totalPosIters = 0, totalVelIters = 0;
    // Per-pair surface materials (no reference counterpart — the reference
    // hard-wires Friction 0.9 / Restitution 0.1 at every `c.Friction = 0.9` FIXME site,
    // colliders.go:199-202, :246-249, :358-361, :435-438, :509-512, :702-705).  nMaterials == 0
    // keeps the constants.  A contact produced by CheckForCollisions(one, two) takes
    // table[material(one)][material(two)], the operands as the schedule names them.
    int nMaterials = 0;
    std::vector<R> matFriction, matRestitution;   // [nMaterials * nMaterials], row = one, column = two
    std::vector<int32_t> bodyMaterial, planeMaterial;

    void tag_materials(Contacts<R> &cs, size_t from, int a, int b) const {
        if (nMaterials <= 0) return;
        const int ma = a >= 0 ? bodyMaterial[a] : planeMaterial[-a - 1];
        const int mb = b >= 0 ? bodyMaterial[b] : planeMaterial[-b - 1];
        for (size_t k = from; k < cs.size(); k++) {
            cs[k]->friction = matFriction[(size_t)ma * nMaterials + mb];
            cs[k]->restitution = matRestitution[(size_t)ma * nMaterials + mb];
        }
    }

    void fix_pointers() { for (size_t i = 0; i < colliders.size(); i++) colliders[i].body = &bodies[i]; }
    bool active(int i) const { return stepIndex >= activeFrom[i]; }

    void step(R dt) {
        const int n = (int)bodies.size();
        if (episodeLength > 0 && (episodePhase0 + (stepIndex - episodeStep0)) % episodeLength == 0) {
            bodies = episodeBodies;
            colliders = episodeColliders;
            // (new API, no reference counterpart) a reset also drops forces that were still waiting in the accumulators —
            // of sleeping bodies, or added just before this frame: the episode starts from the snapshot, at rest
            for (auto &b : bodies) { v_clear(b.forceAccum); v_clear(b.torqueAccum); }
        }
        fix_pointers();
        // updateObjects — cubedrop.go:29-39 / ballistic.go:27-44
        for (int i = 0; i < n; i++) {
            if (!active(i) || !integrateFlag[i]) continue;
            body_integrate<R>(bodies[i], dt);
            if (colliders[i].shape != SHAPE_NONE) collider_derive(colliders[i]);
        }
        // generateContacts — cubedrop.go:42-67 / ballistic.go:47-97
        Contacts<R> contacts;
        if (schedule == SCHED_ALL_PAIRS_ORDERED) {
            for (int i = 0; i < n; i++) {
                if (!active(i) || colliders[i].shape == SHAPE_NONE) continue;
                for (size_t p = 0; p < planes.size(); p++) {
                    const size_t from = contacts.size();
                    check_for_collisions<R>(&colliders[i], nullptr, nullptr, &planes[p], contacts);
                    tag_materials(contacts, from, i, -(int)p - 1);
                }
                for (int j = 0; j < n; j++) {
                    if (j == i || !active(j) || colliders[j].shape == SHAPE_NONE) continue;
                    const size_t from = contacts.size();
                    check_for_collisions<R>(&colliders[i], nullptr, &colliders[j], nullptr, contacts);
                    tag_materials(contacts, from, i, j);
                }
            }
        } else {
            for (size_t k = 0; k < checkOne.size(); k++) {
                int a = checkOne[k], b = checkTwo[k];
                const Collider<R> *ca = nullptr, *cb = nullptr;
                const Plane<R> *pa = nullptr, *pb = nullptr;
                if (a >= 0) { if (!active(a) || colliders[a].shape == SHAPE_NONE) continue; ca = &colliders[a]; } else pa = &planes[-a - 1];
                if (b >= 0) { if (!active(b) || colliders[b].shape == SHAPE_NONE) continue; cb = &colliders[b]; } else pb = &planes[-b - 1];
                const size_t from = contacts.size();
                check_for_collisions<R>(ca, pa, cb, pb, contacts);
                tag_materials(contacts, from, a, b);
            }
        }
        lastContacts.clear();
        for (Contact<R> *c : contacts) {
            ContactRecord<R> r;
            r.body[0] = c->bodies[0] ? (int32_t)(c->bodies[0] - bodies.data()) : -1;
            r.body[1] = c->bodies[1] ? (int32_t)(c->bodies[1] - bodies.data()) : -1;
            for (int k = 0; k < 3; k++) { r.point[k] = c->contactPoint[k]; r.normal[k] = c->contactNormal[k]; }
            r.penetration = c->penetration;
            r.friction = c->friction; r.restitution = c->restitution;
            lastContacts.push_back(r);
        }
        int iters[2] = {0, 0};
        if (!contacts.empty()) resolve_contacts<R>((int)contacts.size() * 8, contacts, dt, iters, &status);   // cubedrop.go:72-74
        posIters = iters[0]; velIters = iters[1];
        totalContacts += (int64_t)contacts.size(); totalPosIters += iters[0]; totalVelIters += iters[1];
        for (Contact<R> *c : contacts) delete c;
        stepIndex++;
    }
};

// FNV-1a-64 over the raw IEEE bits of Position, Orientation, Velocity, Rotation, IsAwake in
// body order (SURVEY §8d "Checksum / energy" — a new definition, no reference counterpart).
template <class R> static uint64_t world_checksum(const World<R> &w) {
    uint64_t h = 0xcbf29ce484222325ull;
    auto eat = [&h](const void *p, size_t n) {
        const unsigned char *b = (const unsigned char *)p;
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    };
    for (const Body<R> &b : w.bodies) {
        eat(b.position.c, sizeof(R) * 3); eat(b.orientation.c, sizeof(R) * 4);
        eat(b.velocity.c, sizeof(R) * 3); eat(b.rotation.c, sizeof(R) * 3);
        unsigned char a = b.isAwake ? 1 : 0;
        eat(&a, 1);
    }
    return h;
}

}  // namespace czo
