// cz_body.cuh — HBM layout of rigid-body state and the integration device functions (K1).
//
// Layout ("chunked SoA"): every body's reals are grouped into 16-byte pairs (real2 =
// double2 in f64 / float2 in f32) and each pair is its own array over bodies, so a thread that
// owns body i issues one 128-bit LDG/STG per chunk and a warp touches 512 contiguous bytes
// per chunk array.  Chunk arrays live in one allocation: chunk k starts at base + k*stride.
//
// K1 traffic per awake body: reads 14 chunks (28 reals) + 2 flag bytes, writes
// 19 chunks (38 reals) + 1 flag byte = 531 B (f64) / 267 B (f32).
#pragma once
#include "cz_math.cuh"

namespace czb {
using namespace czm;

enum Chunk : int {
    // read + written by Integrate
    C_P01 = 0,  // position.x, position.y
    C_P2M,      // position.z, motion
    C_Q01,      // orientation.w, .x
    C_Q23,      // orientation.y, .z
    C_V01,      // velocity.x, .y
    C_V2R0,     // velocity.z, rotation.x
    C_R12,      // rotation.y, .z
    // read-only for Integrate
    C_A01,      // acceleration.x, .y
    C_A2LP,     // acceleration.z, linPow  (= Pow(LinearDamping, dt), rigidbody.go:233)
    C_APW0,     // angPow (= Pow(AngularDamping, dt), :234), inverseInertiaTensor[0]
    C_I12, C_I34, C_I56, C_I78,   // inverseInertiaTensor[1..8] (body space)
    // written by Integrate
    C_L01,      // lastFrameAcceleration.x, .y
    C_L2T0,     // lastFrameAcceleration.z, transform[0]
    C_T12, C_T34, C_T56, C_T78, C_T910,   // transform[1..10]
    C_T11W0,    // transform[11], inverseInertiaTensorWorld[0]
    C_W12, C_W34, C_W56, C_W78,           // inverseInertiaTensorWorld[1..8]
    // collider (colliders.go:39-71); collider i belongs to body i in the world layout
    C_H01,      // halfSize.x, .y
    C_H2R,      // halfSize.z, radius
    C_O01, C_O23, C_O45, C_O67, C_O89, C_O1011,   // collider Offset[0..11]
    C_X01, C_X23, C_X45, C_X67, C_X89, C_X1011,   // collider transform[0..11] (derived)
    C_MD,       // inverseMass, linearDamping   (damping kept for download / pow refresh)
    C_AD,       // angularDamping, unused
    N_CHUNKS
};

// Device view of a batch of bodies (all worlds of a handle concatenated).
struct BodyStore {
    real2 *base;          // N_CHUNKS arrays of `stride` real2 each
    int64_t stride;       // elements per chunk array (>= n, multiple of 64)
    uint8_t *awake;       // IsAwake
    uint8_t *can_sleep;   // CanSleep
    uint8_t *integ;       // 0: never integrated (ballistic backboard)
    uint8_t *shape;       // CZ_SHAPE_*
    uint8_t *ident;       // 1: collider Offset is the identity (skip the 12-real read)
    int32_t *active_from; // takes part from this step on
    int64_t n;
    // forceAccum / torqueAccum (rigidbody.go:86-92), [n*3] each, or NULL: identically zero — the reference has no
    // writer for them, so they only exist once the host has called cz_world_add_forces
    real *force, *torque;

    CZD real2 *chunk(int k) const { return base + (int64_t)k * stride; }
    CZD real2 ld(int k, int64_t i) const { return chunk(k)[i]; }
    CZD void st(int k, int64_t i, real2 v) const { chunk(k)[i] = v; }
};

// --- whole-field accessors (used by the gather-style kernels and the resolver) ----------
CZD V3 ld_position(const BodyStore &s, int64_t i) { real2 a = s.ld(C_P01, i), b = s.ld(C_P2M, i); return mk3(a.x, a.y, b.x); }
CZD Q4 ld_orientation(const BodyStore &s, int64_t i) { real2 a = s.ld(C_Q01, i), b = s.ld(C_Q23, i); Q4 q; q.c[0] = a.x; q.c[1] = a.y; q.c[2] = b.x; q.c[3] = b.y; return q; }
CZD V3 ld_velocity(const BodyStore &s, int64_t i) { real2 a = s.ld(C_V01, i), b = s.ld(C_V2R0, i); return mk3(a.x, a.y, b.x); }
CZD V3 ld_rotation(const BodyStore &s, int64_t i) { real2 a = s.ld(C_V2R0, i), b = s.ld(C_R12, i); return mk3(a.y, b.x, b.y); }
CZD V3 ld_last_acc(const BodyStore &s, int64_t i) { real2 a = s.ld(C_L01, i), b = s.ld(C_L2T0, i); return mk3(a.x, a.y, b.x); }
CZD M3 ld_iit_body(const BodyStore &s, int64_t i) {
    M3 m; real2 a = s.ld(C_APW0, i), b = s.ld(C_I12, i), c = s.ld(C_I34, i), d = s.ld(C_I56, i), e = s.ld(C_I78, i);
    m.c[0] = a.y; m.c[1] = b.x; m.c[2] = b.y; m.c[3] = c.x; m.c[4] = c.y; m.c[5] = d.x; m.c[6] = d.y; m.c[7] = e.x; m.c[8] = e.y;
    return m;
}
CZD M3 ld_iit_world(const BodyStore &s, int64_t i) {
    M3 m; real2 a = s.ld(C_T11W0, i), b = s.ld(C_W12, i), c = s.ld(C_W34, i), d = s.ld(C_W56, i), e = s.ld(C_W78, i);
    m.c[0] = a.y; m.c[1] = b.x; m.c[2] = b.y; m.c[3] = c.x; m.c[4] = c.y; m.c[5] = d.x; m.c[6] = d.y; m.c[7] = e.x; m.c[8] = e.y;
    return m;
}
CZD M34 ld_transform(const BodyStore &s, int64_t i) {
    M34 m; real2 a = s.ld(C_L2T0, i), b = s.ld(C_T12, i), c = s.ld(C_T34, i), d = s.ld(C_T56, i), e = s.ld(C_T78, i), f = s.ld(C_T910, i), g = s.ld(C_T11W0, i);
    m.c[0] = a.y; m.c[1] = b.x; m.c[2] = b.y; m.c[3] = c.x; m.c[4] = c.y; m.c[5] = d.x; m.c[6] = d.y; m.c[7] = e.x; m.c[8] = e.y; m.c[9] = f.x; m.c[10] = f.y; m.c[11] = g.x;
    return m;
}
CZD M34 ld_m34(const BodyStore &s, int first_chunk, int64_t i) {
    M34 m;
#pragma unroll
    for (int k = 0; k < 6; k++) { real2 v = s.ld(first_chunk + k, i); m.c[2 * k] = v.x; m.c[2 * k + 1] = v.y; }
    return m;
}
CZD void st_m34(const BodyStore &s, int first_chunk, int64_t i, const M34 &m) {
#pragma unroll
    for (int k = 0; k < 6; k++) s.st(first_chunk + k, i, make_real2(m.c[2 * k], m.c[2 * k + 1]));
}
CZD M34 identity34() {
    M34 m;
#pragma unroll
    for (int k = 0; k < 12; k++) m.c[k] = R_(0);
    m.c[0] = R_(1); m.c[4] = R_(1); m.c[8] = R_(1);
    return m;
}
// transform and world inertia share chunks with their neighbours; these helpers write the
// derived block C_L2T0.y .. C_W78 given lastAcc.z (which shares C_L2T0).
CZD void st_derived(const BodyStore &s, int64_t i, real lastAccZ, const M34 &t, const M3 &w) {
    s.st(C_L2T0, i, make_real2(lastAccZ, t.c[0]));
    s.st(C_T12, i, make_real2(t.c[1], t.c[2]));
    s.st(C_T34, i, make_real2(t.c[3], t.c[4]));
    s.st(C_T56, i, make_real2(t.c[5], t.c[6]));
    s.st(C_T78, i, make_real2(t.c[7], t.c[8]));
    s.st(C_T910, i, make_real2(t.c[9], t.c[10]));
    s.st(C_T11W0, i, make_real2(t.c[11], w.c[0]));
    s.st(C_W12, i, make_real2(w.c[1], w.c[2]));
    s.st(C_W34, i, make_real2(w.c[3], w.c[4]));
    s.st(C_W56, i, make_real2(w.c[5], w.c[6]));
    s.st(C_W78, i, make_real2(w.c[7], w.c[8]));
}

// rigidbody.go:268-272 on values
CZD void calculate_derived(const V3 &pos, Q4 &q, const M3 &iitBody, M34 &transform, M3 &iitWorld) {
    q_normalize(q);
    m34_set_as_transform(transform, pos, q);
    transform_inertia_tensor(iitWorld, iitBody, transform);
}

// Result of integrating one body; everything Integrate writes (rigidbody.go:213-259).
struct Integrated {
    V3 pos, vel, rot, lastAcc;
    Q4 q;
    M34 transform;
    M3 iitWorld;
    real motion;
    bool awake;
};

// rigidbody.go:213-259 for an awake body.  linPow/angPow/bias are the host-evaluated
// math.Pow results (:233, :234, :250).  forceAccum and torqueAccum have no writer anywhere in
// the reference (only read :220,:223 and cleared :207-208), so they are the constant +0 here:
// `x + 0` is kept because it turns a -0 component into +0 exactly as the Go code does.
// rigidbody.go:233-258: damping, position / orientation update, derived data, sleep test (o.vel, o.rot, o.lastAcc set by the caller)
CZD void integrate_body_tail(Integrated &o, const V3 &pos, const Q4 &q, const M3 &iitBody, real motion, bool canSleep, real dt,
                             real linPow, real angPow, real bias) {
    v_mul(o.vel, linPow);                                                        // :233
    v_mul(o.rot, angPow);                                                        // :234
    o.pos = pos;
    v_add_scaled(o.pos, o.vel, dt);                                              // :238
    o.q = q;
    q_add_scaled_vector(o.q, o.rot, dt);                                         // :241
    calculate_derived(o.pos, o.q, iitBody, o.transform, o.iitWorld);             // :244
    o.motion = motion;
    o.awake = true;
    if (canSleep) {                                                              // :248-258
        real cur = v_dot(o.vel, o.vel) + v_dot(o.rot, o.rot);
        o.motion = bias * motion + (R_(1.0) - bias) * cur;
        if (o.motion < R_(0.3)) {                                                // SetAwake(false) :187-190
            o.awake = false;
            o.vel = zero3();
            o.rot = zero3();
        } else if (o.motion > R_(3.0)) {
            o.motion = R_(3.0);
        }
    }
}


CZD void integrate_body(Integrated &o, const V3 &pos, const Q4 &q, const V3 &vel, const V3 &rot, const V3 &acc,
                        const M3 &iitBody, real motion, bool canSleep, real dt, real linPow, real angPow, real bias) {
    o.lastAcc = acc;
    o.lastAcc.c[0] += R_(0); o.lastAcc.c[1] += R_(0); o.lastAcc.c[2] += R_(0);   // AddScaled(forceAccum = 0, inverseMass) :220
    o.vel = vel;
    v_add_scaled(o.vel, o.lastAcc, dt);                                          // :227
    o.rot = rot;
    { real z = R_(0) * dt; o.rot.c[0] += z; o.rot.c[1] += z; o.rot.c[2] += z; }   // AddScaled(iitWorld*torque = 0, dt) :230
    integrate_body_tail(o, pos, q, iitBody, motion, canSleep, dt, linPow, angPow, bias);
}
// The same with live accumulators (cz_world_add_forces): lastFrameAcceleration = Acceleration + forceAccum * inverseMass
// (:219-220), angularAcceleration = inverseInertiaTensorWorld (of the previous frame) * torqueAccum (:223).
CZD void integrate_body_forces(Integrated &o, const V3 &pos, const Q4 &q, const V3 &vel, const V3 &rot, const V3 &acc,
                               const M3 &iitBody, real motion, bool canSleep, real dt, real linPow, real angPow, real bias,
                               const V3 &force, const V3 &torque, real inverseMass, const M3 &iitWorldPrev) {
    o.lastAcc = acc;
    v_add_scaled(o.lastAcc, force, inverseMass);                                 // :220
    const V3 angAcc = m3_mul_v(iitWorldPrev, torque);                            // :223
    o.vel = vel;
    v_add_scaled(o.vel, o.lastAcc, dt);                                          // :227
    o.rot = rot;
    v_add_scaled(o.rot, angAcc, dt);                                             // :230
    integrate_body_tail(o, pos, q, iitBody, motion, canSleep, dt, linPow, angPow, bias);
}

}  // namespace czb
