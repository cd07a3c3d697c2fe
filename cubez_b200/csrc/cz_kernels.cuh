// cz_kernels.cuh — the __global__ kernels of libcubezcuda (sm_100a) for the multi-kernel
// world step: K1 integrate+derive, K3 narrowphase (count / emit with order-preserving
// compaction), K4 resolver, checksum/energy, and the pack/unpack kernels of the C ABI.
#pragma once
#include "cz_body.cuh"
#include "cz_narrow.cuh"
#include "cz_resolve.cuh"

namespace czk {
using namespace czm;
using czb::BodyStore;
using czn::ColliderView;
using czn::GenContact;
using czn::PlaneView;

#define CZ_MAX_PLANES 8

// Everything a world kernel needs, passed by value.
struct WorldParams {
    BodyStore st;
    int W, B, P, Cc;          // worlds, bodies per world, planes, contact capacity per world
    int wFirst, wCount;       // world range a fused launch works on (chunked host pipeline)
    const int *order;         // optional: order[wFirst + k] = k-th world to process (expensive worlds first), else identity
    int nchk;                 // checks per world
    int schedule;
    const int *chk_one, *chk_two;   // explicit schedule (shared by all worlds)
    PlaneView planes[CZ_MAX_PLANES];
    long long step_index;
    real *gen;                // as-generated contacts [G_NF][W*Cc]
    int *gb0, *gb1;           // [W*Cc]
    int *nContacts, *posIters, *velIters;   // [W]
    unsigned long long *stats;   // [0] contacts [1] pos iters [2] vel iters [3] checks [4] max contacts [5] status
    // RL-style episodes: world w is restored from `snap` at the start of every frame where
    // (phase0[w] + step - episodeStep0) % episodeLen == 0
    int episodeLen;
    long long episodeStep0;
    const int *phase0;
    BodyStore snap;
    // Per-pair surface materials (cz_world_set_materials).  matFric == NULL: the
    // reference's constants 0.9 / 0.1 (colliders.go:199-202 and the five other FIXME sites).
    const real *matFric, *matRest;   // [nMat * nMat], row = material of check operand `one`, column = `two`
    const uint8_t *bodyMat;          // [W * B]
    int nMat;
    uint8_t planeMat[CZ_MAX_PLANES];
};
// Friction / Restitution of the contacts produced by check (a, b) of the world whose body 0 is `base`
// (a, b: >= 0 collider, < 0 plane -(p+1), as the schedule names them)
CZD void check_material(const WorldParams &p, long long base, int a, int b, real &fric, real &rest) {
    if (!p.matFric) { fric = R_(0.9); rest = R_(0.1); return; }
    const int ma = a >= 0 ? (int)p.bodyMat[base + a] : (int)p.planeMat[-a - 1];
    const int mb = b >= 0 ? (int)p.bodyMat[base + b] : (int)p.planeMat[-b - 1];
    fric = p.matFric[ma * p.nMat + mb];
    rest = p.matRest[ma * p.nMat + mb];
}
CZD bool episode_wraps(const WorldParams &p, int w, long long step) {
    return p.episodeLen > 0 && (p.phase0[w] + (step - p.episodeStep0)) % p.episodeLen == 0;
}
enum { ST_CONTACTS = 0, ST_POS = 1, ST_VEL = 2, ST_CHECKS = 3, ST_MAXC = 4, ST_STATUS = 5, ST_N = 8 };

CZD int cz_popc(unsigned v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

__device__ __forceinline__ void raise_status(unsigned long long *stats, int code) {
    atomicCAS(&stats[ST_STATUS], 0ull, (unsigned long long)(unsigned)(-code));
}

// --------------------------------------------------------------------------------------
// K1  integrate_derive.   One thread per body, grid-stride, 128-bit accesses only.
// WORLD=false: the free-body form (cfg5): Integrate + CalculateDerivedData, nothing else.
// WORLD=true : also honours the per-body integrate/active flags of the world handle and
//              refreshes the collider transform (colliders.go:173-176/302-304), which the
//              reference's loop does right after Integrate, sleeping bodies included.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ real2 ldg_stream(const real2 *p) { return __ldcs(p); }
__device__ __forceinline__ void stg_stream(real2 *p, real2 v) { __stcs(p, v); }

template <bool WORLD>
__global__ void __launch_bounds__(256, 2) k_integrate(BodyStore s, real dt, real bias, long long step) {
    using namespace czb;
    const long long n = s.n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        bool awake = s.awake[i] != 0;
        bool run = true;
        if (WORLD) run = s.integ[i] != 0 && step >= (long long)s.active_from[i];
        if (!run) continue;
        M34 tr;
        bool haveTr = false;
        if (awake) {
            // 14 independent 128-bit loads, issued before any use
            real2 p01 = ldg_stream(s.chunk(C_P01) + i), p2m = ldg_stream(s.chunk(C_P2M) + i);
            real2 q01 = ldg_stream(s.chunk(C_Q01) + i), q23 = ldg_stream(s.chunk(C_Q23) + i);
            real2 v01 = ldg_stream(s.chunk(C_V01) + i), v2r0 = ldg_stream(s.chunk(C_V2R0) + i), r12 = ldg_stream(s.chunk(C_R12) + i);
            real2 a01 = ldg_stream(s.chunk(C_A01) + i), a2lp = ldg_stream(s.chunk(C_A2LP) + i), apw0 = ldg_stream(s.chunk(C_APW0) + i);
            real2 i12 = ldg_stream(s.chunk(C_I12) + i), i34 = ldg_stream(s.chunk(C_I34) + i), i56 = ldg_stream(s.chunk(C_I56) + i), i78 = ldg_stream(s.chunk(C_I78) + i);
            bool canSleep = s.can_sleep[i] != 0;
            V3 pos = mk3(p01.x, p01.y, p2m.x), vel = mk3(v01.x, v01.y, v2r0.x), rot = mk3(v2r0.y, r12.x, r12.y), acc = mk3(a01.x, a01.y, a2lp.x);
            Q4 q; q.c[0] = q01.x; q.c[1] = q01.y; q.c[2] = q23.x; q.c[3] = q23.y;
            M3 ib; ib.c[0] = apw0.y; ib.c[1] = i12.x; ib.c[2] = i12.y; ib.c[3] = i34.x; ib.c[4] = i34.y; ib.c[5] = i56.x; ib.c[6] = i56.y; ib.c[7] = i78.x; ib.c[8] = i78.y;
            Integrated o;
            if (s.force) {   // live accumulators (cz_world_add_forces): consumed and cleared by this Integrate (ClearAccumulators, rigidbody.go:206)
                const V3 f = mk3(s.force[i * 3], s.force[i * 3 + 1], s.force[i * 3 + 2]), tq = mk3(s.torque[i * 3], s.torque[i * 3 + 1], s.torque[i * 3 + 2]);
                integrate_body_forces(o, pos, q, vel, rot, acc, ib, p2m.y, canSleep, dt, a2lp.y, apw0.x, bias, f, tq, s.ld(C_MD, i).x, ld_iit_world(s, i));
#pragma unroll
                for (int k = 0; k < 3; k++) { s.force[i * 3 + k] = R_(0); s.torque[i * 3 + k] = R_(0); }
            } else {
                integrate_body(o, pos, q, vel, rot, acc, ib, p2m.y, canSleep, dt, a2lp.y, apw0.x, bias);
            }
            stg_stream(s.chunk(C_P01) + i, make_real2(o.pos.c[0], o.pos.c[1]));
            stg_stream(s.chunk(C_P2M) + i, make_real2(o.pos.c[2], o.motion));
            stg_stream(s.chunk(C_Q01) + i, make_real2(o.q.c[0], o.q.c[1]));
            stg_stream(s.chunk(C_Q23) + i, make_real2(o.q.c[2], o.q.c[3]));
            stg_stream(s.chunk(C_V01) + i, make_real2(o.vel.c[0], o.vel.c[1]));
            stg_stream(s.chunk(C_V2R0) + i, make_real2(o.vel.c[2], o.rot.c[0]));
            stg_stream(s.chunk(C_R12) + i, make_real2(o.rot.c[1], o.rot.c[2]));
            stg_stream(s.chunk(C_L01) + i, make_real2(o.lastAcc.c[0], o.lastAcc.c[1]));
            stg_stream(s.chunk(C_L2T0) + i, make_real2(o.lastAcc.c[2], o.transform.c[0]));
            stg_stream(s.chunk(C_T12) + i, make_real2(o.transform.c[1], o.transform.c[2]));
            stg_stream(s.chunk(C_T34) + i, make_real2(o.transform.c[3], o.transform.c[4]));
            stg_stream(s.chunk(C_T56) + i, make_real2(o.transform.c[5], o.transform.c[6]));
            stg_stream(s.chunk(C_T78) + i, make_real2(o.transform.c[7], o.transform.c[8]));
            stg_stream(s.chunk(C_T910) + i, make_real2(o.transform.c[9], o.transform.c[10]));
            stg_stream(s.chunk(C_T11W0) + i, make_real2(o.transform.c[11], o.iitWorld.c[0]));
            stg_stream(s.chunk(C_W12) + i, make_real2(o.iitWorld.c[1], o.iitWorld.c[2]));
            stg_stream(s.chunk(C_W34) + i, make_real2(o.iitWorld.c[3], o.iitWorld.c[4]));
            stg_stream(s.chunk(C_W56) + i, make_real2(o.iitWorld.c[5], o.iitWorld.c[6]));
            stg_stream(s.chunk(C_W78) + i, make_real2(o.iitWorld.c[7], o.iitWorld.c[8]));
            if (!o.awake) s.awake[i] = 0;
            if (WORLD) { tr = o.transform; haveTr = true; }
        }
        if (WORLD && s.shape[i] != CZ_SHAPE_NONE) {
            if (!haveTr) tr = ld_transform(s, i);
            M34 off = s.ident[i] ? identity34() : ld_m34(s, C_O01, i);
            st_m34(s, C_X01, i, m34_mul_m34(tr, off));
        }
    }
}

// Episode reset (multi-kernel path): restore the bodies of every world whose phase wraps.
__global__ void k_episode_reset(WorldParams p) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.st.n) return;
    int w = (int)(i / p.B);
    if (!episode_wraps(p, w, p.step_index)) return;
#pragma unroll 4
    for (int k = 0; k < czb::N_CHUNKS; k++) p.st.st(k, i, p.snap.ld(k, i));
    p.st.awake[i] = p.snap.awake[i];
    if (p.st.force) {   // forces still waiting in the accumulators do not survive a reset
#pragma unroll
        for (int k = 0; k < 3; k++) { p.st.force[i * 3 + k] = R_(0); p.st.torque[i * 3 + k] = R_(0); }
    }
}

// CalculateDerivedData for n bodies + collider transforms (upload with derive=1, and the
// cz_calculate_derived_data shim).  rigidbody.go:268-272, colliders.go:173-176.
__global__ void k_derive(BodyStore s, long long first, long long n, int bodies, int colliders) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    long long i = first + t;
    M34 tr;
    if (bodies) {
        V3 pos = ld_position(s, i);
        Q4 q = ld_orientation(s, i);
        M3 ib = ld_iit_body(s, i), iw;
        calculate_derived(pos, q, ib, tr, iw);
        s.st(C_Q01, i, make_real2(q.c[0], q.c[1]));
        s.st(C_Q23, i, make_real2(q.c[2], q.c[3]));
        real laz = s.ld(C_L2T0, i).x;
        st_derived(s, i, laz, tr, iw);
    } else {
        tr = ld_transform(s, i);
    }
    if (colliders && s.shape[i] != CZ_SHAPE_NONE) {
        M34 off = s.ident[i] ? identity34() : ld_m34(s, C_O01, i);
        st_m34(s, C_X01, i, m34_mul_m34(tr, off));
    }
}

// --------------------------------------------------------------------------------------
// pack / unpack between the ABI's "array of small vectors" and the chunked SoA.
// A field is a run of `comps` consecutive real slots starting at `first_slot`
// (slot = chunk*2 + lane).
// --------------------------------------------------------------------------------------
__global__ void k_pack(real2 *base, long long stride, long long first_body, long long n, const real *src, int first_slot, int comps) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * comps) return;
    long long i = t / comps;
    int sl = first_slot + (int)(t % comps);
    ((real *)(base + (long long)(sl >> 1) * stride))[2 * (first_body + i) + (sl & 1)] = src[t];
}
__global__ void k_unpack(const real2 *base, long long stride, long long first_body, long long n, real *dst, int first_slot, int comps) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * comps) return;
    long long i = t / comps;
    int sl = first_slot + (int)(t % comps);
    dst[t] = ((const real *)(base + (long long)(sl >> 1) * stride))[2 * (first_body + i) + (sl & 1)];
}
// ---- one-kernel pack / unpack of the per-frame host state (cz_world_step_host) ----------------
// Inputs: everything the host may have edited (the K1 read set); outputs: everything a frame
// writes.  Arrays are the ABI's flat "array of small vectors", device copies of the host's.
struct HostIn { const real *pos, *ori, *vel, *rot, *acc, *iitb, *motion; const uint8_t *awake, *can_sleep; };
struct HostOut { real *pos, *ori, *vel, *rot, *motion, *lacc, *tr, *iitw; uint8_t *awake; };

__global__ void k_pack_all(czb::BodyStore s, long long first, long long n, HostIn h) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = first + t;
    s.st(C_P01, i, make_real2(h.pos[i * 3], h.pos[i * 3 + 1]));
    s.st(C_P2M, i, make_real2(h.pos[i * 3 + 2], h.motion[i]));
    s.st(C_Q01, i, make_real2(h.ori[i * 4], h.ori[i * 4 + 1]));
    s.st(C_Q23, i, make_real2(h.ori[i * 4 + 2], h.ori[i * 4 + 3]));
    s.st(C_V01, i, make_real2(h.vel[i * 3], h.vel[i * 3 + 1]));
    s.st(C_V2R0, i, make_real2(h.vel[i * 3 + 2], h.rot[i * 3]));
    s.st(C_R12, i, make_real2(h.rot[i * 3 + 1], h.rot[i * 3 + 2]));
    s.st(C_A01, i, make_real2(h.acc[i * 3], h.acc[i * 3 + 1]));
    s.st(C_A2LP, i, make_real2(h.acc[i * 3 + 2], s.ld(C_A2LP, i).y));        // keeps linPow
    s.st(C_APW0, i, make_real2(s.ld(C_APW0, i).x, h.iitb[i * 9]));           // keeps angPow
    s.st(C_I12, i, make_real2(h.iitb[i * 9 + 1], h.iitb[i * 9 + 2]));
    s.st(C_I34, i, make_real2(h.iitb[i * 9 + 3], h.iitb[i * 9 + 4]));
    s.st(C_I56, i, make_real2(h.iitb[i * 9 + 5], h.iitb[i * 9 + 6]));
    s.st(C_I78, i, make_real2(h.iitb[i * 9 + 7], h.iitb[i * 9 + 8]));
    s.awake[i] = h.awake[i];
    s.can_sleep[i] = h.can_sleep[i];
}
// every output array is optional (NULL = not wanted): cz_world_step_host asks for all of them,
// cz_world_step_rl for the observation the caller named
__global__ void k_unpack_all(czb::BodyStore s, long long first, long long n, HostOut h) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = first + t;
    if (h.pos) { V3 v = ld_position(s, i); for (int k = 0; k < 3; k++) h.pos[i * 3 + k] = v.c[k]; }
    if (h.vel) { V3 v = ld_velocity(s, i); for (int k = 0; k < 3; k++) h.vel[i * 3 + k] = v.c[k]; }
    if (h.rot) { V3 v = ld_rotation(s, i); for (int k = 0; k < 3; k++) h.rot[i * 3 + k] = v.c[k]; }
    if (h.lacc) { V3 v = ld_last_acc(s, i); for (int k = 0; k < 3; k++) h.lacc[i * 3 + k] = v.c[k]; }
    if (h.ori) { Q4 q = ld_orientation(s, i); for (int k = 0; k < 4; k++) h.ori[i * 4 + k] = q.c[k]; }
    if (h.tr) {
        M34 tr = ld_transform(s, i);
#pragma unroll
        for (int k = 0; k < 12; k++) h.tr[i * 12 + k] = tr.c[k];
    }
    if (h.iitw) {
        M3 iw = ld_iit_world(s, i);
#pragma unroll
        for (int k = 0; k < 9; k++) h.iitw[i * 9 + k] = iw.c[k];
    }
    if (h.motion) h.motion[i] = s.ld(C_P2M, i).y;
    if (h.awake) h.awake[i] = s.awake[i];
}

// Batched RigidBody.AddVelocity / AddRotation (rigidbody.go:195-202): Velocity.Add(v), Rotation.Add(v),
// one rounding per component, applied to every body of the range (sleeping bodies are not woken — the
// reference does not either).  Either array may be NULL.
__global__ void k_apply_actions(czb::BodyStore s, long long first, long long n, const real *addVel, const real *addRot) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = first + t;
    real2 v01 = s.ld(C_V01, i), v2r0 = s.ld(C_V2R0, i), r12 = s.ld(C_R12, i);
    if (addVel) { v01.x = v01.x + addVel[i * 3]; v01.y = v01.y + addVel[i * 3 + 1]; v2r0.x = v2r0.x + addVel[i * 3 + 2]; }
    if (addRot) { v2r0.y = v2r0.y + addRot[i * 3]; r12.x = r12.x + addRot[i * 3 + 1]; r12.y = r12.y + addRot[i * 3 + 2]; }
    s.st(C_V01, i, v01); s.st(C_V2R0, i, v2r0); s.st(C_R12, i, r12);
}

// Force / torque accumulator input: forceAccum += f, torqueAccum += t per body, one rounding per
// component — cyclone's addForce / addTorque, which the reference keeps the accumulators for (rigidbody.go:86-92) but
// never exposes.  The next Integrate of an awake body consumes and clears them (:219-223, :206).
__global__ void k_add_forces(czb::BodyStore s, long long first, long long n, const real *force, const real *torque) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * 3) return;
    const long long i = first * 3 + t;
    if (force) s.force[i] = s.force[i] + force[t];
    if (torque) s.torque[i] = s.torque[i] + torque[t];
}

// Renderer-side export: what the example loop does per body and frame on the host,
// SetGlVector3(&Node.Location, &body.Position) / SetGlQuat(&Node.LocalRotation, &body.Orientation)
// (examples/cubedrop.go:35-37, examples/exampleapp.go:146-159): float32(x) of every component, quaternion as
// (W, V[0], V[1], V[2]).  model (optional): the body transform (rigidbody.go:88) as a column-major 4x4.
__global__ void k_export_gl(czb::BodyStore s, long long first, long long n, float *loc, float *rot, float *model) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = first + t;
    if (loc) { V3 v = ld_position(s, i); for (int k = 0; k < 3; k++) loc[t * 3 + k] = (float)v.c[k]; }
    if (rot) { Q4 q = ld_orientation(s, i); for (int k = 0; k < 4; k++) rot[t * 4 + k] = (float)q.c[k]; }
    if (model) {
        M34 tr = ld_transform(s, i);
#pragma unroll
        for (int c = 0; c < 4; c++) {
#pragma unroll
            for (int r = 0; r < 3; r++) model[t * 16 + c * 4 + r] = (float)tr.c[c * 3 + r];
            model[t * 16 + c * 4 + 3] = c == 3 ? 1.0f : 0.0f;
        }
    }
}

__global__ void k_fill_u8(uint8_t *p, long long n, uint8_t v) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}
__global__ void k_fill_i32(int *p, long long n, int v) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) p[t] = v;
}
// ident[i] = 1 when the collider Offset is exactly the identity matrix
__global__ void k_detect_identity(BodyStore s, long long first, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    M34 o = czb::ld_m34(s, czb::C_O01, first + t);
    bool id = true;
#pragma unroll
    for (int k = 0; k < 12; k++) id = id && (o.c[k] == ((k == 0 || k == 4 || k == 8) ? R_(1) : R_(0)));
    s.ident[first + t] = id ? 1 : 0;
}

// --------------------------------------------------------------------------------------
// K3  narrowphase over the world's check schedule with order-preserving compaction.
// --------------------------------------------------------------------------------------
struct CheckEval {
    int count;          // contacts this check produces (0..8)
    unsigned mask;      // cube-plane: vertex mask
    int kind;           // 0 none, 1 single contact in gc, 2 cube-plane (use mask)
    GenContact gc;
    int cubeLocal;      // for kind 2
    int plane;
};

CZD ColliderView load_collider(const BodyStore &s, long long gi, int local) {
    ColliderView v;
    v.shape = s.shape[gi];
    v.body = local;
    v.t = czb::ld_m34(s, czb::C_X01, gi);
    real2 h01 = s.ld(czb::C_H01, gi), h2r = s.ld(czb::C_H2R, gi);
    v.half = mk3(h01.x, h01.y, h2r.x);
    v.radius = h2r.y;
    return v;
}

// decode check k of the world's schedule into (a, b): >= 0 collider, < 0 plane -(p+1); returns
// false when the slot holds no check (i == j).
CZD bool decode_check(const WorldParams &p, int k, int &a, int &b) {
    if (p.schedule == CZ_SCHED_ALL_PAIRS_ORDERED) {
        int per = p.P + p.B;
#ifdef __CUDA_ARCH__
        // k / per without the ~20-instruction integer division: (k + 0.5) / per is at least 0.5 / per away from an
        // integer, far more than the float rounding for k < 2^16, per < 2^10 (exact; checked exhaustively in the tests)
        int i = (k < 65536 && per < 1024) ? (int)(((float)k + 0.5f) * __frcp_rn((float)per)) : k / per;
        int slot = k - i * per;
#else
        int i = k / per, slot = k - i * per;
#endif
        a = i;
        if (slot < p.P) { b = -(slot + 1); return true; }
        b = slot - p.P;
        return b != i;
    }
    a = p.chk_one[k];
    b = p.chk_two[k];
    return true;
}

CZD bool body_active(const WorldParams &p, long long gi) {
    return p.step_index >= (long long)p.st.active_from[gi] && p.st.shape[gi] != CZ_SHAPE_NONE;
}

// CheckForCollisions (colliders.go:720-747) for check (a, b) of world with body base `base`.
CZD void eval_check(const WorldParams &p, long long base, int a, int b, CheckEval &e) {
    e.count = 0; e.kind = 0; e.mask = 0;
    if (a < 0 && b < 0) return;                                  // plane-plane :111-113
    if (a >= 0 && !body_active(p, base + a)) return;
    if (b >= 0 && !body_active(p, base + b)) return;
    if (a < 0 || b < 0) {
        int ci = a < 0 ? b : a, pi = a < 0 ? -a - 1 : -b - 1;
        ColliderView c = load_collider(p.st, base + ci, ci);
        const PlaneView &pl = p.planes[pi];
        if (c.shape == CZ_SHAPE_SPHERE) {
            if (czn::sphere_halfspace(c, pl, e.gc)) { e.count = 1; e.kind = 1; }
        } else if (c.shape == CZ_SHAPE_CUBE) {
            e.mask = czn::cube_halfspace_mask(c, pl);
            if (e.mask) { e.count = cz_popc(e.mask); e.kind = 2; e.cubeLocal = ci; e.plane = pi; }
        }
        return;
    }
    ColliderView one = load_collider(p.st, base + a, a), two = load_collider(p.st, base + b, b);
    V3 v1 = zero3(), v2 = zero3();
    if (one.shape != two.shape) {   // cube-sphere fallback normal needs the sphere body's velocity
        v1 = czb::ld_velocity(p.st, base + a);
        v2 = czb::ld_velocity(p.st, base + b);
    }
    if (czn::check_pair(one, two, v1, v2, e.gc)) { e.count = 1; e.kind = 1; }
}

CZD void store_gen(real *gen, long long gs, int *gb0, int *gb1, long long slot, const GenContact &c, real fric, real rest) {
    using namespace czr;
#pragma unroll
    for (int k = 0; k < 3; k++) { gen[(G_POINT + k) * gs + slot] = c.point.c[k]; gen[(G_NORMAL + k) * gs + slot] = c.normal.c[k]; }
    gen[G_PEN * gs + slot] = c.pen;
    gen[G_FRIC * gs + slot] = fric;   // 0.9 / 0.1 (the test constants of colliders.go:201-202 etc.) unless materials are set
    gen[G_REST * gs + slot] = rest;
    gb0[slot] = c.b0;
    gb1[slot] = c.b1;
}

// block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32); returns the exclusive
// prefix and the block total.  warpTotals: shared int[32].
__device__ __forceinline__ int block_exclusive_scan(int v, int *warpTotals, int &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpTotals[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nw ? warpTotals[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warpTotals[lane] = wi - w;   // exclusive warp offsets
        if (lane == 31) warpTotals[32] = wi;
    }
    __syncthreads();
    total = warpTotals[32];
    int r = warpTotals[warp] + incl - v;
    __syncthreads();
    return r;
}

enum { NARROW_SINGLE = 0, NARROW_COUNT = 1, NARROW_EMIT = 2 };

// One CTA per (world, tile of blockDim.x checks).
//  SINGLE: the world's whole schedule fits one tile: evaluate, scan, emit.
//  COUNT : evaluate, store per-check hit counts and the tile total.
//  EMIT  : re-evaluate only the checks that hit, emit at tileBase + in-tile prefix.
template <int MODE>
__global__ void __launch_bounds__(256) k_narrow(WorldParams p, int tiles, int *tileCount, const int *tileBase, uint8_t *hitCount) {
    __shared__ int warpTotals[33];
    const int w = blockIdx.x / tiles, t = blockIdx.x - w * tiles;
    const int k = t * blockDim.x + threadIdx.x;
    const long long base = (long long)w * p.B;
    CheckEval e;
    e.count = 0; e.kind = 0; e.mask = 0;
    int a = 0, b = 0;
    bool valid = k < p.nchk && decode_check(p, k, a, b);
    if (MODE == NARROW_EMIT) {
        if (valid && hitCount[(long long)w * p.nchk + k] != 0) eval_check(p, base, a, b, e);
    } else if (valid) {
        eval_check(p, base, a, b, e);
    }
    if (MODE == NARROW_COUNT) {
        if (k < p.nchk) hitCount[(long long)w * p.nchk + k] = (uint8_t)e.count;
    }
    int total;
    int off = block_exclusive_scan(e.count, warpTotals, total);
    if (MODE == NARROW_COUNT) {
        if (threadIdx.x == 0) tileCount[blockIdx.x] = total;
        return;
    }
    int tb = MODE == NARROW_EMIT ? tileBase[blockIdx.x] : 0;
    if (MODE == NARROW_SINGLE && threadIdx.x == 0) {
        p.nContacts[w] = total;
        if (total > p.Cc) raise_status(p.stats, CZ_ERR_CAPACITY);
        atomicAdd(&p.stats[ST_CONTACTS], (unsigned long long)total);
        atomicMax(&p.stats[ST_MAXC], (unsigned long long)total);
    }
    if (e.count == 0) return;
    const long long gs = (long long)p.W * p.Cc;
    real *gen = p.gen + (long long)w * p.Cc;
    int *gb0 = p.gb0 + (long long)w * p.Cc, *gb1 = p.gb1 + (long long)w * p.Cc;
    int slot = tb + off;
    real fric, rest;
    check_material(p, base, a, b, fric, rest);
    if (e.kind == 1) {
        if (slot < p.Cc) store_gen(gen, gs, gb0, gb1, slot, e.gc, fric, rest);
    } else {
        ColliderView c = load_collider(p.st, base + e.cubeLocal, e.cubeLocal);
#pragma unroll 1
        for (int v = 0; v < 8; v++) {
            if (e.mask & (1u << v)) {
                GenContact gc;
                czn::cube_halfspace_contact(c, p.planes[e.plane], v, gc);
                if (slot < p.Cc) store_gen(gen, gs, gb0, gb1, slot, gc, fric, rest);
                slot++;
            }
        }
    }
}

// Per world: exclusive scan of its tile totals -> tileBase, nContacts.  One CTA per world.
__global__ void __launch_bounds__(256) k_scan_tiles(WorldParams p, int tiles, const int *tileCount, int *tileBase) {
    __shared__ int warpTotals[33];
    const int w = blockIdx.x;
    int running = 0;
    for (int t0 = 0; t0 < tiles; t0 += blockDim.x) {
        int t = t0 + threadIdx.x;
        int v = t < tiles ? tileCount[(long long)w * tiles + t] : 0;
        int total;
        int off = block_exclusive_scan(v, warpTotals, total);
        if (t < tiles) tileBase[(long long)w * tiles + t] = running + off;
        running += total;
    }
    if (threadIdx.x == 0) {
        p.nContacts[w] = running;
        if (running > p.Cc) raise_status(p.stats, CZ_ERR_CAPACITY);
        atomicAdd(&p.stats[ST_CONTACTS], (unsigned long long)running);
        atomicMax(&p.stats[ST_MAXC], (unsigned long long)running);
    }
}

// --------------------------------------------------------------------------------------
// K4  resolver: one thread group (CTA of NT threads) per world.
// Work records live in dynamic shared memory when they fit (useSmem), otherwise in the global
// scratch `gscratch` (L2-resident for the sizes involved).
// --------------------------------------------------------------------------------------
// contact work reals per contact in k_resolve's staging: 18 cold + pen, ddv, fric, rest
#define CW_NREAL (czr::CW_NCOLD + 4)
struct ResolveScratch {
    real *bw;   // [W][BW_NF*B]
    real *cw;   // [W][CW_NREAL*Cc]
    int *cb;    // [W][2*Cc]
    real *pre;  // [W][VP_NF*Cc] per-contact velocity response of the large-world loop, or NULL
    unsigned short *adj;   // [W][2*Cc] adjacency lists (body -> contacts) of the large-world loop, or NULL
};

CZD void load_body_work(const czr::Ctx &x, const BodyStore &s, long long gi, int b) {
    using namespace czr;
    V3 pos = czb::ld_position(s, gi), vel = czb::ld_velocity(s, gi), rot = czb::ld_rotation(s, gi), la = czb::ld_last_acc(s, gi);
    Q4 q = czb::ld_orientation(s, gi);
    M3 iw = czb::ld_iit_world(s, gi);
    bw3_set(x, BW_POS, b, pos); bw3_set(x, BW_VEL, b, vel); bw3_set(x, BW_ROT, b, rot); bw3_set(x, BW_LACC, b, la);
#pragma unroll
    for (int k = 0; k < 4; k++) x.bw[(BW_Q + k) * x.bs + b] = q.c[k];
#pragma unroll
    for (int k = 0; k < 9; k++) x.bw[(BW_IITW + k) * x.bs + b] = iw.c[k];
    x.bw[BW_INVM * x.bs + b] = s.ld(czb::C_MD, gi).x;
    x.bw[BW_MOTION * x.bs + b] = s.ld(czb::C_P2M, gi).y;
    x.bw[BW_AWAKE * x.bs + b] = s.awake[gi] ? R_(1) : R_(0);
}
CZD void store_body_work(const czr::Ctx &x, const BodyStore &s, long long gi, int b) {
    using namespace czr;
    V3 pos = bw3(x, BW_POS, b), vel = bw3(x, BW_VEL, b), rot = bw3(x, BW_ROT, b);
    s.st(czb::C_P01, gi, make_real2(pos.c[0], pos.c[1]));
    s.st(czb::C_P2M, gi, make_real2(pos.c[2], x.bw[BW_MOTION * x.bs + b]));
    s.st(czb::C_Q01, gi, make_real2(x.bw[(BW_Q + 0) * x.bs + b], x.bw[(BW_Q + 1) * x.bs + b]));
    s.st(czb::C_Q23, gi, make_real2(x.bw[(BW_Q + 2) * x.bs + b], x.bw[(BW_Q + 3) * x.bs + b]));
    s.st(czb::C_V01, gi, make_real2(vel.c[0], vel.c[1]));
    s.st(czb::C_V2R0, gi, make_real2(vel.c[2], rot.c[0]));
    s.st(czb::C_R12, gi, make_real2(rot.c[1], rot.c[2]));
    s.awake[gi] = x.bw[BW_AWAKE * x.bs + b] != R_(0) ? 1 : 0;
}

template <int NT>
__global__ void __launch_bounds__(NT) k_resolve(WorldParams p, ResolveScratch rs, int useSmem, int maxIterOverride, real dt, int hotCap) {
    using namespace czr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GroupScratch gs;
    const int w = blockIdx.x, tid = threadIdx.x;
    int nC = p.nContacts[w];
    if (nC <= 0 || nC > p.Cc) {   // no contact: ResolveContacts is not called (cubedrop.go:72); overflow: error already raised
        if (tid == 0) { p.posIters[w] = 0; p.velIters[w] = 0; }
        return;
    }
    Ctx x;
    x.bs = p.B; x.nC = nC; x.dt = dt; x.store = p.st; x.body_base = (long long)w * p.B;
    x.xb = nullptr; x.xbs = 0; x.mlist = nullptr; x.bmask = nullptr;
    real *cwbase;
    if (useSmem == 1) {
        x.bw = (real *)smem_raw;
        cwbase = x.bw + BW_NF * p.B;
        x.cb0 = (int *)(cwbase + CW_NREAL * p.Cc);
    } else {
        x.bw = rs.bw + (long long)w * BW_NF * p.B;
        cwbase = rs.cw + (long long)w * CW_NREAL * p.Cc;
        x.cb0 = rs.cb + (long long)w * 2 * p.Cc;
    }
    x.cb1 = x.cb0 + p.Cc;
    x.cold = cwbase; x.cfs = p.Cc; x.ccs = 1;                 // SoA
    x.pen = cwbase + (size_t)CW_NCOLD * p.Cc; x.ddv = x.pen + p.Cc; x.fric = x.ddv + p.Cc; x.rest = x.fric + p.Cc;
    for (int b = tid; b < p.B; b += NT) load_body_work(x, p.st, x.body_base + b, b);
    __syncthreads();
    const long long gstride = (long long)p.W * p.Cc;
    GenView g;
    g.pn = p.gen + (long long)w * p.Cc; g.fs = (int)gstride; g.cs = 1;
    g.pen = p.gen + G_PEN * gstride + (long long)w * p.Cc;
    g.fric = p.gen + G_FRIC * gstride + (long long)w * p.Cc;
    g.rest = p.gen + G_REST * gstride + (long long)w * p.Cc;
    g.b0 = p.gb0 + (long long)w * p.Cc; g.b1 = p.gb1 + (long long)w * p.Cc;
    for (int c = tid; c < nC; c += NT) prepare_contact(x, c, g);
    __syncthreads();
    int maxIter = maxIterOverride >= 0 ? maxIterOverride : nC * 8;   // cubedrop.go:73
    int status = 0;
    int pi, vi;
    if (NT == 32) {
        pi = resolve_loop<32, false>(x, true, maxIter, tid, &status);
        __syncthreads();
        vi = resolve_loop<32, true>(x, true, maxIter, tid, &status);
    } else if ((useSmem == 3 || useSmem == 4) && nC <= hotCap) {   // one large world: adjacency lists + shared arg-max caches + prefetched propagation (4: list offsets in global memory too)
        __shared__ BigBroadcast bb;
        __shared__ int scanScratch[NT / 32 + 1];
        unsigned short *adjW = rs.adj + (size_t)w * (2 * (size_t)p.Cc + p.B + 2);
        const BigShared sh = big_carve(smem_raw, (NT > 32 ? NT : 64), hotCap, p.B, adjW, useSmem == 4 ? adjW + 2 * (size_t)p.Cc : nullptr);
        big_build_adjacency<(NT > 32 ? NT : 64)>(x, p.B, tid, sh, scanScratch);
        real *velPre = rs.pre ? rs.pre + (size_t)w * czr::VP_NF * p.Cc : nullptr;
        pi = resolve_loop_big<(NT > 32 ? NT : 64), false>(x, maxIter, &gs, &bb, tid, &status, sh, nullptr);
        __syncthreads();
        vi = resolve_loop_big<(NT > 32 ? NT : 64), true>(x, maxIter, &gs, &bb, tid, &status, sh, velPre);
    } else if (useSmem == 2 && nC <= hotCap) {   // one large world: hot value + 16-bit body ids in shared memory, cached arg-max
        real *sHot = (real *)smem_raw;
        unsigned short *sB0 = (unsigned short *)(sHot + hotCap), *sB1 = sB0 + hotCap;
        pi = resolve_loop_cta_cached<(NT > 32 ? NT : 64), false>(x, maxIter, &gs, tid, &status, sHot, sB0, sB1);
        __syncthreads();
        vi = resolve_loop_cta_cached<(NT > 32 ? NT : 64), true>(x, maxIter, &gs, tid, &status, sHot, sB0, sB1);
    } else {
        pi = resolve_loop_cta<(NT > 32 ? NT : 64), false>(x, maxIter, &gs, tid, &status);
        __syncthreads();
        vi = resolve_loop_cta<(NT > 32 ? NT : 64), true>(x, maxIter, &gs, tid, &status);
    }
    __syncthreads();
    for (int b = tid; b < p.B; b += NT) store_body_work(x, p.st, x.body_base + b, b);
    if (tid == 0) {
        p.posIters[w] = pi; p.velIters[w] = vi;
        atomicAdd(&p.stats[ST_POS], (unsigned long long)pi);
        atomicAdd(&p.stats[ST_VEL], (unsigned long long)vi);
        if (status) raise_status(p.stats, status);
    }
}

// --------------------------------------------------------------------------------------
// Contact islands of ONE large world (north star: "one CTA per world or contact island").
// Contacts that share no body — directly or through a chain of contacts — never read each other's writes, and the
// worst-first loop picks, inside an island, exactly the subsequence of the global pick order that belongs to it (the
// global arg-max restricted to an island is the island's arg-max; ties go to the lowest GLOBAL index either way).
// So each island may run its own loop, concurrently, and every body / contact ends in the state the reference's single
// loop leaves — PROVIDED the reference's loop runs to convergence, i.e. the sum of the islands' iterations stays
// within its cap of 8*len(contacts) per phase (examples/cubedrop.go:73).  When the cap cuts the global loop mid-way,
// which island had its turn last depends on the interleaving: that case is DETECTED here (sums against the cap) and the
// host re-runs the frame's resolve on the exact single-CTA path from a snapshot.  Connectivity runs through every
// non-nil body, static ones included (a zero-mass body still carries the IsAwake flag the velocity loop re-reads,
// contact.go:438); plane contacts link nothing (hard part 3).
//   k_island_labels   connected components by min-label propagation with pointer jumping (one CTA, labels in shared memory)
//   radix sort        contacts by island label, stable: ascending contact index inside an island (cz_sort.cuh)
//   k_island_ranges   island boundaries in the sorted list
//   k_resolve_islands one CTA per island on a renumbered, order-preserving slice of the contact arrays
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_island_labels(WorldParams p, unsigned long long *keys, unsigned *vals) {
    extern __shared__ int label[];   // [B]
    const int nC = min(p.nContacts[0], p.Cc);
    for (int b = threadIdx.x; b < p.B; b += blockDim.x) label[b] = b;
    __syncthreads();
    for (int round = 0; round < 4096; round++) {
        int changed = 0;
        for (int c = threadIdx.x; c < nC; c += blockDim.x) {
            const int b0 = p.gb0[c], b1 = p.gb1[c];
            if (b0 < 0 || b1 < 0) continue;
            const int l0 = label[b0], l1 = label[b1];
            if (l0 != l1) {
                const int m = min(l0, l1);
                atomicMin(&label[b0], m);
                atomicMin(&label[b1], m);
                atomicMin(&label[max(l0, l1)], m);   // hook the larger root too: components merge in O(log) rounds
                changed = 1;
            }
        }
        __syncthreads();
        for (int b = threadIdx.x; b < p.B; b += blockDim.x) {   // pointer jumping
            int l = label[b];
            while (label[l] != l) l = label[l];
            label[b] = l;
        }
        if (!__syncthreads_or(changed)) break;
    }
    for (int c = threadIdx.x; c < nC; c += blockDim.x) {
        const int b0 = p.gb0[c], b1 = p.gb1[c];
        keys[c] = (unsigned long long)label[b0 >= 0 ? b0 : b1];
        vals[c] = (unsigned)c;
    }
}
struct IslandTable {
    int *start;      // [Cc + 1] first sorted position of island k
    int *count;      // [4]: number of islands | sum of position iterations | sum of velocity iterations | largest island
};
// one CTA: flags -> exclusive scan -> island starts
__global__ void __launch_bounds__(1024) k_island_ranges(WorldParams p, const unsigned long long *sortedKeys, IslandTable t) {
    __shared__ int warpTotals[33];
    const int nC = min(p.nContacts[0], p.Cc);
    int running = 0;
    for (int c0 = 0; c0 < nC; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        const int flag = (c < nC && (c == 0 || sortedKeys[c] != sortedKeys[c - 1])) ? 1 : 0;
        int total;
        const int off = block_exclusive_scan(flag, warpTotals, total);
        if (flag) t.start[running + off] = c;
        running += total;
    }
    if (threadIdx.x == 0) { t.start[running] = nC; t.count[0] = running; t.count[1] = 0; t.count[2] = 0; t.count[3] = 0; }
}
__global__ void k_bw_load(WorldParams p, ResolveScratch rs) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    czr::Ctx x;
    x.bw = rs.bw; x.bs = p.B;
    load_body_work(x, p.st, b, b);
}
__global__ void k_bw_store(WorldParams p, ResolveScratch rs, IslandTable t) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) {
        p.posIters[0] = t.count[1]; p.velIters[0] = t.count[2];
        atomicAdd(&p.stats[ST_POS], (unsigned long long)t.count[1]);
        atomicAdd(&p.stats[ST_VEL], (unsigned long long)t.count[2]);
    }
    if (b >= p.B) return;
    czr::Ctx x;
    x.bw = rs.bw; x.bs = p.B;
    store_body_work(x, p.st, b, b);
}
template <int NT>
__global__ void __launch_bounds__(NT) k_resolve_islands(WorldParams p, ResolveScratch rs, const unsigned *perm, IslandTable t, real dt, int hotCap) {
    using namespace czr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ GroupScratch gs;
    __shared__ BigBroadcast bb;
    __shared__ int scanScratch[NT / 32 + 1];
    const int tid = threadIdx.x;
    const int nIslands = t.count[0], nAll = min(p.nContacts[0], p.Cc);
    for (int isl = blockIdx.x; isl < nIslands; isl += gridDim.x) {
        const int s = t.start[isl], nC = t.start[isl + 1] - s;
        Ctx x;
        x.bs = p.B; x.nC = nC; x.dt = dt; x.store = p.st; x.body_base = 0;
        x.xb = nullptr; x.xbs = 0; x.mlist = nullptr; x.bmask = nullptr;
        x.bw = rs.bw;
        real *cwbase = rs.cw + s;                               // SoA over the whole capacity: a slice keeps the field stride
        x.cb0 = rs.cb + s; x.cb1 = rs.cb + p.Cc + s;
        x.cold = cwbase; x.cfs = p.Cc; x.ccs = 1;
        x.pen = cwbase + (size_t)CW_NCOLD * p.Cc; x.ddv = x.pen + p.Cc; x.fric = x.ddv + p.Cc; x.rest = x.fric + p.Cc;
        const long long gstride = (long long)p.W * p.Cc;
        GenView g;
        g.pn = p.gen; g.fs = (int)gstride; g.cs = 1;
        g.pen = p.gen + G_PEN * gstride; g.fric = p.gen + G_FRIC * gstride; g.rest = p.gen + G_REST * gstride;
        g.b0 = p.gb0; g.b1 = p.gb1;
        for (int j = tid; j < nC; j += NT) prepare_contact(x, j, g, (int)perm[s + j]);   // slot j <- the island's j-th contact in global order
        __syncthreads();
        const int maxIter = nAll * 8;          // the reference's cap is global (examples/cubedrop.go:73); the sums are checked against it afterwards
        int status = 0, pi, vi;
        if (nC <= hotCap) {
            const BigShared sh = big_carve(smem_raw, NT, hotCap, p.B, rs.adj + (size_t)2 * s);   // (W = 1 on this path)
            big_build_adjacency<NT>(x, p.B, tid, sh, scanScratch);
            real *velPre = rs.pre ? rs.pre + (size_t)s * VP_NF : nullptr;
            pi = resolve_loop_big<NT, false>(x, maxIter, &gs, &bb, tid, &status, sh, nullptr);
            __syncthreads();
            vi = resolve_loop_big<NT, true>(x, maxIter, &gs, &bb, tid, &status, sh, velPre);
        } else {
            pi = resolve_loop_cta<NT, false>(x, maxIter, &gs, tid, &status);
            __syncthreads();
            vi = resolve_loop_cta<NT, true>(x, maxIter, &gs, tid, &status);
        }
        __syncthreads();
        if (tid == 0) {
            atomicAdd(&t.count[1], pi);
            atomicAdd(&t.count[2], vi);
            atomicMax(&t.count[3], nC);
            if (status) raise_status(p.stats, status);
        }
        __syncthreads();
    }
}

// After cz_resolve_contacts: copy the resolver's view of the contacts (possibly swapped
// bodies / negated normal, updated penetration) back into the as-generated arrays so the
// shim can mirror contact.go:61-65 and :274 into the caller's Contact structs.
__global__ void k_contacts_writeback(WorldParams p, ResolveScratch rs) {
    using namespace czr;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int nC = p.nContacts[0];
    if (c >= nC) return;
    long long gs = (long long)p.W * p.Cc;
#pragma unroll
    for (int k = 0; k < 3; k++) p.gen[(G_NORMAL + k) * gs + c] = rs.cw[(CW_N + k) * p.Cc + c];
    p.gen[G_PEN * gs + c] = rs.cw[CW_NCOLD * p.Cc + c];
    p.gb0[c] = rs.cb[c];
    p.gb1[c] = rs.cb[p.Cc + c];
}

// --------------------------------------------------------------------------------------
// K6  checksum + energy per world (new definitions, no reference counterpart).
// --------------------------------------------------------------------------------------
__device__ __forceinline__ void fnv_eat(unsigned long long &h, real v) {
#ifdef CUBEZ_REAL_FLOAT
    unsigned int bits = __float_as_uint(v);
    for (int k = 0; k < 4; k++) { h ^= (bits >> (8 * k)) & 0xffu; h *= 0x100000001b3ull; }
#else
    unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    for (int k = 0; k < 8; k++) { h ^= (bits >> (8 * k)) & 0xffull; h *= 0x100000001b3ull; }
#endif
}
__global__ void k_checksum_energy(BodyStore s, int W, int B, unsigned long long *hash, double *energy) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    unsigned long long h = 0xcbf29ce484222325ull;
    double e = 0.0;
    for (int b = 0; b < B; b++) {
        long long i = (long long)w * B + b;
        V3 pos = czb::ld_position(s, i), vel = czb::ld_velocity(s, i), rot = czb::ld_rotation(s, i);
        Q4 q = czb::ld_orientation(s, i);
        for (int k = 0; k < 3; k++) fnv_eat(h, pos.c[k]);
        for (int k = 0; k < 4; k++) fnv_eat(h, q.c[k]);
        for (int k = 0; k < 3; k++) fnv_eat(h, vel.c[k]);
        for (int k = 0; k < 3; k++) fnv_eat(h, rot.c[k]);
        h ^= (unsigned long long)(s.awake[i] ? 1 : 0);
        h *= 0x100000001b3ull;
        real invMass = s.ld(czb::C_MD, i).x;
        if (invMass > R_(0)) {
            double m = 1.0 / (double)invMass;
            real2 a01 = s.ld(czb::C_A01, i), a2 = s.ld(czb::C_A2LP, i);
            V3 acc = mk3(a01.x, a01.y, a2.x);
            M3 iw = m3_invert(czb::ld_iit_world(s, i));
            V3 Iw = m3_mul_v(iw, rot);
            e += 0.5 * m * (double)v_dot(vel, vel) + 0.5 * (double)v_dot(rot, Iw) - m * (double)v_dot(acc, pos);
        }
    }
    hash[w] = h;
    energy[w] = e;
}

// NaN / overflow watch: the reference propagates NaN silently (0/0 when both bodies of a contact are
// static, contact.go:286-386; a zero fallback normal in cube-sphere, colliders.go:417-421).  Counts the bodies whose
// position, orientation, velocity or rotation holds a non-finite component.
__global__ void k_count_nonfinite(BodyStore s, long long n, unsigned long long *count) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n) {
        V3 pos = czb::ld_position(s, i), vel = czb::ld_velocity(s, i), rot = czb::ld_rotation(s, i);
        Q4 q = czb::ld_orientation(s, i);
        real sum = R_(0);
        for (int k = 0; k < 3; k++) sum += pos.c[k] * R_(0) + vel.c[k] * R_(0) + rot.c[k] * R_(0);
        for (int k = 0; k < 4; k++) sum += q.c[k] * R_(0);
        bad = !(sum == R_(0));      // x * 0 is NaN exactly when x is NaN or infinite
    }
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(count, (unsigned long long)__popc(m));
}

// --------------------------------------------------------------------------------------
// cfg5 initial state on the device: the same splitmix64 stream as scenes.free_bodies().
// --------------------------------------------------------------------------------------
__device__ __forceinline__ double sm64_next(unsigned long long &state) {
    state += 0x9E3779B97F4A7C15ull;
    unsigned long long z = state;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double uni(double u, double a, double b) { return a + (b - a) * u; }

__global__ void k_init_free_bodies(BodyStore s, unsigned long long seed, real dt) {
    using namespace czb;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    unsigned long long st = seed * 0x100000001B3ull + (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    double u[21];
    for (int k = 0; k < 21; k++) u[k] = sm64_next(st);
    V3 pos = mk3((real)uni(u[0], -100, 100), (real)uni(u[1], -100, 100), (real)uni(u[2], -100, 100));
    double qd[4] = {uni(u[3], -1, 1), uni(u[4], -1, 1), uni(u[5], -1, 1), uni(u[6], -1, 1)};
    double l2 = qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3];
    if (sqrt(l2) < 0.1) { qd[0] = 1; qd[1] = 0; qd[2] = 0; qd[3] = 0; }
    Q4 q; for (int k = 0; k < 4; k++) q.c[k] = (real)qd[k];
    real ln = rsqrt_(((q.c[0] * q.c[0] + q.c[1] * q.c[1]) + q.c[2] * q.c[2]) + q.c[3] * q.c[3]);
    for (int k = 0; k < 4; k++) q.c[k] = rdiv(q.c[k], ln);
    V3 vel = mk3((real)uni(u[7], -5, 5), (real)uni(u[8], -5, 5), (real)uni(u[9], -5, 5));
    V3 rot = mk3((real)uni(u[10], -3, 3), (real)uni(u[11], -3, 3), (real)uni(u[12], -3, 3));
    real ld = (real)uni(u[13], 0.90, 0.99), ad = (real)uni(u[14], 0.90, 0.99);
    M3 ib;
    ib.c[0] = (real)uni(u[15], 0.5, 2); ib.c[4] = (real)uni(u[16], 0.5, 2); ib.c[8] = (real)uni(u[17], 0.5, 2);
    ib.c[1] = ib.c[3] = (real)uni(u[18], -0.1, 0.1);
    ib.c[2] = ib.c[6] = (real)uni(u[19], -0.1, 0.1);
    ib.c[5] = ib.c[7] = (real)uni(u[20], -0.1, 0.1);
    // Pow(damping, dt) evaluated in float64 and rounded to Real (rigidbody.go:233-234).  The
    // bench hoists it out of the timed loop exactly like the world handle does.
    real lp = (real)pow((double)ld, (double)dt), ap = (real)pow((double)ad, (double)dt);
    s.st(C_P01, i, make_real2(pos.c[0], pos.c[1]));
    s.st(C_P2M, i, make_real2(pos.c[2], R_(0.6)));
    s.st(C_Q01, i, make_real2(q.c[0], q.c[1]));
    s.st(C_Q23, i, make_real2(q.c[2], q.c[3]));
    s.st(C_V01, i, make_real2(vel.c[0], vel.c[1]));
    s.st(C_V2R0, i, make_real2(vel.c[2], rot.c[0]));
    s.st(C_R12, i, make_real2(rot.c[1], rot.c[2]));
    s.st(C_A01, i, make_real2(R_(0.0), R_(-9.78)));
    s.st(C_A2LP, i, make_real2(R_(0.0), lp));
    s.st(C_APW0, i, make_real2(ap, ib.c[0]));
    s.st(C_I12, i, make_real2(ib.c[1], ib.c[2]));
    s.st(C_I34, i, make_real2(ib.c[3], ib.c[4]));
    s.st(C_I56, i, make_real2(ib.c[5], ib.c[6]));
    s.st(C_I78, i, make_real2(ib.c[7], ib.c[8]));
    s.st(C_MD, i, make_real2(R_(1.0), ld));
    s.st(C_AD, i, make_real2(ad, R_(0)));
    s.awake[i] = 1;
    s.can_sleep[i] = 1;
}

}  // namespace czk
