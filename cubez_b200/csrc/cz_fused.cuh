// cz_fused.cuh — K5 world_step_small: the whole frame (integrate -> collider derive ->
// narrowphase -> contact compaction -> resolver) fused in one kernel for batches of small
// independent worlds (the RL-environment case, cfg4), persistent over n_steps.
//
// Mapping: an aligned group of G lanes (G = 8, 16 or 32) owns one world; 128/G worlds share
// a CTA and 32/G worlds share a warp and run in SIMD — the serial winner-resolve of the
// worst-first loops then serves 32/G worlds per issued instruction.  A world's full state
// (82 reals + 5 ints per body) and its contact records (22 reals + 2 ints per contact) are
// staged in shared memory for the whole call; HBM is touched once at entry and once at exit
// (plus the optional dump of the last step's generated contacts).
// No __syncthreads is used: groups are independent and only synchronise their own lanes.
#pragma once
#include "cz_kernels.cuh"

namespace czf {
using namespace czm;
using namespace czk;

// staged body record: the first BW_NF fields are exactly the resolver's body work record
enum FusedBody : int {
    FB_ACC = czr::BW_NF,        // 28..30 acceleration
    FB_LINPOW = FB_ACC + 3,     // 31
    FB_ANGPOW = FB_LINPOW + 1,  // 32
    FB_IITB = FB_ANGPOW + 1,    // 33..41  body-space inverse inertia   } czr::XB_IITB
    FB_TR = FB_IITB + 9,        // 42..53  body transform              } czr::XB_TR
    FB_CTR = FB_TR + 12,        // 54..65  collider transform
    FB_HALF = FB_CTR + 12,      // 66..68
    FB_RADIUS = FB_HALF + 3,    // 69
    FB_OFFSET = FB_RADIUS + 1,  // 70..81  collider Offset
    FB_NF = FB_OFFSET + 12      // 82
};
enum FusedInt : int { FI_SHAPE = 0, FI_CANSLEEP, FI_INTEG, FI_IDENT, FI_ACTIVE, FI_NF };

struct FusedPlan {
    int G;               // lanes per world
    int threads;         // CTA size
    int worldsPerBlock;
    int bs, cs;          // field strides (bodies, contacts)
    size_t worldBytes;   // shared memory per world
    size_t smemBytes;    // per CTA
    int keepContacts;
};

static inline size_t world_bytes(int B, int Cc) {
    size_t reals = (size_t)FB_NF * B + (size_t)czr::CW_NF * Cc;
    size_t ints = (size_t)FI_NF * B + 2 * (size_t)Cc;
    size_t bytes = reals * sizeof(real) + ints * sizeof(int);
    return (bytes + 15) / 16 * 16;
}

// Decide whether (and how) a world shape runs on the fused kernel.
static inline bool plan(FusedPlan &fp, int B, int P, int Cc, int nchk, int schedule, size_t smemOptin) {
    (void)P; (void)schedule;
    if (B > 64 || nchk > 8192 || nchk <= 0) return false;
    size_t wb = world_bytes(B, Cc);
    size_t limit = smemOptin > 1024 ? smemOptin - 1024 : 0;
    if (wb > limit) return false;
    int G = 32;
    if (const char *e = getenv("CUBEZ_FUSED_G")) {
        int g = atoi(e);
        if (g == 8 || g == 16 || g == 32) G = g;
    } else if (B <= 8) {
        G = 8;
    } else if (B <= 16) {
        G = 16;
    }
    int threads = 128;
    if (const char *e = getenv("CUBEZ_FUSED_THREADS")) {
        int t = atoi(e);
        if (t == 32 || t == 64 || t == 128 || t == 256) threads = t;
    }
    int wpb = threads / G;
    while (wpb > 1 && wb * wpb > limit) { wpb >>= 1; }
    if (wpb * G < 32) {   // keep whole warps
        wpb = 32 / G;
        if (wb * wpb > limit) { G = 32; wpb = 1; }
    }
    fp.G = G;
    fp.worldsPerBlock = wpb;
    fp.threads = wpb * G;
    fp.bs = B;
    fp.cs = Cc;
    fp.worldBytes = wb;
    fp.smemBytes = wb * wpb;
    fp.keepContacts = 1;
    if (const char *e = getenv("CUBEZ_FUSED_KEEP_CONTACTS")) fp.keepContacts = atoi(e);
    return true;
}

// staged-world view of one group
struct Staged {
    real *fb; int bs;
    int *fi;
    real *cw; int cs;
    int *cb0, *cb1;
};

__device__ __forceinline__ ColliderView staged_collider(const Staged &s, int i) {
    ColliderView v;
    v.shape = s.fi[FI_SHAPE * s.bs + i];
    v.body = i;
#pragma unroll
    for (int k = 0; k < 12; k++) v.t.c[k] = s.fb[(FB_CTR + k) * s.bs + i];
    v.half = mk3(s.fb[(FB_HALF + 0) * s.bs + i], s.fb[(FB_HALF + 1) * s.bs + i], s.fb[(FB_HALF + 2) * s.bs + i]);
    v.radius = s.fb[FB_RADIUS * s.bs + i];
    return v;
}
__device__ __forceinline__ bool staged_active(const Staged &s, int i, long long step) {
    return step >= (long long)s.fi[FI_ACTIVE * s.bs + i] && s.fi[FI_SHAPE * s.bs + i] != CZ_SHAPE_NONE;
}

// CheckForCollisions on staged colliders (same logic as czk::eval_check)
__device__ __forceinline__ void eval_check_staged(const WorldParams &p, const Staged &s, long long step, int a, int b, CheckEval &e) {
    e.count = 0; e.kind = 0; e.mask = 0;
    if (a < 0 && b < 0) return;
    if (a >= 0 && !staged_active(s, a, step)) return;
    if (b >= 0 && !staged_active(s, b, step)) return;
    if (a < 0 || b < 0) {
        int ci = a < 0 ? b : a, pi = a < 0 ? -a - 1 : -b - 1;
        ColliderView c = staged_collider(s, ci);
        const PlaneView &pl = p.planes[pi];
        if (c.shape == CZ_SHAPE_SPHERE) {
            if (czn::sphere_halfspace(c, pl, e.gc)) { e.count = 1; e.kind = 1; }
        } else if (c.shape == CZ_SHAPE_CUBE) {
            e.mask = czn::cube_halfspace_mask(c, pl);
            if (e.mask) { e.count = cz_popc(e.mask); e.kind = 2; e.cubeLocal = ci; e.plane = pi; }
        }
        return;
    }
    ColliderView one = staged_collider(s, a), two = staged_collider(s, b);
    V3 v1 = zero3(), v2 = zero3();
    if (one.shape != two.shape) {
        v1 = mk3(s.fb[(czr::BW_VEL + 0) * s.bs + a], s.fb[(czr::BW_VEL + 1) * s.bs + a], s.fb[(czr::BW_VEL + 2) * s.bs + a]);
        v2 = mk3(s.fb[(czr::BW_VEL + 0) * s.bs + b], s.fb[(czr::BW_VEL + 1) * s.bs + b], s.fb[(czr::BW_VEL + 2) * s.bs + b]);
    }
    if (czn::check_pair(one, two, v1, v2, e.gc)) { e.count = 1; e.kind = 1; }
}

__device__ __forceinline__ void stage_gen(const Staged &s, int slot, const GenContact &c) {
    using namespace czr;
#pragma unroll
    for (int k = 0; k < 3; k++) { s.cw[(G_POINT + k) * s.cs + slot] = c.point.c[k]; s.cw[(G_NORMAL + k) * s.cs + slot] = c.normal.c[k]; }
    s.cw[G_PEN * s.cs + slot] = c.pen;
    s.cw[G_FRIC * s.cs + slot] = R_(0.9);
    s.cw[G_REST * s.cs + slot] = R_(0.1);
    s.cb0[slot] = c.b0;
    s.cb1[slot] = c.b1;
}

template <int G>
__global__ void __launch_bounds__(256) k_world_fused(WorldParams p, FusedPlan fp, real dt, real bias, int nSteps) {
    using namespace czr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / G, tid = threadIdx.x % G;
    const int w = blockIdx.x * fp.worldsPerBlock + grp;
    if (w >= p.W) return;
    const unsigned mask = group_mask<G>();
    const int B = p.B, Cc = p.Cc;
    Staged s;
    unsigned char *base = smem_raw + (size_t)grp * fp.worldBytes;
    s.fb = (real *)base; s.bs = B;
    s.cw = s.fb + (size_t)FB_NF * B; s.cs = Cc;
    s.fi = (int *)(s.cw + (size_t)CW_NF * Cc);
    s.cb0 = s.fi + FI_NF * B;
    s.cb1 = s.cb0 + Cc;
    const BodyStore &st = p.st;
    const long long gbase = (long long)w * B;

    // ---- stage the world's state --------------------------------------------------------
    for (int b = tid; b < B; b += G) {
        const long long gi = gbase + b;
        // slots 0..55 of the chunked layout map 1:1 onto record fields (see table below)
        real2 c;
        c = st.ld(czb::C_P01, gi); s.fb[(BW_POS + 0) * B + b] = c.x; s.fb[(BW_POS + 1) * B + b] = c.y;
        c = st.ld(czb::C_P2M, gi); s.fb[(BW_POS + 2) * B + b] = c.x; s.fb[BW_MOTION * B + b] = c.y;
        c = st.ld(czb::C_Q01, gi); s.fb[(BW_Q + 0) * B + b] = c.x; s.fb[(BW_Q + 1) * B + b] = c.y;
        c = st.ld(czb::C_Q23, gi); s.fb[(BW_Q + 2) * B + b] = c.x; s.fb[(BW_Q + 3) * B + b] = c.y;
        c = st.ld(czb::C_V01, gi); s.fb[(BW_VEL + 0) * B + b] = c.x; s.fb[(BW_VEL + 1) * B + b] = c.y;
        c = st.ld(czb::C_V2R0, gi); s.fb[(BW_VEL + 2) * B + b] = c.x; s.fb[(BW_ROT + 0) * B + b] = c.y;
        c = st.ld(czb::C_R12, gi); s.fb[(BW_ROT + 1) * B + b] = c.x; s.fb[(BW_ROT + 2) * B + b] = c.y;
        c = st.ld(czb::C_A01, gi); s.fb[(FB_ACC + 0) * B + b] = c.x; s.fb[(FB_ACC + 1) * B + b] = c.y;
        c = st.ld(czb::C_A2LP, gi); s.fb[(FB_ACC + 2) * B + b] = c.x; s.fb[FB_LINPOW * B + b] = c.y;
        c = st.ld(czb::C_APW0, gi); s.fb[FB_ANGPOW * B + b] = c.x; s.fb[(FB_IITB + 0) * B + b] = c.y;
#pragma unroll
        for (int k = 0; k < 4; k++) { c = st.ld(czb::C_I12 + k, gi); s.fb[(FB_IITB + 1 + 2 * k) * B + b] = c.x; s.fb[(FB_IITB + 2 + 2 * k) * B + b] = c.y; }
        V3 la = czb::ld_last_acc(st, gi);
        M34 tr = czb::ld_transform(st, gi);
        M3 iw = czb::ld_iit_world(st, gi);
#pragma unroll
        for (int k = 0; k < 3; k++) s.fb[(BW_LACC + k) * B + b] = la.c[k];
#pragma unroll
        for (int k = 0; k < 12; k++) s.fb[(FB_TR + k) * B + b] = tr.c[k];
#pragma unroll
        for (int k = 0; k < 9; k++) s.fb[(BW_IITW + k) * B + b] = iw.c[k];
        M34 ctr = czb::ld_m34(st, czb::C_X01, gi), off = czb::ld_m34(st, czb::C_O01, gi);
#pragma unroll
        for (int k = 0; k < 12; k++) { s.fb[(FB_CTR + k) * B + b] = ctr.c[k]; s.fb[(FB_OFFSET + k) * B + b] = off.c[k]; }
        c = st.ld(czb::C_H01, gi); s.fb[(FB_HALF + 0) * B + b] = c.x; s.fb[(FB_HALF + 1) * B + b] = c.y;
        c = st.ld(czb::C_H2R, gi); s.fb[(FB_HALF + 2) * B + b] = c.x; s.fb[FB_RADIUS * B + b] = c.y;
        s.fb[BW_INVM * B + b] = st.ld(czb::C_MD, gi).x;
        s.fb[BW_AWAKE * B + b] = st.awake[gi] ? R_(1) : R_(0);
        s.fi[FI_SHAPE * B + b] = st.shape[gi];
        s.fi[FI_CANSLEEP * B + b] = st.can_sleep[gi];
        s.fi[FI_INTEG * B + b] = st.integ[gi];
        s.fi[FI_IDENT * B + b] = st.ident[gi];
        s.fi[FI_ACTIVE * B + b] = st.active_from[gi];
    }
    __syncwarp(mask);

    Ctx x;
    x.bw = s.fb; x.bs = B; x.cw = s.cw; x.cs = Cc; x.cb0 = s.cb0; x.cb1 = s.cb1; x.nC = 0; x.dt = dt;
    x.xb = s.fb + (size_t)FB_IITB * B; x.xbs = B;
    x.store = st; x.body_base = gbase;

    unsigned long long accContacts = 0, accPos = 0, accVel = 0;
    int maxC = 0, lastC = 0, lastPos = 0, lastVel = 0, status = 0;

    for (int stepNo = 0; stepNo < nSteps; stepNo++) {
        const long long step = p.step_index + stepNo;
        // ---- updateObjects: Integrate + collider derive (cubedrop.go:29-39) -------------
        for (int b = tid; b < B; b += G) {
            if (!(s.fi[FI_INTEG * B + b] != 0 && step >= (long long)s.fi[FI_ACTIVE * B + b])) continue;
            M34 tr;
            if (s.fb[BW_AWAKE * B + b] != R_(0)) {
                V3 pos = bw3(x, BW_POS, b), vel = bw3(x, BW_VEL, b), rot = bw3(x, BW_ROT, b), acc = bw3(x, FB_ACC, b);
                Q4 q = bw_q(x, b);
                M3 ib;
#pragma unroll
                for (int k = 0; k < 9; k++) ib.c[k] = s.fb[(FB_IITB + k) * B + b];
                czb::Integrated o;
                czb::integrate_body(o, pos, q, vel, rot, acc, ib, s.fb[BW_MOTION * B + b], s.fi[FI_CANSLEEP * B + b] != 0, dt,
                                    s.fb[FB_LINPOW * B + b], s.fb[FB_ANGPOW * B + b], bias);
                bw3_set(x, BW_POS, b, o.pos); bw3_set(x, BW_VEL, b, o.vel); bw3_set(x, BW_ROT, b, o.rot); bw3_set(x, BW_LACC, b, o.lastAcc);
#pragma unroll
                for (int k = 0; k < 4; k++) s.fb[(BW_Q + k) * B + b] = o.q.c[k];
#pragma unroll
                for (int k = 0; k < 12; k++) s.fb[(FB_TR + k) * B + b] = o.transform.c[k];
#pragma unroll
                for (int k = 0; k < 9; k++) s.fb[(BW_IITW + k) * B + b] = o.iitWorld.c[k];
                s.fb[BW_MOTION * B + b] = o.motion;
                if (!o.awake) s.fb[BW_AWAKE * B + b] = R_(0);
                tr = o.transform;
            } else {
#pragma unroll
                for (int k = 0; k < 12; k++) tr.c[k] = s.fb[(FB_TR + k) * B + b];
            }
            if (s.fi[FI_SHAPE * B + b] != CZ_SHAPE_NONE) {
                M34 off;
                if (s.fi[FI_IDENT * B + b]) off = czb::identity34();
                else {
#pragma unroll
                    for (int k = 0; k < 12; k++) off.c[k] = s.fb[(FB_OFFSET + k) * B + b];
                }
                M34 ctr = m34_mul_m34(tr, off);
#pragma unroll
                for (int k = 0; k < 12; k++) s.fb[(FB_CTR + k) * B + b] = ctr.c[k];
            }
        }
        __syncwarp(mask);

        // ---- generateContacts with order-preserving compaction ---------------------------
        int nC = 0;
        for (int k0 = 0; k0 < p.nchk; k0 += G) {
            const int k = k0 + tid;
            CheckEval e;
            e.count = 0; e.kind = 0; e.mask = 0;
            int a = 0, b2 = 0;
            if (k < p.nchk && decode_check(p, k, a, b2)) eval_check_staged(p, s, step, a, b2, e);
            int incl = e.count;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) {
                int t = __shfl_up_sync(mask, incl, o, G);
                if (tid >= o) incl += t;
            }
            const int total = __shfl_sync(mask, incl, G - 1, G);
            int slot = nC + incl - e.count;
            if (e.kind == 1) {
                if (slot < Cc) stage_gen(s, slot, e.gc);
            } else if (e.kind == 2) {
                ColliderView c = staged_collider(s, e.cubeLocal);
#pragma unroll
                for (int v = 0; v < 8; v++) {
                    if (e.mask & (1u << v)) {
                        GenContact gc;
                        czn::cube_halfspace_contact(c, p.planes[e.plane], v, gc);
                        if (slot < Cc) stage_gen(s, slot, gc);
                        slot++;
                    }
                }
            }
            nC += total;
        }
        __syncwarp(mask);
        accContacts += (unsigned long long)nC;
        maxC = max(maxC, nC);
        lastC = nC;
        lastPos = lastVel = 0;
        if (nC > Cc) { status = CZ_ERR_CAPACITY; nC = 0; }
        if (fp.keepContacts && stepNo == nSteps - 1 && nC > 0) {
            // dump the as-generated contacts of the last step for cz_world_download_contacts
            const long long gs = (long long)p.W * Cc;
            real *gen = p.gen + (long long)w * Cc;
            for (int c = tid; c < nC; c += G) {
#pragma unroll
                for (int f = 0; f < G_NF; f++) gen[f * gs + c] = s.cw[f * Cc + c];
                p.gb0[(long long)w * Cc + c] = s.cb0[c];
                p.gb1[(long long)w * Cc + c] = s.cb1[c];
            }
            __syncwarp(mask);
        }
        // ---- ResolveContacts(8*len) (cubedrop.go:72-74) -----------------------------------
        if (nC > 0) {
            x.nC = nC;
            for (int c = tid; c < nC; c += G) prepare_contact(x, c, s.cw, Cc, s.cb0, s.cb1);
            __syncwarp(mask);
            int st2 = 0;
            lastPos = resolve_loop<G, false>(x, nC * 8, nullptr, tid, &st2);
            lastVel = resolve_loop<G, true>(x, nC * 8, nullptr, tid, &st2);
            if (st2) status = st2;
            accPos += (unsigned long long)lastPos;
            accVel += (unsigned long long)lastVel;
        }
        __syncwarp(mask);
    }

    // ---- write the state back ------------------------------------------------------------
    for (int b = tid; b < B; b += G) {
        const long long gi = gbase + b;
        st.st(czb::C_P01, gi, make_real2(s.fb[(BW_POS + 0) * B + b], s.fb[(BW_POS + 1) * B + b]));
        st.st(czb::C_P2M, gi, make_real2(s.fb[(BW_POS + 2) * B + b], s.fb[BW_MOTION * B + b]));
        st.st(czb::C_Q01, gi, make_real2(s.fb[(BW_Q + 0) * B + b], s.fb[(BW_Q + 1) * B + b]));
        st.st(czb::C_Q23, gi, make_real2(s.fb[(BW_Q + 2) * B + b], s.fb[(BW_Q + 3) * B + b]));
        st.st(czb::C_V01, gi, make_real2(s.fb[(BW_VEL + 0) * B + b], s.fb[(BW_VEL + 1) * B + b]));
        st.st(czb::C_V2R0, gi, make_real2(s.fb[(BW_VEL + 2) * B + b], s.fb[(BW_ROT + 0) * B + b]));
        st.st(czb::C_R12, gi, make_real2(s.fb[(BW_ROT + 1) * B + b], s.fb[(BW_ROT + 2) * B + b]));
        st.st(czb::C_L01, gi, make_real2(s.fb[(BW_LACC + 0) * B + b], s.fb[(BW_LACC + 1) * B + b]));
        M34 tr;
        M3 iw;
#pragma unroll
        for (int k = 0; k < 12; k++) tr.c[k] = s.fb[(FB_TR + k) * B + b];
#pragma unroll
        for (int k = 0; k < 9; k++) iw.c[k] = s.fb[(BW_IITW + k) * B + b];
        czb::st_derived(st, gi, s.fb[(BW_LACC + 2) * B + b], tr, iw);
        M34 ctr;
#pragma unroll
        for (int k = 0; k < 12; k++) ctr.c[k] = s.fb[(FB_CTR + k) * B + b];
        czb::st_m34(st, czb::C_X01, gi, ctr);
        st.awake[gi] = s.fb[BW_AWAKE * B + b] != R_(0) ? 1 : 0;
    }
    if (tid == 0) {
        p.nContacts[w] = lastC;
        p.posIters[w] = lastPos;
        p.velIters[w] = lastVel;
        atomicAdd(&p.stats[ST_CONTACTS], accContacts);
        atomicAdd(&p.stats[ST_POS], accPos);
        atomicAdd(&p.stats[ST_VEL], accVel);
        atomicMax(&p.stats[ST_MAXC], (unsigned long long)maxC);
        if (status) raise_status(p.stats, status);
    }
}

// returns 0 or a cudaError_t
static inline int launch(const FusedPlan &fp, const WorldParams &p, real dt, real bias, int nSteps, cudaStream_t stream, int smCount) {
    (void)smCount;
    const int grid = (p.W + fp.worldsPerBlock - 1) / fp.worldsPerBlock;
    cudaError_t e = cudaSuccess;
#define CZF_LAUNCH(GG)                                                                                              \
    do {                                                                                                            \
        if (fp.smemBytes > 48 * 1024)                                                                               \
            e = cudaFuncSetAttribute(k_world_fused<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fp.smemBytes); \
        if (e == cudaSuccess) k_world_fused<GG><<<grid, fp.threads, fp.smemBytes, stream>>>(p, fp, dt, bias, nSteps); \
    } while (0)
    if (fp.G == 8) CZF_LAUNCH(8);
    else if (fp.G == 16) CZF_LAUNCH(16);
    else CZF_LAUNCH(32);
#undef CZF_LAUNCH
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

}  // namespace czf
