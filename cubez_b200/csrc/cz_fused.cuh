// cz_fused.cuh — K5 world_step_small: the whole frame (integrate -> collider derive ->
// narrowphase -> contact compaction -> resolver) fused in one persistent kernel for batches of
// small independent worlds (the RL-environment case, cfg4), over n_steps frames per launch.
//
// Mapping: an aligned group of G lanes (G = 8, 16 or 32) owns one world at a time; 32/G worlds
// share a warp and run in SIMD, so the serial winner-resolve of the worst-first loops serves
// 32/G worlds per issued instruction.  Groups are persistent and fetch worlds from a global
// counter (worlds differ a lot in cost: falling / colliding / sleeping).
//
// Memory plan (the kernel is FP64-latency bound, so occupancy — worlds in flight per SM — is
// what buys throughput; shared memory is spent only on what every iteration touches):
//   shared, per world : body work record (28 reals/body: position, orientation, velocities,
//                       last-frame acceleration, world inverse inertia, inverse mass, motion,
//                       awake), collider transform + half sizes + radius (16 reals/body), the
//                       HOT contact fields (penetration, desired delta-v, body ids: 24 B/contact)
//   global, L2-resident, per resident group : the COLD contact fields (18 reals/contact, AoS —
//                       the winner's record is two cache lines), reused by every world the group
//                       processes
//   global body store : constants (acceleration, Pow factors, body-space inertia, Offset) are
//                       read where used; the body transform is written through.
// HBM is touched once at entry and once at exit per world (plus the optional dump of the last
// frame's generated contacts).  No __syncthreads: groups only synchronise their own lanes.
#pragma once
#include "cz_kernels.cuh"

namespace czf {
using namespace czm;
using namespace czk;

// shared-memory record per body: the resolver's body work record followed by collider data
enum FusedBody : int {
    FB_CTR = czr::BW_NF,       // 28..39 collider transform
    FB_HALF = FB_CTR + 12,     // 40..42 half sizes
    FB_RADIUS = FB_HALF + 3,   // 43
    FB_NF = FB_RADIUS + 1      // 44
};
enum FusedFlag : int { FF_SHAPE_MASK = 3, FF_CANSLEEP = 4, FF_INTEG = 8, FF_IDENT = 16 };

struct FusedPlan {
    int G;               // lanes per world
    int threads;         // CTA size
    int groupsPerBlock;
    int blocksPerSM;
    int minb;            // __launch_bounds__ min blocks (register budget) of the instantiation used
    int grid;
    size_t worldBytes;   // shared memory per group
    size_t smemBytes;    // per CTA
    int keepContacts;
    int bodyMasks;       // per-body contact bitmasks drive the propagation (contact capacity <= 64)
    int lockstep;        // CTA barriers between the phases of a frame (instruction-cache locality)
    int velPre;          // the velocity loop evaluates every contact's response matrix once per frame (czr::friction_response)
    real *cold;          // [grid*groupsPerBlock][Cc*CW_NCOLD]
    size_t coldReals;    // per group
    // split mode: one launch per phase of the frame (A: integrate + narrowphase + prepare, B: position
    // loop, C: velocity loop); contact state lives in per-WORLD global arrays between the launches
    int split, splitMinb;
    // the loop phases (B, C) of split mode stage only what the loops touch (no collider data, no narrowphase
    // queues): a smaller world record -> more resident CTAs per SM (the loops are latency-bound: 168 registers at
    // 3 CTAs of 128 threads per SM without spills, 128 at 4)
    size_t loopWorldBytes, loopSmemBytes;
    size_t phaseAWorldBytes, phaseASmemBytes;   // split mode: phase A keeps the hot contact fields in the per-world global arrays
    int phaseAGrid;
    int maxGrid;         // largest grid of any launch: sizes the per-group scratch
    int loopBlocksPerSM, loopGrid;
    int phaseAMinb;      // register budget of the phase-A launch
    real *coldW;         // [W][Cc*CW_NCOLD]
    real *preW;          // [W][Cc*VP_NF] per-contact velocity response of the velocity loop (split mode), or NULL
    real *hotPen, *hotDdv;   // [W][Cc]
    int *hotCb0, *hotCb1;    // [W][Cc]
};

// Every size below is in BYTES and mirrors the carve at the top of k_world_fused field by field.  The per-body contact
// masks are 64-bit whatever `real` is (the float32 build once sized them as one real per body and overran the record).
static inline size_t world_bytes(int B, int Cc, int nchk) {
    size_t bytes = ((size_t)FB_NF * B + 2 * (size_t)Cc) * sizeof(real);   // body + collider record, penetration, desired delta-v
    bytes += (size_t)B * sizeof(unsigned long long);                          // one 64-bit contact mask per body
    bytes += (2 * (size_t)Cc + 2 * (size_t)B) * sizeof(int);                  // contact body ids, flags, activation
    bytes += 3 * (size_t)nchk * sizeof(unsigned short);                       // per-check info + the queues of checks that need a full test + plane-check base slots
    bytes += 2 * (size_t)nchk + (size_t)Cc;                                   // per-check counts, vertex masks, the match list
    return (bytes + 15) / 16 * 16;
}

// world record of a loop-phase launch: body work record, the phase's ONE hot contact field (penetration for the
// position loop, desired delta-v for the velocity loop), body masks, contact body ids; the match list only when
// the masks are off.  2.8 KB for an 8-body world with 64 contacts: three CTAs need 138 KB of shared memory, which
// leaves the SM a 92 KB L1 for the cold contact records (164 KB carve-out instead of 196 KB).
static inline size_t world_bytes_loops(int B, int Cc, bool masks) {
    size_t bytes = ((size_t)czr::BW_NF * B + (((size_t)Cc + 1) & ~(size_t)1)) * sizeof(real);   // hot field padded to an even count: the 64-bit masks follow it
    bytes += (size_t)B * sizeof(unsigned long long);
    bytes += 2 * (size_t)Cc * sizeof(int) + (masks ? 0 : (size_t)Cc);
    return (bytes + 15) / 16 * 16;
}

// world record of the phase-A launch of split mode: body + collider record, flags, check queues
static inline size_t world_bytes_phase_a(int B, int nchk) {
    size_t bytes = (size_t)FB_NF * B * sizeof(real) + 2 * (size_t)B * sizeof(int) + 3 * (size_t)nchk * sizeof(unsigned short) + 2 * (size_t)nchk;
    return (bytes + 15) / 16 * 16;
}

// End offset (bytes) of the shared-memory carve k_world_fused performs for one world, computed with the kernel's own
// pointer arithmetic: mode 0 = full record (PH_ALL), 1 = phase A only, 2 = a loop phase.  plan() refuses a shape whose
// carve would run past the record the helpers above size (a host-side guard for the two to stay in step).
static inline size_t carve_end(int B, int Cc, int nchk, int mode, bool masks) {
    const bool full = mode != 2, aOnly = mode == 1;
    const size_t ch = aOnly ? 0 : (size_t)Cc;
    size_t off = (size_t)(full ? (int)FB_NF : (int)czr::BW_NF) * B * sizeof(real);   // s.pen
    off += (full ? 2 * ch : ((ch + 1) & ~(size_t)1)) * sizeof(real);                   // bmask
    off += (aOnly ? 0 : (size_t)B) * sizeof(unsigned long long);                       // s.cb0
    off += 2 * ch * sizeof(int);                                                       // s.flags
    off += (full ? 2 * (size_t)B : 0) * sizeof(int);                                   // s.info
    const size_t nq = full ? (size_t)nchk : 0;
    off += 3 * nq * sizeof(unsigned short) + 2 * nq;                                   // mlist
    const bool hasMlist = Cc <= 256 && !aOnly && (full || !masks);
    return off + (hasMlist ? (size_t)Cc : 0);
}

static inline int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}

// Decide whether (and how) a world shape runs on the fused kernel.
static inline bool plan(FusedPlan &fp, int B, int P, int Cc, int nchk, int schedule, size_t smemOptin, int smCount, int W) {
    (void)P; (void)schedule;
    if (B > 64 || nchk > 1024 || nchk <= 0) return false;
    const size_t wb = world_bytes(B, Cc, nchk);
    const size_t perSM = 227 * 1024;
    if (wb > smemOptin) return false;
    int G = B <= 8 ? 8 : (B <= 16 ? 16 : 32);
    int g = env_int("CUBEZ_FUSED_G", 0);
    if (g == 8 || g == 16 || g == 32) G = g;
    else {
        // Small batches do not fill the GPU with 8 lanes per world: a frame is then one world's serial chain (~0.4 ms),
        // and wider groups shorten its parallel parts.  Measured on B200, 8-body worlds, us per frame at G = 8 / 16 / 32:
        // 1 024 worlds 389 / 278 / 272, 2 048: 402 / 318 / 314, 4 096: 455 / 376 / 511, 6 144: 478 / 527 / 707
        // (tools/split_threshold_probe.py) -> the widest group while the batch is within about two waves of it.
        const long long r32 = (long long)smCount * 2 * 4, r16 = (long long)smCount * 2 * 8;   // resident worlds at 2 CTAs of 128 threads per SM
        if ((long long)W * 4 <= r32 * 7) G = 32;
        else if (G < 16 && (long long)W * 5 <= r16 * 11) G = 16;
    }
    int threads = env_int("CUBEZ_FUSED_THREADS", 128);
    if (threads != 32 && threads != 64 && threads != 128) threads = 128;
    int gpb = threads / G;
    while (gpb > 32 / G && wb * gpb > smemOptin) gpb >>= 1;
    if (wb * gpb > smemOptin) { G = 32; gpb = 1; }
    threads = gpb * G;
    size_t smem = wb * gpb;
    int bps = (int)(perSM / (smem + 1024));
    if (bps < 1) bps = 1;
    int maxByThreads = 2048 / threads;
    if (bps > maxByThreads) bps = maxByThreads;
    int minb = env_int("CUBEZ_FUSED_MINB", 2);     // register budget: 65536 / (128 * MINB) per thread
    if (minb < 2) minb = 2;
    if (minb > 3) minb = 3;
    if (bps > minb * (128 / threads)) bps = minb * (128 / threads);
    fp.minb = minb;
    fp.G = G;
    fp.threads = threads;
    fp.groupsPerBlock = gpb;
    fp.blocksPerSM = bps;
    fp.grid = smCount * bps;
    fp.worldBytes = wb;
    fp.smemBytes = smem;
    fp.keepContacts = env_int("CUBEZ_FUSED_KEEP_CONTACTS", 1);
    fp.bodyMasks = env_int("CUBEZ_FUSED_BODY_MASKS", 1);
    // Large batches run one launch per phase of the frame (every warp of the GPU is then in the same code region and
    // each phase stages a short record: 3 CTAs per SM); small batches keep the single persistent launch.  Measured
    // crossover on B200 with the split step running as independent slices on their own streams (cz_world_step, "lanes";
    // tools/lanes_probe.py, staggered episodes, us per frame persistent vs split): 8 192 worlds 511 vs 573, 12 288 worlds
    // 737 vs 593, 16 384 worlds 887 vs 675; episodes in step from t = 0 (tools/strong_probe.py): 8 192 worlds 262 vs
    // 265 ms per 600 frames, 16 384 worlds 474 vs 380; 10 240 worlds 602 vs 572 us -> split from 10 240 worlds on.
    fp.split = env_int("CUBEZ_FUSED_SPLIT", W >= 10240 ? 1 : 0);
    fp.splitMinb = env_int("CUBEZ_FUSED_SPLIT_MINB", 3);
    if (fp.splitMinb < 2 || fp.splitMinb > 4 || G != 8) fp.splitMinb = 2;   // instantiated for G = 8 only
    fp.phaseAMinb = env_int("CUBEZ_FUSED_PHASE_A_MINB", 3);
    if (fp.phaseAMinb < 2 || fp.phaseAMinb > 4 || G != 8) fp.phaseAMinb = 2;
    fp.loopWorldBytes = world_bytes_loops(B, Cc, Cc <= 64 && fp.bodyMasks);
    fp.loopSmemBytes = fp.loopWorldBytes * gpb;
    {
        int lb = (int)(perSM / (fp.loopSmemBytes + 1024));
        if (lb > maxByThreads) lb = maxByThreads;
        if (lb > fp.splitMinb * (128 / threads)) lb = fp.splitMinb * (128 / threads);
        if (lb < 1) lb = 1;
        fp.loopBlocksPerSM = lb;
        fp.loopGrid = smCount * lb;
    }
    fp.phaseAWorldBytes = world_bytes_phase_a(B, nchk);
    fp.phaseASmemBytes = fp.phaseAWorldBytes * gpb;
    {
        int ab = (int)(perSM / (fp.phaseASmemBytes + 1024));
        if (ab > maxByThreads) ab = maxByThreads;
        if (ab > fp.phaseAMinb * (128 / threads)) ab = fp.phaseAMinb * (128 / threads);
        if (ab < 1) ab = 1;
        fp.phaseAGrid = smCount * ab;
    }
    fp.maxGrid = fp.grid > fp.phaseAGrid ? fp.grid : fp.phaseAGrid;
    if (fp.loopGrid > fp.maxGrid) fp.maxGrid = fp.loopGrid;
    fp.coldW = nullptr; fp.preW = nullptr; fp.hotPen = fp.hotDdv = nullptr; fp.hotCb0 = fp.hotCb1 = nullptr;
    fp.lockstep = env_int("CUBEZ_FUSED_LOCKSTEP", 3);   // 0 none, 1 frame start, 2 + before narrowphase/resolve, 3 + between the two loops
    fp.velPre = env_int("CUBEZ_FUSED_VEL_PRE", 1);
    fp.coldReals = (size_t)Cc * czr::CW_NCOLD + (size_t)nchk * 8 + (fp.velPre ? (size_t)Cc * czr::VP_NF : 0);   // + staging of pair-test contacts + velocity responses
    fp.cold = nullptr;
    const bool lm = Cc <= 64 && fp.bodyMasks;
    if (carve_end(B, Cc, nchk, 0, lm) > fp.worldBytes || carve_end(B, Cc, nchk, 1, lm) > fp.phaseAWorldBytes ||
        carve_end(B, Cc, nchk, 2, lm) > fp.loopWorldBytes) {
        fprintf(stderr, "libcubezcuda: fused shared-memory carve exceeds its record (B %d Cc %d nchk %d) — not using the fused kernel\n", B, Cc, nchk);
        return false;
    }
    return true;
}

// staged-world view of one group
struct Staged {
    real *fb; int bs;     // fb[field*bs + body], FB_NF fields
    real *pen, *ddv;      // [Cc]
    int *cb0, *cb1;       // [Cc]
    int *flags, *active;  // [B]
    real *cold;           // global scratch, AoS [contact][CW_NCOLD]
    real *pairGen;        // global scratch: contact of queued pair test q, [q][8]
    unsigned short *info, *queue;   // [nchk] per-check queue slot; queue of check ids (pair tests from the front, plane checks from the back)
    unsigned short *pbase;          // [nchk] first contact slot of queued plane check q
    unsigned char *cnt;   // [nchk] contacts produced by check k
    unsigned char *pmask; // [nchk] vertex mask of queued plane check q
};

__device__ __forceinline__ ColliderView staged_collider(const Staged &s, int i) {
    ColliderView v;
    v.shape = s.flags[i] & FF_SHAPE_MASK;
    v.body = i;
#pragma unroll
    for (int k = 0; k < 12; k++) v.t.c[k] = s.fb[(FB_CTR + k) * s.bs + i];
    v.half = mk3(s.fb[(FB_HALF + 0) * s.bs + i], s.fb[(FB_HALF + 1) * s.bs + i], s.fb[(FB_HALF + 2) * s.bs + i]);
    v.radius = s.fb[FB_RADIUS * s.bs + i];
    return v;
}
__device__ __forceinline__ bool staged_active(const Staged &s, int i, long long step) {
    return step >= (long long)s.active[i] && (s.flags[i] & FF_SHAPE_MASK) != CZ_SHAPE_NONE;
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

// as-generated contact -> point/normal into the cold record (slots 0..5), penetration and
// body ids into the hot arrays.  Friction / restitution are the constants of colliders.go.
__device__ __forceinline__ void stage_gen(const Staged &s, int slot, const GenContact &c) {
    real *r = s.cold + (size_t)slot * czr::CW_NCOLD;
#pragma unroll
    for (int k = 0; k < 3; k++) { r[czr::G_POINT + k] = c.point.c[k]; r[czr::G_NORMAL + k] = c.normal.c[k]; }
    s.pen[slot] = c.pen;
    s.cb0[slot] = c.b0;
    s.cb1[slot] = c.b1;
}

// Split mode: the as-generated record goes straight into the world's download arrays (point, normal,
// penetration, body ids: what cz_world_download_contacts returns) and prepare_contact reads it from there —
// the copy into those arrays after generation cost a dependent L2 round trip per world and frame
// (ncu: 21 % of the stall samples of the phase-A launch).
__device__ __forceinline__ void stage_gen_direct(const czr::GenView &g, int slot, const GenContact &c) {
    real *r = g.pn + (size_t)slot * g.cs;
#pragma unroll
    for (int k = 0; k < 3; k++) { r[(size_t)(czr::G_POINT + k) * g.fs] = c.point.c[k]; r[(size_t)(czr::G_NORMAL + k) * g.fs] = c.normal.c[k]; }
    g.pen[slot] = c.pen;
    g.b0[slot] = c.b0;
    g.b1[slot] = c.b1;
}

// copy one world's hot state from a chunked store into the staged record
template <int G, bool FULL>
__device__ __forceinline__ void stage_world(const Staged &s, const BodyStore &st, long long gbase, int B, int tid) {
    using namespace czr;
    for (int b = tid; b < B; b += G) {
        const long long gi = gbase + b;
        real2 c;
        c = st.ld(czb::C_P01, gi); s.fb[(BW_POS + 0) * B + b] = c.x; s.fb[(BW_POS + 1) * B + b] = c.y;
        c = st.ld(czb::C_P2M, gi); s.fb[(BW_POS + 2) * B + b] = c.x; s.fb[BW_MOTION * B + b] = c.y;
        c = st.ld(czb::C_Q01, gi); s.fb[(BW_Q + 0) * B + b] = c.x; s.fb[(BW_Q + 1) * B + b] = c.y;
        c = st.ld(czb::C_Q23, gi); s.fb[(BW_Q + 2) * B + b] = c.x; s.fb[(BW_Q + 3) * B + b] = c.y;
        c = st.ld(czb::C_V01, gi); s.fb[(BW_VEL + 0) * B + b] = c.x; s.fb[(BW_VEL + 1) * B + b] = c.y;
        c = st.ld(czb::C_V2R0, gi); s.fb[(BW_VEL + 2) * B + b] = c.x; s.fb[(BW_ROT + 0) * B + b] = c.y;
        c = st.ld(czb::C_R12, gi); s.fb[(BW_ROT + 1) * B + b] = c.x; s.fb[(BW_ROT + 2) * B + b] = c.y;
        V3 la = czb::ld_last_acc(st, gi);
        M3 iw = czb::ld_iit_world(st, gi);
#pragma unroll
        for (int k = 0; k < 3; k++) s.fb[(BW_LACC + k) * B + b] = la.c[k];
#pragma unroll
        for (int k = 0; k < 9; k++) s.fb[(BW_IITW + k) * B + b] = iw.c[k];
        if (FULL) {
            M34 ctr = czb::ld_m34(st, czb::C_X01, gi);
#pragma unroll
            for (int k = 0; k < 12; k++) s.fb[(FB_CTR + k) * B + b] = ctr.c[k];
            c = st.ld(czb::C_H01, gi); s.fb[(FB_HALF + 0) * B + b] = c.x; s.fb[(FB_HALF + 1) * B + b] = c.y;
            c = st.ld(czb::C_H2R, gi); s.fb[(FB_HALF + 2) * B + b] = c.x; s.fb[FB_RADIUS * B + b] = c.y;
        }
        s.fb[BW_INVM * B + b] = st.ld(czb::C_MD, gi).x;
        s.fb[BW_AWAKE * B + b] = st.awake[gi] ? R_(1) : R_(0);
        if (FULL) {
            s.flags[b] = (int)st.shape[gi] | (st.can_sleep[gi] ? FF_CANSLEEP : 0) | (st.integ[gi] ? FF_INTEG : 0) | (st.ident[gi] ? FF_IDENT : 0);
            s.active[b] = st.active_from[gi];
        }
    }
}

// LOCKSTEP: the groups of a CTA run the phases of a frame (integrate | narrowphase | resolve
// position | resolve velocity) between CTA barriers, so the warps of a CTA execute the same
// code region at the same time (the kernel is ~7 k SASS instructions, far larger than the
// instruction caches; without this every warp is in a different region and fetch-bound).
enum { PH_A = 1, PH_B = 2, PH_C = 4, PH_ALL = 7 };
// MAT: per-pair surface materials (cz_world_set_materials): Friction / Restitution of every contact live in the
// world's as-generated arrays (global, read for the winner and the contacts it touches); without MAT they are
// the compile-time constants 0.9 / 0.1 of the reference.
template <int G, int MINB, bool LOCKSTEP, int PH, bool MAT>
__global__ void __launch_bounds__(128, MINB) k_world_fused(WorldParams p, FusedPlan fp, real dt, real bias, int nSteps, unsigned int *nextWorld) {
    using namespace czr;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / G, tid = threadIdx.x % G;
    const unsigned mask = 0xffffffffu;   // all collectives are warp-wide with width G: the groups of a warp stay converged
    const int B = p.B, Cc = p.Cc;
    Staged s;
    // FULL: the launch runs phase A (integrate + narrowphase) and needs collider data and the check queues; the loop
    // phases of split mode use the short record of world_bytes_loops
    constexpr bool FULL = (PH & PH_A) != 0;
    // A_ONLY: the phase-A launch of split mode hands the hot contact fields to the loop launches through the
    // per-world global arrays anyway, so it writes them there directly (each is touched two or three times per
    // contact, never scanned) and keeps no copy in shared memory: a 3.4 KB record instead of 5 KB, 3 CTAs per SM
    constexpr bool A_ONLY = PH == PH_A;
    constexpr bool DIRECT = PH != PH_ALL;   // as-generated contacts are written in place (stage_gen_direct)
    unsigned char *base = smem_raw + (size_t)grp * (A_ONLY ? fp.phaseAWorldBytes : (FULL ? fp.worldBytes : fp.loopWorldBytes));
    s.fb = (real *)base; s.bs = B;
    const int ch = A_ONLY ? 0 : Cc;   // hot contact fields held in shared memory
    s.pen = s.fb + (size_t)(FULL ? (int)FB_NF : (int)BW_NF) * B;
    s.ddv = FULL ? s.pen + ch : s.pen;   // a loop-phase launch holds one hot field: B scans pen, C scans ddv
    unsigned long long *const bmask = (unsigned long long *)(s.ddv + (FULL ? ch : ((ch + 1) & ~1)));   // [B]
    s.cb0 = (int *)(bmask + (A_ONLY ? 0 : B));
    s.cb1 = s.cb0 + ch;
    s.flags = s.cb1 + ch;                // flags, activation and the check queues exist only where phase A runs
    s.active = s.flags + (FULL ? B : 0);
    const int nq = FULL ? p.nchk : 0;
    s.info = (unsigned short *)(s.active + (FULL ? B : 0));
    s.queue = s.info + nq;
    s.pbase = s.queue + nq;
    s.cnt = (unsigned char *)(s.pbase + nq);
    s.pmask = s.cnt + nq;
    unsigned char *mlist = s.pmask + nq;
    real *const groupScratch = fp.cold + ((size_t)blockIdx.x * fp.groupsPerBlock + grp) * fp.coldReals;
    s.cold = groupScratch;
    s.pairGen = groupScratch + (size_t)Cc * CW_NCOLD;
    const BodyStore &st = p.st;

    Ctx x;
    x.bw = s.fb; x.bs = B;
    x.cold = s.cold; x.cfs = 1; x.ccs = CW_NCOLD;       // AoS
    x.pen = s.pen; x.ddv = s.ddv; x.fric = nullptr; x.rest = nullptr;
    x.cb0 = s.cb0; x.cb1 = s.cb1; x.nC = 0; x.dt = dt;
    x.mlist = (Cc <= 256 && !A_ONLY && (FULL || !(Cc <= 64 && fp.bodyMasks))) ? mlist : nullptr;   // the loops' scratch: absent from the phase-A record, and from a loop record that has masks
    x.bmask = (Cc <= 64 && fp.bodyMasks && !A_ONLY) ? bmask : nullptr;
    x.xb = nullptr; x.xbs = 0; x.store = st;
    x.pre = (fp.velPre && PH == PH_ALL) ? groupScratch + (size_t)Cc * CW_NCOLD + (size_t)p.nchk * 8 : nullptr;
    GenView gv;
    gv.pn = s.cold; gv.fs = 1; gv.cs = CW_NCOLD; gv.pen = s.pen; gv.fric = nullptr; gv.rest = nullptr; gv.b0 = s.cb0; gv.b1 = s.cb1;

    unsigned long long accContacts = 0, accPos = 0, accVel = 0;
    int maxC = 0, status = 0;

    while (true) {
        // ---- fetch the next world (dynamic scheduling over persistent groups) ---------------
        unsigned int wu = 0;
        if (tid == 0) wu = atomicAdd(nextWorld, 1u);
        wu = __shfl_sync(mask, wu, 0, G);
        const bool live = wu < (unsigned)p.wCount;
        if (LOCKSTEP) {
            if (!__syncthreads_or(live ? 1 : 0)) break;     // every group of the CTA is out of worlds
        } else if (!__any_sync(mask, live ? 1 : 0)) {
            break;
        }
        const int w = live ? (p.order ? p.order[p.wFirst + (int)wu] : p.wFirst + (int)wu) : 0;
        const long long gbase = (long long)w * B;
        x.body_base = gbase;
        if (MAT) {
            const long long gsAll = (long long)p.W * Cc;
            x.fric = p.gen + G_FRIC * gsAll + (long long)w * Cc;
            x.rest = p.gen + G_REST * gsAll + (long long)w * Cc;
            gv.fric = x.fric; gv.rest = x.rest;
        }
        if (live) stage_world<G, FULL>(s, st, gbase, B, tid);
        int lastC = 0, lastPos = 0, lastVel = 0;
        if (PH != PH_ALL) {   // split mode: contact state of this world lives in global memory between launches
            s.cold = fp.coldW + (size_t)w * Cc * CW_NCOLD;
            x.cold = s.cold; gv.pn = s.cold;
            if ((PH & PH_C) && fp.preW) x.pre = fp.preW + (size_t)w * Cc * VP_NF;
            if (DIRECT && (PH & PH_A)) {
                const long long gsAll = (long long)p.W * Cc, o = (long long)w * Cc;
                gv.pn = p.gen + o; gv.fs = (int)gsAll; gv.cs = 1;
                gv.pen = p.gen + G_PEN * gsAll + o; gv.b0 = p.gb0 + o; gv.b1 = p.gb1 + o;
            }
            if (A_ONLY) {
                const size_t o = (size_t)w * Cc;
                s.pen = fp.hotPen + o; s.ddv = fp.hotDdv + o; s.cb0 = fp.hotCb0 + o; s.cb1 = fp.hotCb1 + o;
                x.pen = s.pen; x.ddv = s.ddv; x.cb0 = s.cb0; x.cb1 = s.cb1;
            }
            if (!(PH & PH_A) && live) {
                lastC = p.nContacts[w];
                const int nL = lastC > Cc ? 0 : lastC;
                const size_t o = (size_t)w * Cc;
                for (int c = tid; c < nL; c += G) {
                    s.cb0[c] = fp.hotCb0[o + c]; s.cb1[c] = fp.hotCb1[o + c];
                    if (PH & PH_B) s.pen[c] = fp.hotPen[o + c];
                    if (PH & PH_C) s.ddv[c] = fp.hotDdv[o + c];
                }
            }
        }
        __syncwarp(mask);

        for (int stepNo = 0; stepNo < nSteps; stepNo++) {
            const long long step = p.step_index + stepNo;
            if (LOCKSTEP) __syncthreads();
            int nC = 0;
            if (PH & PH_A) {
            if (live && episode_wraps(p, w, step)) {   // RL-style episode reset: restore the snapshot
                stage_world<G, FULL>(s, p.snap, gbase, B, tid);
                for (int b = tid; b < B; b += G) {   // the body transform lives in the global store
#pragma unroll
                    for (int k = czb::C_L2T0; k <= czb::C_T11W0; k++) st.st(k, gbase + b, p.snap.ld(k, gbase + b));
                    if (st.force) {   // forces still waiting in the accumulators do not survive a reset
#pragma unroll
                        for (int k = 0; k < 3; k++) { st.force[(gbase + b) * 3 + k] = R_(0); st.torque[(gbase + b) * 3 + k] = R_(0); }
                    }
                }
            }
            __syncwarp(mask);   // (warp-wide collectives only at warp-uniform points)
            // ---- updateObjects: Integrate + collider derive (cubedrop.go:29-39) ---------
            for (int b = live ? tid : B; b < B; b += G) {
                const int fl = s.flags[b];
                if (!((fl & FF_INTEG) && step >= (long long)s.active[b])) continue;
                const long long gi = gbase + b;
                M34 tr;
                if (s.fb[BW_AWAKE * B + b] != R_(0)) {
                    real2 a01 = st.ld(czb::C_A01, gi), a2lp = st.ld(czb::C_A2LP, gi), apw0 = st.ld(czb::C_APW0, gi);
                    real2 i12 = st.ld(czb::C_I12, gi), i34 = st.ld(czb::C_I34, gi), i56 = st.ld(czb::C_I56, gi), i78 = st.ld(czb::C_I78, gi);
                    V3 pos = bw3(x, BW_POS, b), vel = bw3(x, BW_VEL, b), rot = bw3(x, BW_ROT, b), acc = mk3(a01.x, a01.y, a2lp.x);
                    Q4 q = bw_q(x, b);
                    M3 ib;
                    ib.c[0] = apw0.y; ib.c[1] = i12.x; ib.c[2] = i12.y; ib.c[3] = i34.x; ib.c[4] = i34.y; ib.c[5] = i56.x; ib.c[6] = i56.y; ib.c[7] = i78.x; ib.c[8] = i78.y;
                    czb::Integrated o;
                    if (st.force) {   // live accumulators (cz_world_add_forces); the staged world inertia is still the previous frame's (rigidbody.go:223)
                        const V3 f = mk3(st.force[gi * 3], st.force[gi * 3 + 1], st.force[gi * 3 + 2]), tq = mk3(st.torque[gi * 3], st.torque[gi * 3 + 1], st.torque[gi * 3 + 2]);
                        czb::integrate_body_forces(o, pos, q, vel, rot, acc, ib, s.fb[BW_MOTION * B + b], (fl & FF_CANSLEEP) != 0, dt, a2lp.y, apw0.x, bias,
                                                   f, tq, s.fb[BW_INVM * B + b], bw_iitw(x, b));
#pragma unroll
                        for (int k = 0; k < 3; k++) { st.force[gi * 3 + k] = R_(0); st.torque[gi * 3 + k] = R_(0); }
                    } else {
                        czb::integrate_body(o, pos, q, vel, rot, acc, ib, s.fb[BW_MOTION * B + b], (fl & FF_CANSLEEP) != 0, dt, a2lp.y, apw0.x, bias);
                    }
                    bw3_set(x, BW_POS, b, o.pos); bw3_set(x, BW_VEL, b, o.vel); bw3_set(x, BW_ROT, b, o.rot); bw3_set(x, BW_LACC, b, o.lastAcc);
#pragma unroll
                    for (int k = 0; k < 4; k++) s.fb[(BW_Q + k) * B + b] = o.q.c[k];
#pragma unroll
                    for (int k = 0; k < 9; k++) s.fb[(BW_IITW + k) * B + b] = o.iitWorld.c[k];
                    s.fb[BW_MOTION * B + b] = o.motion;
                    if (!o.awake) s.fb[BW_AWAKE * B + b] = R_(0);
                    tr = o.transform;
                    // body transform: written through to the store (chunks C_L2T0 .. C_T11W0)
                    st.st(czb::C_L2T0, gi, make_real2(o.lastAcc.c[2], tr.c[0]));
                    st.st(czb::C_T12, gi, make_real2(tr.c[1], tr.c[2]));
                    st.st(czb::C_T34, gi, make_real2(tr.c[3], tr.c[4]));
                    st.st(czb::C_T56, gi, make_real2(tr.c[5], tr.c[6]));
                    st.st(czb::C_T78, gi, make_real2(tr.c[7], tr.c[8]));
                    st.st(czb::C_T910, gi, make_real2(tr.c[9], tr.c[10]));
                    st.st(czb::C_T11W0, gi, make_real2(tr.c[11], o.iitWorld.c[0]));
                } else {
                    tr = czb::ld_transform(st, gi);
                }
                if ((fl & FF_SHAPE_MASK) != CZ_SHAPE_NONE) {
                    M34 off = (fl & FF_IDENT) ? czb::identity34() : czb::ld_m34(st, czb::C_O01, gi);
                    M34 ctr = m34_mul_m34(tr, off);
#pragma unroll
                    for (int k = 0; k < 12; k++) s.fb[(FB_CTR + k) * B + b] = ctr.c[k];
                }
            }
            __syncwarp(mask);
            if (LOCKSTEP && fp.lockstep >= 2) __syncthreads();

            // ---- generateContacts (cubedrop.go:42-67) -----------------------------------------------------
            // A: every check is classified with cheap tests only: plane checks are queued (from the back of
            //    the queue), pair checks that survive the bounding-sphere rejection are queued (from the
            //    front).  B1 / B2: the queued plane checks (8 vertex transforms) and pair tests (15-axis SAT
            //    etc.) run densely, one per lane — in the all-pairs schedule one check in nine is a plane
            //    check and a few pair checks reach the full test; evaluated in place they left 1 of 8 lanes
            //    busy (ncu: 3.5 of 32 lanes active on those lines).  C: order-preserving slot assignment in
            //    check order (the reference's append order): prefix scan of the per-check contact counts; pair
            //    contacts are copied to their slot.  D: plane contacts are produced densely, one (check,
            //    vertex) per lane, at slot = base of the check + rank of the vertex among the set bits.
            int nQ = 0, nP = 0;
            const unsigned gshift = (threadIdx.x & 31u) & ~(unsigned)(G - 1);
            const unsigned gbits = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
            unsigned short *const queueP = s.queue + (p.nchk - 1);   // queueP[-q]
            for (int k0 = 0; k0 < p.nchk; k0 += G) {
                const int k = k0 + tid;
                bool need = false, plane = false;
                if (live && k < p.nchk) {
                    int a = 0, b2 = 0;
                    if (decode_check(p, k, a, b2) && !(a < 0 && b2 < 0) && (a < 0 || staged_active(s, a, step)) && (b2 < 0 || staged_active(s, b2, step))) {
                        if (a < 0 || b2 < 0) {
                            plane = (s.flags[a < 0 ? b2 : a] & FF_SHAPE_MASK) != CZ_SHAPE_NONE;
                        } else {
                            // bounding-sphere rejection needs only the centres and sizes
                            ColliderView one, two;
                            one.shape = s.flags[a] & FF_SHAPE_MASK; two.shape = s.flags[b2] & FF_SHAPE_MASK;
#pragma unroll
                            for (int j = 0; j < 3; j++) {
                                one.t.c[9 + j] = s.fb[(FB_CTR + 9 + j) * B + a]; two.t.c[9 + j] = s.fb[(FB_CTR + 9 + j) * B + b2];
                                one.half.c[j] = s.fb[(FB_HALF + j) * B + a]; two.half.c[j] = s.fb[(FB_HALF + j) * B + b2];
                            }
                            one.radius = s.fb[FB_RADIUS * B + a]; two.radius = s.fb[FB_RADIUS * B + b2];
                            need = !czn::bounding_reject(one, two);
                        }
                    }
                    s.cnt[k] = 0;
                    s.info[k] = 0;
                }
                const unsigned ballQ = (__ballot_sync(mask, need) >> gshift) & gbits;
                const unsigned ballP = (__ballot_sync(mask, plane) >> gshift) & gbits;
                if (need) s.queue[nQ + __popc(ballQ & ((1u << tid) - 1u))] = (unsigned short)k;
                if (plane) {
                    const int qp = nP + __popc(ballP & ((1u << tid) - 1u));
                    queueP[-qp] = (unsigned short)k;
                    s.info[k] = (unsigned short)qp;
                }
                nQ += __popc(ballQ);
                nP += __popc(ballP);
            }
            __syncwarp(mask);
            for (int qp = tid; qp < nP; qp += G) {   // pass B1: dense plane checks
                const int k = queueP[-qp];
                int a = 0, b2 = 0;
                decode_check(p, k, a, b2);
                const int ci = a < 0 ? b2 : a, pi = a < 0 ? -a - 1 : -b2 - 1;
                ColliderView c = staged_collider(s, ci);
                unsigned vm = 0;
                if (c.shape == CZ_SHAPE_SPHERE) {
                    GenContact gc;
                    vm = czn::sphere_halfspace(c, p.planes[pi], gc) ? 1u : 0u;
                } else if (c.shape == CZ_SHAPE_CUBE) {
                    vm = czn::cube_halfspace_mask(c, p.planes[pi]);
                }
                s.cnt[k] = (unsigned char)cz_popc(vm);
                s.pmask[qp] = (unsigned char)vm;
            }
            for (int q = tid; q < nQ; q += G) {   // pass B2: dense pair tests
                const int k = s.queue[q];
                int a = 0, b2 = 0;
                decode_check(p, k, a, b2);
                ColliderView one = staged_collider(s, a), two = staged_collider(s, b2);
                V3 v1 = zero3(), v2 = zero3();
                if (one.shape != two.shape) { v1 = bw3(x, BW_VEL, a); v2 = bw3(x, BW_VEL, b2); }
                GenContact gc;
                const bool hit = czn::check_pair(one, two, v1, v2, gc);
                s.cnt[k] = hit ? 1 : 0;
                s.info[k] = (unsigned short)q;
                if (hit) {
                    real *r = s.pairGen + (size_t)q * 8;
#pragma unroll
                    for (int j = 0; j < 3; j++) { r[j] = gc.point.c[j]; r[3 + j] = gc.normal.c[j]; }
                    r[6] = gc.pen;
                    r[7] = (real)(gc.b0 * 128 + (gc.b1 + 1));   // bodies (one,two) may be swapped by the test
                }
            }
            __syncwarp(mask);
            for (int k0 = 0; k0 < p.nchk; k0 += G) {   // pass C: ordered slots (uniform trip count: warp-wide scan)
                const int k = k0 + tid;
                const int cnt = (live && k < p.nchk) ? (int)s.cnt[k] : 0;
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < G; o <<= 1) {
                    int t = __shfl_up_sync(mask, incl, o, G);
                    if (tid >= o) incl += t;
                }
                const int total = __shfl_sync(mask, incl, G - 1, G);
                const int slot = nC + incl - cnt;
                if (cnt > 0) {
                    int a = 0, b2 = 0;
                    decode_check(p, k, a, b2);
                    if (a < 0 || b2 < 0) {
                        s.pbase[s.info[k]] = (unsigned short)min(slot, 0xffff);
                    } else if (slot < Cc) {
                        const real *r = s.pairGen + (size_t)s.info[k] * 8;
                        GenContact gc;
#pragma unroll
                        for (int j = 0; j < 3; j++) { gc.point.c[j] = r[j]; gc.normal.c[j] = r[3 + j]; }
                        gc.pen = r[6];
                        const int code = (int)r[7];
                        gc.b0 = code / 128; gc.b1 = code % 128 - 1;
                        if (DIRECT) stage_gen_direct(gv, slot, gc);
                        else stage_gen(s, slot, gc);
                        if (MAT) check_material(p, gbase, a, b2, x.fric[slot], x.rest[slot]);
                    }
                }
                nC += total;
            }
            __syncwarp(mask);
            for (int item = tid; item < nP * 8; item += G) {   // pass D: dense plane contacts, one (check, vertex) per lane
                const int qp = item >> 3, v = item & 7;
                const unsigned vm = s.pmask[qp];
                if (!((vm >> v) & 1u)) continue;
                const int slot = (int)s.pbase[qp] + cz_popc(vm & ((1u << v) - 1u));
                if (slot >= Cc) continue;
                const int k = queueP[-qp];
                int a = 0, b2 = 0;
                decode_check(p, k, a, b2);
                const int ci = a < 0 ? b2 : a, pi = a < 0 ? -a - 1 : -b2 - 1;
                ColliderView c = staged_collider(s, ci);
                GenContact gc;
                if (c.shape == CZ_SHAPE_SPHERE) czn::sphere_halfspace(c, p.planes[pi], gc);
                else czn::cube_halfspace_contact(c, p.planes[pi], v, gc);
                if (DIRECT) stage_gen_direct(gv, slot, gc);
                else stage_gen(s, slot, gc);
                if (MAT) check_material(p, gbase, a, b2, x.fric[slot], x.rest[slot]);
            }
            __syncwarp(mask);
            accContacts += (unsigned long long)nC;
            maxC = max(maxC, nC);
            lastC = nC;
            lastPos = lastVel = 0;
            if (nC > Cc) { status = CZ_ERR_CAPACITY; nC = 0; }
            if (!DIRECT && fp.keepContacts && stepNo == nSteps - 1 && nC > 0) {
                // dump the as-generated contacts of the last frame for cz_world_download_contacts
                const long long gs = (long long)p.W * Cc;
                real *gen = p.gen + (long long)w * Cc;
                for (int c = tid; c < nC; c += G) {
                    const real *r = s.cold + (size_t)c * CW_NCOLD;
#pragma unroll
                    for (int f = 0; f < 6; f++) gen[f * gs + c] = r[f];
                    gen[G_PEN * gs + c] = s.pen[c];
                    if (!MAT) { gen[G_FRIC * gs + c] = R_(0.9); gen[G_REST * gs + c] = R_(0.1); }
                    p.gb0[(long long)w * Cc + c] = s.cb0[c];
                    p.gb1[(long long)w * Cc + c] = s.cb1[c];
                }
            }
            __syncwarp(mask);
            // ---- ResolveContacts(8*len) (cubedrop.go:72-74) -----------------------------------
            if (LOCKSTEP && fp.lockstep >= 2) __syncthreads();
            x.nC = nC;
            for (int c = tid; c < nC; c += G) prepare_contact(x, c, gv);
            __syncwarp(mask);
            } else {   // phases B / C of split mode: the contacts were prepared by the phase-A launch
                nC = lastC > Cc ? 0 : lastC;
                x.nC = nC;
            }
            int st2 = 0;
            if ((PH & (PH_B | PH_C)) && x.bmask) build_body_masks<G>(x, B, tid);   // group-uniform: every lane of a group has the same x.nC
            if (PH & PH_B) {
                lastPos = resolve_loop<G, false>(x, nC > 0, nC * 8, tid, &st2);
                accPos += (unsigned long long)lastPos;
            }
            if (LOCKSTEP && fp.lockstep >= 3) __syncthreads();
            if (PH & PH_C) {
                lastVel = resolve_loop<G, true>(x, nC > 0, nC * 8, tid, &st2);
                accVel += (unsigned long long)lastVel;
            }
            if (st2) status = st2;
            __syncwarp(mask);
        }

        // ---- write the world's state back ----------------------------------------------------
        for (int b = live ? tid : B; b < B; b += G) {
            const long long gi = gbase + b;
            st.st(czb::C_P01, gi, make_real2(s.fb[(BW_POS + 0) * B + b], s.fb[(BW_POS + 1) * B + b]));
            st.st(czb::C_P2M, gi, make_real2(s.fb[(BW_POS + 2) * B + b], s.fb[BW_MOTION * B + b]));
            st.st(czb::C_Q01, gi, make_real2(s.fb[(BW_Q + 0) * B + b], s.fb[(BW_Q + 1) * B + b]));
            st.st(czb::C_Q23, gi, make_real2(s.fb[(BW_Q + 2) * B + b], s.fb[(BW_Q + 3) * B + b]));
            st.st(czb::C_V01, gi, make_real2(s.fb[(BW_VEL + 0) * B + b], s.fb[(BW_VEL + 1) * B + b]));
            st.st(czb::C_V2R0, gi, make_real2(s.fb[(BW_VEL + 2) * B + b], s.fb[(BW_ROT + 0) * B + b]));
            st.st(czb::C_R12, gi, make_real2(s.fb[(BW_ROT + 1) * B + b], s.fb[(BW_ROT + 2) * B + b]));
            st.st(czb::C_L01, gi, make_real2(s.fb[(BW_LACC + 0) * B + b], s.fb[(BW_LACC + 1) * B + b]));
            M34 tr = czb::ld_transform(st, gi);   // written through during the frames
            M3 iw;
#pragma unroll
            for (int k = 0; k < 9; k++) iw.c[k] = s.fb[(BW_IITW + k) * B + b];
            czb::st_derived(st, gi, s.fb[(BW_LACC + 2) * B + b], tr, iw);
            if (FULL) {   // the collider transform changes in updateObjects only
                M34 ctr;
#pragma unroll
                for (int k = 0; k < 12; k++) ctr.c[k] = s.fb[(FB_CTR + k) * B + b];
                czb::st_m34(st, czb::C_X01, gi, ctr);
            }
            st.awake[gi] = s.fb[BW_AWAKE * B + b] != R_(0) ? 1 : 0;
        }
        if (PH != PH_ALL && live) {   // hand the contact state to the next phase's launch
            const int nS = lastC > Cc ? 0 : lastC;
            const size_t o = (size_t)w * Cc;
            for (int c = tid; c < nS; c += G) {
                if (A_ONLY) break;   // written in place
                if (PH & PH_A) { fp.hotCb0[o + c] = s.cb0[c]; fp.hotCb1[o + c] = s.cb1[c]; fp.hotDdv[o + c] = s.ddv[c]; }
                if (PH & (PH_A | PH_B)) fp.hotPen[o + c] = s.pen[c];
            }
        }
        if (tid == 0 && live) {
            if (PH & PH_A) p.nContacts[w] = lastC;
            if (PH & PH_B) p.posIters[w] = lastPos;
            if (PH & PH_C) p.velIters[w] = lastVel;
        }
        __syncwarp(mask);
    }
    if (tid == 0) {
        if (accContacts) atomicAdd(&p.stats[ST_CONTACTS], accContacts);
        if (accPos) atomicAdd(&p.stats[ST_POS], accPos);
        if (accVel) atomicAdd(&p.stats[ST_VEL], accVel);
        atomicMax(&p.stats[ST_MAXC], (unsigned long long)maxC);
        if (status) raise_status(p.stats, status);
    }
}


// Processing order of the worlds of one launch: the counters of the previous frame (contacts for the
// integrate + narrowphase + prepare phase, position / velocity iterations for the two loops) predict
// this frame's cost well, so worlds are fetched in descending key order.  The expensive worlds start
// first (a short tail at the end of every launch — per-chunk launches of the host pipelines are
// only a few world-rounds long), and the worlds that share a warp run for a similar number of
// iterations (the groups of a warp iterate in lockstep to the longest of them).
// Counting sort of world ids into 64 buckets, one CTA per key; result-neutral: worlds are independent.
// ORDER_SLICES CTAs per key: CTA j sorts the worlds first + j, first + j + S, ... (a round-robin slice: the
// slices are statistically alike) and writes its r-th world to position r*S + j, so the merged order is sorted up
// to the differences between slices and no CTA waits for another.  The same launch zeroes the world counters of
// the phase kernels (counters[0..3]: phases A, B, C, all), saving one memset node per phase launch.
constexpr int ORDER_SLICES = 8;
__global__ void __launch_bounds__(1024) k_order_worlds(const int *kContacts, const int *kPos, const int *kVel, int first, int count, int W, int *order3,
                                                       unsigned int *counters) {
    // per-warp histograms (a few buckets hold most worlds: one shared counter per bucket would serialise),
    // warp w owns a contiguous run of the slice -> stable, deterministic order
    __shared__ unsigned hist[32][64];
    __shared__ unsigned bucketStart[64];
    const int which = blockIdx.x / ORDER_SLICES, slice = blockIdx.x % ORDER_SLICES;
    if (slice == 0 && threadIdx.x == 0) { counters[which] = 0; if (which == 2) counters[3] = 0; }
    const int *key = which == 0 ? kContacts : (which == 1 ? kPos : kVel);
    const int shift = which == 2 ? 1 : 0;   // contacts 0..63, position iterations 0..63+, velocity iterations in twos
    int *order = order3 + (size_t)which * W;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int len = count > slice ? (count - slice + ORDER_SLICES - 1) / ORDER_SLICES : 0;   // worlds of this slice
    const int per = ((len + 31) / 32 + 31) / 32 * 32;   // elements per warp, a multiple of 32
    const int lo = warp * per, hi = min(len, lo + per);
    for (int b = lane; b < 64; b += 32) hist[warp][b] = 0;
    __syncwarp();
    for (int k0 = lo; k0 < hi; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < hi;
        const unsigned b = valid ? 63u - min(63u, (unsigned)max(key[first + slice + k * ORDER_SLICES], 0) >> shift) : 64u;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (valid && lane == __ffs(peers) - 1) hist[warp][b] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x < 64) {   // per bucket: exclusive prefix over the warps, and the bucket total
        unsigned run = 0;
        for (int w2 = 0; w2 < 32; w2++) { const unsigned c = hist[w2][threadIdx.x]; hist[w2][threadIdx.x] = run; run += c; }
        bucketStart[threadIdx.x] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned run = 0;
        for (int b = 0; b < 64; b++) { const unsigned c = bucketStart[b]; bucketStart[b] = run; run += c; }
    }
    __syncthreads();
    for (int k0 = lo; k0 < hi; k0 += 32) {
        const int k = k0 + lane;
        const bool valid = k < hi;
        const unsigned b = valid ? 63u - min(63u, (unsigned)max(key[first + slice + k * ORDER_SLICES], 0) >> shift) : 64u;
        const unsigned peers = __match_any_sync(0xffffffffu, b);
        if (valid) {
            const unsigned r = bucketStart[b] + hist[warp][b] + __popc(peers & ((1u << lane) - 1u));   // rank inside the slice
            order[first + (int)r * ORDER_SLICES + slice] = first + slice + k * ORDER_SLICES;
        }
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) hist[warp][b] += __popc(peers);
        __syncwarp();
    }
}

// returns 0 or a cudaError_t.  phases = PH_ALL (persistent over nSteps frames) or one of PH_A/B/C.
// `counters` holds four world counters (phases A, B, C, all); countersZeroed: k_order_worlds of this frame cleared them.
static inline int launch(const FusedPlan &fp, const WorldParams &p, real dt, real bias, int nSteps, unsigned int *counters, cudaStream_t stream,
                         int phases = PH_ALL, bool countersZeroed = false) {
    unsigned int *nextWorld = counters + (phases == PH_A ? 0 : (phases == PH_B ? 1 : (phases == PH_C ? 2 : 3)));
    cudaError_t e = cudaSuccess;
    if (!countersZeroed) e = cudaMemsetAsync(nextWorld, 0, sizeof(unsigned int), stream);   // stream-ordered before the launch
    if (e != cudaSuccess) return (int)e;
    const bool loops = phases == PH_B || phases == PH_C;
    int grid = loops ? fp.loopGrid : (phases == PH_A ? fp.phaseAGrid : fp.grid);
    const size_t smemBytes = loops ? fp.loopSmemBytes : (phases == PH_A ? fp.phaseASmemBytes : fp.smemBytes);
    const int needed = (p.wCount + fp.groupsPerBlock - 1) / fp.groupsPerBlock;
    if (grid > needed) grid = needed;
#define CZF_LAUNCH(GG, MB, LS, PHS, MT)                                                                               \
    do {                                                                                                            \
        if (smemBytes > 48 * 1024)                                                                                  \
            e = cudaFuncSetAttribute(k_world_fused<GG, MB, LS, PHS, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes); \
        if (e == cudaSuccess) k_world_fused<GG, MB, LS, PHS, MT><<<grid, fp.threads, smemBytes, stream>>>(p, fp, dt, bias, nSteps, nextWorld); \
    } while (0)
#define CZF_LAUNCH_G(GG)                                                                                            \
    do {                                                                                                            \
        if (p.matFric) {   /* materials: the default register budget only */                                        \
            if (phases == PH_A) CZF_LAUNCH(GG, 2, false, PH_A, true);                                               \
            else if (phases == PH_B) CZF_LAUNCH(GG, 2, false, PH_B, true);                                          \
            else if (phases == PH_C) CZF_LAUNCH(GG, 2, false, PH_C, true);                                          \
            else CZF_LAUNCH(GG, 2, true, PH_ALL, true);                                                             \
        } else if (phases == PH_A && GG == 8 && fp.phaseAMinb == 3) { CZF_LAUNCH(8, 3, false, PH_A, false);             \
        } else if (phases == PH_A && GG == 8 && fp.phaseAMinb == 4) { CZF_LAUNCH(8, 4, false, PH_A, false);             \
        } else if (loops && GG == 8 && fp.splitMinb == 3) {                                                         \
            if (phases == PH_B) CZF_LAUNCH(8, 3, false, PH_B, false);                                               \
            else CZF_LAUNCH(8, 3, false, PH_C, false);                                                              \
        } else if (loops && GG == 8 && fp.splitMinb == 4) {                                                         \
            if (phases == PH_B) CZF_LAUNCH(8, 4, false, PH_B, false);                                               \
            else CZF_LAUNCH(8, 4, false, PH_C, false);                                                              \
        } else if (phases == PH_A) CZF_LAUNCH(GG, 2, false, PH_A, false);                                           \
        else if (phases == PH_B) CZF_LAUNCH(GG, 2, false, PH_B, false);                                             \
        else if (phases == PH_C) CZF_LAUNCH(GG, 2, false, PH_C, false);                                             \
        else if (fp.lockstep) {                                                                                     \
            if (fp.minb <= 2) CZF_LAUNCH(GG, 2, true, PH_ALL, false);                                               \
            else CZF_LAUNCH(GG, 3, true, PH_ALL, false);                                                            \
        } else {                                                                                                    \
            if (fp.minb <= 2) CZF_LAUNCH(GG, 2, false, PH_ALL, false);                                              \
            else CZF_LAUNCH(GG, 3, false, PH_ALL, false);                                                           \
        }                                                                                                           \
    } while (0)
if (fp.G == 8) CZF_LAUNCH_G(8);
    else if (fp.G == 16) CZF_LAUNCH_G(16);
    else CZF_LAUNCH_G(32);
#undef CZF_LAUNCH_G
#undef CZF_LAUNCH
    if (e == cudaSuccess) e = cudaGetLastError();
    return (int)e;
}

}  // namespace czf
