// cz_resolve.cuh — the contact resolver (K4): prepare + worst-first adjustPositions /
// adjustVelocities, one thread group (a warp, or a whole CTA) per world.
//
// The algorithm is inherently sequential across iterations (Gauss-Seidel, worst first, each
// iteration reads the previous one's writes — contact.go:233-283, :390-445).  Parallelism
// inside one iteration: (1) arg-max over the contacts by warp shuffles, lowest index wins ties,
// NaN never wins; (2) the scalar resolve of the winner is executed by the lanes of the first
// warp redundantly from staged operands (same issue cost as one lane, no divergence), lane 0
// commits; (3) every thread propagates the change to the contacts it owns, in the
// reference's (b, d) order.  Iteration caps are per phase.
//
// Staging: per world a "body work record" bw[field*bs + body] (28 reals) and a contact work
// record split into hot fields (penetration, desired delta-v, body ids: scanned every
// iteration — shared memory) and cold fields (18 reals: only the winner and its neighbours
// are touched — shared memory when it fits, else an L2-resident global scratch).  The code
// only sees pointers and strides.
#pragma once
#include "cz_body.cuh"
#include "cz_narrow.cuh"

namespace czr {
using namespace czm;

enum BodyWork : int { BW_POS = 0, BW_Q = 3, BW_VEL = 7, BW_ROT = 10, BW_LACC = 13, BW_IITW = 16, BW_INVM = 25, BW_MOTION = 26, BW_AWAKE = 27, BW_NF = 28 };
// Contact work record.  "Cold" fields (read only for the winner and for the contacts that
// share a body with it): normal, the two tangents, the two relative positions, the contact
// velocity.  "Hot" fields (scanned by every arg-max): penetration, desiredDeltaVelocity, body ids.
enum ContactCold : int { CW_N = 0, CW_TY = 3, CW_TZ = 6, CW_RP0 = 9, CW_RP1 = 12, CW_CV = 15, CW_NCOLD = 18 };
// as-generated contact: point and normal live in cold slots 0..5 until prepare overwrites them
enum GenField : int { G_POINT = 0, G_NORMAL = 3, G_PEN = 6, G_FRIC = 7, G_REST = 8, G_NF = 9 };

struct Ctx {
    real *bw; int bs;           // body work record bw[field*bs + body]
    real *cold;                 // cold contact fields: cold[field*cfs + contact*ccs]
    int cfs, ccs;               //   SoA: cfs = capacity, ccs = 1;  AoS: cfs = 1, ccs = CW_NCOLD
    real *pen, *ddv;            // hot: Penetration, desiredDeltaVelocity  [contact]
    real *fric, *rest;          // per-contact Friction / Restitution, or NULL: the constants 0.9 / 0.1
    int *cb0, *cb1;             // contact body indices (world-local, -1 = nil)
    unsigned char *mlist;       // optional [capacity <= 256]: compacted list of contacts touched by a resolve
    unsigned long long *bmask;  // optional [bodies], capacity <= 64: bit c set <=> contact c touches the body (build_body_masks)
    int nC;
    real dt;
    // The rare "resolved body is asleep" path (contact.go:380-382) reads the body-space inverse
    // inertia and rewrites transform + world inertia.  xb != NULL: staged copy
    // xb[(XB_IITB+k)*xbs + b], xb[(XB_TR+k)*xbs + b]; else the global store.
    real *xb; int xbs;
    czb::BodyStore store;
    int64_t body_base;          // global index of the world's body 0 in `store`
    real *pre = nullptr;        // optional [capacity][VP_NF] scratch: per-contact velocity response, evaluated once per frame by
                                // the velocity loop (see friction_response) instead of once per pick
};
enum ExtraBody : int { XB_IITB = 0, XB_TR = 9, XB_NF = 21 };

// as-generated contacts handed to prepare_contact (may alias the cold record, see above)
struct GenView {
    real *pn; int fs, cs;       // point/normal: pn[(G_POINT+k)*fs + c*cs]
    real *pen;                  // [contact]
    real *fric, *rest;          // [contact] or NULL (0.9 / 0.1)
    int *b0, *b1;
};

CZD V3 bw3(const Ctx &x, int f, int b) { return mk3(x.bw[(f + 0) * x.bs + b], x.bw[(f + 1) * x.bs + b], x.bw[(f + 2) * x.bs + b]); }
CZD void bw3_set(const Ctx &x, int f, int b, const V3 &v) { x.bw[(f + 0) * x.bs + b] = v.c[0]; x.bw[(f + 1) * x.bs + b] = v.c[1]; x.bw[(f + 2) * x.bs + b] = v.c[2]; }
CZD M3 bw_iitw(const Ctx &x, int b) { M3 m;
#pragma unroll
    for (int k = 0; k < 9; k++) m.c[k] = x.bw[(BW_IITW + k) * x.bs + b];
    return m; }
CZD Q4 bw_q(const Ctx &x, int b) { Q4 q;
#pragma unroll
    for (int k = 0; k < 4; k++) q.c[k] = x.bw[(BW_Q + k) * x.bs + b];
    return q; }
CZD bool bw_awake(const Ctx &x, int b) { return x.bw[BW_AWAKE * x.bs + b] != R_(0); }
CZD V3 cw3(const Ctx &x, int f, int c) {
    const real *p = x.cold + (size_t)f * x.cfs + (size_t)c * x.ccs;
    return mk3(p[0], p[x.cfs], p[2 * x.cfs]);
}
CZD void cw3_set(const Ctx &x, int f, int c, const V3 &v) {
    real *p = x.cold + (size_t)f * x.cfs + (size_t)c * x.ccs;
    p[0] = v.c[0]; p[x.cfs] = v.c[1]; p[2 * x.cfs] = v.c[2];
}
CZD real ctx_friction(const Ctx &x, int c) { return x.fric ? x.fric[c] : R_(0.9); }
CZD real ctx_restitution(const Ctx &x, int c) { return x.rest ? x.rest[c] : R_(0.1); }

// contact.go:87-112
CZD real desired_delta_velocity(const Ctx &x, int b0, int b1, const V3 &n, real cvx, real restitution) {
    real vfa = R_(0);
    if (bw_awake(x, b0)) {
        V3 t = bw3(x, BW_LACC, b0);
        v_mul(t, x.dt);
        vfa += v_dot(t, n);
    }
    if (b1 >= 0 && bw_awake(x, b1)) {
        V3 t = bw3(x, BW_LACC, b1);
        v_mul(t, x.dt);
        vfa -= v_dot(t, n);
    }
    real rest = restitution;
    if (rabs(cvx) < R_(0.25)) rest = R_(0);
    return -cvx - rest * (cvx - vfa);
}

// contact.go:159-182
CZD V3 local_velocity(const Ctx &x, int b, const V3 &rp, const V3 &n, const V3 &ty, const V3 &tz) {
    V3 vel = v_cross(bw3(x, BW_ROT, b), rp);
    v_add(vel, bw3(x, BW_VEL, b));
    // contactToWorld columns are (n, ty, tz); TransformTranspose (math/matrix.go:157) = dots with the columns
    V3 cv = mk3(v_dot(vel, n), v_dot(vel, ty), v_dot(vel, tz));
    V3 acc = bw3(x, BW_LACC, b);
    v_mul(acc, x.dt);
    V3 a2 = mk3(v_dot(acc, n), v_dot(acc, ty), v_dot(acc, tz));
    a2.c[0] = R_(0);
    v_add(cv, a2);
    return cv;
}

// contact.go:59-85 + :118-156 for contact c.  All as-generated fields are read before anything
// is written: `g` may alias the cold record (same contact column).
CZD void prepare_contact(const Ctx &x, int c, const GenView &g, int src = -1) {
    // src: index of the as-generated contact when it differs from the work slot c (island slices renumber the contacts)
    const int cs = src >= 0 ? src : c;
    int b0 = g.b0[cs], b1 = g.b1[cs];
    const real *gp = g.pn + (size_t)cs * g.cs;
    V3 point = mk3(gp[(G_POINT + 0) * g.fs], gp[(G_POINT + 1) * g.fs], gp[(G_POINT + 2) * g.fs]);
    V3 n = mk3(gp[(G_NORMAL + 0) * g.fs], gp[(G_NORMAL + 1) * g.fs], gp[(G_NORMAL + 2) * g.fs]);
    const real restitution = g.rest ? g.rest[cs] : R_(0.1), friction = g.fric ? g.fric[cs] : R_(0.9), pen0 = g.pen[cs];
    if (b0 < 0) {   // :61-65
        v_mul(n, R_(-1.0));
        b0 = b1;
        b1 = -1;
    }
    V3 ty, tz;      // calculateContactBasis :118-156
    if (rabs(n.c[0]) > rabs(n.c[1])) {
        real s = rdiv(R_(1.0), rsqrt_(n.c[2] * n.c[2] + n.c[0] * n.c[0]));
        ty.c[0] = n.c[2] * s; ty.c[1] = R_(0); ty.c[2] = n.c[0] * -s;
        tz.c[0] = n.c[1] * ty.c[0];
        tz.c[1] = n.c[2] * ty.c[0] - n.c[0] * ty.c[2];
        tz.c[2] = -n.c[1] * ty.c[0];
    } else {
        real s = rdiv(R_(1.0), rsqrt_(n.c[2] * n.c[2] + n.c[1] * n.c[1]));
        ty.c[0] = R_(0); ty.c[1] = -n.c[2] * s; ty.c[2] = n.c[1] * s;
        tz.c[0] = n.c[1] * ty.c[2] - n.c[2] * ty.c[1];
        tz.c[1] = -n.c[0] * ty.c[2];
        tz.c[2] = n.c[0] * ty.c[1];
    }
    V3 rp0 = point;
    v_sub(rp0, bw3(x, BW_POS, b0));
    V3 cv = local_velocity(x, b0, rp0, n, ty, tz);
    V3 rp1 = zero3();
    if (b1 >= 0) {
        rp1 = point;
        v_sub(rp1, bw3(x, BW_POS, b1));
        V3 cv1 = local_velocity(x, b1, rp1, n, ty, tz);
        v_sub(cv, cv1);
    }
    real ddv = desired_delta_velocity(x, b0, b1, n, cv.c[0], restitution);
    x.cb0[c] = b0; x.cb1[c] = b1;
    cw3_set(x, CW_N, c, n); cw3_set(x, CW_TY, c, ty); cw3_set(x, CW_TZ, c, tz);
    cw3_set(x, CW_RP0, c, rp0); cw3_set(x, CW_RP1, c, rp1); cw3_set(x, CW_CV, c, cv);
    x.ddv[c] = ddv;
    x.pen[c] = pen0;
    if (x.fric) x.fric[c] = friction;
    if (x.rest) x.rest[c] = restitution;
}

// What one resolve hands to the propagation step.
struct Change {
    V3 lin[2], ang[2];   // linearChange/angularChange (position) or velocityChange/rotationChange (velocity)
    int b[2];
};

// contact.go:185-202 on staged flags: returns which body (0/1) must be woken, or -1.
CZD int match_awake(bool a0, bool a1, int b1) {
    if (b1 < 0) return -1;
    if ((a0 || a1) && !(a0 && a1)) return a0 ? 1 : 0;
    return -1;
}

// What a resolve must write back to the bodies; committed by one lane after the whole group
// has finished reading the old state.
struct PosCommit {
    int b[2];
    V3 pos[2];
    Q4 q[2];
    bool asleep[2];   // body is (still) asleep: CalculateDerivedData() follows (contact.go:380-382)
    int wake;         // body slot woken by matchAwakeState, or -1
};

// contact.go:286-386 for the winner `c` with penetration `penetration`.  Pure: reads the staged
// state, returns the change (for the propagation) and the values to commit.  Executed by every
// lane of the group redundantly (same issue cost as one lane, no divergence).
CZD void resolve_position(const Ctx &x, int c, real penetration, Change &ch, PosCommit &pc) {
    const real angularLimit = R_(0.2);
    int b[2] = {x.cb0[c], x.cb1[c]};
    V3 n = cw3(x, CW_N, c);
    V3 rp[2] = {cw3(x, CW_RP0, c), cw3(x, CW_RP1, c)};
    bool awake[2] = {bw_awake(x, b[0]), b[1] >= 0 ? bw_awake(x, b[1]) : false};
    int wake = match_awake(awake[0], awake[1], b[1]);   // :252
    if (wake == 0) awake[0] = true;
    if (wake == 1) awake[1] = true;
    pc.wake = wake;
    real angularInertia[2] = {R_(0), R_(0)}, linearInertia[2] = {R_(0), R_(0)}, angularMove[2], linearMove[2];
    real totalInertia = R_(0);
    M3 iit[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        ch.lin[i] = zero3(); ch.ang[i] = zero3(); ch.b[i] = b[i];
        pc.b[i] = b[i]; pc.asleep[i] = false;
        if (b[i] < 0) continue;
        iit[i] = bw_iitw(x, b[i]);
        pc.pos[i] = bw3(x, BW_POS, b[i]);
        pc.q[i] = bw_q(x, b[i]);
        V3 aiw = v_cross(rp[i], n);
        aiw = m3_mul_v(iit[i], aiw);
        aiw = v_cross(aiw, rp[i]);
        angularInertia[i] = v_dot(aiw, n);
        linearInertia[i] = x.bw[BW_INVM * x.bs + b[i]];
        totalInertia += linearInertia[i] + angularInertia[i];
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (b[i] < 0) continue;
        real sign = i == 0 ? R_(1.0) : R_(-1.0);
        angularMove[i] = sign * penetration * rdiv(angularInertia[i], totalInertia);
        linearMove[i] = sign * penetration * rdiv(linearInertia[i], totalInertia);
        V3 proj = rp[i];
        v_add_scaled(proj, n, -v_dot(rp[i], n));
        real maxMag = angularLimit * v_mag(proj);
        if (angularMove[i] < -maxMag) {
            real total = angularMove[i] + linearMove[i];
            angularMove[i] = -maxMag;
            linearMove[i] = total - angularMove[i];
        } else if (angularMove[i] > maxMag) {
            real total = angularMove[i] + linearMove[i];
            angularMove[i] = maxMag;
            linearMove[i] = total - angularMove[i];
        }
        if (angularMove[i] == R_(0.0)) {
            ch.ang[i] = zero3();
        } else {
            V3 target = v_cross(rp[i], n);
            ch.ang[i] = m3_mul_v(iit[i], target);
            v_mul(ch.ang[i], rdiv(angularMove[i], angularInertia[i]));
        }
        ch.lin[i] = n;
        v_mul(ch.lin[i], linearMove[i]);
        v_add_scaled(pc.pos[i], n, linearMove[i]);
        q_add_scaled_vector(pc.q[i], ch.ang[i], R_(1.0));
        q_normalize(pc.q[i]);
        pc.asleep[i] = !awake[i];
    }
}

// The body writes of applyPositionChange (one lane).  A body that is (still) asleep gets
// CalculateDerivedData() so the move shows up in its transform / world inertia (:380-382).
CZD void commit_position(const Ctx &x, PosCommit &pc) {
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const int b = pc.b[i];
        if (b < 0) continue;
        if (pc.asleep[i]) {
            M3 iitBody;
            if (x.xb) {
#pragma unroll
                for (int k = 0; k < 9; k++) iitBody.c[k] = x.xb[(XB_IITB + k) * x.xbs + b];
            } else {
                iitBody = czb::ld_iit_body(x.store, x.body_base + b);
            }
            M34 tr;
            M3 iw;
            czb::calculate_derived(pc.pos[i], pc.q[i], iitBody, tr, iw);
            if (x.xb) {
#pragma unroll
                for (int k = 0; k < 12; k++) x.xb[(XB_TR + k) * x.xbs + b] = tr.c[k];
            } else {
                real laz = x.bw[(BW_LACC + 2) * x.bs + b];
                czb::st_derived(x.store, x.body_base + b, laz, tr, iw);
            }
#pragma unroll
            for (int k = 0; k < 9; k++) x.bw[(BW_IITW + k) * x.bs + b] = iw.c[k];
        }
        bw3_set(x, BW_POS, b, pc.pos[i]);
#pragma unroll
        for (int k = 0; k < 4; k++) x.bw[(BW_Q + k) * x.bs + b] = pc.q[i].c[k];
    }
    if (pc.wake >= 0) {   // SetAwake(true) rigidbody.go:183-186
        const int wb = pc.wake == 0 ? pc.b[0] : pc.b[1];
        x.bw[BW_AWAKE * x.bs + wb] = R_(1);
        x.bw[BW_MOTION * x.bs + wb] = R_(0.6);
    }
}

// contact.go:259-279 for one owned contact
CZD void propagate_position(const Ctx &x, int c, const Change &ch) {
    int cb[2] = {x.cb0[c], x.cb1[c]};
    if (cb[0] != ch.b[0] && cb[0] != ch.b[1] && cb[1] != ch.b[0] && (cb[1] != ch.b[1] || cb[1] < 0)) return;
    V3 n = cw3(x, CW_N, c);
    real pen = x.pen[c];
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (cb[b] < 0) continue;
        V3 rp = cw3(x, b == 0 ? CW_RP0 : CW_RP1, c);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (cb[b] == ch.b[d]) {
                V3 dp = v_cross(ch.ang[d], rp);
                v_add(dp, ch.lin[d]);
                real sign = b == 0 ? R_(-1.0) : R_(1.0);
                pen += v_dot(dp, n) * sign;
            }
        }
    }
    x.pen[c] = pen;
}

// contact.go:448-494 (+ :498-606) for the winner `c`.
struct VelCommit {
    int b[2];
    V3 vel[2], rot[2];
    int wake;
    int status;
};

// calculateFrictionImpulse, contact.go:535-577: the velocity change per unit impulse in contact coordinates (dv) and its
// inverse (im).  Functions of the relative contact positions, the bodies' world inverse inertia, their inverse masses
// and the contact basis only — none of which changes during adjustVelocities (contact.go:390-445 moves velocities, not
// positions), so a caller may evaluate this once per contact and frame (VEL_PRE record below) instead of once per pick:
// same expressions, same operands, same bits.
enum VelPre : int { VP_IM = 0, VP_DV0 = 9, VP_DV3 = 10, VP_DV6 = 11, VP_NF = 12 };
CZD void friction_response(const V3 rp[2], const M3 iit[2], const real invMass[2], bool two, const M3 &c2w, M3 &dv, M3 &im) {
    real inverseMass = invMass[0];
    M3 itt;
    set_skew(itt, rp[0]);
    M3 dvw = m3_mul_m(itt, iit[0]);
    dvw = m3_mul_m(dvw, itt);
#pragma unroll
    for (int k = 0; k < 9; k++) dvw.c[k] *= R_(-1.0);
    if (two) {
        set_skew(itt, rp[1]);
        M3 dvw2 = m3_mul_m(itt, iit[1]);
        dvw2 = m3_mul_m(dvw2, itt);
#pragma unroll
        for (int k = 0; k < 9; k++) dvw2.c[k] *= R_(-1.0);
#pragma unroll
        for (int k = 0; k < 9; k++) dvw.c[k] += dvw2.c[k];
        inverseMass += invMass[1];
    }
    dv = m3_transpose(c2w);
    dv = m3_mul_m(dv, dvw);
    dv = m3_mul_m(dv, c2w);
    dv.c[0] += inverseMass; dv.c[4] += inverseMass; dv.c[8] += inverseMass;
    im = m3_invert(dv);
}
CZD M3 contact_to_world(const V3 &n, const V3 &ty, const V3 &tz) {   // columns (n, ty, tz)  (SetComponents math/matrix.go:52)
    M3 c2w;
    c2w.c[0] = n.c[0]; c2w.c[1] = n.c[1]; c2w.c[2] = n.c[2];
    c2w.c[3] = ty.c[0]; c2w.c[4] = ty.c[1]; c2w.c[5] = ty.c[2];
    c2w.c[6] = tz.c[0]; c2w.c[7] = tz.c[1]; c2w.c[8] = tz.c[2];
    return c2w;
}
// the VEL_PRE record of contact c (12 reals): friction_response evaluated ahead of the loop
CZD void precompute_velocity_response(const Ctx &x, int c, real *pre) {
    const int b0 = x.cb0[c], b1 = x.cb1[c];
    const V3 rp[2] = {cw3(x, CW_RP0, c), cw3(x, CW_RP1, c)};
    M3 iit[2];
    real invMass[2] = {R_(0), R_(0)};
    iit[0] = bw_iitw(x, b0);
    invMass[0] = x.bw[BW_INVM * x.bs + b0];
#pragma unroll
    for (int k = 0; k < 9; k++) iit[1].c[k] = R_(0);
    if (b1 >= 0) { iit[1] = bw_iitw(x, b1); invMass[1] = x.bw[BW_INVM * x.bs + b1]; }
    const M3 c2w = contact_to_world(cw3(x, CW_N, c), cw3(x, CW_TY, c), cw3(x, CW_TZ, c));
    M3 dv, im;
    friction_response(rp, iit, invMass, b1 >= 0, c2w, dv, im);
#pragma unroll
    for (int k = 0; k < 9; k++) pre[VP_IM + k] = im.c[k];
    pre[VP_DV0] = dv.c[0]; pre[VP_DV3] = dv.c[3]; pre[VP_DV6] = dv.c[6];
}

// pre: the contact's VEL_PRE record or NULL (evaluate the response here)
CZD void resolve_velocity(const Ctx &x, int c, Change &ch, VelCommit &vc, const real *pre = nullptr) {
    int b[2] = {x.cb0[c], x.cb1[c]};
    V3 n = cw3(x, CW_N, c), ty = cw3(x, CW_TY, c), tz = cw3(x, CW_TZ, c);
    V3 rp[2] = {cw3(x, CW_RP0, c), cw3(x, CW_RP1, c)};
    V3 cv = cw3(x, CW_CV, c);
    real ddv = x.ddv[c];
    real friction = ctx_friction(x, c);
    bool awake0 = bw_awake(x, b[0]), awake1 = b[1] >= 0 ? bw_awake(x, b[1]) : false;
    vc.wake = match_awake(awake0, awake1, b[1]);   // :409
    vc.status = 0;
    M3 iit[2];
    real invMass[2] = {R_(0), R_(0)};
#pragma unroll
    for (int i = 0; i < 2; i++) {
        ch.lin[i] = zero3(); ch.ang[i] = zero3(); ch.b[i] = b[i]; vc.b[i] = b[i];
#pragma unroll
        for (int k = 0; k < 9; k++) iit[i].c[k] = R_(0);
        if (b[i] < 0) continue;
        iit[i] = bw_iitw(x, b[i]);
        invMass[i] = x.bw[BW_INVM * x.bs + b[i]];
        vc.vel[i] = bw3(x, BW_VEL, b[i]);
        vc.rot[i] = bw3(x, BW_ROT, b[i]);
    }
    const M3 c2w = contact_to_world(n, ty, tz);
    V3 ic;
    if (friction == R_(0.0)) {
        // calculateFrictionlessImpulse :498-531 (second-body block is guarded by Bodies[1]==nil
        // in the reference: skipped for two-body contacts, a nil dereference otherwise)
        V3 dvw = v_cross(rp[0], n);
        dvw = m3_mul_v(iit[0], dvw);
        dvw = v_cross(dvw, rp[0]);
        real dv = v_dot(dvw, n);
        dv += invMass[0];
        if (b[1] < 0) vc.status = CZ_ERR_NIL_BODY;
        ic = mk3(rdiv(ddv, dv), R_(0), R_(0));
    } else {
        // calculateFrictionImpulse :535-606
        M3 im;
        real dv0, dv3, dv6;
        if (pre) {
#pragma unroll
            for (int k = 0; k < 9; k++) im.c[k] = pre[VP_IM + k];
            dv0 = pre[VP_DV0]; dv3 = pre[VP_DV3]; dv6 = pre[VP_DV6];
        } else {
            M3 dv;
            friction_response(rp, iit, invMass, b[1] >= 0, c2w, dv, im);
            dv0 = dv.c[0]; dv3 = dv.c[3]; dv6 = dv.c[6];
        }
        V3 velKill = mk3(ddv, -cv.c[1], -cv.c[2]);
        ic = m3_mul_v(im, velKill);
        real planar = rsqrt_(ic.c[1] * ic.c[1] + ic.c[2] * ic.c[2]);
        if (planar > ic.c[0] * friction) {
            ic.c[1] = rdiv(ic.c[1], planar);
            ic.c[2] = rdiv(ic.c[2], planar);
            ic.c[0] = dv0 + dv3 * friction * ic.c[1] + dv6 * friction * ic.c[2];
            ic.c[0] = rdiv(ddv, ic.c[0]);
            ic.c[1] *= friction * ic.c[0];
            ic.c[2] *= friction * ic.c[0];
        }
    }
    V3 impulse = m3_mul_v(c2w, ic);
    V3 torque = v_cross(rp[0], impulse);
    ch.ang[0] = m3_mul_v(iit[0], torque);
    ch.lin[0] = zero3();
    v_add_scaled(ch.lin[0], impulse, invMass[0]);
    v_add(vc.vel[0], ch.lin[0]);
    v_add(vc.rot[0], ch.ang[0]);
    if (b[1] >= 0) {
        torque = v_cross(impulse, rp[1]);
        ch.ang[1] = m3_mul_v(iit[1], torque);
        ch.lin[1] = zero3();
        v_add_scaled(ch.lin[1], impulse, -invMass[1]);
        v_add(vc.vel[1], ch.lin[1]);
        v_add(vc.rot[1], ch.ang[1]);
    }
}

// The body writes of applyVelocityChange (AddVelocity / AddRotation, contact.go:478-490) (one lane).
CZD void commit_velocity(const Ctx &x, const VelCommit &vc) {
    bw3_set(x, BW_VEL, vc.b[0], vc.vel[0]);
    bw3_set(x, BW_ROT, vc.b[0], vc.rot[0]);
    if (vc.b[1] >= 0) {
        bw3_set(x, BW_VEL, vc.b[1], vc.vel[1]);
        bw3_set(x, BW_ROT, vc.b[1], vc.rot[1]);
    }
    if (vc.wake >= 0) {
        const int wb = vc.wake == 0 ? vc.b[0] : vc.b[1];
        x.bw[BW_AWAKE * x.bs + wb] = R_(1);
        x.bw[BW_MOTION * x.bs + wb] = R_(0.6);
    }
}

// contact.go:416-442 for one owned contact
CZD void propagate_velocity(const Ctx &x, int c, const Change &ch) {
    int cb[2] = {x.cb0[c], x.cb1[c]};
    if (cb[0] != ch.b[0] && cb[0] != ch.b[1] && cb[1] != ch.b[0] && (cb[1] != ch.b[1] || cb[1] < 0)) return;
    V3 n = cw3(x, CW_N, c), ty = cw3(x, CW_TY, c), tz = cw3(x, CW_TZ, c);
    V3 cv = cw3(x, CW_CV, c);
    real restitution = ctx_restitution(x, c);
    real ddv;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (cb[b] < 0) continue;
        V3 rp = cw3(x, b == 0 ? CW_RP0 : CW_RP1, c);
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (cb[b] == ch.b[d]) {
                V3 dv = v_cross(ch.ang[d], rp);
                v_add(dv, ch.lin[d]);
                real sign = b == 1 ? R_(-1.0) : R_(1.0);
                V3 t = mk3(v_dot(dv, n), v_dot(dv, ty), v_dot(dv, tz));
                v_mul(t, sign);
                v_add(cv, t);
                // calculateDesiredDeltaVelocity (:438) is a pure function of the final
                // contactVelocity and of flags that do not change here: evaluated once below.
            }
        }
    }
    ddv = desired_delta_velocity(x, cb[0], cb[1], n, cv.c[0], restitution);
    cw3_set(x, CW_CV, c, cv);
    x.ddv[c] = ddv;
}

// Arg-max of (value, index): larger value wins, equal values -> lower index (the reference's
// linear scan keeps the first strictly-greater element, contact.go:240-245 / :396-402).
CZD void argmax_combine(real &v, int &i, real ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
template <int WIDTH>
__device__ __forceinline__ void warp_argmax(real &v, int &i, unsigned mask) {
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) {
        real ov = __shfl_xor_sync(mask, v, o);
        int oi = __shfl_xor_sync(mask, i, o);
        argmax_combine(v, i, ov, oi);
    }
}
// The same arg-max over a full warp for values that are positive and not NaN (the owners' caches of the large-world
// loop: a cached value is > epsilon, or the epsilon sentinel itself): IEEE order equals the order of the raw bits, so
// three (float: two) redux.sync replace five rounds of 64-bit shuffles and compares.
__device__ __forceinline__ void warp_argmax_positive(real &v, int &i) {
    const unsigned full = 0xffffffffu;
#ifdef CUBEZ_REAL_FLOAT
    const unsigned key = __float_as_uint(v);
    const unsigned mk = __reduce_max_sync(full, key);
    i = __reduce_min_sync(full, key == mk ? i : 0x7fffffff);
    v = __uint_as_float(mk);
#else
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_max_sync(full, hi);
    const unsigned ml = __reduce_max_sync(full, hi == mh ? lo : 0u);
    i = __reduce_min_sync(full, (hi == mh && lo == ml) ? i : 0x7fffffff);
    v = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
#endif
}
// Per-body contact bitmasks for worlds with at most 64 contacts: mask[b] has bit c set when contact c
// touches body b.  Called by the NT lanes that own the world, after the contact body ids are final
// (prepare_contact may swap them).  Contacts do not change bodies inside the loops.
template <int NT>
__device__ __forceinline__ void build_body_masks(const Ctx &x, int nBodies, int tid, unsigned lanes = 0xffffffffu) {
    for (int b = tid; b < nBodies; b += NT) x.bmask[b] = 0ull;
    __syncwarp(lanes);
    for (int c = tid; c < x.nC; c += NT) {
        atomicOr(&x.bmask[x.cb0[c]], 1ull << c);
        const int b1 = x.cb1[c];
        if (b1 >= 0) atomicOr(&x.bmask[b1], 1ull << c);
    }
    __syncwarp(lanes);
}

// Position of the k-th set bit (k = 0 for the lowest) of a 64-bit mask, k < popc(m): a branch-free binary
// search on population counts (the __fns intrinsic is a software loop: ~9 % of the velocity loop's instructions).
__device__ __forceinline__ int select64(unsigned long long m, int k) {
    const unsigned lo = (unsigned)m, hi = (unsigned)(m >> 32);
    const int nLo = __popc(lo);
    const bool up = k >= nLo;
    unsigned w = up ? hi : lo;
    int pos = up ? 32 : 0, t;
    k = up ? k - nLo : k;
    t = __popc(w & 0xffffu); if (k >= t) { w >>= 16; k -= t; pos += 16; }
    t = __popc(w & 0xffu);   if (k >= t) { w >>= 8;  k -= t; pos += 8; }
    t = __popc(w & 0xfu);    if (k >= t) { w >>= 4;  k -= t; pos += 4; }
    t = __popc(w & 0x3u);    if (k >= t) { w >>= 2;  k -= t; pos += 2; }
    t = (int)(w & 1u);       if (k >= t) pos += 1;
    return pos;
}

// The worst-first loop of one phase for worlds owned by (sub-)warp groups of NT <= 32 lanes.
// Called by all 32 lanes of a warp together: the 32/NT worlds of the warp iterate in lock step
// (a world that is done idles) so every collective uses the full-warp mask and the groups stay
// converged — sub-warp masks make the hardware run each group's instruction stream separately.
// `enabled` and every branch condition are uniform within a group.
template <int NT, bool VELOCITY>
__device__ __forceinline__ int resolve_loop(const Ctx &x, bool enabled, int maxIterations, int tid, int *status) {
    static_assert(NT <= 32, "warp-level loop");
    const real *hot = VELOCITY ? x.ddv : x.pen;
    const unsigned full = 0xffffffffu;
    const unsigned gshift = (threadIdx.x & 31u) & ~(unsigned)(NT - 1);
    const unsigned gbits = NT == 32 ? 0xffffffffu : ((1u << NT) - 1u);
    bool done = !enabled || maxIterations <= 0;
    int maxC = enabled ? x.nC : 0;   // longest contact list among the worlds of this warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxC = max(maxC, __shfl_xor_sync(full, maxC, o));
    int used = 0;
    bool havePre = false;
    while (true) {
        real best = R_(0.01);   // positionEpsilon / velocityEpsilon (contact.go:12-13)
        int idx = 0x7fffffff;
        if (!done) {
            for (int c = tid; c < x.nC; c += NT) {
                real v = hot[c];
                if (v > best) { best = v; idx = c; }
            }
        }
        warp_argmax<NT>(best, idx, full);
        if (idx == 0x7fffffff) done = true;
        if (__all_sync(full, done)) break;
        if (VELOCITY && x.pre && !havePre) {
            // First pick of the frame: the worlds of this warp that have anything to resolve evaluate the 3x3 velocity
            // response of every contact (4 matrix products, determinant, inverse: two thirds of the serial resolve) ONCE,
            // one contact per lane, instead of redundantly in all lanes of the group at every pick.  Constants of this loop
            // (friction_response); worlds at rest never get here.
            for (int c0 = 0; c0 < maxC; c0 += NT) {
                const int c = c0 + tid;
                if (!done && c < x.nC && ctx_friction(x, c) != R_(0.0)) precompute_velocity_response(x, c, x.pre + (size_t)c * VP_NF);
            }
            __syncwarp();
            havePre = true;
        }
        // The contacts the winner will touch are the set bits of the masks of its two bodies.  (Prefetching their
        // cold records into L1 here, under the winner's FP64 chain, was tried: 1.5 % slower — the L1 left beside
        // the shared-memory carve-out is too small for the lines to survive.)
        unsigned long long m64 = 0;
        int nM = 0;
        if (x.bmask) {
            if (!done) {
                const int wb0 = x.cb0[idx], wb1 = x.cb1[idx];
                m64 = x.bmask[wb0];
                if (wb1 >= 0) m64 |= x.bmask[wb1];
            }
            nM = __popcll(m64);
        }
        Change ch;
        PosCommit pc;
        VelCommit vc;
        if (!done) {
            if (VELOCITY) resolve_velocity(x, idx, ch, vc, (x.pre && havePre) ? x.pre + (size_t)idx * VP_NF : nullptr);
            else resolve_position(x, idx, best, ch, pc);
        }
        __syncwarp();   // every lane has read the old body state before one lane overwrites it
        if (!done && tid == 0) {
            if (VELOCITY) { commit_velocity(x, vc); if (vc.status) *status = vc.status; }
            else commit_position(x, pc);
        }
        __syncwarp();
        if (x.bmask) {
            // Propagation from per-body contact bitmasks (at most 64 contacts): the contacts that share a body
            // with the winner are the set bits of two masks — no scan of the contact list per iteration
            // (the ballot-compaction below was 12 % of the velocity loop's instructions).
            for (int k = tid; k < nM; k += NT) {
                const int c = select64(m64, k);
                if (VELOCITY) propagate_velocity(x, c, ch);
                else propagate_position(x, c, ch);
            }
        } else if (x.mlist) {
            // Propagation, compacted: only the contacts that share a body with the winner change
            // (typically the 4 plane contacts of the same cube and a few pair contacts).  Pass 1 marks
            // them with a cheap id compare and ballot-compacts their indices; pass 2 runs the update
            // densely, one touched contact per lane.  (In place, ~3 of 4 lanes idled through the update.)
            int nM = 0;
            for (int c0 = 0; c0 < maxC; c0 += NT) {
                const int c = c0 + tid;
                bool m = false;
                if (!done && c < x.nC) {
                    const int c0b = x.cb0[c], c1b = x.cb1[c];
                    m = c0b == ch.b[0] || c0b == ch.b[1] || (c1b >= 0 && (c1b == ch.b[0] || c1b == ch.b[1]));
                }
                const unsigned ball = (__ballot_sync(full, m) >> gshift) & gbits;
                if (m) x.mlist[nM + __popc(ball & ((1u << tid) - 1u))] = (unsigned char)c;
                nM += __popc(ball);
            }
            __syncwarp();
            if (!done) {
                for (int k = tid; k < nM; k += NT) {
                    if (VELOCITY) propagate_velocity(x, x.mlist[k], ch);
                    else propagate_position(x, x.mlist[k], ch);
                }
            }
        } else if (!done) {
            for (int c = tid; c < x.nC; c += NT) {
                if (VELOCITY) propagate_velocity(x, c, ch);
                else propagate_position(x, c, ch);
            }
        }
        if (!done) {
            used++;
            if (used >= maxIterations) done = true;
        }
        __syncwarp();
    }
    return used;
}

// Shared scratch of the CTA-wide loop.
struct GroupScratch {
    real redv[32];
    int redi[32];
    real chg[12];
    int chb[2];
};

// The same loop for one world owned by a whole CTA of NT threads (large worlds).  Only the first
// warp runs the scalar resolve; the change is broadcast through shared memory.
template <int NT, bool VELOCITY>
__device__ __forceinline__ int resolve_loop_cta(const Ctx &x, int maxIterations, GroupScratch *gs, int tid, int *status) {
    const int lane = tid & 31, warp = tid >> 5;
    const real *hot = VELOCITY ? x.ddv : x.pen;
    const unsigned full = 0xffffffffu;
    int used = 0;
    while (used < maxIterations) {
        real best = R_(0.01);
        int idx = 0x7fffffff;
        for (int c = tid; c < x.nC; c += NT) {
            real v = hot[c];
            if (v > best) { best = v; idx = c; }
        }
        warp_argmax<32>(best, idx, full);
        if (lane == 0) { gs->redv[warp] = best; gs->redi[warp] = idx; }
        __syncthreads();
        best = lane < NT / 32 ? gs->redv[lane] : R_(0.01);
        idx = lane < NT / 32 ? gs->redi[lane] : 0x7fffffff;
        warp_argmax<32>(best, idx, full);
        if (idx == 0x7fffffff) break;
        Change ch;
        if (warp == 0) {
            PosCommit pc;
            VelCommit vc;
            if (VELOCITY) resolve_velocity(x, idx, ch, vc);
            else resolve_position(x, idx, best, ch, pc);
            __syncwarp();
            if (lane == 0) {
                if (VELOCITY) { commit_velocity(x, vc); if (vc.status) *status = vc.status; }
                else commit_position(x, pc);
#pragma unroll
                for (int k = 0; k < 3; k++) { gs->chg[k] = ch.lin[0].c[k]; gs->chg[3 + k] = ch.lin[1].c[k]; gs->chg[6 + k] = ch.ang[0].c[k]; gs->chg[9 + k] = ch.ang[1].c[k]; }
                gs->chb[0] = ch.b[0]; gs->chb[1] = ch.b[1];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; k++) { ch.lin[0].c[k] = gs->chg[k]; ch.lin[1].c[k] = gs->chg[3 + k]; ch.ang[0].c[k] = gs->chg[6 + k]; ch.ang[1].c[k] = gs->chg[9 + k]; }
        ch.b[0] = gs->chb[0]; ch.b[1] = gs->chb[1];
        for (int c = tid; c < x.nC; c += NT) {
            if (VELOCITY) propagate_velocity(x, c, ch);
            else propagate_position(x, c, ch);
        }
        __syncthreads();
        used++;
    }
    return used;
}


// The CTA-wide loop for ONE LARGE world (thousands of contacts; cfg3).  Every iteration of
// resolve_loop_cta scans all contacts twice from L2 (arg-max, then the body-id match of the update):
// at 14 k contacts that was ~13 of the 16 us of an iteration.  Here the phase's hot value
// (penetration or desired delta-v) and the body ids (16 bit) of every contact are staged in shared
// memory, and every thread keeps the arg-max of the contacts it owns (c = tid, tid+NT, ...) in
// registers: an iteration is a block reduce of the cached maxima, the scalar resolve by warp 0,
// and an update in which a thread rescans its own contacts only if one of them was touched.
// Same worst-first order as the reference: ties go to the lowest index in every reduce.
template <int NT, bool VELOCITY>
__device__ __forceinline__ int resolve_loop_cta_cached(const Ctx &x0, int maxIterations, GroupScratch *gs, int tid, int *status, real *sHot,
                                                       unsigned short *sB0, unsigned short *sB1) {
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    Ctx x = x0;
    real *gHot = VELOCITY ? x0.ddv : x0.pen;
    for (int c = tid; c < x.nC; c += NT) {
        sHot[c] = gHot[c];
        const int b0 = x.cb0[c], b1 = x.cb1[c];
        sB0[c] = (unsigned short)b0;
        sB1[c] = b1 < 0 ? (unsigned short)0xffffu : (unsigned short)b1;
    }
    if (VELOCITY) x.ddv = sHot; else x.pen = sHot;
    __syncthreads();
    real myBest = R_(0.01);   // positionEpsilon / velocityEpsilon (contact.go:12-13)
    int myIdx = 0x7fffffff;
    for (int c = tid; c < x.nC; c += NT) {
        const real v = sHot[c];
        if (v > myBest) { myBest = v; myIdx = c; }
    }
    int used = 0;
    while (used < maxIterations) {
        real best = myBest;
        int idx = myIdx;
        warp_argmax<32>(best, idx, full);
        if (lane == 0) { gs->redv[warp] = best; gs->redi[warp] = idx; }
        __syncthreads();
        best = lane < NT / 32 ? gs->redv[lane] : R_(0.01);
        idx = lane < NT / 32 ? gs->redi[lane] : 0x7fffffff;
        warp_argmax<32>(best, idx, full);
        if (idx == 0x7fffffff) break;
        Change ch;
        if (warp == 0) {
            PosCommit pc;
            VelCommit vc;
            if (VELOCITY) resolve_velocity(x, idx, ch, vc);
            else resolve_position(x, idx, best, ch, pc);
            __syncwarp();
            if (lane == 0) {
                if (VELOCITY) { commit_velocity(x, vc); if (vc.status) *status = vc.status; }
                else commit_position(x, pc);
#pragma unroll
                for (int k = 0; k < 3; k++) { gs->chg[k] = ch.lin[0].c[k]; gs->chg[3 + k] = ch.lin[1].c[k]; gs->chg[6 + k] = ch.ang[0].c[k]; gs->chg[9 + k] = ch.ang[1].c[k]; }
                gs->chb[0] = ch.b[0]; gs->chb[1] = ch.b[1];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; k++) { ch.lin[0].c[k] = gs->chg[k]; ch.lin[1].c[k] = gs->chg[3 + k]; ch.ang[0].c[k] = gs->chg[6 + k]; ch.ang[1].c[k] = gs->chg[9 + k]; }
        ch.b[0] = gs->chb[0]; ch.b[1] = gs->chb[1];
        // -1 (no second body) never matches: body ids are < 0xffff and 0xffff marks "nil" in sB1
        const unsigned m0 = (unsigned)ch.b[0] & 0xffffu, m1 = ch.b[1] < 0 ? 0x10000u : (unsigned)ch.b[1];
        bool touched = false;
        for (int c = tid; c < x.nC; c += NT) {
            const unsigned c0 = sB0[c], c1 = sB1[c];
            if (c0 == m0 || c0 == m1 || c1 == m0 || c1 == m1) {
                if (VELOCITY) propagate_velocity(x, c, ch);
                else propagate_position(x, c, ch);
                touched = true;
            }
        }
        if (touched) {
            myBest = R_(0.01); myIdx = 0x7fffffff;
            for (int c = tid; c < x.nC; c += NT) {
                const real v = sHot[c];
                if (v > myBest) { myBest = v; myIdx = c; }
            }
        }
        used++;
        // the next iteration's first __syncthreads (after the warp reduce) orders these shared-memory
        // writes before warp 0 reads the winner's hot value
    }
    __syncthreads();
    for (int c = tid; c < x.nC; c += NT) gHot[c] = sHot[c];
    __syncthreads();
    return used;
}


// ------------------------------------------------------------------------------------------------
// ONE LARGE world per CTA, third design (cfg3: 14 k contacts, both loops at their 8*len cap = 120 k strictly
// sequential iterations per frame).  ncu of the loop above at 14 k contacts (profiles/r02_resolve_large_world_before.txt):
// 14.9 k warp-instructions per iteration, 55 % of them the body-id scan of every contact, 11 % the re-scan of the
// owners' hot values, 25 % of the stall samples the barrier at which 7 warps wait for warp 0's scalar resolve.
// Here an iteration touches only what the winner touches:
//   * adjacency lists body -> contacts (CSR, 16-bit ids, shared memory), built once per frame: the contacts the
//     winner's change reaches are the two lists of its bodies — O(degree), no scan;
//   * the per-owner arg-max cache (owner o = contacts o, o+NT, ...) lives in shared memory; an owner is re-scanned
//     only when its cached best was touched or overtaken, by its whole warp (2 loads per lane);
//   * while warp 0 computes the winner's resolve (one L2 round trip + the FP64 chain), the other warps already hold
//     the cold records of the touched contacts in registers (same L2 round trip): the propagation is arithmetic only.
// Same worst-first order as the reference (contact.go:240-245 / :396-402): ties to the lowest index in every reduce,
// NaN never wins.  The order in which touched contacts are updated is free: each is updated exactly once per
// iteration with the reference's (b, d) order inside (contact.go:259-279, :416-442).
// ------------------------------------------------------------------------------------------------
struct BigShared {
    real *hot;                  // [cap] staged hot value of the running phase (penetration | desired delta-v)
    unsigned short *adjList;    // [2*cap] contact ids grouped by body — GLOBAL memory (L2): only the warps that prefetch the
                                // touched contacts read it, under warp 0's resolve, so its latency is off the critical path
                                // and shared memory is left to the hot values (26 k contacts instead of 18 k)
    unsigned short *adjOff;     // [B+1] first entry of every body's list
    real *cacheV;               // [NT] best hot value among the owner's contacts (> epsilon) ...
    int *cacheI;                // [NT] ... and its contact (0x7fffffff: none)
    unsigned char *dirty;       // [NT] owner must be re-scanned
};
// shared-memory bytes of the plan for `cap` contacts (cap <= 32767) and B bodies (B < 65535)
// offsetsInGlobal: the list offsets live in global memory too (used when keeping them would push the CTA over the next
// shared-memory carve-out step and take the L1 away from the cold records: 196 -> 228 KB leaves 28 KB of L1)
static inline size_t big_shared_bytes(int NT, long long cap, long long B, bool offsetsInGlobal = false) {
    size_t hotBytes = (size_t)cap * sizeof(real);
    if (hotBytes < (size_t)B * sizeof(int)) hotBytes = (size_t)B * sizeof(int);   // the CSR build borrows the region for its counters
    size_t bytes = (hotBytes + 15) / 16 * 16;
    if (!offsetsInGlobal) bytes += ((size_t)(B + 1) * sizeof(unsigned short) + 15) / 16 * 16;
    bytes += (size_t)NT * (sizeof(real) + sizeof(int) + 1) + 16;
    return (bytes + 15) / 16 * 16;
}
__device__ __forceinline__ BigShared big_carve(unsigned char *base, int NT, int cap, int B, unsigned short *adjListGlobal, unsigned short *adjOffGlobal = nullptr) {
    BigShared s;
    size_t hotBytes = (size_t)cap * sizeof(real);
    if (hotBytes < (size_t)B * sizeof(int)) hotBytes = (size_t)B * sizeof(int);
    size_t off = 0;
    s.hot = (real *)base; off += (hotBytes + 15) / 16 * 16;
    s.adjList = adjListGlobal;
    if (adjOffGlobal) s.adjOff = adjOffGlobal;
    else { s.adjOff = (unsigned short *)(base + off); off += ((size_t)(B + 1) * sizeof(unsigned short) + 15) / 16 * 16; }
    s.cacheV = (real *)(base + off); off += (size_t)NT * sizeof(real);
    s.cacheI = (int *)(base + off); off += (size_t)NT * sizeof(int);
    s.dirty = base + off;
    return s;
}

// CSR adjacency of the contact graph: counting sort of (body, contact) incidences in shared memory.
template <int NT>
__device__ __forceinline__ void big_build_adjacency(const Ctx &x, int nBodies, int tid, const BigShared &sh, int *scanScratch /* shared int[NT/32 + 1] */) {
    int *deg = (int *)sh.hot;    // counters, then fill cursors
    for (int b = tid; b < nBodies; b += NT) deg[b] = 0;
    __syncthreads();
    for (int c = tid; c < x.nC; c += NT) {
        const int b0 = x.cb0[c], b1 = x.cb1[c];
        atomicAdd(&deg[b0], 1);
        if (b1 >= 0 && b1 != b0) atomicAdd(&deg[b1], 1);
    }
    __syncthreads();
    const int per = (nBodies + NT - 1) / NT, lo = min(tid * per, nBodies), hi = min(lo + per, nBodies);
    int sum = 0;
    for (int b = lo; b < hi; b++) sum += deg[b];
    const int lane = tid & 31, warp = tid >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) scanScratch[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w2 = 0; w2 < warp; w2++) base += scanScratch[w2];
    int run = base + incl - sum;
    for (int b = lo; b < hi; b++) {
        const int d = deg[b];
        sh.adjOff[b] = (unsigned short)run;
        deg[b] = run;
        run += d;
    }
    if (hi == nBodies && lo < nBodies) sh.adjOff[nBodies] = (unsigned short)run;
    if (nBodies == 0 && tid == 0) sh.adjOff[0] = 0;
    __syncthreads();
    for (int c = tid; c < x.nC; c += NT) {
        const int b0 = x.cb0[c], b1 = x.cb1[c];
        sh.adjList[atomicAdd(&deg[b0], 1)] = (unsigned short)c;
        if (b1 >= 0 && b1 != b0) sh.adjList[atomicAdd(&deg[b1], 1)] = (unsigned short)c;
    }
    __syncthreads();
}

// a touched contact held in registers while warp 0 resolves the winner
struct Touched {
    int c, b0, b1;
    V3 n, ty, tz, rp0, rp1, cv, la0, la1;
    real restitution;
    bool aw0, aw1;
};

template <bool VELOCITY>
__device__ __forceinline__ void big_prefetch(const Ctx &x, int c, Touched &t) {
    t.c = c;
    t.b0 = x.cb0[c]; t.b1 = x.cb1[c];
    t.n = cw3(x, CW_N, c);
    t.rp0 = cw3(x, CW_RP0, c);
    t.rp1 = cw3(x, CW_RP1, c);
    if (VELOCITY) {
        t.ty = cw3(x, CW_TY, c); t.tz = cw3(x, CW_TZ, c); t.cv = cw3(x, CW_CV, c);
        t.restitution = ctx_restitution(x, c);
        t.aw0 = bw_awake(x, t.b0);
        t.la0 = bw3(x, BW_LACC, t.b0);
        t.aw1 = false; t.la1 = zero3();
        if (t.b1 >= 0) { t.aw1 = bw_awake(x, t.b1); t.la1 = bw3(x, BW_LACC, t.b1); }
    }
}

// contact.go:259-279 on a prefetched contact; returns the new penetration
__device__ __forceinline__ real big_propagate_position(const Touched &t, real pen, const Change &ch) {
    const int cb[2] = {t.b0, t.b1};
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (cb[b] < 0) continue;
        const V3 rp = b == 0 ? t.rp0 : t.rp1;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (cb[b] == ch.b[d]) {
                V3 dp = v_cross(ch.ang[d], rp);
                v_add(dp, ch.lin[d]);
                const real sign = b == 0 ? R_(-1.0) : R_(1.0);
                pen += v_dot(dp, t.n) * sign;
            }
        }
    }
    return pen;
}
// contact.go:416-442 on a prefetched contact (flags already fixed up for a body woken by the winner); returns the new
// desiredDeltaVelocity, `cv` receives the new contactVelocity
__device__ __forceinline__ real big_propagate_velocity(const Touched &t, const Change &ch, real dt, V3 &cv) {
    const int cb[2] = {t.b0, t.b1};
    cv = t.cv;
#pragma unroll
    for (int b = 0; b < 2; b++) {
        if (cb[b] < 0) continue;
        const V3 rp = b == 0 ? t.rp0 : t.rp1;
#pragma unroll
        for (int d = 0; d < 2; d++) {
            if (cb[b] == ch.b[d]) {
                V3 dv = v_cross(ch.ang[d], rp);
                v_add(dv, ch.lin[d]);
                const real sign = b == 1 ? R_(-1.0) : R_(1.0);
                V3 tt = mk3(v_dot(dv, t.n), v_dot(dv, t.ty), v_dot(dv, t.tz));
                v_mul(tt, sign);
                v_add(cv, tt);
            }
        }
    }
    // calculateDesiredDeltaVelocity (contact.go:87-112) from the prefetched flags and last-frame accelerations
    real vfa = R_(0);
    if (t.aw0) { V3 a = t.la0; v_mul(a, dt); vfa += v_dot(a, t.n); }
    if (t.b1 >= 0 && t.aw1) { V3 a = t.la1; v_mul(a, dt); vfa -= v_dot(a, t.n); }
    real rest = t.restitution;
    if (rabs(cv.c[0]) < R_(0.25)) rest = R_(0);
    return -cv.c[0] - rest * (cv.c[0] - vfa);
}

// Owner of a contact.  A resolve touches the contacts of one or two bodies, and those were generated together
// (consecutive indices: the 4 or 8 plane contacts of a cube, the two checks of a pair), so owners are dealt round-robin
// over the WARPS: consecutive contacts belong to owners in different warps and their re-scans run side by side instead of
// one after the other in a single warp (ncu before: a quarter of an iteration was the other warps waiting for that warp).
template <int NT> __device__ __forceinline__ int big_owner(int c) { return (c % (NT / 32)) * 32 + (c / (NT / 32)) % 32; }
template <int NT> __device__ __forceinline__ int big_owner_first(int o) { return (o >> 5) + (o & 31) * (NT / 32); }   // its contacts: first, first + NT, ...

// warp-cooperative re-scan of owner `o`: lanes stride the owner's contacts, arg-max with ties to the lowest index
template <int NT>
__device__ __forceinline__ void big_rescan_owner(const BigShared &sh, int nC, int o, int lane) {
    real v = R_(0.01);    // positionEpsilon / velocityEpsilon (contact.go:12-13)
    int i = 0x7fffffff;
    for (int c = big_owner_first<NT>(o) + lane * NT; c < nC; c += 32 * NT) {
        const real hv = sh.hot[c];
        if (hv > v) { v = hv; i = c; }
    }
    warp_argmax_positive(v, i);
    if (lane == 0) { sh.cacheV[o] = v; sh.cacheI[o] = i; sh.dirty[o] = 0; }
}

struct BigBroadcast {   // warp 0 -> everyone, once per iteration
    real chg[12];
    int chb[2];
    int wake;           // body woken by matchAwakeState, or -1
};

template <int NT, bool VELOCITY>
__device__ __forceinline__ int resolve_loop_big(const Ctx &x0, int maxIterations, GroupScratch *gs, BigBroadcast *bb, int tid, int *status, const BigShared &sh,
                                                real *velPre /* [nC][VP_NF] scratch or NULL */) {
    static_assert(NT >= 64 && (NT & (NT - 1)) == 0, "owner = contact mod NT");
    const int lane = tid & 31, warp = tid >> 5;
    const unsigned full = 0xffffffffu;
    Ctx x = x0;
    real *gHot = VELOCITY ? x0.ddv : x0.pen;
    for (int c = tid; c < x.nC; c += NT) sh.hot[c] = gHot[c];
    if (VELOCITY) x.ddv = sh.hot; else x.pen = sh.hot;
    sh.dirty[tid] = 0;
    if (VELOCITY && velPre) {
        // the 3x3 response matrix of every contact and its inverse are constants of this loop (see friction_response):
        // evaluated here by all threads, once per contact, instead of by warp 0 once per pick (4 matrix products, a
        // determinant, a division — the longest dependent chain of the serial resolve)
        for (int c = tid; c < x.nC; c += NT)
            if (ctx_friction(x, c) != R_(0.0)) precompute_velocity_response(x, c, velPre + (size_t)c * VP_NF);
    }
    __syncthreads();
    {   // every thread is an owner: initial cache
        real v = R_(0.01);
        int i = 0x7fffffff;
        for (int c = big_owner_first<NT>(tid); c < x.nC; c += NT) {
            const real hv = sh.hot[c];
            if (hv > v) { v = hv; i = c; }
        }
        sh.cacheV[tid] = v; sh.cacheI[tid] = i;
    }
    __syncthreads();
    int used = 0;
    while (used < maxIterations) {
        // ---- worst contact: block arg-max of the owners' caches ------------------------------------------------
        real best = sh.cacheV[tid];
        int idx = sh.cacheI[tid];
        warp_argmax_positive(best, idx);
        if (lane == 0) { gs->redv[warp] = best; gs->redi[warp] = idx; }
        __syncthreads();
        best = lane < NT / 32 ? gs->redv[lane] : R_(0.01);
        idx = lane < NT / 32 ? gs->redi[lane] : 0x7fffffff;
        warp_argmax_positive(best, idx);
        if (idx == 0x7fffffff) break;
        // ---- the contacts the winner's change will reach: the adjacency lists of its bodies --------------------
        const int wb0 = x.cb0[idx], wb1 = x.cb1[idx];
        const int s0 = sh.adjOff[wb0], n0 = (int)sh.adjOff[wb0 + 1] - s0;
        const int s1 = wb1 >= 0 && wb1 != wb0 ? (int)sh.adjOff[wb1] : 0, n1 = wb1 >= 0 && wb1 != wb0 ? (int)sh.adjOff[wb1 + 1] - s1 : 0;
        const int nT = n0 + n1;
        // warps 1.. take one touched contact per thread and fetch its record now, under warp 0's resolve
        Touched t;
        const int item = tid - 32;
        bool mine = false;
        if (item >= 0 && item < nT) {
            const int c = item < n0 ? sh.adjList[s0 + item] : sh.adjList[s1 + item - n0];
            big_prefetch<VELOCITY>(x, c, t);
            // a contact between the winner's two bodies is in both lists: it is updated once, from the first
            mine = item < n0 || (t.b0 != wb0 && t.b1 != wb0);
        }
        Change ch;
        if (warp == 0) {
            PosCommit pc;
            VelCommit vc;
            if (VELOCITY) resolve_velocity(x, idx, ch, vc, velPre ? velPre + (size_t)idx * VP_NF : nullptr);
            else resolve_position(x, idx, best, ch, pc);
            __syncwarp();
            if (lane == 0) {
                if (VELOCITY) { commit_velocity(x, vc); if (vc.status) *status = vc.status; }
                else commit_position(x, pc);
#pragma unroll
                for (int k = 0; k < 3; k++) { bb->chg[k] = ch.lin[0].c[k]; bb->chg[3 + k] = ch.lin[1].c[k]; bb->chg[6 + k] = ch.ang[0].c[k]; bb->chg[9 + k] = ch.ang[1].c[k]; }
                bb->chb[0] = ch.b[0]; bb->chb[1] = ch.b[1];
                const int wk = VELOCITY ? vc.wake : pc.wake;
                bb->wake = wk == 0 ? ch.b[0] : (wk == 1 ? ch.b[1] : -1);
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 3; k++) { ch.lin[0].c[k] = bb->chg[k]; ch.lin[1].c[k] = bb->chg[3 + k]; ch.ang[0].c[k] = bb->chg[6 + k]; ch.ang[1].c[k] = bb->chg[9 + k]; }
        ch.b[0] = bb->chb[0]; ch.b[1] = bb->chb[1];
        const int woken = bb->wake;
        // ---- propagation: arithmetic on the prefetched records ------------------------------------------------------
        auto publish = [&](int c, real v) {
            sh.hot[c] = v;
            const int o = big_owner<NT>(c);
            const real cv0 = sh.cacheV[o];
            const int ci = sh.cacheI[o];
            if (c == ci || v > cv0 || (v == cv0 && c < ci)) sh.dirty[o] = 1;
        };
        if (mine) {
            if (VELOCITY) {
                if (woken >= 0) { if (t.b0 == woken) t.aw0 = true; if (t.b1 == woken) t.aw1 = true; }
                V3 cv;
                const real ddv = big_propagate_velocity(t, ch, x.dt, cv);
                cw3_set(x, CW_CV, t.c, cv);
                publish(t.c, ddv);
            } else {
                publish(t.c, big_propagate_position(t, sh.hot[t.c], ch));
            }
        }
        // more touched contacts than prefetching threads (a body with hundreds of contacts): the rest in place
        for (int it2 = NT - 32 + tid; it2 < nT; it2 += NT) {
            const int c = it2 < n0 ? sh.adjList[s0 + it2] : sh.adjList[s1 + it2 - n0];
            if (it2 >= n0 && (x.cb0[c] == wb0 || x.cb1[c] == wb0)) continue;
            if (VELOCITY) propagate_velocity(x, c, ch);
            else propagate_position(x, c, ch);
            publish(c, sh.hot[c]);
        }
        __syncthreads();
        // ---- owners whose cached best was touched or overtaken are re-scanned by their warp ----------------------
        unsigned dm = __ballot_sync(full, sh.dirty[tid] != 0);
        while (dm) {
            const int l2 = __ffs(dm) - 1;
            dm &= dm - 1;
            big_rescan_owner<NT>(sh, x.nC, warp * 32 + l2, lane);
        }
        __syncwarp();
        used++;
        // the next iteration's first barrier (inside the reduce) orders this iteration's shared-memory writes of other
        // warps before warp 0 reads the winner's hot value
    }
    __syncthreads();
    for (int c = tid; c < x.nC; c += NT) gHot[c] = sh.hot[c];
    __syncthreads();
    return used;
}

}  // namespace czr
