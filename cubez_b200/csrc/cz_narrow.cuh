// cz_narrow.cuh — narrowphase device functions (K3): sphere/cube/plane primitive tests.
//
// Each routine restates the arithmetic and the comparison strictness of the reference
// routine it replaces (colliders.go, cited per function) and returns contacts by value;
// ordering/compaction into the reference's append order is the caller's job.
#pragma once
#include "cz_math.cuh"

namespace czn {
using namespace czm;

struct ColliderView {   // colliders.go:39-71
    int shape;          // CZ_SHAPE_*
    int body;           // world-local body index
    M34 t;              // collider transform (body.transform x Offset)
    V3 half;            // cube half sizes
    real radius;        // sphere radius
};
struct PlaneView { V3 n; real offset; };   // colliders.go:29-35

struct GenContact {     // public part of contact.go:17-36 as generated
    int b0, b1;         // world-local body indices, -1 = nil
    V3 point, normal;
    real pen;
};

// dynamic column select without local-memory indexing
CZD V3 axis_dyn(const M34 &m, int i) {
    return mk3(i == 0 ? m.c[0] : (i == 1 ? m.c[3] : (i == 2 ? m.c[6] : m.c[9])),
               i == 0 ? m.c[1] : (i == 1 ? m.c[4] : (i == 2 ? m.c[7] : m.c[10])),
               i == 0 ? m.c[2] : (i == 1 ? m.c[5] : (i == 2 ? m.c[8] : m.c[11])));
}
CZD real sel3(const V3 &v, int i) { return i == 0 ? v.c[0] : (i == 1 ? v.c[1] : v.c[2]); }

// colliders.go:762-770
CZD real transform_to_axis(const M34 &t, const V3 &half, const V3 &axis) {
    return half.c[0] * rabs(v_dot(axis, m34_axis(t, 0))) + half.c[1] * rabs(v_dot(axis, m34_axis(t, 1))) +
           half.c[2] * rabs(v_dot(axis, m34_axis(t, 2)));
}

// colliders.go:180-207.  Bodies (sphere, nil).
CZD bool sphere_halfspace(const ColliderView &s, const PlaneView &p, GenContact &c) {
    V3 pos = m34_axis(s.t, 3);
    real distance = v_dot(p.n, pos) - s.radius;
    if (!(distance <= p.offset)) return false;
    c.point = p.n;
    v_mul(c.point, distance + s.radius * R_(-1.0));
    v_add(c.point, pos);
    c.normal = p.n;
    c.pen = -distance;
    c.b0 = s.body; c.b1 = -1;
    return true;
}

// colliders.go:216-254.  Bodies (this, other).
CZD bool sphere_sphere(const ColliderView &s, const ColliderView &o, GenContact &c) {
    V3 p1 = m34_axis(s.t, 3), p2 = m34_axis(o.t, 3);
    V3 mid = p1;
    v_sub(mid, p2);
    real size = v_mag(mid);
    if (size <= R_(0.0) || size >= s.radius + o.radius) return false;
    c.point = mid;
    v_mul(c.point, R_(0.5));
    v_add(c.point, p1);
    c.normal = mid;
    v_mul(c.normal, rdiv(R_(1.0), size));
    c.pen = s.radius + o.radius - size;
    c.b0 = s.body; c.b1 = o.body;
    return true;
}

// colliders.go:308-366 + :750-760.  Returns the bit mask of the vertices (in the fixed sign
// order of :319-327) that are in contact; 0 = no contact.
CZD V3 cube_vertex(const ColliderView &cube, int k) {
    V3 v = mk3((k & 1) ? R_(-1.0) : R_(1.0), (k & 2) ? R_(-1.0) : R_(1.0), (k & 4) ? R_(-1.0) : R_(1.0));
    v_component_product(v, cube.half);
    return m34_mul_v(cube.t, v);
}
CZD unsigned cube_halfspace_mask(const ColliderView &cube, const PlaneView &p) {
    real pr = transform_to_axis(cube.t, cube.half, p.n);
    real cd = v_dot(p.n, m34_axis(cube.t, 3)) - pr;
    if (!(cd <= p.offset)) return 0u;
    unsigned mask = 0;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
        V3 vp = cube_vertex(cube, k);
        real vd = v_dot(vp, p.n);
        if (vd <= p.offset) mask |= 1u << k;
    }
    return mask;
}
// contact for vertex k of a cube-plane check (k must be set in the mask)
CZD void cube_halfspace_contact(const ColliderView &cube, const PlaneView &p, int k, GenContact &c) {
    V3 vp = cube_vertex(cube, k);
    real vd = v_dot(vp, p.n);
    c.point = p.n;
    v_mul(c.point, vd - p.offset);
    v_add(c.point, vp);
    c.normal = p.n;
    c.pen = p.offset - vd;
    c.b0 = cube.body; c.b1 = -1;
}

// colliders.go:369-441.  Bodies (cube, sphere) whatever the call direction (:210-213).
// sphereVel = sphere.Body.Velocity, the fallback normal of :417-421.
CZD bool cube_sphere(const ColliderView &cube, const ColliderView &sphere, const V3 &sphereVel, GenContact &c) {
    V3 position = m34_axis(sphere.t, 3);
    V3 rel = m34_transform_inverse(cube.t, position);
    if (rabs(rel.c[0]) - sphere.radius > cube.half.c[0] || rabs(rel.c[1]) - sphere.radius > cube.half.c[1] ||
        rabs(rel.c[2]) - sphere.radius > cube.half.c[2])
        return false;
    V3 closest;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        real dist = rel.c[i];
        if (dist > cube.half.c[i]) dist = cube.half.c[i];
        else if (dist < -cube.half.c[i]) dist = -cube.half.c[i];
        closest.c[i] = dist;
    }
    V3 dc = closest;
    v_sub(dc, rel);
    real dist = v_sqmag(dc);
    if (dist > sphere.radius * sphere.radius) return false;
    V3 cw = m34_mul_v(cube.t, closest);
    c.point = cw;
    c.normal = cw;
    v_sub(c.normal, position);
    if (real_equal(v_mag(c.normal), R_(0.0))) c.normal = sphereVel;
    v_normalize(c.normal);
    c.pen = sphere.radius;
    if (!real_equal(dist, R_(0.0))) c.pen -= rsqrt_(dist);
    else c.pen = R_(0.0);
    c.b0 = cube.body; c.b1 = sphere.body;
    return true;
}

// colliders.go:458-477 (+ :445-456).  Returns false on a separating axis.
CZD bool try_axis(const ColliderView &one, const ColliderView &two, V3 axis, const V3 &toCenter, int index, real &smallest, int &smallestCase) {
    if (v_sqmag(axis) < CZ_EPSILON) return true;
    v_normalize(axis);
    real p1 = transform_to_axis(one.t, one.half, axis), p2 = transform_to_axis(two.t, two.half, axis);
    real distance = rabs(v_dot(toCenter, axis));
    real pen = p1 + p2 - distance;
    if (pen < R_(0)) return false;
    if (pen < smallest) { smallest = pen; smallestCase = index; }
    return true;
}

// colliders.go:481-517
CZD void fill_point_face(const ColliderView &one, const ColliderView &two, const V3 &toCenter, int best, real pen, GenContact &c) {
    V3 normal = axis_dyn(one.t, best);
    if (v_dot(normal, toCenter) > R_(0)) v_mul(normal, R_(-1.0));
    V3 v = two.half;
    if (v_dot(m34_axis(two.t, 0), normal) < R_(0)) v.c[0] = -v.c[0];
    if (v_dot(m34_axis(two.t, 1), normal) < R_(0)) v.c[1] = -v.c[1];
    if (v_dot(m34_axis(two.t, 2), normal) < R_(0)) v.c[2] = -v.c[2];
    c.normal = normal;
    c.pen = pen;
    c.point = m34_mul_v(two.t, v);
    c.b0 = one.body; c.b1 = two.body;
}

// colliders.go:519-573
CZD V3 contact_point(const V3 &pOne, const V3 &dOne, real oneSize, const V3 &pTwo, const V3 &dTwo, real twoSize, bool useOne) {
    real smOne = v_sqmag(dOne), smTwo = v_sqmag(dTwo);
    real dpOneTwo = v_dot(dTwo, dOne);
    V3 toSt = pOne;
    v_sub(toSt, pTwo);
    real dpStaOne = v_dot(dOne, toSt), dpStaTwo = v_dot(dTwo, toSt);
    real denom = smOne * smTwo - dpOneTwo * dpOneTwo;
    if (rabs(denom) < CZ_EPSILON) return useOne ? pOne : pTwo;
    real mua = rdiv(dpOneTwo * dpStaTwo - smTwo * dpStaOne, denom);
    real mub = rdiv(smOne * dpStaTwo - dpOneTwo * dpStaOne, denom);
    if (mua > oneSize || mua < -oneSize || mub > twoSize || mub < -twoSize) return useOne ? pOne : pTwo;
    V3 cOne = dOne; v_mul(cOne, mua); v_add(cOne, pOne);
    V3 cTwo = dTwo; v_mul(cTwo, mub); v_add(cTwo, pTwo);
    v_mul(cOne, R_(0.5)); v_mul(cTwo, R_(0.5));
    v_add(cOne, cTwo);
    return cOne;
}

// colliders.go:576-710.  15-axis SAT; one contact.
CZD bool cube_cube(const ColliderView &one, const ColliderView &two, GenContact &c) {
    V3 toCenter = m34_axis(two.t, 3);
    v_sub(toCenter, m34_axis(one.t, 3));
    real pen = CZ_REAL_MAX;
    int best = 0xffffff;
    int bestSingleAxis = best;
    // The 15 candidate axes in the reference's order (:588-626): one's 3 face axes, two's 3, then
    // one.axis(i) x two.axis(j) for i, j in 0..2.  Kept as a rolled loop: the fully unrolled form
    // is ~3 000 SASS instructions and the fused kernel is instruction-cache bound.
#pragma unroll 1
    for (int idx = 0; idx < 15; idx++) {
        if (idx == 6) bestSingleAxis = best;   // :602
        V3 axis;
        if (idx < 3) axis = axis_dyn(one.t, idx);
        else if (idx < 6) axis = axis_dyn(two.t, idx - 3);
        else axis = v_cross(axis_dyn(one.t, (idx - 6) / 3), axis_dyn(two.t, (idx - 6) % 3));
        if (!try_axis(one, two, axis, toCenter, idx, pen, best)) return false;
    }
    if (best < 3) {
        fill_point_face(one, two, toCenter, best, pen, c);
        return true;
    } else if (best < 6) {
        V3 nc = toCenter;
        v_mul(nc, R_(-1.0));
        fill_point_face(two, one, nc, best - 3, pen, c);
        return true;
    }
    best -= 6;
    int oneIdx = best / 3, twoIdx = best % 3;
    V3 oneAxis = axis_dyn(one.t, oneIdx), twoAxis = axis_dyn(two.t, twoIdx);
    V3 axis = v_cross(oneAxis, twoAxis);
    v_normalize(axis);
    if (v_dot(axis, toCenter) > R_(0)) v_mul(axis, R_(-1.0));
    V3 ptOne = one.half, ptTwo = two.half;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        if (i == oneIdx) ptOne.c[i] = R_(0);
        else if (v_dot(m34_axis(one.t, i), axis) > R_(0)) ptOne.c[i] = -ptOne.c[i];
        if (i == twoIdx) ptTwo.c[i] = R_(0);
        else if (v_dot(m34_axis(two.t, i), axis) < R_(0)) ptTwo.c[i] = -ptTwo.c[i];
    }
    ptOne = m34_mul_v(one.t, ptOne);
    ptTwo = m34_mul_v(two.t, ptTwo);
    bool useOne = bestSingleAxis > 2;
    c.point = contact_point(ptOne, oneAxis, sel3(one.half, oneIdx), ptTwo, twoAxis, sel3(two.half, twoIdx), useOne);
    c.normal = axis;
    c.pen = pen;
    c.b0 = one.body; c.b1 = two.body;
    return true;
}

// Conservative bounding-sphere rejection (new; not in the reference, which runs the full test
// for every pair).  A pair is skipped only when the centre distance exceeds the sum of the
// bounding radii by a 0.5 % margin — then every routine below returns "no contact" as well
// (sphere-sphere :227, cube-sphere :397, SAT :468), so the contact set is unchanged.
// (R1+R2)^2 <= 2(R1^2+R2^2) avoids the square roots.
CZD real bounding_r2(const ColliderView &c) {
    return c.shape == CZ_SHAPE_SPHERE ? c.radius * c.radius : v_dot(c.half, c.half);
}
CZD bool bounding_reject(const ColliderView &one, const ColliderView &two) {
    V3 d = m34_axis(two.t, 3);
    v_sub(d, m34_axis(one.t, 3));
    return v_sqmag(d) > R_(2.02) * (bounding_r2(one) + bounding_r2(two)) + R_(1e-9);
}

// CheckForCollisions for two body colliders (colliders.go:720-747 with the forwarding of
// :210-213).  velOne/velTwo are the bodies' velocities (only used by cube-sphere).
CZD bool check_pair(const ColliderView &one, const ColliderView &two, const V3 &velOne, const V3 &velTwo, GenContact &c) {
    if (bounding_reject(one, two)) return false;
    if (two.shape == CZ_SHAPE_SPHERE) {
        if (one.shape == CZ_SHAPE_SPHERE) return sphere_sphere(one, two, c);
        if (one.shape == CZ_SHAPE_CUBE) return cube_sphere(one, two, velTwo, c);
    } else if (two.shape == CZ_SHAPE_CUBE) {
        if (one.shape == CZ_SHAPE_SPHERE) return cube_sphere(two, one, velOne, c);
        if (one.shape == CZ_SHAPE_CUBE) return cube_cube(one, two, c);
    }
    return false;
}

}  // namespace czn
