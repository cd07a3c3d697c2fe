// oracle_capi.cpp — C ABI over the CPU restatement (cubez_oracle.hpp) for ctypes.
//
// TEST INFRASTRUCTURE ONLY (see the header of cubez_oracle.hpp).  Mirrors the product ABI of
// include/cubezcuda.h with a czo_ prefix so the parity tests can drive both with the same
// numpy arrays.  Parity status: pinned against the mechanically translated reference (see cubez_oracle.hpp).
//
// Build: g++ -O2 -ffp-contract=off -fPIC -shared [-DCUBEZ_REAL_FLOAT] oracle_capi.cpp
#include "../include/cubezcuda.h"
#include "cubez_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <atomic>
#include <thread>

using namespace czo;
typedef cz_real R;

namespace {

void load_body(Body<R> &b, const cz_bodies *s, int i) {
    body_init(b);
    auto get = [](const R *p, int idx, R dflt) { return p ? p[idx] : dflt; };
    for (int k = 0; k < 3; k++) {
        if (s->position) b.position[k] = s->position[i * 3 + k];
        if (s->velocity) b.velocity[k] = s->velocity[i * 3 + k];
        if (s->rotation) b.rotation[k] = s->rotation[i * 3 + k];
        if (s->acceleration) b.acceleration[k] = s->acceleration[i * 3 + k];
        if (s->last_frame_acceleration) b.lastFrameAcc[k] = s->last_frame_acceleration[i * 3 + k];
    }
    if (s->orientation) for (int k = 0; k < 4; k++) b.orientation[k] = s->orientation[i * 4 + k];
    b.linearDamping = get(s->linear_damping, i, b.linearDamping);
    b.angularDamping = get(s->angular_damping, i, b.angularDamping);
    if (s->inverse_inertia_tensor) for (int k = 0; k < 9; k++) b.inverseInertiaTensor[k] = s->inverse_inertia_tensor[i * 9 + k];
    b.inverseMass = get(s->inverse_mass, i, (R)0);
    b.mass = b.inverseMass != 0 ? (R)1 / b.inverseMass : (R)0;
    b.motion = get(s->motion, i, b.motion);
    if (s->is_awake) b.isAwake = s->is_awake[i] != 0;
    if (s->can_sleep) b.canSleep = s->can_sleep[i] != 0;
    if (s->transform) for (int k = 0; k < 12; k++) b.transform[k] = s->transform[i * 12 + k];
    if (s->inverse_inertia_tensor_world) for (int k = 0; k < 9; k++) b.iitWorld[k] = s->inverse_inertia_tensor_world[i * 9 + k];
}

void store_body(const Body<R> &b, cz_bodies *s, int i) {
    for (int k = 0; k < 3; k++) {
        if (s->position) s->position[i * 3 + k] = b.position[k];
        if (s->velocity) s->velocity[i * 3 + k] = b.velocity[k];
        if (s->rotation) s->rotation[i * 3 + k] = b.rotation[k];
        if (s->acceleration) s->acceleration[i * 3 + k] = b.acceleration[k];
        if (s->last_frame_acceleration) s->last_frame_acceleration[i * 3 + k] = b.lastFrameAcc[k];
    }
    if (s->orientation) for (int k = 0; k < 4; k++) s->orientation[i * 4 + k] = b.orientation[k];
    if (s->linear_damping) s->linear_damping[i] = b.linearDamping;
    if (s->angular_damping) s->angular_damping[i] = b.angularDamping;
    if (s->inverse_inertia_tensor) for (int k = 0; k < 9; k++) s->inverse_inertia_tensor[i * 9 + k] = b.inverseInertiaTensor[k];
    if (s->inverse_mass) s->inverse_mass[i] = b.inverseMass;
    if (s->motion) s->motion[i] = b.motion;
    if (s->is_awake) s->is_awake[i] = b.isAwake ? 1 : 0;
    if (s->can_sleep) s->can_sleep[i] = b.canSleep ? 1 : 0;
    if (s->transform) for (int k = 0; k < 12; k++) s->transform[i * 12 + k] = b.transform[k];
    if (s->inverse_inertia_tensor_world) for (int k = 0; k < 9; k++) s->inverse_inertia_tensor_world[i * 9 + k] = b.iitWorld[k];
}

void load_collider(Collider<R> &c, const cz_colliders *s, int i, Body<R> *body) {
    std::memset(&c, 0, sizeof(c));
    c.shape = s->shape ? s->shape[i] : SHAPE_NONE;
    c.body = body;
    m34_identity(c.offset);
    m34_identity(c.transform);
    if (s->offset) for (int k = 0; k < 12; k++) c.offset[k] = s->offset[i * 12 + k];
    if (s->transform) for (int k = 0; k < 12; k++) c.transform[k] = s->transform[i * 12 + k];
    if (s->half_size) for (int k = 0; k < 3; k++) c.halfSize[k] = s->half_size[i * 3 + k];
    if (s->radius) c.radius = s->radius[i];
}

struct OracleWorlds {
    cz_world_desc desc;
    std::vector<World<R>> worlds;
    std::vector<Plane<R>> planes;
};

}  // namespace

extern "C" {

int czo_real_size(void) { return (int)sizeof(R); }

int czo_math_op(int32_t op, const R *in, R *out) {
    auto v3 = [&](int o) { return V3<R>{{in[o], in[o + 1], in[o + 2]}}; };
    auto q4 = [&](int o) { return Q4<R>{{in[o], in[o + 1], in[o + 2], in[o + 3]}}; };
    auto m3 = [&](int o) { M3<R> m; for (int k = 0; k < 9; k++) m[k] = in[o + k]; return m; };
    auto m34 = [&](int o) { M34<R> m; for (int k = 0; k < 12; k++) m[k] = in[o + k]; return m; };
    auto put3 = [&](const V3<R> &v) { out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; };
    switch (op) {
    case CZ_OP_VEC_ADD: { V3<R> a = v3(0); v_add(a, v3(3)); put3(a); return 0; }
    case CZ_OP_VEC_ADD_SCALED: { V3<R> a = v3(0); v_add_scaled(a, v3(3), in[6]); put3(a); return 0; }
    case CZ_OP_VEC_COMPONENT_PRODUCT: { V3<R> a = v3(0); v_component_product(a, v3(3)); put3(a); return 0; }
    case CZ_OP_VEC_CROSS: put3(v_cross(v3(0), v3(3))); return 0;
    case CZ_OP_VEC_DOT: out[0] = v_dot(v3(0), v3(3)); return 0;
    case CZ_OP_VEC_MAGNITUDE: out[0] = v_mag(v3(0)); return 0;
    case CZ_OP_VEC_SQUARE_MAGNITUDE: out[0] = v_sqmag(v3(0)); return 0;
    case CZ_OP_VEC_MUL_WITH: { V3<R> a = v3(0); v_mul(a, in[3]); put3(a); return 0; }
    case CZ_OP_VEC_NORMALIZE: { V3<R> a = v3(0); v_normalize(a); put3(a); return 0; }
    case CZ_OP_VEC_SUB: { V3<R> a = v3(0); v_sub(a, v3(3)); put3(a); return 0; }
    case CZ_OP_QUAT_MUL: { Q4<R> a = q4(0); q_mul(a, q4(4)); for (int k = 0; k < 4; k++) out[k] = a[k]; return 0; }
    case CZ_OP_QUAT_LEN: out[0] = q_len(q4(0)); return 0;
    case CZ_OP_QUAT_NORMALIZE: { Q4<R> a = q4(0); q_normalize(a); for (int k = 0; k < 4; k++) out[k] = a[k]; return 0; }
    case CZ_OP_QUAT_ROTATE: put3(q_rotate(q4(0), v3(4))); return 0;
    case CZ_OP_QUAT_ADD_SCALED_VECTOR: { Q4<R> a = q4(0); q_add_scaled_vector(a, v3(4), in[7]); for (int k = 0; k < 4; k++) out[k] = a[k]; return 0; }
    case CZ_OP_M3_MUL_M3: { M3<R> r = m3_mul_m(m3(0), m3(9)); for (int k = 0; k < 9; k++) out[k] = r[k]; return 0; }
    case CZ_OP_M3_INVERT: { M3<R> r = m3_invert(m3(0)); for (int k = 0; k < 9; k++) out[k] = r[k]; return 0; }
    case CZ_OP_M3_MUL_V: put3(m3_mul_v(m3(0), v3(9))); return 0;
    case CZ_OP_M3_TRANSFORM_TRANSPOSE: put3(m3_transform_transpose(m3(0), v3(9))); return 0;
    case CZ_OP_M3_DETERMINANT: out[0] = m3_det(m3(0)); return 0;
    case CZ_OP_M34_MUL_M34: { M34<R> r = m34_mul_m34(m34(0), m34(12)); for (int k = 0; k < 12; k++) out[k] = r[k]; return 0; }
    case CZ_OP_M34_MUL_V: put3(m34_mul_v(m34(0), v3(12))); return 0;
    case CZ_OP_M34_TRANSFORM_INVERSE: put3(m34_transform_inverse(m34(0), v3(12))); return 0;
    case CZ_OP_M34_SET_AS_TRANSFORM: { M34<R> r; m34_set_as_transform(r, v3(0), q4(3)); for (int k = 0; k < 12; k++) out[k] = r[k]; return 0; }
    case CZ_OP_REAL_EQUAL: out[0] = real_equal<R>(in[0], in[1]) ? 1 : 0; return 0;
    case CZ_OP_TRANSFORM_INERTIA: { M3<R> w; transform_inertia_tensor(w, m3(0), m34(9)); for (int k = 0; k < 9; k++) out[k] = w[k]; return 0; }
    }
    return CZ_ERR_INVALID;
}

// ---- object API -------------------------------------------------------------------------
int czo_integrate(cz_bodies *io, R dt, const R *lin_pow, const R *ang_pow, const R *bias) {
    for (int i = 0; i < io->n; i++) {
        Body<R> b;
        load_body(b, io, i);
        R lp = lin_pow ? lin_pow[i] : pow_factor<R>(b.linearDamping, dt);
        R ap = ang_pow ? ang_pow[i] : pow_factor<R>(b.angularDamping, dt);
        R bs = bias ? *bias : pow_factor<R>((R)0.5, dt);
        body_integrate_pows<R>(b, dt, lp, ap, bs);
        store_body(b, io, i);
    }
    return 0;
}

int czo_calculate_derived_data(cz_bodies *io) {
    for (int i = 0; i < io->n; i++) {
        Body<R> b;
        load_body(b, io, i);
        body_calculate_derived(b);
        store_body(b, io, i);
    }
    return 0;
}

int czo_collider_derive(int32_t n, const R *body_transform, const R *offset, R *out) {
    for (int i = 0; i < n; i++) {
        M34<R> t, o;
        for (int k = 0; k < 12; k++) { t[k] = body_transform[i * 12 + k]; o[k] = offset[i * 12 + k]; }
        M34<R> r = m34_mul_m34(t, o);
        for (int k = 0; k < 12; k++) out[i * 12 + k] = r[k];
    }
    return 0;
}

static int export_contacts(const Contacts<R> &cs, const Body<R> *base, cz_contacts *out, const std::vector<int32_t> *checkOf) {
    if ((int)cs.size() > out->capacity) { out->n = (int32_t)cs.size(); return CZ_ERR_CAPACITY; }
    out->n = (int32_t)cs.size();
    for (size_t i = 0; i < cs.size(); i++) {
        const Contact<R> *c = cs[i];
        if (out->body0) out->body0[i] = c->bodies[0] ? (int32_t)(c->bodies[0] - base) : -1;
        if (out->body1) out->body1[i] = c->bodies[1] ? (int32_t)(c->bodies[1] - base) : -1;
        if (out->friction) out->friction[i] = c->friction;
        if (out->restitution) out->restitution[i] = c->restitution;
        for (int k = 0; k < 3; k++) {
            if (out->point) out->point[i * 3 + k] = c->contactPoint[k];
            if (out->normal) out->normal[i * 3 + k] = c->contactNormal[k];
        }
        if (out->penetration) out->penetration[i] = c->penetration;
        if (out->check && checkOf) out->check[i] = (*checkOf)[i];
    }
    return 0;
}

int czo_narrowphase(const cz_colliders *colliders, const cz_planes *planes, const cz_bodies *bodies, int32_t n_checks,
                    const int32_t *one, const int32_t *two, cz_contacts *out, uint8_t *found) {
    std::vector<Body<R>> bs(bodies ? bodies->n : 0);
    for (size_t i = 0; i < bs.size(); i++) load_body(bs[i], bodies, (int)i);
    std::vector<Collider<R>> cs(colliders->n);
    for (int i = 0; i < colliders->n; i++) {
        int bi = colliders->body ? colliders->body[i] : i;
        load_collider(cs[i], colliders, i, (bi >= 0 && bi < (int)bs.size()) ? &bs[bi] : nullptr);
    }
    std::vector<Plane<R>> ps(planes ? planes->n : 0);
    for (size_t p = 0; p < ps.size(); p++) {
        for (int k = 0; k < 3; k++) ps[p].normal[k] = planes->normal[p * 3 + k];
        ps[p].offset = planes->offset[p];
    }
    Contacts<R> contacts;
    std::vector<int32_t> checkOf;
    for (int k = 0; k < n_checks; k++) {
        int a = one[k], b = two[k];
        const Collider<R> *ca = a >= 0 ? &cs[a] : nullptr, *cb = b >= 0 ? &cs[b] : nullptr;
        const Plane<R> *pa = a < 0 ? &ps[-a - 1] : nullptr, *pb = b < 0 ? &ps[-b - 1] : nullptr;
        size_t before = contacts.size();
        bool f = false;
        if ((ca == nullptr || ca->shape != SHAPE_NONE) && (cb == nullptr || cb->shape != SHAPE_NONE))
            f = check_for_collisions<R>(ca, pa, cb, pb, contacts);
        if (found) found[k] = f ? 1 : 0;
        for (size_t i = before; i < contacts.size(); i++) checkOf.push_back(k);
    }
    int rc = export_contacts(contacts, bs.data(), out, &checkOf);
    for (Contact<R> *c : contacts) delete c;
    return rc;
}

int czo_resolve_contacts(int32_t max_iterations, cz_contacts *io, cz_bodies *bodies_io, R dt, int32_t *iterations_used) {
    std::vector<Body<R>> bs(bodies_io->n);
    for (int i = 0; i < bodies_io->n; i++) load_body(bs[i], bodies_io, i);
    Contacts<R> contacts;
    for (int i = 0; i < io->n; i++) {
        Contact<R> *c = new_contact<R>();
        c->bodies[0] = io->body0[i] >= 0 ? &bs[io->body0[i]] : nullptr;
        c->bodies[1] = io->body1[i] >= 0 ? &bs[io->body1[i]] : nullptr;
        c->friction = io->friction ? io->friction[i] : (R)0.9;
        c->restitution = io->restitution ? io->restitution[i] : (R)0.1;
        for (int k = 0; k < 3; k++) { c->contactPoint[k] = io->point[i * 3 + k]; c->contactNormal[k] = io->normal[i * 3 + k]; }
        c->penetration = io->penetration[i];
        contacts.push_back(c);
    }
    int iters[2] = {0, 0}, status = 0;
    resolve_contacts<R>(max_iterations, contacts, dt, iters, &status);
    if (iterations_used) { iterations_used[0] = iters[0]; iterations_used[1] = iters[1]; }
    export_contacts(contacts, bs.data(), io, nullptr);
    for (Contact<R> *c : contacts) delete c;
    for (int i = 0; i < bodies_io->n; i++) store_body(bs[i], bodies_io, i);
    return status ? CZ_ERR_NIL_BODY : 0;
}

// ---- world API --------------------------------------------------------------------------
void *czo_world_create(const cz_world_desc *d) {
    OracleWorlds *w = new OracleWorlds;
    w->desc = *d;
    w->worlds.resize(d->n_worlds);
    for (auto &wd : w->worlds) {
        wd.bodies.resize(d->bodies_per_world);
        wd.colliders.resize(d->bodies_per_world);
        wd.activeFrom.assign(d->bodies_per_world, 0);
        wd.integrateFlag.assign(d->bodies_per_world, 1);
        wd.schedule = d->schedule;
        for (int i = 0; i < d->bodies_per_world; i++) {
            body_init(wd.bodies[i]);
            std::memset(&wd.colliders[i], 0, sizeof(Collider<R>));
            m34_identity(wd.colliders[i].offset);
            m34_identity(wd.colliders[i].transform);
        }
    }
    return w;
}
int czo_world_destroy(void *h) { delete (OracleWorlds *)h; return 0; }

int czo_world_upload_bodies(void *h, int32_t first, int32_t n, const cz_bodies *b, int32_t derive) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < B; i++) {
            Body<R> &body = w->worlds[first + k].bodies[i];
            const V3<R> fAcc = body.forceAccum, tAcc = body.torqueAccum;   // a state upload does not touch pending accumulators
            load_body(body, b, k * B + i);                                  // (they are not part of cz_bodies; cz_world_add_forces owns them)
            body.forceAccum = fAcc; body.torqueAccum = tAcc;
            if (derive) body_calculate_derived(body);
        }
    return 0;
}
int czo_world_upload_colliders(void *h, int32_t first, int32_t n, const cz_colliders *c, int32_t derive) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < B; i++) {
            World<R> &wd = w->worlds[first + k];
            load_collider(wd.colliders[i], c, k * B + i, &wd.bodies[i]);
            if (derive && wd.colliders[i].shape != SHAPE_NONE) collider_derive(wd.colliders[i]);
        }
    return 0;
}
int czo_world_upload_planes(void *h, const cz_planes *p) {
    OracleWorlds *w = (OracleWorlds *)h;
    w->planes.resize(p->n);
    for (int i = 0; i < p->n; i++) {
        for (int k = 0; k < 3; k++) w->planes[i].normal[k] = p->normal[i * 3 + k];
        w->planes[i].offset = p->offset[i];
    }
    for (auto &wd : w->worlds) wd.planes = w->planes;
    return 0;
}
int czo_world_upload_schedule(void *h, int32_t n_checks, const int32_t *one, const int32_t *two) {
    OracleWorlds *w = (OracleWorlds *)h;
    for (auto &wd : w->worlds) { wd.checkOne.assign(one, one + n_checks); wd.checkTwo.assign(two, two + n_checks); }
    return 0;
}
int czo_world_set_activation(void *h, int32_t first, int32_t n, const int32_t *active_from, const uint8_t *integrate) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < B; i++) {
            if (active_from) w->worlds[first + k].activeFrom[i] = active_from[k * B + i];
            if (integrate) w->worlds[first + k].integrateFlag[i] = integrate[k * B + i];
        }
    return 0;
}
int czo_world_set_episodes(void *h, int32_t length, const int32_t *phase0) {
    OracleWorlds *w = (OracleWorlds *)h;
    for (size_t i = 0; i < w->worlds.size(); i++) {
        World<R> &wd = w->worlds[i];
        wd.episodeLength = length;
        wd.episodePhase0 = phase0 ? phase0[i] : 0;
        wd.episodeStep0 = wd.stepIndex;
        wd.episodeBodies = wd.bodies;
        wd.episodeColliders = wd.colliders;
    }
    return 0;
}
// per-pair surface materials (cz_world_set_materials); n_materials == 0 restores the constants
int czo_world_set_materials(void *h, int32_t n_materials, const R *friction, const R *restitution, int32_t first, int32_t n,
                            const int32_t *body_material, const int32_t *plane_material) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (auto &wd : w->worlds) {
        if (wd.nMaterials != n_materials) { wd.bodyMaterial.assign(B, 0); wd.planeMaterial.assign(16, 0); }
        wd.nMaterials = n_materials;
        if (n_materials > 0) {
            wd.matFriction.assign(friction, friction + (size_t)n_materials * n_materials);
            wd.matRestitution.assign(restitution, restitution + (size_t)n_materials * n_materials);
            if (plane_material) for (size_t p = 0; p < w->planes.size(); p++) wd.planeMaterial[p] = plane_material[p];
        }
    }
    if (n_materials > 0 && body_material)
        for (int k = 0; k < n; k++)
            for (int i = 0; i < B; i++) w->worlds[first + k].bodyMaterial[i] = body_material[k * B + i];
    return 0;
}
// forceAccum += force, torqueAccum += torque (the writer of rigidbody.go:86-92's accumulators; mirrors cz_world_add_forces)
int czo_world_add_forces(void *h, int32_t first, int32_t n, const R *force, const R *torque) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int b = 0; b < B; b++) {
            Body<R> &body = w->worlds[first + k].bodies[b];
            const long long i = ((long long)k * B + b) * 3;
            for (int c = 0; c < 3; c++) {
                if (force) body.forceAccum[c] = body.forceAccum[c] + force[i + c];
                if (torque) body.torqueAccum[c] = body.torqueAccum[c] + torque[i + c];
            }
        }
    return 0;
}
int czo_world_set_step_index(void *h, int64_t s) { for (auto &wd : ((OracleWorlds *)h)->worlds) wd.stepIndex = s; return 0; }

// n_threads > 1 partitions the worlds over std::threads (worlds are independent).
int czo_world_step(void *h, R dt, int32_t n_steps, int32_t n_threads, cz_step_stats *stats) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int W = (int)w->worlds.size();
    auto t0 = std::chrono::steady_clock::now();
    // worlds are handed out one at a time from a shared counter (they differ in cost: a static split left threads idle)
    std::atomic<int> next{0};
    auto run = [&]() {
        for (int i = next.fetch_add(1); i < W; i = next.fetch_add(1)) {
            w->worlds[i].totalContacts = w->worlds[i].totalPosIters = w->worlds[i].totalVelIters = 0;
            for (int s = 0; s < n_steps; s++) w->worlds[i].step(dt);
        }
    };
    if (n_threads <= 1 || W == 1) run();
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < std::min<int>(n_threads, W); t++) th.emplace_back(run);
        for (auto &t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->steps = (int64_t)W * n_steps;
        for (auto &wd : w->worlds) {
            stats->contacts += wd.totalContacts; stats->pos_iterations += wd.totalPosIters; stats->vel_iterations += wd.totalVelIters;
            stats->max_contacts = std::max<int32_t>(stats->max_contacts, (int32_t)wd.lastContacts.size());
            if (wd.status) stats->status = CZ_ERR_NIL_BODY;
        }
        stats->device_ms = (float)std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    return 0;
}
int czo_world_download_bodies(void *h, int32_t first, int32_t n, cz_bodies *out) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < B; i++) store_body(w->worlds[first + k].bodies[i], out, k * B + i);
    return 0;
}
int czo_world_download_colliders(void *h, int32_t first, int32_t n, cz_colliders *out) {
    OracleWorlds *w = (OracleWorlds *)h;
    const int B = w->desc.bodies_per_world;
    for (int k = 0; k < n; k++)
        for (int i = 0; i < B; i++) {
            const Collider<R> &c = w->worlds[first + k].colliders[i];
            if (out->transform) for (int j = 0; j < 12; j++) out->transform[(k * B + i) * 12 + j] = c.transform[j];
        }
    return 0;
}
int czo_world_download_contacts(void *h, int32_t world, cz_contacts *out) {
    OracleWorlds *w = (OracleWorlds *)h;
    const auto &cs = w->worlds[world].lastContacts;
    out->n = (int32_t)cs.size();
    if ((int)cs.size() > out->capacity) return CZ_ERR_CAPACITY;
    for (size_t i = 0; i < cs.size(); i++) {
        if (out->body0) out->body0[i] = cs[i].body[0];
        if (out->body1) out->body1[i] = cs[i].body[1];
        if (out->friction) out->friction[i] = cs[i].friction;
        if (out->restitution) out->restitution[i] = cs[i].restitution;
        for (int k = 0; k < 3; k++) {
            if (out->point) out->point[i * 3 + k] = cs[i].point[k];
            if (out->normal) out->normal[i * 3 + k] = cs[i].normal[k];
        }
        if (out->penetration) out->penetration[i] = cs[i].penetration;
    }
    return 0;
}
int czo_world_last_step_counts(void *h, int32_t *n_contacts, int32_t *pos_it, int32_t *vel_it) {
    OracleWorlds *w = (OracleWorlds *)h;
    for (size_t i = 0; i < w->worlds.size(); i++) {
        if (n_contacts) n_contacts[i] = (int32_t)w->worlds[i].lastContacts.size();
        if (pos_it) pos_it[i] = w->worlds[i].posIters;
        if (vel_it) vel_it[i] = w->worlds[i].velIters;
    }
    return 0;
}
// energy: sum over finite-mass bodies of 1/2 m v.v + 1/2 w^T (R I_body R^T) w - m (a . p)   
int czo_world_checksum_energy(void *h, uint64_t *checksum, double *energy) {
    OracleWorlds *w = (OracleWorlds *)h;
    uint64_t sum = 0;
    double e = 0;
    for (auto &wd : w->worlds) {
        sum += world_checksum(wd);
        for (size_t i = 0; i < wd.bodies.size(); i++) {
            const Body<R> &b = wd.bodies[i];
            if (!(b.inverseMass > 0)) continue;
            double m = 1.0 / (double)b.inverseMass;
            double ke = 0.5 * m * (double)v_dot(b.velocity, b.velocity);
            // rotational: w^T I_world w with I_world = (iitWorld)^-1
            M3<R> iw = m3_invert(b.iitWorld);
            V3<R> Iw = m3_mul_v(iw, b.rotation);
            double kr = 0.5 * (double)v_dot(b.rotation, Iw);
            double pe = -m * (double)v_dot(b.acceleration, b.position);
            e += ke + kr + pe;
        }
    }
    if (checksum) *checksum = sum;
    if (energy) *energy = e;
    return 0;
}

// Integrate+CalculateDerivedData microbench over AoS bodies (cfg5 CPU baseline).
int czo_bench_integrate(cz_bodies *io, R dt, int32_t steps, int32_t n_threads, double *seconds) {
    std::vector<Body<R>> bs(io->n);
    for (int i = 0; i < io->n; i++) load_body(bs[i], io, i);
    auto t0 = std::chrono::steady_clock::now();
    auto run = [&](int lo, int hi) {
        for (int s = 0; s < steps; s++)
            for (int i = lo; i < hi; i++) body_integrate<R>(bs[i], dt);
    };
    if (n_threads <= 1) run(0, io->n);
    else {
        std::vector<std::thread> th;
        int per = (io->n + n_threads - 1) / n_threads;
        for (int t = 0; t < n_threads; t++) { int lo = t * per, hi = std::min<int>(io->n, lo + per); if (lo < hi) th.emplace_back(run, lo, hi); }
        for (auto &t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    for (int i = 0; i < io->n; i++) store_body(bs[i], io, i);
    return 0;
}

}  // extern "C"
