// hostemu.cu — TEST INFRASTRUCTURE ONLY.
// Runs the product's device functions (cz_math/cz_body/cz_narrow/cz_resolve .cuh — the very
// code the CUDA kernels inline) sequentially on the CPU, so that their arithmetic and control
// flow can be compared bit-for-bit with the oracle in a container that has no GPU.  It is not
// part of libcubezcuda, is never loaded by cubez_b200, and is not a fallback: only
// tests/test_hostemu_parity.py uses it.  The warp/CTA orchestration (shuffles, scans,
// barriers) is NOT covered here — that is what the -m gpu tests are for.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../cubez_b200/csrc/cz_kernels.cuh"

using namespace czk;
using namespace czr;

namespace {
struct HostStore {
    czb::BodyStore st;
    std::vector<real2> chunks;
    std::vector<uint8_t> awake, can_sleep, integ, shape, ident;
    std::vector<int32_t> active;
    void alloc(long long n) {
        long long stride = (n + 63) / 64 * 64;
        chunks.assign((size_t)stride * czb::N_CHUNKS, make_real2(0, 0));
        awake.assign(stride, 1); can_sleep.assign(stride, 1); integ.assign(stride, 1); shape.assign(stride, 0); ident.assign(stride, 1);
        active.assign(stride, 0);
        st.base = chunks.data(); st.stride = stride; st.n = n;
        st.awake = awake.data(); st.can_sleep = can_sleep.data(); st.integ = integ.data(); st.shape = shape.data(); st.ident = ident.data();
        st.active_from = active.data();
        st.force = st.torque = nullptr;
    }
    real &slot(int s, long long i) { return ((real *)(st.base + (long long)(s >> 1) * st.stride))[2 * i + (s & 1)]; }
};
void put(HostStore &h, int first, int comps, const real *src, long long n) {
    if (!src) return;
    for (long long i = 0; i < n; i++) for (int c = 0; c < comps; c++) h.slot(first + c, i) = src[i * comps + c];
}
void get(HostStore &h, int first, int comps, real *dst, long long n) {
    if (!dst) return;
    for (long long i = 0; i < n; i++) for (int c = 0; c < comps; c++) dst[i * comps + c] = h.slot(first + c, i);
}

template <bool VELOCITY> int host_loop(const Ctx &x, int maxIter, int *status) {
    const real *hot = VELOCITY ? x.ddv : x.pen;
    int used = 0;
    while (used < maxIter) {
        real best = R_(0.01);
        int idx = 0x7fffffff;
        for (int c = 0; c < x.nC; c++) {
            real v = hot[c];
            if (v > best) { best = v; idx = c; }
        }
        if (idx == 0x7fffffff) break;
        Change ch;
        if (VELOCITY) {
            VelCommit vc;
            resolve_velocity(x, idx, ch, vc);
            commit_velocity(x, vc);
            if (vc.status) *status = vc.status;
        } else {
            PosCommit pc;
            resolve_position(x, idx, best, ch, pc);
            commit_position(x, pc);
        }
        for (int c = 0; c < x.nC; c++) {
            if (VELOCITY) propagate_velocity(x, c, ch);
            else propagate_position(x, c, ch);
        }
        used++;
    }
    return used;
}
}  // namespace

extern "C" {

// One world.  Steps n_steps frames; per step writes the contact count, the two iteration
// counts and an FNV hash of the (body0, body1) sequence.  State is returned in `io`.
// materials for the next cze_run calls (n == 0: the constants); copies are kept
static std::vector<real> g_matFric, g_matRest;
static std::vector<uint8_t> g_bodyMat, g_planeMat;
int cze_set_materials(int n, const cz_real *fric, const cz_real *rest, int n_bodies, const int32_t *body_mat, int n_planes, const int32_t *plane_mat) {
    g_matFric.clear(); g_matRest.clear(); g_bodyMat.clear(); g_planeMat.assign(CZ_MAX_PLANES, 0);
    if (n <= 0) return 0;
    g_matFric.assign(fric, fric + n * n); g_matRest.assign(rest, rest + n * n);
    g_bodyMat.assign(n_bodies, 0);
    if (body_mat) for (int i = 0; i < n_bodies; i++) g_bodyMat[i] = (uint8_t)body_mat[i];
    if (plane_mat) for (int i = 0; i < n_planes && i < CZ_MAX_PLANES; i++) g_planeMat[i] = (uint8_t)plane_mat[i];
    return 0;
}

int cze_run(cz_bodies *io, const cz_colliders *col, const cz_planes *planes, int schedule, int n_checks, const int32_t *one,
            const int32_t *two, const int32_t *active_from, const uint8_t *integ, cz_real dt, int n_steps, int contact_cap,
            int32_t *out_counts, int32_t *out_pos, int32_t *out_vel, uint64_t *out_pairhash, cz_contacts *last_contacts) {
    const int B = io->n;
    HostStore h;
    h.alloc(B);
    put(h, czb::C_P01 * 2, 3, io->position, B); put(h, czb::C_P2M * 2 + 1, 1, io->motion, B);
    put(h, czb::C_Q01 * 2, 4, io->orientation, B); put(h, czb::C_V01 * 2, 3, io->velocity, B);
    put(h, czb::C_V2R0 * 2 + 1, 3, io->rotation, B); put(h, czb::C_A01 * 2, 3, io->acceleration, B);
    put(h, czb::C_APW0 * 2 + 1, 9, io->inverse_inertia_tensor, B); put(h, czb::C_MD * 2, 1, io->inverse_mass, B);
    put(h, czb::C_L01 * 2, 3, io->last_frame_acceleration, B);
    put(h, czb::C_L2T0 * 2 + 1, 12, io->transform, B); put(h, czb::C_T11W0 * 2 + 1, 9, io->inverse_inertia_tensor_world, B);
    put(h, czb::C_H01 * 2, 3, col->half_size, B); put(h, czb::C_H2R * 2 + 1, 1, col->radius, B);
    put(h, czb::C_O01 * 2, 12, col->offset, B); put(h, czb::C_X01 * 2, 12, col->transform, B);
    for (int i = 0; i < B; i++) {
        h.awake[i] = io->is_awake ? io->is_awake[i] : 1;
        h.can_sleep[i] = io->can_sleep ? io->can_sleep[i] : 1;
        h.shape[i] = (uint8_t)col->shape[i];
        h.integ[i] = integ ? integ[i] : 1;
        h.active[i] = active_from ? active_from[i] : 0;
        bool id = true;
        for (int k = 0; k < 12; k++) id = id && col->offset[i * 12 + k] == ((k == 0 || k == 4 || k == 8) ? (real)1 : (real)0);
        h.ident[i] = id;
        h.slot(czb::C_A2LP * 2 + 1, i) = (real)pow((double)io->linear_damping[i], (double)dt);
        h.slot(czb::C_APW0 * 2, i) = (real)pow((double)io->angular_damping[i], (double)dt);
    }
    const real bias = (real)pow(0.5, (double)dt);
    WorldParams p;
    memset(&p, 0, sizeof(p));
    p.st = h.st; p.W = 1; p.B = B; p.P = planes ? planes->n : 0; p.Cc = contact_cap;
    p.schedule = schedule; p.chk_one = one; p.chk_two = two;
    p.nchk = schedule == CZ_SCHED_ALL_PAIRS_ORDERED ? B * (p.P + B) : n_checks;
    for (int i = 0; i < p.P; i++) { p.planes[i].n = mk3(planes->normal[i * 3], planes->normal[i * 3 + 1], planes->normal[i * 3 + 2]); p.planes[i].offset = planes->offset[i]; }
    if (!g_matFric.empty() && (int)g_bodyMat.size() == B) {
        p.matFric = g_matFric.data(); p.matRest = g_matRest.data(); p.bodyMat = g_bodyMat.data();
        p.nMat = 1; while (p.nMat * p.nMat < (int)g_matFric.size()) p.nMat++;
        for (int i = 0; i < CZ_MAX_PLANES; i++) p.planeMat[i] = g_planeMat[i];
    }
    std::vector<real> gen((size_t)G_NF * contact_cap), bw((size_t)BW_NF * B), cw((size_t)CW_NREAL * contact_cap);
    std::vector<int> gb0(contact_cap), gb1(contact_cap), cb(2 * (size_t)contact_cap);
    int lastC = 0;
    for (int s = 0; s < n_steps; s++) {
        p.step_index = s;
        // K1 (same statements as k_integrate<true>)
        for (long long i = 0; i < B; i++) {
            if (!(h.integ[i] != 0 && s >= h.active[i])) continue;
            M34 tr;
            bool haveTr = false;
            if (h.awake[i]) {
                V3 pos = czb::ld_position(h.st, i), vel = czb::ld_velocity(h.st, i), rot = czb::ld_rotation(h.st, i);
                real2 a01 = h.st.ld(czb::C_A01, i), a2lp = h.st.ld(czb::C_A2LP, i), apw0 = h.st.ld(czb::C_APW0, i);
                V3 acc = mk3(a01.x, a01.y, a2lp.x);
                Q4 q = czb::ld_orientation(h.st, i);
                M3 ib = czb::ld_iit_body(h.st, i);
                czb::Integrated o;
                czb::integrate_body(o, pos, q, vel, rot, acc, ib, h.st.ld(czb::C_P2M, i).y, h.can_sleep[i] != 0, dt, a2lp.y, apw0.x, bias);
                h.st.st(czb::C_P01, i, make_real2(o.pos.c[0], o.pos.c[1])); h.st.st(czb::C_P2M, i, make_real2(o.pos.c[2], o.motion));
                h.st.st(czb::C_Q01, i, make_real2(o.q.c[0], o.q.c[1])); h.st.st(czb::C_Q23, i, make_real2(o.q.c[2], o.q.c[3]));
                h.st.st(czb::C_V01, i, make_real2(o.vel.c[0], o.vel.c[1])); h.st.st(czb::C_V2R0, i, make_real2(o.vel.c[2], o.rot.c[0]));
                h.st.st(czb::C_R12, i, make_real2(o.rot.c[1], o.rot.c[2]));
                h.st.st(czb::C_L01, i, make_real2(o.lastAcc.c[0], o.lastAcc.c[1]));
                czb::st_derived(h.st, i, o.lastAcc.c[2], o.transform, o.iitWorld);
                if (!o.awake) h.awake[i] = 0;
                tr = o.transform; haveTr = true;
            }
            if (h.shape[i] != CZ_SHAPE_NONE) {
                if (!haveTr) tr = czb::ld_transform(h.st, i);
                M34 off = h.ident[i] ? czb::identity34() : czb::ld_m34(h.st, czb::C_O01, i);
                czb::st_m34(h.st, czb::C_X01, i, m34_mul_m34(tr, off));
            }
        }
        // K3
        int nC = 0;
        for (int k = 0; k < p.nchk; k++) {
            int a, b;
            if (!decode_check(p, k, a, b)) continue;
            CheckEval e;
            eval_check(p, 0, a, b, e);
            real fric, rest;
            check_material(p, 0, a, b, fric, rest);
            if (e.kind == 1) { if (nC < contact_cap) store_gen(gen.data(), contact_cap, gb0.data(), gb1.data(), nC, e.gc, fric, rest); nC++; }
            else if (e.kind == 2) {
                ColliderView c = load_collider(p.st, e.cubeLocal, e.cubeLocal);
                for (int v = 0; v < 8; v++) if (e.mask & (1u << v)) {
                    GenContact gc;
                    czn::cube_halfspace_contact(c, p.planes[e.plane], v, gc);
                    if (nC < contact_cap) store_gen(gen.data(), contact_cap, gb0.data(), gb1.data(), nC, gc, fric, rest);
                    nC++;
                }
            }
        }
        if (nC > contact_cap) return CZ_ERR_CAPACITY;
        uint64_t hsh = 0xcbf29ce484222325ull;
        for (int c = 0; c < nC; c++) { hsh ^= (uint32_t)gb0[c]; hsh *= 0x100000001b3ull; hsh ^= (uint32_t)gb1[c]; hsh *= 0x100000001b3ull; }
        out_counts[s] = nC; out_pairhash[s] = hsh; out_pos[s] = out_vel[s] = 0;
        lastC = nC;
        if (last_contacts && s == n_steps - 1) {
            last_contacts->n = nC;
            for (int c = 0; c < nC && c < last_contacts->capacity; c++) {
                last_contacts->body0[c] = gb0[c]; last_contacts->body1[c] = gb1[c];
                for (int k = 0; k < 3; k++) { last_contacts->point[c * 3 + k] = gen[(G_POINT + k) * contact_cap + c]; last_contacts->normal[c * 3 + k] = gen[(G_NORMAL + k) * contact_cap + c]; }
                last_contacts->penetration[c] = gen[G_PEN * contact_cap + c];
                if (last_contacts->friction) last_contacts->friction[c] = gen[G_FRIC * contact_cap + c];
                if (last_contacts->restitution) last_contacts->restitution[c] = gen[G_REST * contact_cap + c];
            }
        }
        // K4
        if (nC > 0) {
            Ctx x;
            x.bw = bw.data(); x.bs = B; x.cb0 = cb.data(); x.cb1 = cb.data() + contact_cap;
            // cold fields laid out AoS here (the fused kernel's layout); k_resolve uses SoA
            x.cold = cw.data(); x.cfs = 1; x.ccs = CW_NCOLD;
            x.pen = cw.data() + (size_t)CW_NCOLD * contact_cap; x.ddv = x.pen + contact_cap; x.fric = x.ddv + contact_cap; x.rest = x.fric + contact_cap;
            x.nC = nC; x.dt = dt; x.xb = nullptr; x.xbs = 0; x.mlist = nullptr; x.bmask = nullptr; x.store = h.st; x.body_base = 0;
            GenView g;
            g.pn = gen.data(); g.fs = contact_cap; g.cs = 1; g.pen = gen.data() + (size_t)G_PEN * contact_cap;
            g.fric = gen.data() + (size_t)G_FRIC * contact_cap; g.rest = gen.data() + (size_t)G_REST * contact_cap;
            g.b0 = gb0.data(); g.b1 = gb1.data();
            for (int b = 0; b < B; b++) load_body_work(x, h.st, b, b);
            for (int c = 0; c < nC; c++) prepare_contact(x, c, g);
            int status = 0;
            out_pos[s] = host_loop<false>(x, nC * 8, &status);
            out_vel[s] = host_loop<true>(x, nC * 8, &status);
            for (int b = 0; b < B; b++) store_body_work(x, h.st, b, b);
            if (status) return status;
        }
    }
    (void)lastC;
    get(h, czb::C_P01 * 2, 3, io->position, B); get(h, czb::C_P2M * 2 + 1, 1, io->motion, B);
    get(h, czb::C_Q01 * 2, 4, io->orientation, B); get(h, czb::C_V01 * 2, 3, io->velocity, B);
    get(h, czb::C_V2R0 * 2 + 1, 3, io->rotation, B); get(h, czb::C_L01 * 2, 3, io->last_frame_acceleration, B);
    get(h, czb::C_L2T0 * 2 + 1, 12, io->transform, B); get(h, czb::C_T11W0 * 2 + 1, 9, io->inverse_inertia_tensor_world, B);
    if (io->is_awake) for (int i = 0; i < B; i++) io->is_awake[i] = h.awake[i];
    return 0;
}

}  // extern "C"
