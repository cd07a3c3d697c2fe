// libcubezcuda.cu — host runtime + C ABI (include/cubezcuda.h) of the B200 rigid-body step.
//
// One translation unit: the kernels are in cz_kernels.cuh / cz_fused.cuh.  Build (see
// __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true
//        -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off -shared
//        [-DCUBEZ_REAL_FLOAT] libcubezcuda.cu -o libcubezcuda[_f32].so
// -fmad=false is part of the correctness contract: Go on amd64 never contracts
// a*b+c, and contact existence is a chain of float comparisons.
//
// No CPU fallback exists in this file: every compute entry point launches CUDA kernels and
// returns CZ_ERR_CUDA when the runtime reports an error (e.g. no device).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <cuda_runtime.h>

// ------------------------------------------------------------------------------------------
// Small-block device memory cache.  The object-API shims (cz_integrate, cz_narrowphase,
// cz_resolve_contacts: one call per RigidBody.Integrate / CheckForCollisions / ResolveContacts of a
// drop-in caller) build a scratch batch or a one-world handle per call: ~30 cudaMalloc/cudaFree pairs
// of a few KB each, which dominated the call.  Blocks of up to 1 MiB are kept per (device, size class)
// after their release and handed out again; larger blocks go straight to the runtime.  A release keeps
// cudaFree's contract (the device is idle when the block becomes reusable).
// ------------------------------------------------------------------------------------------
namespace czp {
static std::mutex g_mu;
static std::unordered_map<void *, std::pair<int, size_t>> g_live;          // block -> (device, size class)
static std::map<std::pair<int, size_t>, std::vector<void *>> g_free;
static size_t g_cached = 0;
constexpr size_t kSmall = 1u << 20, kMaxCached = 256u << 20;
static inline cudaError_t pool_malloc(void **p, size_t bytes) {
    if (bytes > kSmall) return ::cudaMalloc(p, bytes);
    size_t cls = 256;
    while (cls < bytes) cls <<= 1;
    int dev = 0;
    ::cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_free.find({dev, cls});
        if (it != g_free.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_cached -= cls;
            g_live[*p] = {dev, cls};
            return cudaSuccess;
        }
    }
    cudaError_t e = ::cudaMalloc(p, cls);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(g_mu); g_live[*p] = {dev, cls}; }
    return e;
}
static inline cudaError_t pool_free(void *p) {
    if (!p) return cudaSuccess;
    std::pair<int, size_t> key{0, 0};
    bool cache = false;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_live.find(p);
        if (it != g_live.end()) {
            key = it->second;
            cache = g_cached + key.second <= kMaxCached;
            g_live.erase(it);
            if (cache) g_cached += key.second;   // reserved now, published below
        }
    }
    if (!cache) return ::cudaFree(p);
    // as cudaFree: nothing in flight on the block's OWN device may still use it.  The wait happens outside the lock.
    int cur = 0;
    ::cudaGetDevice(&cur);
    if (cur != key.first) ::cudaSetDevice(key.first);
    ::cudaDeviceSynchronize();
    if (cur != key.first) ::cudaSetDevice(cur);
    std::lock_guard<std::mutex> lk(g_mu);
    g_free[key].push_back(p);
    return cudaSuccess;
}
// release every cached block of one device (cz_shutdown)
static inline void pool_trim(int device) {
    std::vector<void *> blocks;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto &kv : g_free) {
            if (kv.first.first != device) continue;
            for (void *q : kv.second) { blocks.push_back(q); g_cached -= kv.first.second; }
            kv.second.clear();
        }
    }
    for (void *q : blocks) ::cudaFree(q);
}
template <class T> static inline cudaError_t pool_malloc_t(T **p, size_t bytes) { return pool_malloc((void **)p, bytes); }
// page-locked host blocks of up to 64 KiB (the per-world status mirror): cudaHostAlloc / cudaFreeHost cost ~0.5 ms each
static std::unordered_map<void *, size_t> g_hostLive;
static std::map<size_t, std::vector<void *>> g_hostFree;
static inline cudaError_t pool_host_alloc(void **p, size_t bytes, unsigned flags) {
    if (bytes > (64u << 10) || flags != cudaHostAllocDefault) return ::cudaHostAlloc(p, bytes, flags);
    size_t cls = 256;
    while (cls < bytes) cls <<= 1;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_hostFree.find(cls);
        if (it != g_hostFree.end() && !it->second.empty()) {
            *p = it->second.back();
            it->second.pop_back();
            g_hostLive[*p] = cls;
            return cudaSuccess;
        }
    }
    cudaError_t e = ::cudaHostAlloc(p, cls, flags);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lk(g_mu); g_hostLive[*p] = cls; }
    return e;
}
static inline cudaError_t pool_host_free(void *p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_hostLive.find(p);
        if (it != g_hostLive.end()) {
            const size_t cls = it->second;
            g_hostLive.erase(it);
            if (g_hostFree[cls].size() < 64) { ::cudaDeviceSynchronize(); g_hostFree[cls].push_back(p); return cudaSuccess; }
        }
    }
    return ::cudaFreeHost(p);
}
template <class T> static inline cudaError_t pool_host_alloc_t(T **p, size_t bytes, unsigned flags) { return pool_host_alloc((void **)p, bytes, flags); }
}  // namespace czp
#define cudaMalloc(p, n) czp::pool_malloc_t((p), (size_t)(n))
#define cudaFree(p) czp::pool_free((void *)(p))
#define cudaHostAlloc(p, n, f) czp::pool_host_alloc_t((p), (size_t)(n), (f))
#define cudaFreeHost(p) czp::pool_host_free((void *)(p))

#include "cz_kernels.cuh"
#include "cz_fused.cuh"
#include "cz_broadphase.cuh"

// The host pipelines move a chunk as one copy per field (the caller's arrays are separate allocations).  Submitted one
// by one, every copy costs ~5 us of copy-engine time on top of its bytes: 6 chunks x 9 fields x 2 directions = 108
// copies = 0.4-1.0 ms of a 4.6 ms frame (tools/probes/memcpy_batch_probe.cu: 71.7 -> 80.6 GB/s in both directions at
// once).  cudaMemcpyBatchAsync (CUDA 12.8) submits a chunk's fields as ONE operation; it needs page-locked or
// CUDA-allocated operands, so a batch with a pageable pointer in it, a driver without the entry point, or
// CUBEZ_COPY_BATCH=0 goes out as plain cudaMemcpyAsync calls — same bytes, same stream order.
struct CopyBatch {
    static constexpr int MAXN = 16;
    void *dst[MAXN], *src[MAXN];
    size_t size[MAXN];
    int n = 0;
    template <class T> void add(T *d, const T *s, long long first, long long count, int comps) {
        if (!d || !s || count <= 0 || n >= MAXN) return;
        dst[n] = (void *)(d + first * comps); src[n] = (void *)(s + first * comps); size[n] = sizeof(T) * (size_t)count * comps;
        n++;
    }
};
static std::atomic<int> g_copyBatchState{-1};   // -1 not probed, 0 off, 1 on
static bool host_ptr_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged || a.type == cudaMemoryTypeDevice;
}
static cudaError_t copy_batch_submit(CopyBatch &b, cudaMemcpyKind kind, cudaStream_t stream, bool pinned) {
    if (b.n == 0) return cudaSuccess;
    int st = g_copyBatchState.load();
    if (st < 0) { st = czf::env_int("CUBEZ_COPY_BATCH", 1) ? 1 : 0; g_copyBatchState.store(st); }
    if (st == 1 && pinned && b.n > 1) {
        cudaMemcpyAttributes at{};
        at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
        size_t idx0 = 0, failIdx = 0;
        const cudaError_t e = cudaMemcpyBatchAsync(b.dst, b.src, b.size, (size_t)b.n, &at, &idx0, 1, &failIdx, stream);
        if (e == cudaSuccess) { b.n = 0; return e; }
        if (e != cudaErrorNotSupported && e != cudaErrorCallRequiresNewerDriver && e != cudaErrorInvalidValue && e != cudaErrorSymbolNotFound) return e;
        cudaGetLastError();
        g_copyBatchState.store(0);   // this driver cannot: plain copies from now on
    }
    for (int i = 0; i < b.n; i++) {
        const cudaError_t e = cudaMemcpyAsync(b.dst[i], b.src[i], b.size[i], kind, stream);
        if (e != cudaSuccess) return e;
    }
    b.n = 0;
    return cudaSuccess;
}


using namespace czk;

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

struct cz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    size_t smem_optin = 0;
    std::string err;
};

static int fail(cz_ctx *ctx, int code, const std::string &msg) {
    g_err = msg;
    if (ctx) ctx->err = msg;
    return code;
}
#define CK(ctx, call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(ctx, CZ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)
#define CKL(ctx) CK(ctx, cudaGetLastError())

static inline unsigned nblk(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// field -> (first real slot, components) in the chunked SoA (slot = chunk*2 + lane)
enum Field { F_POS, F_MOTION, F_ORI, F_VEL, F_ROT, F_ACC, F_LINPOW, F_ANGPOW, F_IITB, F_LACC, F_TRANSFORM, F_IITW,
             F_HALF, F_RADIUS, F_OFFSET, F_CTRANSFORM, F_INVM, F_LIND, F_ANGD, F_COUNT };
struct FieldSlot { int first, comps; };
static const FieldSlot kField[F_COUNT] = {
    {czb::C_P01 * 2, 3}, {czb::C_P2M * 2 + 1, 1}, {czb::C_Q01 * 2, 4}, {czb::C_V01 * 2, 3}, {czb::C_V2R0 * 2 + 1, 3},
    {czb::C_A01 * 2, 3}, {czb::C_A2LP * 2 + 1, 1}, {czb::C_APW0 * 2, 1}, {czb::C_APW0 * 2 + 1, 9}, {czb::C_L01 * 2, 3},
    {czb::C_L2T0 * 2 + 1, 12}, {czb::C_T11W0 * 2 + 1, 9}, {czb::C_H01 * 2, 3}, {czb::C_H2R * 2 + 1, 1},
    {czb::C_O01 * 2, 12}, {czb::C_X01 * 2, 12}, {czb::C_MD * 2, 1}, {czb::C_MD * 2 + 1, 1}, {czb::C_AD * 2, 1}};

// A device-resident batch of bodies + staging, shared by the world handle and the shims.
struct Batch {
    cz_ctx *ctx = nullptr;
    czb::BodyStore st{};
    long long n = 0;
    real *stage = nullptr;       // device staging for pack/unpack, n*12 reals
    std::vector<real> h_lind, h_angd;   // host shadow of the damping (for the Pow refresh)
    real pow_dt = (real)NAN;
    bool pow_user = false;
};

static int batch_alloc(cz_ctx *ctx, Batch &b, long long n) {
    b.ctx = ctx;
    b.n = n;
    long long stride = (n + 63) / 64 * 64;
    if (stride == 0) stride = 64;
    b.st.n = n;
    b.st.stride = stride;
    CK(ctx, cudaMalloc(&b.st.base, sizeof(real2) * stride * czb::N_CHUNKS));
    CK(ctx, cudaMemsetAsync(b.st.base, 0, sizeof(real2) * stride * czb::N_CHUNKS, ctx->stream));
    CK(ctx, cudaMalloc(&b.st.awake, stride));
    CK(ctx, cudaMalloc(&b.st.can_sleep, stride));
    CK(ctx, cudaMalloc(&b.st.integ, stride));
    CK(ctx, cudaMalloc(&b.st.shape, stride));
    CK(ctx, cudaMalloc(&b.st.ident, stride));
    CK(ctx, cudaMalloc(&b.st.active_from, sizeof(int32_t) * stride));
    CK(ctx, cudaMemsetAsync(b.st.awake, 1, stride, ctx->stream));
    CK(ctx, cudaMemsetAsync(b.st.can_sleep, 1, stride, ctx->stream));
    CK(ctx, cudaMemsetAsync(b.st.integ, 1, stride, ctx->stream));
    CK(ctx, cudaMemsetAsync(b.st.shape, 0, stride, ctx->stream));
    CK(ctx, cudaMemsetAsync(b.st.ident, 1, stride, ctx->stream));
    CK(ctx, cudaMemsetAsync(b.st.active_from, 0, sizeof(int32_t) * stride, ctx->stream));
    CK(ctx, cudaMalloc(&b.stage, sizeof(real) * stride * 12));
    b.h_lind.assign(n, (real)0.95);
    b.h_angd.assign(n, (real)0.95);
    return CZ_OK;
}
static void batch_free(Batch &b) {
    if (b.st.base) cudaFree(b.st.base);
    if (b.st.awake) cudaFree(b.st.awake);
    if (b.st.can_sleep) cudaFree(b.st.can_sleep);
    if (b.st.integ) cudaFree(b.st.integ);
    if (b.st.shape) cudaFree(b.st.shape);
    if (b.st.ident) cudaFree(b.st.ident);
    if (b.st.active_from) cudaFree(b.st.active_from);
    if (b.st.force) cudaFree(b.st.force);
    if (b.st.torque) cudaFree(b.st.torque);
    if (b.stage) cudaFree(b.stage);
    b = Batch();
}

static int put_field(Batch &b, Field f, long long first, long long n, const real *host) {
    if (!host || n == 0) return CZ_OK;
    cz_ctx *ctx = b.ctx;
    const FieldSlot fs = kField[f];
    CK(ctx, cudaMemcpyAsync(b.stage, host, sizeof(real) * n * fs.comps, cudaMemcpyHostToDevice, ctx->stream));
    k_pack<<<nblk(n * fs.comps, 256), 256, 0, ctx->stream>>>(b.st.base, b.st.stride, first, n, b.stage, fs.first, fs.comps);
    CKL(ctx);
    // the staging buffer is reused by the next field: order the copy after the kernel
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CZ_OK;
}
static int get_field(Batch &b, Field f, long long first, long long n, real *host) {
    if (!host || n == 0) return CZ_OK;
    cz_ctx *ctx = b.ctx;
    const FieldSlot fs = kField[f];
    k_unpack<<<nblk(n * fs.comps, 256), 256, 0, ctx->stream>>>(b.st.base, b.st.stride, first, n, b.stage, fs.first, fs.comps);
    CKL(ctx);
    CK(ctx, cudaMemcpyAsync(host, b.stage, sizeof(real) * n * fs.comps, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CZ_OK;
}
static int put_u8(Batch &b, uint8_t *dev, long long first, long long n, const uint8_t *host) {
    if (!host || n == 0) return CZ_OK;
    CK(b.ctx, cudaMemcpyAsync(dev + first, host, n, cudaMemcpyHostToDevice, b.ctx->stream));
    CK(b.ctx, cudaStreamSynchronize(b.ctx->stream));
    return CZ_OK;
}
static int get_u8(Batch &b, const uint8_t *dev, long long first, long long n, uint8_t *host) {
    if (!host || n == 0) return CZ_OK;
    CK(b.ctx, cudaMemcpyAsync(host, dev + first, n, cudaMemcpyDeviceToHost, b.ctx->stream));
    CK(b.ctx, cudaStreamSynchronize(b.ctx->stream));
    return CZ_OK;
}

// Evaluate Pow(damping, dt) in float64 on the host and round to Real (rigidbody.go:233-234),
// then upload.  This is the library default; cz_world_set_pow lets the host language supply
// its own bit pattern instead.
static int refresh_pow(Batch &b, long long first, long long n, real dt) {
    std::vector<real> lp(n), ap(n);
    for (long long i = 0; i < n; i++) {
        lp[i] = (real)std::pow((double)b.h_lind[first + i], (double)dt);
        ap[i] = (real)std::pow((double)b.h_angd[first + i], (double)dt);
    }
    int rc = put_field(b, F_LINPOW, first, n, lp.data());
    if (rc) return rc;
    return put_field(b, F_ANGPOW, first, n, ap.data());
}

static int upload_bodies(Batch &b, long long first, long long n, const cz_bodies *s) {
    int rc;
    if ((rc = put_field(b, F_POS, first, n, s->position))) return rc;
    if ((rc = put_field(b, F_ORI, first, n, s->orientation))) return rc;
    if ((rc = put_field(b, F_VEL, first, n, s->velocity))) return rc;
    if ((rc = put_field(b, F_ROT, first, n, s->rotation))) return rc;
    if ((rc = put_field(b, F_ACC, first, n, s->acceleration))) return rc;
    if ((rc = put_field(b, F_LIND, first, n, s->linear_damping))) return rc;
    if ((rc = put_field(b, F_ANGD, first, n, s->angular_damping))) return rc;
    if ((rc = put_field(b, F_IITB, first, n, s->inverse_inertia_tensor))) return rc;
    if ((rc = put_field(b, F_INVM, first, n, s->inverse_mass))) return rc;
    if ((rc = put_field(b, F_MOTION, first, n, s->motion))) return rc;
    if ((rc = put_field(b, F_TRANSFORM, first, n, s->transform))) return rc;
    if ((rc = put_field(b, F_IITW, first, n, s->inverse_inertia_tensor_world))) return rc;
    if ((rc = put_field(b, F_LACC, first, n, s->last_frame_acceleration))) return rc;
    if ((rc = put_u8(b, b.st.awake, first, n, s->is_awake))) return rc;
    if ((rc = put_u8(b, b.st.can_sleep, first, n, s->can_sleep))) return rc;
    bool dampingChanged = false;
    if (s->linear_damping && std::memcmp(b.h_lind.data() + first, s->linear_damping, sizeof(real) * n) != 0) {
        std::memcpy(b.h_lind.data() + first, s->linear_damping, sizeof(real) * n);
        dampingChanged = true;
    }
    if (s->angular_damping && std::memcmp(b.h_angd.data() + first, s->angular_damping, sizeof(real) * n) != 0) {
        std::memcpy(b.h_angd.data() + first, s->angular_damping, sizeof(real) * n);
        dampingChanged = true;
    }
    if (dampingChanged) b.pow_dt = (real)NAN;   // Pow factors are stale
    return CZ_OK;
}
static int download_bodies(Batch &b, long long first, long long n, cz_bodies *s) {
    int rc;
    if ((rc = get_field(b, F_POS, first, n, s->position))) return rc;
    if ((rc = get_field(b, F_ORI, first, n, s->orientation))) return rc;
    if ((rc = get_field(b, F_VEL, first, n, s->velocity))) return rc;
    if ((rc = get_field(b, F_ROT, first, n, s->rotation))) return rc;
    if ((rc = get_field(b, F_ACC, first, n, s->acceleration))) return rc;
    if ((rc = get_field(b, F_LIND, first, n, s->linear_damping))) return rc;
    if ((rc = get_field(b, F_ANGD, first, n, s->angular_damping))) return rc;
    if ((rc = get_field(b, F_IITB, first, n, s->inverse_inertia_tensor))) return rc;
    if ((rc = get_field(b, F_INVM, first, n, s->inverse_mass))) return rc;
    if ((rc = get_field(b, F_MOTION, first, n, s->motion))) return rc;
    if ((rc = get_field(b, F_TRANSFORM, first, n, s->transform))) return rc;
    if ((rc = get_field(b, F_IITW, first, n, s->inverse_inertia_tensor_world))) return rc;
    if ((rc = get_field(b, F_LACC, first, n, s->last_frame_acceleration))) return rc;
    if ((rc = get_u8(b, b.st.awake, first, n, s->is_awake))) return rc;
    if ((rc = get_u8(b, b.st.can_sleep, first, n, s->can_sleep))) return rc;
    return CZ_OK;
}
static int upload_colliders(Batch &b, long long first, long long n, const cz_colliders *c) {
    int rc;
    cz_ctx *ctx = b.ctx;
    if (c->shape) {
        std::vector<uint8_t> sh(n);
        for (long long i = 0; i < n; i++) sh[i] = (uint8_t)c->shape[i];
        if ((rc = put_u8(b, b.st.shape, first, n, sh.data()))) return rc;
    }
    if ((rc = put_field(b, F_OFFSET, first, n, c->offset))) return rc;
    if ((rc = put_field(b, F_CTRANSFORM, first, n, c->transform))) return rc;
    if ((rc = put_field(b, F_HALF, first, n, c->half_size))) return rc;
    if ((rc = put_field(b, F_RADIUS, first, n, c->radius))) return rc;
    if (c->offset) {
        k_detect_identity<<<nblk(n, 256), 256, 0, ctx->stream>>>(b.st, first, n);
        CKL(ctx);
    }
    return CZ_OK;
}

// ------------------------------------------------------------------------------------------
// world handle
// ------------------------------------------------------------------------------------------
struct cz_world {
    cz_ctx *ctx = nullptr;
    cz_world_desc d{};
    Batch b;
    int P = 0;
    PlaneView planes[CZ_MAX_PLANES];
    int nchk = 0;
    int *d_one = nullptr, *d_two = nullptr;
    long long step_index = 0;
    real *gen = nullptr;
    int *gb0 = nullptr, *gb1 = nullptr, *nContacts = nullptr, *posIters = nullptr, *velIters = nullptr;
    unsigned long long *stats = nullptr;   // device ST_N counters
    // narrowphase multi-tile scratch
    int tiles = 1, tileThreads = 32;
    int *tileCount = nullptr, *tileBase = nullptr;
    uint8_t *hitCount = nullptr;
    // resolver
    ResolveScratch rs{};
    int resolveNT = 32;
    int resolveSmem = 0;        // dynamic shared bytes when staged in smem, 0 = global scratch
    // contact islands of one large world (k_resolve_islands): tables, snapshot for the detected-cap fallback
    IslandTable isl{};
    int *h_islCount = nullptr;        // pinned [4]
    real2 *islBackup = nullptr;       // chunks C_L2T0 .. C_W78 of the body store (what a resolve may write besides the work record)
    bool lastCapHit = false;          // the previous frame's loops ended at the iteration cap: islands would only be thrown away
    long long islandFrames = 0, islandFallbacks = 0;
    bool useBP = false;               // sort-based broadphase (K2) instead of the all-pairs tiles
    czbp::Broadphase bp;
    bool useFused = false;
    czf::FusedPlan fused{};
    unsigned int *d_next = nullptr;   // dynamic world counter of the fused kernel
    int *order3 = nullptr;            // [3][W] processing order per phase (czf::k_order_worlds)
    bool useOrder = false;
    real bias = (real)NAN;
    // episodes
    int episodeLen = 0;
    long long episodeStep0 = 0;
    int *d_phase0 = nullptr;
    Batch snap;
    unsigned long long *h_stats = nullptr;   // pinned
    // chunked pipeline of cz_world_step_host: H2D | pack + step + unpack | D2H on three streams
    struct HostPipe {
        bool ready = false;
        int chunks = 1;
        cudaStream_t sUp = nullptr, sDown = nullptr;
        int nComp = 1;                               // compute streams: chunk c runs on stream c % nComp, so kernel tails overlap
        cudaStream_t sComp[16] = {};                 // own streams, earlier chunks at higher priority
        std::vector<cudaEvent_t> evUp, evComp;
        cudaEvent_t evBegin = nullptr, evDownDone = nullptr;
        real *dIn = nullptr, *dOut = nullptr;        // per-field device staging
        uint8_t *dFlags = nullptr;                   // awake_in, can_sleep_in, awake_out
        unsigned int *dNext = nullptr;               // per-chunk world counters
        real *coldX[16] = {};                        // cold-contact scratch per compute stream beyond the first (kernels of different chunks co-run)
        // RL pipeline slots (cz_world_step_rl_async): staging alternates between two sets, so the downloads of one step overlap the frames of the next
        real *dInSlot[2] = {}, *dOutSlot[2] = {};
        float *dObs32[2] = {};
        cudaEvent_t evSlot[2] = {};
        int inFlight = 0, nextTicket = 0;
        long long launchesInFlight = 0, stepsInFlight = 0;
    } pipe;
    // Resident split-mode steps run the batch as a few independent slices on their own streams (see cz_world_step):
    // the tail of one slice's phase launch is filled by the other slices' launches.
    struct StepLanes {
        bool ready = false;
        int n = 1;
        cudaStream_t s[4] = {};
        real *cold[4] = {};                  // group scratch of the launches of lane k (lane 0 uses the plan's)
        unsigned int *dNext = nullptr;       // [4 * n] world counters
        cudaEvent_t evBegin = nullptr, evDone[4] = {};
    } lanes;
    real *h_pin = nullptr;
    size_t h_pin_bytes = 0;
    // per-pair surface materials (cz_world_set_materials); nMat == 0: the reference's constants
    int nMat = 0;
    real *d_matFric = nullptr, *d_matRest = nullptr;   // [nMat * nMat]
    uint8_t *d_bodyMat = nullptr;                      // [W * B]
    uint8_t planeMat[CZ_MAX_PLANES] = {};
    float *d_export = nullptr;                         // staging of cz_world_export_gl
    size_t exportFloats = 0;
};

static WorldParams world_params(cz_world *w) {
    WorldParams p;
    p.st = w->b.st;
    p.W = w->d.n_worlds; p.B = w->d.bodies_per_world; p.P = w->P; p.Cc = w->d.contacts_per_world;
    p.wFirst = 0; p.wCount = w->d.n_worlds;
    p.order = nullptr;
    p.nchk = w->nchk;
    p.schedule = w->d.schedule;
    p.chk_one = w->d_one; p.chk_two = w->d_two;
    for (int i = 0; i < CZ_MAX_PLANES; i++) p.planes[i] = w->planes[i];
    p.step_index = w->step_index;
    p.gen = w->gen; p.gb0 = w->gb0; p.gb1 = w->gb1;
    p.nContacts = w->nContacts; p.posIters = w->posIters; p.velIters = w->velIters;
    p.stats = w->stats;
    p.episodeLen = w->episodeLen; p.episodeStep0 = w->episodeStep0; p.phase0 = w->d_phase0; p.snap = w->snap.st;
    p.nMat = w->nMat;
    p.matFric = w->nMat > 0 ? w->d_matFric : nullptr; p.matRest = w->nMat > 0 ? w->d_matRest : nullptr; p.bodyMat = w->d_bodyMat;
    for (int i = 0; i < CZ_MAX_PLANES; i++) p.planeMat[i] = w->planeMat[i];
    return p;
}


// launch the per-phase world ordering for a world range and return the order pointer of a phase
// (also zeroes the four world counters at `counters`; returns whether it ran)
static inline bool order_worlds(cz_world *w, const WorldParams &p, cudaStream_t st, long long &launches, unsigned int *counters) {
    if (!w->useOrder) return false;
    czf::k_order_worlds<<<3 * czf::ORDER_SLICES, 1024, 0, st>>>(p.nContacts, p.posIters, p.velIters, p.wFirst, p.wCount, p.W, w->order3, counters);
    launches++;
    return true;
}
static inline const int *phase_order(cz_world *w, int ph) {
    if (!w->useOrder) return nullptr;
    return w->order3 + (size_t)(ph == czf::PH_A ? 0 : (ph == czf::PH_B ? 1 : 2)) * w->d.n_worlds;
}

template <class T> static inline void free_and_null(T *&p) { if (p) { cudaFree(p); p = nullptr; } }

// The chunk pipeline of cz_world_step_host / cz_world_step_rl is sized from the fused plan (cold scratch per compute
// stream, chunk count, compute streams).  A re-plan (planes or schedule uploaded later) invalidates it: it is torn
// down here and rebuilt from the current plan by the next host step.
static void host_pipe_destroy(cz_world *w) {
    auto &pp = w->pipe;
    if (!pp.ready) return;
    cudaStreamSynchronize(w->ctx->stream);
    cudaStreamSynchronize(pp.sUp); cudaStreamSynchronize(pp.sDown);
    for (int k = 0; k < pp.nComp; k++) if (pp.sComp[k]) cudaStreamSynchronize(pp.sComp[k]);
    cudaStreamDestroy(pp.sUp); cudaStreamDestroy(pp.sDown);
    for (int k = 0; k < pp.nComp; k++) if (pp.sComp[k] && pp.sComp[k] != w->ctx->stream) cudaStreamDestroy(pp.sComp[k]);
    for (auto e : pp.evUp) cudaEventDestroy(e);
    for (auto e : pp.evComp) cudaEventDestroy(e);
    cudaEventDestroy(pp.evBegin); cudaEventDestroy(pp.evDownDone);
    cudaFree(pp.dIn); cudaFree(pp.dOut); cudaFree(pp.dFlags); cudaFree(pp.dNext);
    for (int k = 1; k < 16; k++) if (pp.coldX[k]) cudaFree(pp.coldX[k]);
    if (pp.dInSlot[1]) cudaFree(pp.dInSlot[1]);
    if (pp.dOutSlot[1]) cudaFree(pp.dOutSlot[1]);
    for (int k = 0; k < 2; k++) { if (pp.dObs32[k]) cudaFree(pp.dObs32[k]); if (pp.evSlot[k]) cudaEventDestroy(pp.evSlot[k]); }
    pp = cz_world::HostPipe();
}

static void step_lanes_destroy(cz_world *w) {
    auto &ln = w->lanes;
    if (!ln.ready) return;
    cudaStreamSynchronize(w->ctx->stream);
    for (int k = 0; k < ln.n; k++) {
        if (ln.s[k]) { cudaStreamSynchronize(ln.s[k]); cudaStreamDestroy(ln.s[k]); }
        if (k > 0 && ln.cold[k]) cudaFree(ln.cold[k]);
        if (ln.evDone[k]) cudaEventDestroy(ln.evDone[k]);
    }
    if (ln.evBegin) cudaEventDestroy(ln.evBegin);
    if (ln.dNext) cudaFree(ln.dNext);
    ln = cz_world::StepLanes();
}
// Lanes of the resident split-mode step.  Measured (tools/lanes_probe.py, us per frame at 1 / 2 / 3 / 4 lanes):
// 65 536 worlds 1920 / 1720 / 1766 / 1795, 49 152: 1532 / 1362 / 1392 / 1374, 32 768: 1153 / 1049 / 1008 / 1002,
// 24 576: 964 / 879 / 851 / 820, 16 384: 809 / 749 / 698 / 675, 12 288: 713 / 647 / 619 / 593 -> two slices for
// large batches, four below 40 960 worlds (CUBEZ_STEP_LANES overrides).
static int step_lanes_init(cz_world *w) {
    cz_ctx *ctx = w->ctx;
    auto &ln = w->lanes;
    if (ln.ready) return CZ_OK;
    int n = czf::env_int("CUBEZ_STEP_LANES", w->d.n_worlds >= 40960 ? 2 : 4);
    n = std::max(1, std::min(4, std::min(n, w->d.n_worlds / std::max(1, czf::env_int("CUBEZ_STEP_LANE_MIN", 2048)))));
    ln.n = n;
    if (n > 1) {
        CK(ctx, cudaMalloc(&ln.dNext, sizeof(unsigned int) * 4 * n));
        CK(ctx, cudaEventCreateWithFlags(&ln.evBegin, cudaEventDisableTiming));
        for (int k = 0; k < n; k++) {
            CK(ctx, cudaStreamCreateWithFlags(&ln.s[k], cudaStreamNonBlocking));
            CK(ctx, cudaEventCreateWithFlags(&ln.evDone[k], cudaEventDisableTiming));
            if (k == 0) ln.cold[0] = w->fused.cold;
            else CK(ctx, cudaMalloc(&ln.cold[k], sizeof(real) * w->fused.coldReals * (size_t)w->fused.maxGrid * w->fused.groupsPerBlock));
        }
    }
    ln.ready = true;
    return CZ_OK;
}

// (Re)derive everything that depends on the schedule size / capacities.
static int world_plan(cz_world *w) {
    cz_ctx *ctx = w->ctx;
    host_pipe_destroy(w);
    step_lanes_destroy(w);
    const int B = w->d.bodies_per_world, Cc = w->d.contacts_per_world, W = w->d.n_worlds;
    if (w->d.schedule == CZ_SCHED_ALL_PAIRS_ORDERED) w->nchk = B * (w->P + B);
    // narrowphase tiling
    int threads = (w->nchk + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    w->tileThreads = threads;
    w->tiles = w->nchk > 0 ? (w->nchk + threads - 1) / threads : 1;
    if (w->tileCount) { cudaFree(w->tileCount); w->tileCount = nullptr; }
    if (w->tileBase) { cudaFree(w->tileBase); w->tileBase = nullptr; }
    if (w->hitCount) { cudaFree(w->hitCount); w->hitCount = nullptr; }
    if (w->tiles > 1) {
        CK(ctx, cudaMalloc(&w->tileCount, sizeof(int) * (size_t)W * w->tiles));
        CK(ctx, cudaMalloc(&w->tileBase, sizeof(int) * (size_t)W * w->tiles));
        CK(ctx, cudaMalloc(&w->hitCount, (size_t)W * w->nchk));
    }
    // resolver staging
    size_t need = sizeof(real) * ((size_t)czr::BW_NF * B + (size_t)CW_NREAL * Cc) + sizeof(int) * 2 * (size_t)Cc;
    w->resolveNT = Cc <= 128 ? 32 : 256;
    size_t limit = ctx->smem_optin > 2048 ? ctx->smem_optin - 2048 : 0;
    if (need <= limit) {
        w->resolveSmem = (int)need;
    } else {
        w->resolveSmem = 0;
        if (!w->rs.bw) {
            CK(ctx, cudaMalloc(&w->rs.bw, sizeof(real) * (size_t)W * czr::BW_NF * B));
            CK(ctx, cudaMalloc(&w->rs.cw, sizeof(real) * (size_t)W * CW_NREAL * Cc));
            CK(ctx, cudaMalloc(&w->rs.cb, sizeof(int) * (size_t)W * 2 * Cc));
            if (!czf::env_int("CUBEZ_RESOLVE_NO_PRE", 0)) CK(ctx, cudaMalloc(&w->rs.pre, sizeof(real) * (size_t)W * czr::VP_NF * Cc));
            CK(ctx, cudaMalloc(&w->rs.adj, sizeof(unsigned short) * (size_t)W * (2 * (size_t)Cc + B + 2)));   // lists, then the list offsets
        }
    }
    // sort-based broadphase for one large world
    w->useBP = false;
    if (w->d.flags & CZ_WORLD_BROADPHASE) {
        if (W != 1 || w->d.schedule != CZ_SCHED_ALL_PAIRS_ORDERED)
            return fail(ctx, CZ_ERR_INVALID, "CZ_WORLD_BROADPHASE needs a single world with the all-pairs-ordered schedule");
        if (!w->bp.bounds) {
            cudaError_t e = czbp::bp_alloc(w->bp, B, (unsigned long long)B * 64ull + 1024ull, (unsigned long long)Cc * 2ull + 1024ull,
                                           std::max<long long>(1ll << 22, 16ll * B));
            if (e != cudaSuccess) return fail(ctx, CZ_ERR_CUDA, std::string("broadphase alloc: ") + cudaGetErrorString(e));
        }
        if (!w->isl.start) {
            CK(ctx, cudaMalloc(&w->isl.start, sizeof(int) * ((size_t)Cc + 2)));
            CK(ctx, cudaMalloc(&w->isl.count, sizeof(int) * 4));
            CK(ctx, cudaHostAlloc(&w->h_islCount, sizeof(int) * 4, cudaHostAllocDefault));
            CK(ctx, cudaMalloc(&w->islBackup, sizeof(real2) * (size_t)w->b.st.stride * 11));
        }
        w->useBP = true;
    }
    // fused small-world kernel
    w->useFused = false;
    if (!(w->d.flags & CZ_WORLD_NO_FUSED)) {
        free_and_null(w->fused.cold);
        free_and_null(w->fused.coldW); free_and_null(w->fused.preW); free_and_null(w->fused.hotPen); free_and_null(w->fused.hotDdv);
        free_and_null(w->fused.hotCb0); free_and_null(w->fused.hotCb1);
        w->useFused = !w->useBP && czf::plan(w->fused, B, w->P, Cc, w->nchk, w->d.schedule, ctx->smem_optin, ctx->sm_count, W);
        if (w->useFused) {
            CK(ctx, cudaMalloc(&w->fused.cold, sizeof(real) * w->fused.coldReals * (size_t)w->fused.maxGrid * w->fused.groupsPerBlock));
            if (!w->d_next) CK(ctx, cudaMalloc(&w->d_next, sizeof(unsigned int) * 4));
            if (!w->order3) CK(ctx, cudaMalloc(&w->order3, sizeof(int) * 3 * (size_t)W));
            w->useOrder = czf::env_int("CUBEZ_FUSED_ORDER", 1) != 0 && W >= 64;
            if (w->fused.split) {
                const size_t WC = (size_t)W * Cc;
                CK(ctx, cudaMalloc(&w->fused.coldW, sizeof(real) * WC * czr::CW_NCOLD));
                if (w->fused.velPre) CK(ctx, cudaMalloc(&w->fused.preW, sizeof(real) * WC * czr::VP_NF));
                CK(ctx, cudaMalloc(&w->fused.hotPen, sizeof(real) * WC));
                CK(ctx, cudaMalloc(&w->fused.hotDdv, sizeof(real) * WC));
                CK(ctx, cudaMalloc(&w->fused.hotCb0, sizeof(int) * WC));
                CK(ctx, cudaMalloc(&w->fused.hotCb1, sizeof(int) * WC));
            }
        }
    }
    return CZ_OK;
}

static int take_status(cz_world *w);

extern "C" {

int cz_real_size(void) { return (int)sizeof(real); }

const char *cz_last_error(cz_ctx *ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int cz_init(int device, cz_ctx **out) {
    if (!out) return fail(nullptr, CZ_ERR_INVALID, "cz_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, CZ_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libcubezcuda has no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, CZ_ERR_INVALID, "cz_init: bad device index");
    cz_ctx *ctx = new cz_ctx;
    ctx->device = device;
    CK(ctx, cudaSetDevice(device));
    CK(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(ctx, cudaEventCreate(&ctx->ev0));
    CK(ctx, cudaEventCreate(&ctx->ev1));
    cudaDeviceProp prop;
    CK(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    if (const char *e = getenv("CUBEZ_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e));
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    *out = ctx;
    return CZ_OK;
}
int cz_shutdown(cz_ctx *ctx) {
    if (!ctx) return CZ_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    czp::pool_trim(ctx->device);
    delete ctx;
    return CZ_OK;
}
void *cz_ctx_stream(cz_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int cz_ctx_synchronize(cz_ctx *ctx) {
    if (!ctx) return CZ_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CZ_OK;
}
int cz_host_alloc(cz_ctx *ctx, uint64_t bytes, void **out) {
    if (!ctx || !out) return CZ_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return CZ_OK;
}
int cz_host_free(cz_ctx *ctx, void *p) {
    if (!ctx) return CZ_ERR_INVALID;
    CK(ctx, cudaFreeHost(p));
    return CZ_OK;
}

// ---- world -----------------------------------------------------------------------------
int cz_world_create(cz_ctx *ctx, const cz_world_desc *desc, cz_world **out) {
    if (!ctx || !desc || !out) return fail(ctx, CZ_ERR_INVALID, "cz_world_create: NULL argument");
    if (desc->n_worlds <= 0 || desc->bodies_per_world <= 0 || desc->contacts_per_world <= 0)
        return fail(ctx, CZ_ERR_INVALID, "cz_world_create: sizes must be positive");
    CK(ctx, cudaSetDevice(ctx->device));
    cz_world *w = new cz_world;
    w->ctx = ctx;
    w->d = *desc;
    const long long NB = (long long)desc->n_worlds * desc->bodies_per_world;
    const size_t NC = (size_t)desc->n_worlds * desc->contacts_per_world;
    int rc = batch_alloc(ctx, w->b, NB);
    if (rc) { cz_world_destroy(w); return rc; }
    auto alloc_all = [&]() -> int {
        CK(ctx, cudaMalloc(&w->gen, sizeof(real) * czr::G_NF * NC));
        CK(ctx, cudaMalloc(&w->gb0, sizeof(int) * NC));
        CK(ctx, cudaMalloc(&w->gb1, sizeof(int) * NC));
        CK(ctx, cudaMalloc(&w->nContacts, sizeof(int) * desc->n_worlds));
        CK(ctx, cudaMalloc(&w->posIters, sizeof(int) * desc->n_worlds));
        CK(ctx, cudaMalloc(&w->velIters, sizeof(int) * desc->n_worlds));
        CK(ctx, cudaMemsetAsync(w->nContacts, 0, sizeof(int) * desc->n_worlds, ctx->stream));
        CK(ctx, cudaMemsetAsync(w->posIters, 0, sizeof(int) * desc->n_worlds, ctx->stream));
        CK(ctx, cudaMemsetAsync(w->velIters, 0, sizeof(int) * desc->n_worlds, ctx->stream));
        CK(ctx, cudaMalloc(&w->stats, sizeof(unsigned long long) * ST_N));
        CK(ctx, cudaMemsetAsync(w->stats, 0, sizeof(unsigned long long) * ST_N, ctx->stream));
        CK(ctx, cudaHostAlloc(&w->h_stats, sizeof(unsigned long long) * ST_N, cudaHostAllocDefault));
        return CZ_OK;
    };
    if ((rc = alloc_all())) { cz_world_destroy(w); return rc; }   // no leak on a failed allocation
    for (int i = 0; i < CZ_MAX_PLANES; i++) { w->planes[i].n = czm::mk3(0, 1, 0); w->planes[i].offset = 0; }
    rc = world_plan(w);
    if (rc) { cz_world_destroy(w); return rc; }
    *out = w;
    return CZ_OK;
}
int cz_world_destroy(cz_world *w) {
    if (!w) return CZ_ERR_INVALID;
    cudaSetDevice(w->ctx->device);
    cudaStreamSynchronize(w->ctx->stream);
    step_lanes_destroy(w);
    batch_free(w->b);
    if (w->bp.bounds) czbp::bp_free(w->bp);
    if (w->snap.st.base) batch_free(w->snap);
    if (w->d_phase0) cudaFree(w->d_phase0);
    if (w->h_islCount) cudaFreeHost(w->h_islCount);
    void *ptrs[] = {w->isl.start, w->isl.count, w->islBackup, w->d_one, w->d_two, w->gen, w->gb0, w->gb1, w->nContacts, w->posIters, w->velIters, w->stats,
                    w->tileCount, w->tileBase, w->hitCount, w->rs.bw, w->rs.cw, w->rs.cb, w->rs.pre, w->rs.adj, w->fused.cold, w->d_next, w->order3,
                    w->fused.coldW, w->fused.preW, w->fused.hotPen, w->fused.hotDdv, w->fused.hotCb0, w->fused.hotCb1, w->d_matFric, w->d_matRest, w->d_bodyMat, w->d_export};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (w->h_stats) cudaFreeHost(w->h_stats);
    if (w->h_pin) cudaFreeHost(w->h_pin);
    host_pipe_destroy(w);
    delete w;
    return CZ_OK;
}
static int world_range(cz_world *w, int first, int n) {
    if (!w) return fail(nullptr, CZ_ERR_INVALID, "NULL world");
    if (first < 0 || n < 0 || first + n > w->d.n_worlds) return fail(w->ctx, CZ_ERR_INVALID, "world range out of bounds");
    return CZ_OK;
}
int cz_world_upload_bodies(cz_world *w, int32_t first, int32_t n, const cz_bodies *b, int32_t derive) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    const long long B = w->d.bodies_per_world;
    if ((rc = upload_bodies(w->b, first * B, n * B, b))) return rc;
    if (derive) {
        k_derive<<<nblk(n * B, 256), 256, 0, ctx->stream>>>(w->b.st, first * B, n * B, 1, 1);
        CKL(ctx);
    }
    return CZ_OK;
}
int cz_world_upload_colliders(cz_world *w, int32_t first, int32_t n, const cz_colliders *c, int32_t derive) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    const long long B = w->d.bodies_per_world;
    if ((rc = upload_colliders(w->b, first * B, n * B, c))) return rc;
    if (derive) {
        k_derive<<<nblk(n * B, 256), 256, 0, ctx->stream>>>(w->b.st, first * B, n * B, 0, 1);
        CKL(ctx);
    }
    return CZ_OK;
}
int cz_world_upload_planes(cz_world *w, const cz_planes *p) {
    if (!w || !p) return CZ_ERR_INVALID;
    if (p->n > CZ_MAX_PLANES) return fail(w->ctx, CZ_ERR_INVALID, "too many planes (max 8)");
    CK(w->ctx, cudaSetDevice(w->ctx->device));   // the re-plan below allocates: on THIS world's device, whatever the calling thread's current one is
    w->P = p->n;
    for (int i = 0; i < p->n; i++) {
        w->planes[i].n = czm::mk3(p->normal[i * 3], p->normal[i * 3 + 1], p->normal[i * 3 + 2]);
        w->planes[i].offset = p->offset[i];
    }
    return world_plan(w);
}
int cz_world_upload_schedule(cz_world *w, int32_t n_checks, const int32_t *one, const int32_t *two) {
    if (!w || n_checks < 0) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    if (w->d.schedule != CZ_SCHED_EXPLICIT) return fail(ctx, CZ_ERR_INVALID, "world was not created with CZ_SCHED_EXPLICIT");
    if (n_checks > 0 && (!one || !two)) return fail(ctx, CZ_ERR_INVALID, "cz_world_upload_schedule: NULL check list");
    for (int i = 0; i < n_checks; i++) {
        // planes are named -(p+1) and must exist: upload the planes before the schedule
        if (one[i] >= w->d.bodies_per_world || two[i] >= w->d.bodies_per_world || one[i] < -w->P || two[i] < -w->P)
            return fail(ctx, CZ_ERR_INVALID, "schedule entry out of range (body id >= bodies_per_world, or a plane that was not uploaded)");
    }
    if (w->d_one) cudaFree(w->d_one);
    if (w->d_two) cudaFree(w->d_two);
    w->d_one = w->d_two = nullptr;
    CK(ctx, cudaMalloc(&w->d_one, sizeof(int) * (n_checks + 1)));
    CK(ctx, cudaMalloc(&w->d_two, sizeof(int) * (n_checks + 1)));
    CK(ctx, cudaMemcpy(w->d_one, one, sizeof(int) * n_checks, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(w->d_two, two, sizeof(int) * n_checks, cudaMemcpyHostToDevice));
    w->nchk = n_checks;
    return world_plan(w);
}
int cz_world_set_activation(cz_world *w, int32_t first, int32_t n, const int32_t *active_from, const uint8_t *integrate) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    const long long B = w->d.bodies_per_world;
    if (active_from) CK(ctx, cudaMemcpy(w->b.st.active_from + first * B, active_from, sizeof(int32_t) * n * B, cudaMemcpyHostToDevice));
    if (integrate) CK(ctx, cudaMemcpy(w->b.st.integ + first * B, integrate, n * B, cudaMemcpyHostToDevice));
    return CZ_OK;
}
int cz_world_set_pow(cz_world *w, cz_real dt, const cz_real *lin_pow, const cz_real *ang_pow, cz_real bias) {
    if (!w || !lin_pow || !ang_pow) return CZ_ERR_INVALID;
    CK(w->ctx, cudaSetDevice(w->ctx->device));
    int rc;
    if ((rc = put_field(w->b, F_LINPOW, 0, w->b.n, lin_pow))) return rc;
    if ((rc = put_field(w->b, F_ANGPOW, 0, w->b.n, ang_pow))) return rc;
    if (w->snap.st.base) {
        if ((rc = put_field(w->snap, F_LINPOW, 0, w->b.n, lin_pow))) return rc;
        if ((rc = put_field(w->snap, F_ANGPOW, 0, w->b.n, ang_pow))) return rc;
    }
    w->b.pow_dt = dt;
    w->b.pow_user = true;
    w->bias = bias;
    return CZ_OK;
}
int cz_world_set_materials(cz_world *w, int32_t n_materials, const cz_real *friction, const cz_real *restitution, int32_t first, int32_t n,
                           const int32_t *body_material, const int32_t *plane_material) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_materials <= 0) { w->nMat = 0; return CZ_OK; }
    if (n_materials > 255 || !friction || !restitution) return fail(ctx, CZ_ERR_INVALID, "cz_world_set_materials: 1..255 materials with both tables");
    const long long B = w->d.bodies_per_world, NB = w->b.n;
    if (body_material)
        for (long long i = 0; i < n * B; i++)
            if (body_material[i] < 0 || body_material[i] >= n_materials) return fail(ctx, CZ_ERR_INVALID, "body material id out of range");
    if (plane_material)
        for (int i = 0; i < w->P; i++)
            if (plane_material[i] < 0 || plane_material[i] >= n_materials) return fail(ctx, CZ_ERR_INVALID, "plane material id out of range");
    if (w->nMat != n_materials) {   // a new table size resets every id to material 0
        if (w->d_matFric) { cudaFree(w->d_matFric); w->d_matFric = nullptr; }
        if (w->d_matRest) { cudaFree(w->d_matRest); w->d_matRest = nullptr; }
        CK(ctx, cudaMalloc(&w->d_matFric, sizeof(real) * n_materials * n_materials));
        CK(ctx, cudaMalloc(&w->d_matRest, sizeof(real) * n_materials * n_materials));
        if (!w->d_bodyMat) CK(ctx, cudaMalloc(&w->d_bodyMat, (size_t)NB));
        CK(ctx, cudaMemset(w->d_bodyMat, 0, (size_t)NB));
        for (int i = 0; i < CZ_MAX_PLANES; i++) w->planeMat[i] = 0;
    }
    CK(ctx, cudaMemcpy(w->d_matFric, friction, sizeof(real) * n_materials * n_materials, cudaMemcpyHostToDevice));
    CK(ctx, cudaMemcpy(w->d_matRest, restitution, sizeof(real) * n_materials * n_materials, cudaMemcpyHostToDevice));
    if (body_material && n > 0) {
        std::vector<uint8_t> ids((size_t)(n * B));
        for (long long i = 0; i < n * B; i++) ids[i] = (uint8_t)body_material[i];
        CK(ctx, cudaMemcpy(w->d_bodyMat + first * B, ids.data(), ids.size(), cudaMemcpyHostToDevice));
    }
    if (plane_material) for (int i = 0; i < w->P && i < CZ_MAX_PLANES; i++) w->planeMat[i] = (uint8_t)plane_material[i];
    w->nMat = n_materials;
    return CZ_OK;
}
int cz_world_add_forces(cz_world *w, int32_t first, int32_t n, const cz_real *force, const cz_real *torque) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    if (n == 0 || (!force && !torque)) return CZ_OK;
    const long long B = w->d.bodies_per_world, NB = w->b.n, nb = (long long)n * B;
    if (!w->b.st.force) {   // first use: the accumulators come into existence, zeroed
        real *f = nullptr, *t = nullptr;
        CK(ctx, cudaMalloc(&f, sizeof(real) * 3 * (size_t)w->b.st.stride));
        CK(ctx, cudaMalloc(&t, sizeof(real) * 3 * (size_t)w->b.st.stride));
        CK(ctx, cudaMemsetAsync(f, 0, sizeof(real) * 3 * (size_t)w->b.st.stride, ctx->stream));
        CK(ctx, cudaMemsetAsync(t, 0, sizeof(real) * 3 * (size_t)w->b.st.stride, ctx->stream));
        w->b.st.force = f; w->b.st.torque = t;
    }
    (void)NB;
    // staging: the pack buffer of the batch holds 12 reals per body
    real *dF = w->b.stage, *dT = w->b.stage + nb * 3;
    if (force) CK(ctx, cudaMemcpyAsync(dF, force, sizeof(real) * nb * 3, cudaMemcpyHostToDevice, ctx->stream));
    if (torque) CK(ctx, cudaMemcpyAsync(dT, torque, sizeof(real) * nb * 3, cudaMemcpyHostToDevice, ctx->stream));
    k_add_forces<<<nblk(nb * 3, 256), 256, 0, ctx->stream>>>(w->b.st, first * B, nb, force ? dF : nullptr, torque ? dT : nullptr);
    CKL(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));   // the staging buffer is shared with the field uploads
    return CZ_OK;
}
int cz_world_set_step_index(cz_world *w, int64_t s) {
    if (!w) return CZ_ERR_INVALID;
    w->step_index = s;
    return CZ_OK;
}
int cz_world_set_episodes(cz_world *w, int32_t length, const int32_t *phase0) {
    if (!w) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    if (length <= 0) { w->episodeLen = 0; return CZ_OK; }
    const int W = w->d.n_worlds;
    if (phase0) for (int i = 0; i < W; i++) if (phase0[i] < 0 || phase0[i] >= length) return fail(ctx, CZ_ERR_INVALID, "phase0 out of range");
    if (!w->snap.st.base) {
        int rc = batch_alloc(ctx, w->snap, w->b.n);
        if (rc) return rc;
    }
    if (!w->d_phase0) CK(ctx, cudaMalloc(&w->d_phase0, sizeof(int) * W));
    if (phase0) CK(ctx, cudaMemcpyAsync(w->d_phase0, phase0, sizeof(int) * W, cudaMemcpyHostToDevice, ctx->stream));
    else CK(ctx, cudaMemsetAsync(w->d_phase0, 0, sizeof(int) * W, ctx->stream));
    const long long stride = w->b.st.stride;
    CK(ctx, cudaMemcpyAsync(w->snap.st.base, w->b.st.base, sizeof(real2) * stride * czb::N_CHUNKS, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.awake, w->b.st.awake, stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.can_sleep, w->b.st.can_sleep, stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.integ, w->b.st.integ, stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.shape, w->b.st.shape, stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.ident, w->b.st.ident, stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaMemcpyAsync(w->snap.st.active_from, w->b.st.active_from, sizeof(int32_t) * stride, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    w->episodeLen = length;
    w->episodeStep0 = w->step_index;
    return CZ_OK;
}
int cz_world_synchronize(cz_world *w) {
    if (!w) return CZ_ERR_INVALID;
    CK(w->ctx, cudaSetDevice(w->ctx->device));
    CK(w->ctx, cudaStreamSynchronize(w->ctx->stream));
    return take_status(w);
}

}  // extern "C"

template <int NT>
static void launch_resolve(cz_world *w, const WorldParams &p, int maxIterOverride, real dt, bool forceGlobal, long long contactsHint = -1) {
    int smem = forceGlobal ? 0 : w->resolveSmem;
    int mode = smem > 0 ? 1 : 0, hotCap = 0;
    // CUBEZ_RESOLVE_MODE: 0 plain CTA loop, 2 staged hot values + cached arg-max (round 1), 3 adjacency lists (default)
    int want_mode = czf::env_int("CUBEZ_RESOLVE_NO_CACHE", 0) ? 0 : czf::env_int("CUBEZ_RESOLVE_MODE", 3);
    if (want_mode == 4) want_mode = 3;   // 4 = mode 3 with the list offsets forced into global memory (test hook)
    if (NT > 32 && mode == 0 && p.B < 0xffff && want_mode != 0) {
        // one large world: the loop's working set in shared memory.  The host knows the contact count on the
        // broadphase path; otherwise size for the capacity.
        long long want = contactsHint >= 0 ? std::min<long long>(contactsHint, p.Cc) : p.Cc;
        want = (want + 255) / 256 * 256;
        const size_t limit = w->ctx->smem_optin > 4096 ? w->ctx->smem_optin - 4096 : 0;
        size_t big = czr::big_shared_bytes(NT, want, p.B);
        const size_t bytes = (size_t)want * (sizeof(real) + 4);
        // shared-memory carve-out steps of the SM: what is not carved stays L1 for the cold contact records
        auto carve = [](size_t b) { for (size_t kb : {8, 16, 32, 64, 100, 132, 164, 196, 228}) if (b + 2048 <= kb * 1024) return kb; return (size_t)256; };
        bool offsetsInGlobal = false;
        if (want_mode == 3 && w->rs.adj) {
            const size_t lean = czr::big_shared_bytes(NT, want, p.B, true);
            if (big > limit || carve(lean) < carve(big) || czf::env_int("CUBEZ_RESOLVE_MODE", 3) == 4) { big = lean; offsetsInGlobal = true; }
        }
        if (want_mode == 3 && want > 0 && want <= 32767 && big <= limit && w->rs.adj) { mode = offsetsInGlobal ? 4 : 3; hotCap = (int)want; smem = (int)big; }
        else if (want > 0 && bytes <= limit) { mode = 2; hotCap = (int)want; smem = (int)bytes; }
    }
    cudaError_t ea = cudaSuccess;
    if (smem + 2048 > 48 * 1024) ea = cudaFuncSetAttribute(k_resolve<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_resolve<NT><<<p.W, NT, smem, w->ctx->stream>>>(p, w->rs, mode, maxIterOverride, dt, hotCap);
    if (getenv("CUBEZ_RESOLVE_TRACE")) fprintf(stderr, "[resolve] NT %d mode %d smem %d hotCap %d hint %lld attr %s launch %s\n", NT, mode, smem, hotCap, contactsHint, cudaGetErrorString(ea), cudaGetErrorString(cudaPeekAtLastError()));
}

// one frame of updateCallback (examples/cubedrop.go:69-75) on the multi-kernel path
static int world_step_multi(cz_world *w, real dt, long long &launches) {
    cz_ctx *ctx = w->ctx;
    WorldParams p = world_params(w);
    const long long NB = w->b.n;
    int grid = (int)std::min<long long>(nblk(NB, 256), (long long)ctx->sm_count * 8);
    if (w->episodeLen > 0) {
        k_episode_reset<<<nblk(NB, 256), 256, 0, ctx->stream>>>(p);
        CKL(ctx);
        launches++;
    }
    k_integrate<true><<<grid, 256, 0, ctx->stream>>>(w->b.st, dt, w->bias, w->step_index);
    CKL(ctx);
    launches++;
    if (w->useBP) {
        // K2: sort-based broadphase -> candidate pairs -> narrowphase -> canonical contact sort
        czbp::Broadphase &bp = w->bp;
        cudaError_t e = czbp::bp_reset_box(bp, ctx->stream);
        if (e != cudaSuccess) return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
        czbp::k_bp_bounds<<<nblk(NB, 256), 256, 0, ctx->stream>>>(w->b.st, NB, w->step_index, bp.bounds, bp.box);
        launches++;
        if ((e = czbp::bp_candidates(bp, ctx->stream, &launches)) != cudaSuccess) return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
        CK(ctx, cudaMemcpyAsync(bp.h_counters, bp.counters, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        const unsigned long long nPairs = bp.h_counters[0];
        bp.lastPairs = nPairs;
        if (nPairs > bp.pairCapacity) return fail(ctx, CZ_ERR_CAPACITY, "broadphase candidate-pair capacity exceeded");
        czbp::KeyedContacts kc{bp.sortContacts.keys[0], bp.sortContacts.vals[0], bp.payload, bp.ids, bp.counters + 1, bp.contactCapacity};
        if (nPairs) czbp::k_bp_narrow<<<nblk((long long)nPairs * 2, 128), 128, 0, ctx->stream>>>(p, bp.pairs, nPairs, kc);
        if (p.P > 0) czbp::k_bp_planes<<<nblk((long long)p.B * p.P, 128), 128, 0, ctx->stream>>>(p, kc);
        launches += 2;
        CK(ctx, cudaMemcpyAsync(bp.h_counters, bp.counters, sizeof(unsigned long long) * 2, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        const unsigned long long nCont = bp.h_counters[1];
        bp.lastContacts = nCont;
        if (nCont > bp.contactCapacity || nCont > (unsigned long long)p.Cc) return fail(ctx, CZ_ERR_CAPACITY, "contact capacity exceeded (contacts_per_world too small)");
        int bits = 8;
        const unsigned long long maxKey = ((unsigned long long)p.B * (unsigned long long)(p.P + p.B) + 1ull) * 8ull;
        while (bits < 64 && (1ull << bits) <= maxKey) bits += 8;
        int cur = czs::radix_sort(bp.sortContacts, (long long)nCont, bits, ctx->stream, &launches);
        czbp::k_bp_emit<<<nblk(std::max<long long>((long long)nCont, 1), 128), 128, 0, ctx->stream>>>(p, bp.sortContacts.keys[cur], bp.sortContacts.vals[cur], bp.payload, bp.ids, bp.counters + 1);
        launches++;
        CKL(ctx);
        // Contact islands: one CTA per connected component of the contact graph, exact while the reference's loop
        // converges within its cap (see k_resolve_islands); the cap case is detected and re-run on the single-CTA path.
        // CUBEZ_RESOLVE_ISLANDS: 0 never, 1 when it can pay (default), 2 always.
        const int islMode = czf::env_int("CUBEZ_RESOLVE_ISLANDS", 1);
        bool resolved = false;
        if (w->resolveNT != 32 && w->resolveSmem == 0 && nCont > 0 && (islMode == 2 || (islMode == 1 && nCont >= 256 && !w->lastCapHit))) {
            long long want = ((long long)nCont + 255) / 256 * 256;
            const size_t limit = ctx->smem_optin > 4096 ? ctx->smem_optin - 4096 : 0;
            const size_t big = czr::big_shared_bytes(256, want, p.B);
            if (want <= 32767 && p.B <= 12000 && big <= limit && nCont <= (unsigned long long)p.Cc) {   // labels of every body in 48 KB of shared memory
                czs::RadixBuffers<unsigned long long> &sb = bp.sortContacts;
                k_island_labels<<<1, 1024, sizeof(int) * (size_t)p.B, ctx->stream>>>(p, sb.keys[0], sb.vals[0]);
                int kb = 8;
                while (kb < 32 && (1ll << kb) < (long long)p.B) kb += 8;
                const int cur2 = czs::radix_sort(sb, (long long)nCont, kb, ctx->stream, &launches);
                k_island_ranges<<<1, 1024, 0, ctx->stream>>>(p, sb.keys[cur2], w->isl);
                CK(ctx, cudaMemcpyAsync(w->islBackup, w->b.st.chunk(czb::C_L2T0), sizeof(real2) * (size_t)w->b.st.stride * 11, cudaMemcpyDeviceToDevice, ctx->stream));
                k_bw_load<<<nblk(p.B, 256), 256, 0, ctx->stream>>>(p, w->rs);
                cudaFuncSetAttribute(k_resolve_islands<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)big);
                const int grid = (int)std::min<unsigned long long>(nCont, (unsigned long long)ctx->sm_count * 4);
                k_resolve_islands<256><<<grid, 256, big, ctx->stream>>>(p, w->rs, sb.vals[cur2], w->isl, dt, (int)want);
                launches += 5;
                CKL(ctx);
                CK(ctx, cudaMemcpyAsync(w->h_islCount, w->isl.count, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(ctx, cudaStreamSynchronize(ctx->stream));
                const long long cap = 8ll * (long long)nCont;
                w->islandFrames++;
                if ((long long)w->h_islCount[1] <= cap && (long long)w->h_islCount[2] <= cap &&
                    // a loop that ends exactly AT the cap may have been cut: only a sum strictly below it proves convergence
                    !((long long)w->h_islCount[1] == cap || (long long)w->h_islCount[2] == cap)) {
                    k_bw_store<<<nblk(p.B, 256), 256, 0, ctx->stream>>>(p, w->rs, w->isl);
                    launches++;
                    resolved = true;
                    w->lastCapHit = false;
                } else {   // the reference's loop would have been cut by its cap: islands cannot reproduce where — restore and run it exactly
                    CK(ctx, cudaMemcpyAsync(w->b.st.chunk(czb::C_L2T0), w->islBackup, sizeof(real2) * (size_t)w->b.st.stride * 11, cudaMemcpyDeviceToDevice, ctx->stream));
                    w->islandFallbacks++;
                }
                if (getenv("CUBEZ_RESOLVE_TRACE")) fprintf(stderr, "[islands] contacts %llu islands %d largest %d pos %d vel %d cap %lld -> %s\n", nCont, w->h_islCount[0], w->h_islCount[3], w->h_islCount[1], w->h_islCount[2], cap, resolved ? "kept" : "fallback");
            }
        }
        if (!resolved) {
            if (w->resolveNT == 32) launch_resolve<32>(w, p, -1, dt, false);
            else launch_resolve<256>(w, p, -1, dt, false, (long long)nCont);
            CKL(ctx);
            launches++;
            // did the loops end at the cap?  (decides whether islands are tried on the next frame; read with the next frame's first sync)
            if (islMode == 1 && w->resolveNT != 32) {
                int it[2] = {0, 0};
                CK(ctx, cudaMemcpyAsync(&it[0], w->posIters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CK(ctx, cudaMemcpyAsync(&it[1], w->velIters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                CK(ctx, cudaStreamSynchronize(ctx->stream));
                w->lastCapHit = nCont > 0 && ((long long)it[0] >= 8ll * (long long)nCont || (long long)it[1] >= 8ll * (long long)nCont);
            }
        }
    } else if (w->nchk > 0) {
        if (w->tiles == 1) {
            k_narrow<NARROW_SINGLE><<<p.W, w->tileThreads, 0, ctx->stream>>>(p, 1, nullptr, nullptr, nullptr);
            CKL(ctx);
            launches++;
        } else {
            k_narrow<NARROW_COUNT><<<p.W * w->tiles, w->tileThreads, 0, ctx->stream>>>(p, w->tiles, w->tileCount, nullptr, w->hitCount);
            CKL(ctx);
            k_scan_tiles<<<p.W, 256, 0, ctx->stream>>>(p, w->tiles, w->tileCount, w->tileBase);
            CKL(ctx);
            k_narrow<NARROW_EMIT><<<p.W * w->tiles, w->tileThreads, 0, ctx->stream>>>(p, w->tiles, nullptr, w->tileBase, w->hitCount);
            CKL(ctx);
            launches += 3;
        }
        if (w->resolveNT == 32) launch_resolve<32>(w, p, -1, dt, false);
        else launch_resolve<256>(w, p, -1, dt, false);
        CKL(ctx);
        launches++;
    }
    return CZ_OK;
}

static int world_prepare_step(cz_world *w, real dt) {
    if (!(w->b.pow_dt == dt)) {   // also true when pow_dt is NaN
        int rc = refresh_pow(w->b, 0, w->b.n, dt);
        if (rc) return rc;
        if (w->snap.st.base) {   // the episode snapshot carries the Pow factors too
            w->snap.h_lind = w->b.h_lind; w->snap.h_angd = w->b.h_angd;
            if ((rc = refresh_pow(w->snap, 0, w->snap.n, dt))) return rc;
        }
        w->b.pow_dt = dt;
        w->b.pow_user = false;
        w->bias = (real)std::pow(0.5, (double)dt);   // rigidbody.go:250
    }
    return CZ_OK;
}

static int status_error(cz_ctx *ctx, int status) {
    if (status == CZ_ERR_CAPACITY) return fail(ctx, status, "contact capacity exceeded (contacts_per_world too small)");
    if (status == CZ_ERR_NIL_BODY) return fail(ctx, status, "frictionless one-body contact: the reference dereferences a nil body (contact.go:512-523)");
    if (status) return fail(ctx, status, "device-side error");
    return CZ_OK;
}
// The device-side status (first error raised by any kernel since it was last observed) is STICKY: a step never clears
// it, so errors raised by asynchronous steps (stats == NULL) are not lost.  Whoever observes it — a step with stats,
// cz_world_synchronize, a download, the counters, the checksum — returns it and clears it.
static int take_status(cz_world *w) {
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(&w->h_stats[ST_STATUS], w->stats + ST_STATUS, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const int status = -(int)w->h_stats[ST_STATUS];
    if (!status) return CZ_OK;
    CK(ctx, cudaMemset(w->stats + ST_STATUS, 0, sizeof(unsigned long long)));
    return status_error(ctx, status);
}

static int read_stats(cz_world *w, cz_step_stats *stats, long long launches, int n_steps, float ms) {
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaMemcpyAsync(w->h_stats, w->stats, sizeof(unsigned long long) * ST_N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    int status = -(int)w->h_stats[ST_STATUS];
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->steps = (int64_t)w->d.n_worlds * n_steps;
        stats->contacts = (int64_t)w->h_stats[ST_CONTACTS];
        stats->pos_iterations = (int64_t)w->h_stats[ST_POS];
        stats->vel_iterations = (int64_t)w->h_stats[ST_VEL];
        stats->checks = (int64_t)w->nchk * w->d.n_worlds * n_steps;
        stats->kernel_launches = launches;
        stats->max_contacts = (int32_t)w->h_stats[ST_MAXC];
        stats->status = status;
        stats->device_ms = ms;
    }
    if (status) CK(ctx, cudaMemsetAsync(w->stats + ST_STATUS, 0, sizeof(unsigned long long), ctx->stream));   // observed: cleared
    return status_error(ctx, status);
}

extern "C" {

int cz_world_step(cz_world *w, cz_real dt, int32_t n_steps, cz_step_stats *stats) {
    if (!w || n_steps < 0) return fail(nullptr, CZ_ERR_INVALID, "cz_world_step: bad argument");
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    if ((w->d.flags & CZ_WORLD_FUSED) && !w->useFused)
        return fail(ctx, CZ_ERR_INVALID, "CZ_WORLD_FUSED requested but the world does not fit the fused kernel");
    int rc = world_prepare_step(w, dt);
    if (rc) return rc;
    long long launches = 0;
    if (stats) {
        CK(ctx, cudaMemsetAsync(w->stats, 0, sizeof(unsigned long long) * ST_STATUS, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    }
    if (w->useFused) {
        WorldParams p = world_params(w);
        if (w->fused.split && (rc = step_lanes_init(w))) return rc;
        if (w->fused.split && w->lanes.n > 1) {
            // Worlds share nothing, so the batch runs as a few slices, each with its own stream, world counters and group
            // scratch and its own order -> A -> B -> C chain per frame.  A phase launch ends with a tail (the last CTAs
            // finish their worlds while the other SMs idle; the next phase cannot start before it); with two chains the
            // other slices' launches fill it.  65 536 worlds: 1.92 -> 1.72 ms per frame (tools/lanes_probe.py).
            auto &ln = w->lanes;
            const int W = w->d.n_worlds;
            CK(ctx, cudaMemsetAsync(ln.dNext, 0, sizeof(unsigned int) * 4 * ln.n, ctx->stream));
            CK(ctx, cudaEventRecord(ln.evBegin, ctx->stream));
            for (int k = 0; k < ln.n; k++) CK(ctx, cudaStreamWaitEvent(ln.s[k], ln.evBegin, 0));
            for (int s = 0; s < n_steps && !rc; s++) {
                for (int k = 0; k < ln.n && !rc; k++) {
                    WorldParams pk = p;
                    pk.step_index = w->step_index + s;
                    pk.wFirst = (int)((long long)W * k / ln.n);
                    pk.wCount = (int)((long long)W * (k + 1) / ln.n) - pk.wFirst;
                    czf::FusedPlan fpl = w->fused;
                    fpl.cold = ln.cold[k];
                    const bool zeroed = order_worlds(w, pk, ln.s[k], launches, ln.dNext + 4 * k);
                    for (int ph : {czf::PH_A, czf::PH_B, czf::PH_C}) {
                        pk.order = phase_order(w, ph);
                        rc = czf::launch(fpl, pk, dt, w->bias, 1, ln.dNext + 4 * k, ln.s[k], ph, zeroed);
                        if (rc) break;
                        launches++;
                    }
                }
            }
            for (int k = 0; k < ln.n; k++) {
                CK(ctx, cudaEventRecord(ln.evDone[k], ln.s[k]));
                CK(ctx, cudaStreamWaitEvent(ctx->stream, ln.evDone[k], 0));
            }
        } else if (w->fused.split) {   // one launch per phase and frame: every warp of the GPU runs the same code region
            for (int s = 0; s < n_steps && !rc; s++) {
                p.step_index = w->step_index + s;
                const bool zeroed = order_worlds(w, p, ctx->stream, launches, w->d_next);
                for (int ph : {czf::PH_A, czf::PH_B, czf::PH_C}) {
                    p.order = phase_order(w, ph);
                    rc = czf::launch(w->fused, p, dt, w->bias, 1, w->d_next, ctx->stream, ph, zeroed);
                    if (rc) break;
                    launches++;
                }
            }
        } else {
            const bool zeroed = order_worlds(w, p, ctx->stream, launches, w->d_next);
            p.order = phase_order(w, czf::PH_C);
            rc = czf::launch(w->fused, p, dt, w->bias, n_steps, w->d_next, ctx->stream, czf::PH_ALL, zeroed);
            launches++;
        }
        if (rc) return fail(ctx, CZ_ERR_CUDA, std::string("fused launch: ") + cudaGetErrorString((cudaError_t)rc));
        w->step_index += n_steps;
    } else {
        for (int s = 0; s < n_steps; s++) {
            if ((rc = world_step_multi(w, dt, launches))) return rc;
            w->step_index++;
        }
    }
    if (stats) {
        CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CK(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        return read_stats(w, stats, launches, n_steps, ms);
    }
    return CZ_OK;
}

int cz_world_download_bodies(cz_world *w, int32_t first, int32_t n, cz_bodies *out) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    CK(w->ctx, cudaSetDevice(w->ctx->device));
    const long long B = w->d.bodies_per_world;
    if ((rc = take_status(w))) return rc;   // an error raised by an asynchronous step surfaces here
    return download_bodies(w->b, first * B, n * B, out);
}
int cz_world_download_colliders(cz_world *w, int32_t first, int32_t n, cz_colliders *out) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    CK(w->ctx, cudaSetDevice(w->ctx->device));
    const long long B = w->d.bodies_per_world;
    if ((rc = get_field(w->b, F_CTRANSFORM, first * B, n * B, out->transform))) return rc;
    if ((rc = get_field(w->b, F_OFFSET, first * B, n * B, out->offset))) return rc;
    if ((rc = get_field(w->b, F_HALF, first * B, n * B, out->half_size))) return rc;
    if ((rc = get_field(w->b, F_RADIUS, first * B, n * B, out->radius))) return rc;
    if (out->shape) {
        std::vector<uint8_t> sh(n * B);
        if ((rc = get_u8(w->b, w->b.st.shape, first * B, n * B, sh.data()))) return rc;
        for (long long i = 0; i < n * B; i++) out->shape[i] = sh[i];
    }
    return CZ_OK;
}
int cz_world_download_contacts(cz_world *w, int32_t world, cz_contacts *out) {
    int rc = world_range(w, world, 1);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    int nC = 0;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(&nC, w->nContacts + world, sizeof(int), cudaMemcpyDeviceToHost));
    out->n = nC;   // reported even on CZ_ERR_CAPACITY: the caller learns the count it must make room for
    if ((rc = take_status(w))) return rc;
    if (nC > out->capacity || nC > w->d.contacts_per_world) return fail(ctx, CZ_ERR_CAPACITY, "cz_world_download_contacts: capacity too small");
    if (nC == 0) return CZ_OK;
    const size_t Cc = w->d.contacts_per_world, gs = (size_t)w->d.n_worlds * Cc, off = (size_t)world * Cc;
    std::vector<real> tmp(nC);
    auto getf = [&](int f, real *dst, int comp, int ncomp) -> int {
        CK(ctx, cudaMemcpy(tmp.data(), w->gen + f * gs + off, sizeof(real) * nC, cudaMemcpyDeviceToHost));
        for (int i = 0; i < nC; i++) dst[i * ncomp + comp] = tmp[i];
        return CZ_OK;
    };
    for (int k = 0; k < 3; k++) {
        if (out->point && (rc = getf(czr::G_POINT + k, out->point, k, 3))) return rc;
        if (out->normal && (rc = getf(czr::G_NORMAL + k, out->normal, k, 3))) return rc;
    }
    if (out->penetration && (rc = getf(czr::G_PEN, out->penetration, 0, 1))) return rc;
    if (w->nMat > 0) {
        if (out->friction && (rc = getf(czr::G_FRIC, out->friction, 0, 1))) return rc;
        if (out->restitution && (rc = getf(czr::G_REST, out->restitution, 0, 1))) return rc;
    } else {   // the reference's constants (colliders.go:201-202 etc.)
        for (int i = 0; i < nC; i++) {
            if (out->friction) out->friction[i] = (real)0.9;
            if (out->restitution) out->restitution[i] = (real)0.1;
        }
    }
    if (out->body0) CK(ctx, cudaMemcpy(out->body0, w->gb0 + off, sizeof(int) * nC, cudaMemcpyDeviceToHost));
    if (out->body1) CK(ctx, cudaMemcpy(out->body1, w->gb1 + off, sizeof(int) * nC, cudaMemcpyDeviceToHost));
    return CZ_OK;
}
int cz_world_export_gl(cz_world *w, int32_t first, int32_t n, float *location, float *rotation, float *model, int32_t dst_on_device) {
    int rc = world_range(w, first, n);
    if (rc) return rc;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    const long long B = w->d.bodies_per_world, nb = (long long)n * B;
    if (nb == 0 || (!location && !rotation && !model)) return CZ_OK;
    if (dst_on_device) {   // e.g. a mapped GL buffer: written in place, asynchronous on the context stream
        k_export_gl<<<nblk(nb, 256), 256, 0, ctx->stream>>>(w->b.st, first * B, nb, location, rotation, model);
        CKL(ctx);
        return CZ_OK;
    }
    const size_t need = (size_t)nb * 23;
    if (w->exportFloats < need) {
        if (w->d_export) { cudaFree(w->d_export); w->d_export = nullptr; }
        CK(ctx, cudaMalloc(&w->d_export, sizeof(float) * need));
        w->exportFloats = need;
    }
    float *dl = w->d_export, *dr = dl + nb * 3, *dm = dr + nb * 4;
    k_export_gl<<<nblk(nb, 256), 256, 0, ctx->stream>>>(w->b.st, first * B, nb, location ? dl : nullptr, rotation ? dr : nullptr, model ? dm : nullptr);
    CKL(ctx);
    if (location) CK(ctx, cudaMemcpyAsync(location, dl, sizeof(float) * nb * 3, cudaMemcpyDeviceToHost, ctx->stream));
    if (rotation) CK(ctx, cudaMemcpyAsync(rotation, dr, sizeof(float) * nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (model) CK(ctx, cudaMemcpyAsync(model, dm, sizeof(float) * nb * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return CZ_OK;
}
int cz_world_last_step_counts(cz_world *w, int32_t *n_contacts, int32_t *pos_it, int32_t *vel_it) {
    if (!w) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    { int rc = take_status(w); if (rc) return rc; }
    const size_t bytes = sizeof(int) * w->d.n_worlds;
    if (n_contacts) CK(ctx, cudaMemcpy(n_contacts, w->nContacts, bytes, cudaMemcpyDeviceToHost));
    if (pos_it) CK(ctx, cudaMemcpy(pos_it, w->posIters, bytes, cudaMemcpyDeviceToHost));
    if (vel_it) CK(ctx, cudaMemcpy(vel_it, w->velIters, bytes, cudaMemcpyDeviceToHost));
    return CZ_OK;
}
int cz_world_count_nonfinite(cz_world *w, int64_t *bodies) {
    if (!w || !bodies) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d = nullptr, h = 0;
    CK(ctx, cudaMalloc(&d, sizeof(unsigned long long)));
    CK(ctx, cudaMemsetAsync(d, 0, sizeof(unsigned long long), ctx->stream));
    k_count_nonfinite<<<nblk(w->b.n, 256), 256, 0, ctx->stream>>>(w->b.st, w->b.n, d);
    CKL(ctx);
    CK(ctx, cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    *bodies = (int64_t)h;
    return CZ_OK;
}
int cz_world_island_stats(cz_world *w, int64_t *island_frames, int64_t *fallbacks) {
    if (!w) return CZ_ERR_INVALID;
    if (island_frames) *island_frames = w->islandFrames;
    if (fallbacks) *fallbacks = w->islandFallbacks;
    return CZ_OK;
}
int cz_world_checksum_energy(cz_world *w, uint64_t *checksum, double *energy) {
    if (!w) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    { int rc = take_status(w); if (rc) return rc; }
    const int W = w->d.n_worlds;
    unsigned long long *dh = nullptr;
    double *de = nullptr;
    CK(ctx, cudaMalloc(&dh, sizeof(unsigned long long) * W));
    CK(ctx, cudaMalloc(&de, sizeof(double) * W));
    k_checksum_energy<<<nblk(W, 128), 128, 0, ctx->stream>>>(w->b.st, W, w->d.bodies_per_world, dh, de);
    CKL(ctx);
    std::vector<unsigned long long> hh(W);
    std::vector<double> he(W);
    CK(ctx, cudaMemcpyAsync(hh.data(), dh, sizeof(unsigned long long) * W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaMemcpyAsync(he.data(), de, sizeof(double) * W, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(dh);
    cudaFree(de);
    unsigned long long sum = 0;
    double e = 0;
    for (int i = 0; i < W; i++) { sum += hh[i]; e += he[i]; }
    if (checksum) *checksum = sum;
    if (energy) *energy = e;
    return CZ_OK;
}

// Host-buffer step: the end-to-end call of a host-resident caller.  Uploads the primary state
// (what the host may have edited: the K1 read set), runs n_steps, downloads everything
// Integrate/ResolveContacts write.  For worlds on the fused kernel the call is a three-stage
// pipeline over world chunks — H2D copies | pack + frames + unpack | D2H copies — on three
// streams, so PCIe traffic in both directions overlaps the kernels.  Host arrays should be pinned
// (cz_host_alloc); pageable memory works but the copies then serialise in the driver.
// chunk boundaries of the host pipelines: short first and last chunks (pipeline fill = first upload,
// drain = last download)
static std::vector<double> host_chunk_weights(int chunks, bool fromEnv = true) {
    std::vector<double> wt;
    if (const char *e = fromEnv ? getenv("CUBEZ_HOST_CHUNK_WEIGHTS") : nullptr) {   // probe hook: "0.1,0.2,0.5,1,1,1,0.5"
        for (const char *q = e; *q;) {
            char *end = nullptr;
            const double v = strtod(q, &end);
            if (end == q) break;
            if (v > 0) wt.push_back(v);
            q = *end == ',' ? end + 1 : end;
        }
        if (!wt.empty()) return wt;
    }
    wt.assign(chunks, 1.0);
    if (chunks >= 4 && !czf::env_int("CUBEZ_HOST_EVEN_CHUNKS", 0)) { wt[0] = wt[chunks - 1] = 0.5; wt[1] = wt[chunks - 2] = 0.85; }
    return wt;
}
static std::vector<int> host_chunk_edges(int chunks, int W) {
    std::vector<double> wt = host_chunk_weights(chunks);
    if ((int)wt.size() != chunks) wt = host_chunk_weights(chunks, false);   // a caller with its own chunk count (the RL step)
    std::vector<int> wEdge(chunks + 1, 0);
    double tot = 0, run = 0;
    for (double v : wt) tot += v;
    for (int c = 0; c < chunks; c++) { run += wt[c]; wEdge[c + 1] = (int)((double)W * run / tot + 0.5); }
    wEdge[chunks] = W;
    return wEdge;
}

static int host_pipe_init(cz_world *w) {
    cz_ctx *ctx = w->ctx;
    auto &pp = w->pipe;
    if (pp.ready) return CZ_OK;
    const long long NB = w->b.n;
    int chunks = (int)host_chunk_weights(czf::env_int("CUBEZ_HOST_CHUNKS", 6)).size();
    if (!w->useFused) chunks = 1;
    if (chunks > w->d.n_worlds) chunks = w->d.n_worlds;
    if (chunks < 1) chunks = 1;
    pp.chunks = chunks;
    CK(ctx, cudaStreamCreateWithFlags(&pp.sUp, cudaStreamNonBlocking));
    CK(ctx, cudaStreamCreateWithFlags(&pp.sDown, cudaStreamNonBlocking));
    // Chunk c computes on stream c % nComp, earlier streams at higher priority.  Measured (profiles/): 3 streams are
    // enough to overlap kernel tails; one stream per chunk or priorities change nothing, because a chunk's four
    // dependent launches (order, A, B, C) take ~1 ms however small the chunk is (per-world latency), and from the
    // first download on the pipeline is bound by the D2H stream.
    pp.nComp = w->useFused ? std::min(16, std::max(1, czf::env_int("CUBEZ_HOST_COMP_STREAMS", 3))) : 1;
    int prLeast = 0, prGreatest = 0;
    CK(ctx, cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));   // numerically lower = higher priority
    const bool usePrio = czf::env_int("CUBEZ_HOST_PRIO", 1) != 0;
    for (int k = 0; k < pp.nComp; k++) {
        if (!w->useFused) { pp.sComp[k] = ctx->stream; continue; }   // the multi-kernel path steps on the context stream
        const int pr = usePrio ? std::min(prLeast, prGreatest + k) : prLeast;
        CK(ctx, cudaStreamCreateWithPriority(&pp.sComp[k], cudaStreamNonBlocking, pr));
    }
    pp.evUp.resize(chunks); pp.evComp.resize(chunks);
    for (int c = 0; c < chunks; c++) {
        CK(ctx, cudaEventCreateWithFlags(&pp.evUp[c], cudaEventDisableTiming));
        CK(ctx, cudaEventCreateWithFlags(&pp.evComp[c], cudaEventDisableTiming));
    }
    CK(ctx, cudaEventCreateWithFlags(&pp.evBegin, cudaEventDisableTiming));
    CK(ctx, cudaEventCreateWithFlags(&pp.evDownDone, cudaEventDisableTiming));
    CK(ctx, cudaMalloc(&pp.dIn, sizeof(real) * NB * 26));    // pos3 ori4 vel3 rot3 acc3 iitb9 motion1
    CK(ctx, cudaMalloc(&pp.dOut, sizeof(real) * NB * 38));   // pos3 ori4 vel3 rot3 motion1 lacc3 tr12 iitw9
    CK(ctx, cudaMalloc(&pp.dFlags, 4 * (size_t)NB));         // awake in, can_sleep in, awake out (one per RL pipeline slot)
    CK(ctx, cudaMalloc(&pp.dNext, sizeof(unsigned int) * 4 * chunks));
    for (int k = 1; k < pp.nComp; k++) CK(ctx, cudaMalloc(&pp.coldX[k], sizeof(real) * w->fused.coldReals * (size_t)w->fused.maxGrid * w->fused.groupsPerBlock));
    pp.dInSlot[0] = pp.dIn; pp.dOutSlot[0] = pp.dOut;          // slot 1 of the RL pipeline is allocated on first asynchronous use
    for (int k = 0; k < 2; k++) {
        CK(ctx, cudaEventCreateWithFlags(&pp.evSlot[k], cudaEventDisableTiming));
        CK(ctx, cudaEventRecord(pp.evSlot[k], ctx->stream));
    }
    pp.ready = true;
    return CZ_OK;
}

int cz_world_step_host(cz_world *w, cz_bodies *io, cz_real dt, int32_t n_steps, cz_step_stats *stats) {
    if (!w || !io) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    if (io->n != w->b.n) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_host: io->n must equal n_worlds*bodies_per_world");
    if (!io->position || !io->orientation || !io->velocity || !io->rotation || !io->acceleration || !io->inverse_inertia_tensor ||
        !io->motion || !io->is_awake || !io->can_sleep || !io->transform || !io->inverse_inertia_tensor_world || !io->last_frame_acceleration)
        return fail(ctx, CZ_ERR_INVALID, "cz_world_step_host: every state array must be given");
    CK(ctx, cudaSetDevice(ctx->device));
    if ((w->d.flags & CZ_WORLD_FUSED) && !w->useFused)
        return fail(ctx, CZ_ERR_INVALID, "CZ_WORLD_FUSED requested but the world does not fit the fused kernel");
    int rc;
    // damping: only touched (and the Pow factors refreshed) when the host changed it.  The comparison
    // runs per chunk inside the pipeline loop (8 MB of memcmp up front cost ~0.5 ms of idle GPU).
    cz_bodies damp{};
    damp.n = io->n; damp.linear_damping = io->linear_damping; damp.angular_damping = io->angular_damping;
    if ((rc = world_prepare_step(w, dt))) return rc;
    if ((rc = host_pipe_init(w))) return rc;
    auto &pp = w->pipe;
    const long long NB = w->b.n, B = w->d.bodies_per_world;
    const int W = w->d.n_worlds;
    // device staging layout
    real *dPos = pp.dIn, *dOri = dPos + NB * 3, *dVel = dOri + NB * 4, *dRot = dVel + NB * 3, *dAcc = dRot + NB * 3, *dIitb = dAcc + NB * 3, *dMot = dIitb + NB * 9;
    real *oPos = pp.dOut, *oOri = oPos + NB * 3, *oVel = oOri + NB * 4, *oRot = oVel + NB * 3, *oMot = oRot + NB * 3, *oLacc = oMot + NB, *oTr = oLacc + NB * 3, *oIitw = oTr + NB * 12;
    uint8_t *fAwake = pp.dFlags, *fSleep = fAwake + NB, *fAwakeOut = fSleep + NB;
    HostIn hin{dPos, dOri, dVel, dRot, dAcc, dIitb, dMot, fAwake, fSleep};
    HostOut hout{oPos, oOri, oVel, oRot, oMot, oLacc, oTr, oIitw, fAwakeOut};
    static const bool hostTrace = getenv("CUBEZ_HOST_TRACE") != nullptr;
    const auto tc0 = std::chrono::steady_clock::now();
    CK(ctx, cudaMemsetAsync(w->stats, 0, sizeof(unsigned long long) * ST_STATUS, ctx->stream));
    CK(ctx, cudaMemsetAsync(pp.dNext, 0, sizeof(unsigned int) * 4 * pp.chunks, ctx->stream));
    CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    CK(ctx, cudaEventRecord(pp.evBegin, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(pp.sUp, pp.evBegin, 0));     // staging buffers of the previous call are free
    CK(ctx, cudaStreamWaitEvent(pp.sDown, pp.evBegin, 0));
    for (int k = 0; k < pp.nComp; k++) CK(ctx, cudaStreamWaitEvent(pp.sComp[k], pp.evBegin, 0));
    long long launches = 0;
    std::vector<cudaEvent_t> trEv;   // trace only: [chunk][up end, compute begin, compute end, down end]
    if (hostTrace) { trEv.resize((size_t)pp.chunks * 4); for (auto &e : trEv) cudaEventCreate(&e); }
    // every array of the call page-locked (cz_host_alloc / cudaHostRegister)?  Asked once per call, on the first arrays
    // of each kind; a caller mixing pinned and pageable arrays gets what plain cudaMemcpyAsync gives for pageable memory.
    const bool pinned = host_ptr_pinned(io->position) && host_ptr_pinned(io->transform) && host_ptr_pinned(io->is_awake);
    const std::vector<int> wEdge = host_chunk_edges(pp.chunks, W);
    for (int c = 0; c < pp.chunks; c++) {
        const int w0 = wEdge[c], w1 = wEdge[c + 1];
        const long long b0 = w0 * B, nb = (w1 - w0) * B;
        if (nb <= 0) continue;
        CopyBatch up;
        up.add(dPos, io->position, b0, nb, 3); up.add(dOri, io->orientation, b0, nb, 4); up.add(dVel, io->velocity, b0, nb, 3); up.add(dRot, io->rotation, b0, nb, 3);
        up.add(dAcc, io->acceleration, b0, nb, 3); up.add(dIitb, io->inverse_inertia_tensor, b0, nb, 9); up.add(dMot, io->motion, b0, nb, 1);
        up.add(fAwake, io->is_awake, b0, nb, 1); up.add(fSleep, io->can_sleep, b0, nb, 1);
        CK(ctx, copy_batch_submit(up, cudaMemcpyHostToDevice, pp.sUp, pinned));
        CK(ctx, cudaEventRecord(pp.evUp[c], pp.sUp));
        if (hostTrace) cudaEventRecord(trEv[c * 4 + 0], pp.sUp);
        // this chunk's damping slice, compared while its upload is in flight
        if ((io->linear_damping && std::memcmp(w->b.h_lind.data() + b0, io->linear_damping + b0, sizeof(real) * nb) != 0) ||
            (io->angular_damping && std::memcmp(w->b.h_angd.data() + b0, io->angular_damping + b0, sizeof(real) * nb) != 0)) {
            // rare: the host changed damping.  Drain the compute streams, refresh every Pow factor, go on.
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            for (int k = 0; k < pp.nComp; k++) CK(ctx, cudaStreamSynchronize(pp.sComp[k]));
            if ((rc = upload_bodies(w->b, 0, io->n, &damp))) return rc;
            if ((rc = world_prepare_step(w, dt))) return rc;
            CK(ctx, cudaStreamSynchronize(ctx->stream));
        }
        cudaStream_t cs = pp.sComp[c % pp.nComp];
        CK(ctx, cudaStreamWaitEvent(cs, pp.evUp[c], 0));
        if (hostTrace) cudaEventRecord(trEv[c * 4 + 1], cs);
        k_pack_all<<<nblk(nb, 256), 256, 0, cs>>>(w->b.st, b0, nb, hin);
        CKL(ctx);
        launches++;
        if (w->useFused) {
            WorldParams p = world_params(w);
            p.wFirst = w0; p.wCount = w1 - w0;
            czf::FusedPlan fpl = w->fused;
            if (c % pp.nComp) fpl.cold = pp.coldX[c % pp.nComp];
            static const int splitMin = czf::env_int("CUBEZ_HOST_SPLIT_MIN", 0);   // chunks below this many worlds: one persistent launch
            if (w->fused.split && !czf::env_int("CUBEZ_HOST_NO_SPLIT", 0) && w1 - w0 >= splitMin) {   // one launch per phase and frame, as cz_world_step does for large batches
                for (int s2 = 0; s2 < n_steps && !rc; s2++) {
                    p.step_index = w->step_index + s2;
                    const bool zeroed = order_worlds(w, p, cs, launches, pp.dNext + 4 * c);
                    for (int ph : {czf::PH_A, czf::PH_B, czf::PH_C}) {
                        p.order = phase_order(w, ph);
                        rc = czf::launch(fpl, p, dt, w->bias, 1, pp.dNext + 4 * c, cs, ph, zeroed);
                        if (rc) break;
                        launches++;
                    }
                }
            } else {
                const bool zeroed = order_worlds(w, p, cs, launches, pp.dNext + 4 * c);
                p.order = phase_order(w, czf::PH_C);
                rc = czf::launch(fpl, p, dt, w->bias, n_steps, pp.dNext + 4 * c, cs, czf::PH_ALL, zeroed);
                launches++;
            }
            if (rc) return fail(ctx, CZ_ERR_CUDA, std::string("fused launch: ") + cudaGetErrorString((cudaError_t)rc));
        } else {
            const long long keep = w->step_index;
            for (int s = 0; s < n_steps; s++) {
                if ((rc = world_step_multi(w, dt, launches))) return rc;
                w->step_index++;
            }
            w->step_index = keep;
        }
        k_unpack_all<<<nblk(nb, 256), 256, 0, cs>>>(w->b.st, b0, nb, hout);
        CKL(ctx);
        launches++;
        CK(ctx, cudaEventRecord(pp.evComp[c], cs));
        if (hostTrace) cudaEventRecord(trEv[c * 4 + 2], cs);
        CK(ctx, cudaStreamWaitEvent(pp.sDown, pp.evComp[c], 0));
        CopyBatch down;
        down.add(io->position, oPos, b0, nb, 3); down.add(io->orientation, oOri, b0, nb, 4); down.add(io->velocity, oVel, b0, nb, 3); down.add(io->rotation, oRot, b0, nb, 3);
        down.add(io->motion, oMot, b0, nb, 1); down.add(io->last_frame_acceleration, oLacc, b0, nb, 3); down.add(io->transform, oTr, b0, nb, 12);
        down.add(io->inverse_inertia_tensor_world, oIitw, b0, nb, 9); down.add(io->is_awake, fAwakeOut, b0, nb, 1);
        CK(ctx, copy_batch_submit(down, cudaMemcpyDeviceToHost, pp.sDown, pinned));
        if (hostTrace) cudaEventRecord(trEv[c * 4 + 3], pp.sDown);
    }
    w->step_index += n_steps;
    CK(ctx, cudaEventRecord(pp.evDownDone, pp.sDown));
    CK(ctx, cudaStreamWaitEvent(ctx->stream, pp.evDownDone, 0));
    CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    const auto tc1 = std::chrono::steady_clock::now();
    CK(ctx, cudaEventSynchronize(ctx->ev1));
    const auto tc2 = std::chrono::steady_clock::now();
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    rc = read_stats(w, stats, launches, n_steps, ms);
    if (hostTrace) {
        const auto tc3 = std::chrono::steady_clock::now();
        auto us = [](auto a, auto b) { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count() * 1e-3; };
        fprintf(stderr, "[step_host] enqueue %.0f us | wait %.0f us | stats %.0f us | device events %.0f us\n", us(tc0, tc1), us(tc1, tc2), us(tc2, tc3), ms * 1e3);
        for (int c = 0; c < pp.chunks; c++) {
            float t[4] = {0, 0, 0, 0};
            for (int k = 0; k < 4; k++) cudaEventElapsedTime(&t[k], ctx->ev0, trEv[c * 4 + k]);
            fprintf(stderr, "   chunk %d (%d worlds): up done %.2f | compute %.2f..%.2f | down done %.2f ms\n", c, wEdge[c + 1] - wEdge[c], t[0], t[1], t[2], t[3]);
        }
        for (auto &e : trEv) cudaEventDestroy(e);
    }
    return rc;
}


}  // extern "C"

// float32 observations of the RL step, converted on the device (half the D2H bytes of the Real arrays)
__global__ void k_unpack_obs32(czb::BodyStore s, long long first, long long n, float *pos, float *ori, float *vel, float *rot) {
    using namespace czb;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long i = first + t;
    if (pos) { V3 v = ld_position(s, i); for (int k = 0; k < 3; k++) pos[i * 3 + k] = (float)v.c[k]; }
    if (ori) { Q4 q = ld_orientation(s, i); for (int k = 0; k < 4; k++) ori[i * 4 + k] = (float)q.c[k]; }
    if (vel) { V3 v = ld_velocity(s, i); for (int k = 0; k < 3; k++) vel[i * 3 + k] = (float)v.c[k]; }
    if (rot) { V3 v = ld_rotation(s, i); for (int k = 0; k < 3; k++) rot[i * 3 + k] = (float)v.c[k]; }
}

// Enqueue one RL step of a fused world on pipeline slot `slot` (0 / 1): actions H2D | apply + frames + unpack | observations
// D2H, chunked over the world range.  Nothing here waits for the device.  Slot s reuses the staging buffers of the call
// two tickets earlier, ordered behind that call's last download by evSlot[s].
static int rl_enqueue(cz_world *w, const cz_real *add_velocity, const cz_real *add_rotation, cz_bodies *obs, cz_obs32 *obs32, cz_real dt,
                      int32_t n_steps, int slot, long long &launches, bool afterContextStream, bool contextStreamWaits, int chunkLimit = 0) {
    cz_ctx *ctx = w->ctx;
    auto &pp = w->pipe;
    int rc = CZ_OK;
    const long long NB = w->b.n, B = w->d.bodies_per_world;
    const int W = w->d.n_worlds;
    real *dVel = pp.dInSlot[slot], *dRot = dVel + NB * 3;
    real *oPos = pp.dOutSlot[slot], *oOri = oPos + NB * 3, *oVel = oOri + NB * 4, *oRot = oVel + NB * 3, *oMot = oRot + NB * 3, *oLacc = oMot + NB, *oTr = oLacc + NB * 3, *oIitw = oTr + NB * 12;
    uint8_t *fAwakeOut = pp.dFlags + (2 + slot) * NB;
    float *fPos = pp.dObs32[slot], *fOri = fPos ? fPos + NB * 3 : nullptr, *fVel = fPos ? fOri + NB * 4 : nullptr, *fRot = fPos ? fVel + NB * 3 : nullptr;
    HostOut hout{};
    if (obs) {
        hout.pos = obs->position ? oPos : nullptr; hout.ori = obs->orientation ? oOri : nullptr; hout.vel = obs->velocity ? oVel : nullptr;
        hout.rot = obs->rotation ? oRot : nullptr; hout.motion = obs->motion ? oMot : nullptr; hout.lacc = obs->last_frame_acceleration ? oLacc : nullptr;
        hout.tr = obs->transform ? oTr : nullptr; hout.iitw = obs->inverse_inertia_tensor_world ? oIitw : nullptr; hout.awake = obs->is_awake ? fAwakeOut : nullptr;
    }
    const bool anyOut = hout.pos || hout.ori || hout.vel || hout.rot || hout.motion || hout.lacc || hout.tr || hout.iitw || hout.awake;
    const bool any32 = obs32 && (obs32->position || obs32->orientation || obs32->velocity || obs32->rotation);
    const bool anyIn = add_velocity || add_rotation;
    if (afterContextStream) {   // order the pipeline behind whatever the caller queued on the context stream (uploads, earlier steps)
        CK(ctx, cudaEventRecord(pp.evBegin, ctx->stream));
        CK(ctx, cudaStreamWaitEvent(pp.sUp, pp.evBegin, 0));
        CK(ctx, cudaStreamWaitEvent(pp.sDown, pp.evBegin, 0));
        for (int k = 0; k < pp.nComp; k++) CK(ctx, cudaStreamWaitEvent(pp.sComp[k], pp.evBegin, 0));
    }
    CK(ctx, cudaStreamWaitEvent(pp.sUp, pp.evSlot[slot], 0));          // the staging of this slot is free again
    for (int k = 0; k < pp.nComp; k++) CK(ctx, cudaStreamWaitEvent(pp.sComp[k], pp.evSlot[slot], 0));
    // observations are a third of the full state: fewer, larger chunks keep the fused kernels efficient
    // (a step enqueued behind another one does not need chunks to hide its transfers — they overlap the neighbour's
    // frames — so it runs as the resident step does, in two slices whose launches fill each other's tails: chunkLimit;
    // 2.28 -> 1.84 ms per step at 65 536 worlds, tools/rl_async_probe.py)
    const int chunks = std::max(1, std::min(pp.chunks, chunkLimit > 0 ? chunkLimit : czf::env_int("CUBEZ_RL_CHUNKS", 4)));
    const std::vector<int> wEdge = host_chunk_edges(chunks, W);
    // page-locked arrays go out as one batched copy per chunk and direction (see CopyBatch)
    const void *firstOut = obs ? (obs->position ? (const void *)obs->position : obs->orientation ? (const void *)obs->orientation : obs->velocity ? (const void *)obs->velocity : (const void *)obs->rotation) : nullptr;
    const void *first32 = obs32 ? (obs32->position ? (const void *)obs32->position : obs32->orientation ? (const void *)obs32->orientation : obs32->velocity ? (const void *)obs32->velocity : (const void *)obs32->rotation) : nullptr;
    const bool pinnedIn = (!add_velocity || host_ptr_pinned(add_velocity)) && (!add_rotation || host_ptr_pinned(add_rotation));
    const bool pinnedOut = (!firstOut || host_ptr_pinned(firstOut)) && (!first32 || host_ptr_pinned(first32)) && (!obs || !obs->is_awake || host_ptr_pinned(obs->is_awake));
    for (int c = 0; c < chunks; c++) {
        const int w0 = wEdge[c], w1 = wEdge[c + 1];
        const long long b0 = w0 * B, nb = (w1 - w0) * B;
        if (nb <= 0) continue;
        cudaStream_t cs = pp.sComp[c % pp.nComp];
        if (anyIn) {
            CopyBatch up;
            up.add(dVel, add_velocity, b0, nb, 3); up.add(dRot, add_rotation, b0, nb, 3);
            CK(ctx, copy_batch_submit(up, cudaMemcpyHostToDevice, pp.sUp, pinnedIn));
            CK(ctx, cudaEventRecord(pp.evUp[c], pp.sUp));
            CK(ctx, cudaStreamWaitEvent(cs, pp.evUp[c], 0));
            k_apply_actions<<<nblk(nb, 256), 256, 0, cs>>>(w->b.st, b0, nb, add_velocity ? dVel : nullptr, add_rotation ? dRot : nullptr);
            CKL(ctx);
            launches++;
        }
        WorldParams p = world_params(w);
        p.wFirst = w0; p.wCount = w1 - w0;
        czf::FusedPlan fpl = w->fused;
        if (c % pp.nComp) fpl.cold = pp.coldX[c % pp.nComp];
        if (w->fused.split) {
            for (int s2 = 0; s2 < n_steps && !rc; s2++) {
                p.step_index = w->step_index + s2;
                const bool zeroed = order_worlds(w, p, cs, launches, pp.dNext + 4 * c);
                for (int ph : {czf::PH_A, czf::PH_B, czf::PH_C}) {
                    p.order = phase_order(w, ph);
                    rc = czf::launch(fpl, p, dt, w->bias, 1, pp.dNext + 4 * c, cs, ph, zeroed);
                    if (rc) break;
                    launches++;
                }
            }
        } else if (n_steps > 0) {
            const bool zeroed = order_worlds(w, p, cs, launches, pp.dNext + 4 * c);
            p.order = phase_order(w, czf::PH_C);
            rc = czf::launch(fpl, p, dt, w->bias, n_steps, pp.dNext + 4 * c, cs, czf::PH_ALL, zeroed);
            launches++;
        }
        if (rc) return fail(ctx, CZ_ERR_CUDA, std::string("fused launch: ") + cudaGetErrorString((cudaError_t)rc));
        if (anyOut) {
            k_unpack_all<<<nblk(nb, 256), 256, 0, cs>>>(w->b.st, b0, nb, hout);
            CKL(ctx);
            launches++;
        }
        if (any32) {
            k_unpack_obs32<<<nblk(nb, 256), 256, 0, cs>>>(w->b.st, b0, nb, obs32->position ? fPos : nullptr, obs32->orientation ? fOri : nullptr,
                                                          obs32->velocity ? fVel : nullptr, obs32->rotation ? fRot : nullptr);
            CKL(ctx);
            launches++;
        }
        CK(ctx, cudaEventRecord(pp.evComp[c], cs));
        CK(ctx, cudaStreamWaitEvent(pp.sDown, pp.evComp[c], 0));
        CopyBatch down;   // add() skips the arrays the caller did not ask for (NULL)
        if (anyOut) {
            down.add(obs->position, oPos, b0, nb, 3); down.add(obs->orientation, oOri, b0, nb, 4); down.add(obs->velocity, oVel, b0, nb, 3); down.add(obs->rotation, oRot, b0, nb, 3);
            down.add(obs->motion, oMot, b0, nb, 1); down.add(obs->last_frame_acceleration, oLacc, b0, nb, 3); down.add(obs->transform, oTr, b0, nb, 12);
            down.add(obs->inverse_inertia_tensor_world, oIitw, b0, nb, 9); down.add(obs->is_awake, fAwakeOut, b0, nb, 1);
        }
        if (any32) {
            down.add(obs32->position, fPos, b0, nb, 3); down.add(obs32->orientation, fOri, b0, nb, 4); down.add(obs32->velocity, fVel, b0, nb, 3); down.add(obs32->rotation, fRot, b0, nb, 3);
        }
        CK(ctx, copy_batch_submit(down, cudaMemcpyDeviceToHost, pp.sDown, pinnedOut));
    }
    w->step_index += n_steps;
    CK(ctx, cudaEventRecord(pp.evSlot[slot], pp.sDown));     // everything of this call — frames and downloads — is behind this event
    if (contextStreamWaits) CK(ctx, cudaStreamWaitEvent(ctx->stream, pp.evSlot[slot], 0));
    return CZ_OK;
}

extern "C" {

int cz_world_step_rl(cz_world *w, const cz_real *add_velocity, const cz_real *add_rotation, cz_bodies *obs, cz_real dt, int32_t n_steps,
                     cz_step_stats *stats) {
    if (!w || n_steps < 0) return fail(nullptr, CZ_ERR_INVALID, "cz_world_step_rl: bad argument");
    cz_ctx *ctx = w->ctx;
    if (obs && obs->n != w->b.n) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_rl: obs->n must equal n_worlds*bodies_per_world");
    CK(ctx, cudaSetDevice(ctx->device));
    if (!w->useFused) {
        // any world shape: actions, resident step, observations (no chunk pipeline: the multi-kernel path steps the whole batch)
        const long long NB0 = w->b.n;
        int rc0;
        if (add_velocity || add_rotation) {
            if ((rc0 = host_pipe_init(w))) return rc0;
            real *dV = w->pipe.dIn, *dR = dV + NB0 * 3;
            if (add_velocity) CK(ctx, cudaMemcpyAsync(dV, add_velocity, sizeof(real) * NB0 * 3, cudaMemcpyHostToDevice, ctx->stream));
            if (add_rotation) CK(ctx, cudaMemcpyAsync(dR, add_rotation, sizeof(real) * NB0 * 3, cudaMemcpyHostToDevice, ctx->stream));
            k_apply_actions<<<nblk(NB0, 256), 256, 0, ctx->stream>>>(w->b.st, 0, NB0, add_velocity ? dV : nullptr, add_rotation ? dR : nullptr);
            CKL(ctx);
        }
        if ((rc0 = cz_world_step(w, dt, n_steps, stats))) return rc0;
        if (obs) {
            cz_bodies o{};
            o.n = obs->n; o.position = obs->position; o.orientation = obs->orientation; o.velocity = obs->velocity; o.rotation = obs->rotation;
            o.motion = obs->motion; o.is_awake = obs->is_awake; o.transform = obs->transform;
            o.inverse_inertia_tensor_world = obs->inverse_inertia_tensor_world; o.last_frame_acceleration = obs->last_frame_acceleration;
            return download_bodies(w->b, 0, NB0, &o);
        }
        return CZ_OK;
    }
    int rc;
    if ((rc = world_prepare_step(w, dt))) return rc;
    if ((rc = host_pipe_init(w))) return rc;
    auto &pp = w->pipe;
    if (pp.inFlight) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_rl: asynchronous RL steps are in flight (cz_world_rl_wait first)");
    CK(ctx, cudaMemsetAsync(w->stats, 0, sizeof(unsigned long long) * ST_STATUS, ctx->stream));
    CK(ctx, cudaMemsetAsync(pp.dNext, 0, sizeof(unsigned int) * 4 * pp.chunks, ctx->stream));
    CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    long long launches = 0;
    if ((rc = rl_enqueue(w, add_velocity, add_rotation, obs, nullptr, dt, n_steps, 0, launches, true, true))) return rc;
    CK(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CK(ctx, cudaEventSynchronize(ctx->ev1));
    float ms = 0;
    CK(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    return read_stats(w, stats, launches, n_steps, ms);
}

int cz_world_step_rl_async(cz_world *w, const cz_real *add_velocity, const cz_real *add_rotation, cz_bodies *obs, cz_obs32 *obs32, cz_real dt,
                           int32_t n_steps, int32_t *ticket) {
    if (!w || n_steps < 0 || !ticket) return fail(nullptr, CZ_ERR_INVALID, "cz_world_step_rl_async: bad argument");
    cz_ctx *ctx = w->ctx;
    if ((obs && obs->n != w->b.n) || (obs32 && obs32->n != w->b.n)) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_rl_async: obs->n must equal n_worlds*bodies_per_world");
    CK(ctx, cudaSetDevice(ctx->device));
    if (!w->useFused) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_rl_async needs a world on the fused small-world kernel (use cz_world_step_rl)");
    int rc;
    if ((rc = world_prepare_step(w, dt))) return rc;
    if ((rc = host_pipe_init(w))) return rc;
    auto &pp = w->pipe;
    if (pp.inFlight >= 2) return fail(ctx, CZ_ERR_INVALID, "cz_world_step_rl_async: two steps are already in flight (cz_world_rl_wait on the older ticket first)");
    if (!pp.dInSlot[1]) {
        CK(ctx, cudaMalloc(&pp.dInSlot[1], sizeof(real) * w->b.n * 6));
        CK(ctx, cudaMalloc(&pp.dOutSlot[1], sizeof(real) * w->b.n * 38));
    }
    if (obs32 && !pp.dObs32[0]) {
        for (int k = 0; k < 2; k++) CK(ctx, cudaMalloc(&pp.dObs32[k], sizeof(float) * 13 * (size_t)w->b.n));
    }
    if (pp.inFlight == 0) {
        CK(ctx, cudaMemsetAsync(pp.dNext, 0, sizeof(unsigned int) * 4 * pp.chunks, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    }
    const int t = pp.nextTicket++;
    long long launches = 0;
    const int behind = pp.inFlight > 0 ? std::max(1, czf::env_int("CUBEZ_RL_ASYNC_CHUNKS", 2)) : 0;
    if ((rc = rl_enqueue(w, add_velocity, add_rotation, obs, obs32, dt, n_steps, t & 1, launches, pp.inFlight == 0, false, behind))) return rc;
    pp.launchesInFlight += launches;
    pp.stepsInFlight += n_steps;
    pp.inFlight++;
    *ticket = t;
    return CZ_OK;
}

int cz_world_rl_wait(cz_world *w, int32_t ticket, cz_step_stats *stats) {
    if (!w) return CZ_ERR_INVALID;
    cz_ctx *ctx = w->ctx;
    CK(ctx, cudaSetDevice(ctx->device));
    auto &pp = w->pipe;
    if (!pp.ready || pp.inFlight <= 0 || ticket < pp.nextTicket - pp.inFlight || ticket >= pp.nextTicket)
        return fail(ctx, CZ_ERR_INVALID, "cz_world_rl_wait: no such step in flight");
    if (ticket != pp.nextTicket - pp.inFlight) return fail(ctx, CZ_ERR_INVALID, "cz_world_rl_wait: wait on the older ticket first");
    CK(ctx, cudaEventSynchronize(pp.evSlot[ticket & 1]));
    pp.inFlight--;
    if (pp.inFlight == 0) CK(ctx, cudaStreamWaitEvent(ctx->stream, pp.evSlot[ticket & 1], 0));   // later calls on the context stream come after the pipeline
    if (!stats) return CZ_OK;
    // counters are cumulative since the last wait with stats (frames of a younger step in flight may already be in them)
    CK(ctx, cudaMemcpy(w->h_stats, w->stats, sizeof(unsigned long long) * ST_N, cudaMemcpyDeviceToHost));
    std::memset(stats, 0, sizeof(*stats));
    stats->steps = (int64_t)w->d.n_worlds * pp.stepsInFlight;
    stats->contacts = (int64_t)w->h_stats[ST_CONTACTS];
    stats->pos_iterations = (int64_t)w->h_stats[ST_POS];
    stats->vel_iterations = (int64_t)w->h_stats[ST_VEL];
    stats->kernel_launches = pp.launchesInFlight;
    stats->max_contacts = (int32_t)w->h_stats[ST_MAXC];
    stats->status = -(int)w->h_stats[ST_STATUS];
    if (pp.inFlight == 0) {   // quiescent: reset the cumulative counters
        CK(ctx, cudaMemset(w->stats, 0, sizeof(unsigned long long) * ST_N));
        pp.launchesInFlight = 0;
        pp.stepsInFlight = 0;
    }
    return status_error(ctx, stats->status);
}

// ---- object-API shims ------------------------------------------------------------------------
int cz_integrate(cz_ctx *ctx, cz_bodies *io, cz_real dt, const cz_real *lin_pow, const cz_real *ang_pow, const cz_real *bias) {
    if (!ctx || !io || io->n <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_integrate: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    Batch b;
    int rc = batch_alloc(ctx, b, io->n);
    if (!rc) rc = upload_bodies(b, 0, io->n, io);
    if (!rc) {
        if (lin_pow && ang_pow) {
            rc = put_field(b, F_LINPOW, 0, io->n, lin_pow);
            if (!rc) rc = put_field(b, F_ANGPOW, 0, io->n, ang_pow);
        } else {
            rc = refresh_pow(b, 0, io->n, dt);
        }
    }
    if (!rc) {
        real bs = bias ? *bias : (real)std::pow(0.5, (double)dt);
        int grid = (int)std::min<long long>(nblk(io->n, 256), (long long)ctx->sm_count * 8);
        k_integrate<false><<<grid, 256, 0, ctx->stream>>>(b.st, dt, bs, 0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
    }
    if (!rc) rc = download_bodies(b, 0, io->n, io);
    batch_free(b);
    return rc;
}

int cz_calculate_derived_data(cz_ctx *ctx, cz_bodies *io) {
    if (!ctx || !io || io->n <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_calculate_derived_data: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    Batch b;
    int rc = batch_alloc(ctx, b, io->n);
    if (!rc) rc = upload_bodies(b, 0, io->n, io);
    if (!rc) {
        k_derive<<<nblk(io->n, 256), 256, 0, ctx->stream>>>(b.st, 0, io->n, 1, 0);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
    }
    if (!rc) rc = download_bodies(b, 0, io->n, io);
    batch_free(b);
    return rc;
}

int cz_collider_derive(cz_ctx *ctx, int32_t n, const cz_real *body_transform, const cz_real *offset, cz_real *out) {
    if (!ctx || n <= 0 || !body_transform || !offset || !out) return fail(ctx, CZ_ERR_INVALID, "cz_collider_derive: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    Batch b;
    int rc = batch_alloc(ctx, b, n);
    if (!rc) rc = put_field(b, F_TRANSFORM, 0, n, body_transform);
    if (!rc) rc = put_field(b, F_OFFSET, 0, n, offset);
    if (!rc) {
        k_fill_u8<<<nblk(n, 256), 256, 0, ctx->stream>>>(b.st.shape, n, CZ_SHAPE_CUBE);
        k_fill_u8<<<nblk(n, 256), 256, 0, ctx->stream>>>(b.st.ident, n, 0);
        k_derive<<<nblk(n, 256), 256, 0, ctx->stream>>>(b.st, 0, n, 0, 1);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
    }
    if (!rc) rc = get_field(b, F_CTRANSFORM, 0, n, out);
    batch_free(b);
    return rc;
}

int cz_narrowphase(cz_ctx *ctx, const cz_colliders *colliders, const cz_planes *planes, const cz_bodies *bodies,
                   int32_t n_checks, const int32_t *one, const int32_t *two, cz_contacts *out, uint8_t *found) {
    if (!ctx || !colliders || !out || n_checks < 0 || colliders->n <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_narrowphase: bad argument");
    // A one-world handle whose "bodies" are the colliders; contact body ids are mapped back
    // through colliders->body on the way out.
    cz_world_desc d{1, colliders->n, std::max(1, out->capacity), CZ_SCHED_EXPLICIT, CZ_WORLD_NO_FUSED};
    cz_world *w = nullptr;
    int rc = cz_world_create(ctx, &d, &w);
    if (rc) return rc;
    const int n = colliders->n;
    do {
        if (planes && planes->n > 0 && (rc = cz_world_upload_planes(w, planes))) break;
        if ((rc = upload_colliders(w->b, 0, n, colliders))) break;
        if (bodies && bodies->velocity) {   // sphere.Body.Velocity for colliders.go:417-421
            std::vector<real> vel((size_t)n * 3, 0);
            for (int i = 0; i < n; i++) {
                int bi = colliders->body ? colliders->body[i] : i;
                if (bi >= 0 && bi < bodies->n) for (int k = 0; k < 3; k++) vel[i * 3 + k] = bodies->velocity[bi * 3 + k];
            }
            if ((rc = put_field(w->b, F_VEL, 0, n, vel.data()))) break;
        }
        if ((rc = cz_world_upload_schedule(w, n_checks, one, two))) break;
        WorldParams p = world_params(w);
        if (n_checks > 0) {
            // always the COUNT/EMIT pair here so that per-check hit counts are available
            if (!w->hitCount) {
                cudaMalloc(&w->tileCount, sizeof(int) * w->tiles);
                cudaMalloc(&w->tileBase, sizeof(int) * w->tiles);
                cudaMalloc(&w->hitCount, n_checks);
            }
            k_narrow<NARROW_COUNT><<<w->tiles, w->tileThreads, 0, ctx->stream>>>(p, w->tiles, w->tileCount, nullptr, w->hitCount);
            k_scan_tiles<<<1, 256, 0, ctx->stream>>>(p, w->tiles, w->tileCount, w->tileBase);
            k_narrow<NARROW_EMIT><<<w->tiles, w->tileThreads, 0, ctx->stream>>>(p, w->tiles, nullptr, w->tileBase, w->hitCount);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
            if (found) {
                cudaStreamSynchronize(ctx->stream);
                cudaMemcpy(found, w->hitCount, n_checks, cudaMemcpyDeviceToHost);
                for (int i = 0; i < n_checks; i++) found[i] = found[i] ? 1 : 0;
            }
        } else {
            cudaMemsetAsync(w->nContacts, 0, sizeof(int), ctx->stream);
        }
        rc = cz_world_download_contacts(w, 0, out);
        if (rc) break;
        if (colliders->body) {
            for (int i = 0; i < out->n; i++) {
                if (out->body0 && out->body0[i] >= 0) out->body0[i] = colliders->body[out->body0[i]];
                if (out->body1 && out->body1[i] >= 0) out->body1[i] = colliders->body[out->body1[i]];
            }
        }
    } while (0);
    cz_world_destroy(w);
    return rc;
}

int cz_resolve_contacts(cz_ctx *ctx, int32_t max_iterations, cz_contacts *io, cz_bodies *bodies, cz_real dt, int32_t *iters) {
    if (!ctx || !io || !bodies || bodies->n <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_resolve_contacts: bad argument");
    if (iters) iters[0] = iters[1] = 0;
    if (!(dt > 0) || io->n <= 0) return CZ_OK;   // contact.go:210-212
    if (!io->body0 || !io->body1 || !io->point || !io->normal || !io->penetration)
        return fail(ctx, CZ_ERR_INVALID, "cz_resolve_contacts: body0, body1, point, normal and penetration must be given");
    for (int i = 0; i < io->n; i++) {
        const int a = io->body0[i], b = io->body1[i];
        if (a >= bodies->n || b >= bodies->n || a < -1 || b < -1) return fail(ctx, CZ_ERR_INVALID, "cz_resolve_contacts: contact body index out of range");
        if (a < 0 && b < 0) return fail(ctx, CZ_ERR_NIL_BODY, "cz_resolve_contacts: contact with two nil bodies (the reference dereferences nil, contact.go:66-70)");
    }
    cz_world_desc d{1, bodies->n, io->n, CZ_SCHED_EXPLICIT, CZ_WORLD_NO_FUSED};
    cz_world *w = nullptr;
    int rc = cz_world_create(ctx, &d, &w);
    if (rc) return rc;
    const int nC = io->n;
    do {
        if ((rc = upload_bodies(w->b, 0, bodies->n, bodies))) break;
        // contacts -> as-generated arrays
        const size_t gs = (size_t)nC;
        std::vector<real> g((size_t)czr::G_NF * nC);
        for (int i = 0; i < nC; i++) {
            for (int k = 0; k < 3; k++) { g[(czr::G_POINT + k) * gs + i] = io->point[i * 3 + k]; g[(czr::G_NORMAL + k) * gs + i] = io->normal[i * 3 + k]; }
            g[czr::G_PEN * gs + i] = io->penetration[i];
            g[czr::G_FRIC * gs + i] = io->friction ? io->friction[i] : (real)0.9;
            g[czr::G_REST * gs + i] = io->restitution ? io->restitution[i] : (real)0.1;
        }
        cudaMemcpy(w->gen, g.data(), sizeof(real) * g.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(w->gb0, io->body0, sizeof(int) * nC, cudaMemcpyHostToDevice);
        cudaMemcpy(w->gb1, io->body1, sizeof(int) * nC, cudaMemcpyHostToDevice);
        cudaMemcpy(w->nContacts, &nC, sizeof(int), cudaMemcpyHostToDevice);
        if (!w->rs.bw) {
            cudaMalloc(&w->rs.bw, sizeof(real) * czr::BW_NF * bodies->n);
            cudaMalloc(&w->rs.cw, sizeof(real) * CW_NREAL * nC);
            cudaMalloc(&w->rs.cb, sizeof(int) * 2 * nC);
        }
        WorldParams p = world_params(w);
        if (w->resolveNT == 32) launch_resolve<32>(w, p, max_iterations, dt, true);
        else launch_resolve<256>(w, p, max_iterations, dt, true);
        k_contacts_writeback<<<nblk(nC, 128), 128, 0, ctx->stream>>>(p, w->rs);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        cz_step_stats st;
        rc = read_stats(w, &st, 2, 1, 0.f);
        if (rc) break;
        if (iters) {
            cudaMemcpy(&iters[0], w->posIters, sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(&iters[1], w->velIters, sizeof(int), cudaMemcpyDeviceToHost);
        }
        int cap = io->capacity;
        io->capacity = std::max(cap, nC);
        rc = cz_world_download_contacts(w, 0, io);
        io->capacity = cap;
        if (rc) break;
        cz_bodies down = *bodies;
        down.acceleration = nullptr; down.linear_damping = nullptr; down.angular_damping = nullptr;
        down.inverse_inertia_tensor = nullptr; down.inverse_mass = nullptr; down.can_sleep = nullptr;
        rc = download_bodies(w->b, 0, bodies->n, &down);
    } while (0);
    cz_world_destroy(w);
    return rc;
}

// ---- multi-GPU runs: worlds sharded over devices, one grouped ncclAllReduce at the end ---------------------------
}  // extern "C"

#include <dlfcn.h>
namespace cznccl {   // the five NCCL entry points the run needs, resolved from libnccl.so.2 at run time (values of nccl.h 2.x)
typedef struct ncclComm *comm_t;
enum { Int64 = 4, Uint64 = 5, Float32 = 7, Float64 = 8, Sum = 0, Max = 2 };
typedef int (*CommInitAll_t)(comm_t *, int, const int *);
typedef int (*CommDestroy_t)(comm_t);
typedef int (*AllReduce_t)(const void *, void *, size_t, int, int, comm_t, cudaStream_t);
typedef int (*Group_t)(void);
typedef const char *(*ErrStr_t)(int);
struct Api {
    void *lib = nullptr;
    CommInitAll_t CommInitAll = nullptr;
    CommDestroy_t CommDestroy = nullptr;
    AllReduce_t AllReduce = nullptr;
    Group_t GroupStart = nullptr, GroupEnd = nullptr;
    ErrStr_t GetErrorString = nullptr;
    bool ok() const { return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd; }
};
static Api &api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) return;
        a.CommInitAll = (CommInitAll_t)dlsym(a.lib, "ncclCommInitAll");
        a.CommDestroy = (CommDestroy_t)dlsym(a.lib, "ncclCommDestroy");
        a.AllReduce = (AllReduce_t)dlsym(a.lib, "ncclAllReduce");
        a.GroupStart = (Group_t)dlsym(a.lib, "ncclGroupStart");
        a.GroupEnd = (Group_t)dlsym(a.lib, "ncclGroupEnd");
        a.GetErrorString = (ErrStr_t)dlsym(a.lib, "ncclGetErrorString");
    });
    return a;
}
}  // namespace cznccl

struct cz_run {
    struct Shard {
        int device = 0;
        cz_ctx *ctx = nullptr;
        cz_world *world = nullptr;
        int first = 0, count = 0;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        bool timing = false;
        long long steps = 0;
        // reduce buffers on the shard's device: [0] u64 checksum | [1] f64 energy | [2..5] i64 world-steps, contacts, pos, vel | [6] f32 ms
        unsigned long long *dRed = nullptr;
    };
    std::vector<Shard> shards;
    cz_world_desc desc{};
    bool distinct = true;                 // every shard on its own device -> NCCL
    std::vector<cznccl::comm_t> comms;
    std::string err;
};
static int run_fail(cz_run *r, int code, const std::string &msg) {
    g_err = msg;
    if (r) r->err = msg;
    return code;
}
// the slice of a whole-batch cz_bodies / cz_colliders that belongs to a shard
template <class T> static inline T *adv(T *p, long long n) { return p ? p + n : nullptr; }
static cz_bodies slice_bodies(const cz_bodies &a, long long first, long long n) {
    cz_bodies s = a;
    s.n = (int32_t)n;
    s.position = adv(a.position, first * 3); s.orientation = adv(a.orientation, first * 4); s.velocity = adv(a.velocity, first * 3);
    s.rotation = adv(a.rotation, first * 3); s.acceleration = adv(a.acceleration, first * 3); s.linear_damping = adv(a.linear_damping, first);
    s.angular_damping = adv(a.angular_damping, first); s.inverse_inertia_tensor = adv(a.inverse_inertia_tensor, first * 9);
    s.inverse_mass = adv(a.inverse_mass, first); s.motion = adv(a.motion, first); s.is_awake = adv(a.is_awake, first); s.can_sleep = adv(a.can_sleep, first);
    s.transform = adv(a.transform, first * 12); s.inverse_inertia_tensor_world = adv(a.inverse_inertia_tensor_world, first * 9);
    s.last_frame_acceleration = adv(a.last_frame_acceleration, first * 3);
    return s;
}
static cz_colliders slice_colliders(const cz_colliders &a, long long first, long long n) {
    cz_colliders s = a;
    s.n = (int32_t)n;
    s.shape = adv(a.shape, first); s.body = adv(a.body, first); s.offset = adv(a.offset, first * 12); s.transform = adv(a.transform, first * 12);
    s.half_size = adv(a.half_size, first * 3); s.radius = adv(a.radius, first);
    return s;
}

extern "C" {

const char *cz_run_last_error(cz_run *r) { return r ? r->err.c_str() : g_err.c_str(); }

int cz_run_destroy(cz_run *r) {
    if (!r) return CZ_ERR_INVALID;
    for (auto &c : r->comms) if (c) { cznccl::api().CommDestroy(c); }
    for (auto &s : r->shards) {
        if (s.ctx) cudaSetDevice(s.device);
        if (s.world) cz_world_destroy(s.world);
        if (s.ev0) cudaEventDestroy(s.ev0);
        if (s.ev1) cudaEventDestroy(s.ev1);
        if (s.dRed) cudaFree(s.dRed);
        if (s.ctx) cz_shutdown(s.ctx);
    }
    delete r;
    return CZ_OK;
}

int cz_run_create(int32_t n_shards, const int32_t *devices, const cz_world_desc *desc, cz_run **out) {
    if (!out || !desc || n_shards <= 0 || n_shards > 64 || desc->n_worlds < n_shards) return run_fail(nullptr, CZ_ERR_INVALID, "cz_run_create: bad argument (1..64 shards, at least one world per shard)");
    *out = nullptr;
    cz_run *r = new cz_run;
    r->desc = *desc;
    r->shards.resize(n_shards);
    const long long W = desc->n_worlds;
    for (int k = 0; k < n_shards; k++) {
        auto &s = r->shards[k];
        s.device = devices ? devices[k] : k;
        for (int j = 0; j < k; j++) if (r->shards[j].device == s.device) r->distinct = false;
        s.first = (int)(W * k / n_shards);
        s.count = (int)(W * (k + 1) / n_shards) - s.first;
        int rc = cz_init(s.device, &s.ctx);
        if (rc) { std::string e = g_err; cz_run_destroy(r); return run_fail(nullptr, rc, "cz_run_create: shard " + std::to_string(k) + ": " + e); }
        cz_world_desc d = *desc;
        d.n_worlds = s.count;
        rc = cz_world_create(s.ctx, &d, &s.world);
        if (rc) { std::string e = s.ctx->err; cz_run_destroy(r); return run_fail(nullptr, rc, "cz_run_create: shard " + std::to_string(k) + ": " + e); }
        if (cudaEventCreate(&s.ev0) != cudaSuccess || cudaEventCreate(&s.ev1) != cudaSuccess || cudaMalloc(&s.dRed, sizeof(unsigned long long) * 8) != cudaSuccess) {
            cz_run_destroy(r);
            return run_fail(nullptr, CZ_ERR_CUDA, "cz_run_create: event / buffer allocation failed");
        }
    }
    if (n_shards > 1 && r->distinct) {
        cznccl::Api &nc = cznccl::api();
        if (!nc.ok()) { cz_run_destroy(r); return run_fail(nullptr, CZ_ERR_CUDA, "cz_run_create: libnccl.so.2 could not be loaded (needed to reduce over more than one device)"); }
        std::vector<int> devs;
        for (auto &s : r->shards) devs.push_back(s.device);
        r->comms.assign(n_shards, nullptr);
        const int e = nc.CommInitAll(r->comms.data(), n_shards, devs.data());
        if (e != 0) {
            std::string msg = std::string("ncclCommInitAll: ") + (nc.GetErrorString ? nc.GetErrorString(e) : "error");
            r->comms.clear();
            cz_run_destroy(r);
            return run_fail(nullptr, CZ_ERR_CUDA, msg);
        }
    }
    *out = r;
    return CZ_OK;
}

int cz_run_shard(cz_run *r, int32_t k, cz_world **world, int32_t *first_world, int32_t *n_worlds) {
    if (!r || k < 0 || k >= (int)r->shards.size()) return run_fail(r, CZ_ERR_INVALID, "cz_run_shard: no such shard");
    if (world) *world = r->shards[k].world;
    if (first_world) *first_world = r->shards[k].first;
    if (n_worlds) *n_worlds = r->shards[k].count;
    return CZ_OK;
}
#define RUN_EACH(call)                                                                                                      \
    for (auto &s : r->shards) {                                                                                             \
        const int rc__ = (call);                                                                                            \
        if (rc__) return run_fail(r, rc__, s.ctx->err);                                                                     \
    }
int cz_run_upload_bodies(cz_run *r, const cz_bodies *all, int32_t derive) {
    if (!r || !all) return CZ_ERR_INVALID;
    const long long B = r->desc.bodies_per_world;
    if (all->n != (long long)r->desc.n_worlds * B) return run_fail(r, CZ_ERR_INVALID, "cz_run_upload_bodies: all->n must cover every world of the run");
    RUN_EACH(([&] { cz_bodies b = slice_bodies(*all, s.first * B, s.count * B); return cz_world_upload_bodies(s.world, 0, s.count, &b, derive); })());
    return CZ_OK;
}
int cz_run_upload_colliders(cz_run *r, const cz_colliders *all, int32_t derive) {
    if (!r || !all) return CZ_ERR_INVALID;
    const long long B = r->desc.bodies_per_world;
    if (all->n != (long long)r->desc.n_worlds * B) return run_fail(r, CZ_ERR_INVALID, "cz_run_upload_colliders: all->n must cover every world of the run");
    RUN_EACH(([&] { cz_colliders c = slice_colliders(*all, s.first * B, s.count * B); return cz_world_upload_colliders(s.world, 0, s.count, &c, derive); })());
    return CZ_OK;
}
int cz_run_upload_planes(cz_run *r, const cz_planes *p) {
    if (!r || !p) return CZ_ERR_INVALID;
    RUN_EACH(cz_world_upload_planes(s.world, p));
    return CZ_OK;
}
int cz_run_set_episodes(cz_run *r, int32_t length, const int32_t *phase0) {
    if (!r) return CZ_ERR_INVALID;
    RUN_EACH(cz_world_set_episodes(s.world, length, phase0 ? phase0 + s.first : nullptr));
    return CZ_OK;
}
int cz_run_step(cz_run *r, cz_real dt, int32_t n_steps) {
    if (!r || n_steps < 0) return CZ_ERR_INVALID;
    for (auto &s : r->shards) {
        if (s.timing) continue;
        if (cudaSetDevice(s.device) != cudaSuccess || cudaEventRecord(s.ev0, s.ctx->stream) != cudaSuccess) return run_fail(r, CZ_ERR_CUDA, "cz_run_step: event record failed");
        s.timing = true;
    }
    // a few frames per shard and round: every device has work queued before the host goes on enqueueing for the first
    const int chunk = 8;
    for (int done = 0; done < n_steps; done += chunk) {
        const int n = std::min(chunk, n_steps - done);
        RUN_EACH(cz_world_step(s.world, dt, n, nullptr));
    }
    for (auto &s : r->shards) {
        s.steps += n_steps;
        if (cudaSetDevice(s.device) != cudaSuccess || cudaEventRecord(s.ev1, s.ctx->stream) != cudaSuccess) return run_fail(r, CZ_ERR_CUDA, "cz_run_step: event record failed");
    }
    return CZ_OK;
}
int cz_run_finish(cz_run *r, cz_run_totals *out) {
    if (!r || !out) return CZ_ERR_INVALID;
    std::memset(out, 0, sizeof(*out));
    const int n = (int)r->shards.size();
    struct Part { unsigned long long checksum; double energy; long long c[4]; float ms; };
    std::vector<Part> part(n);
    for (int k = 0; k < n; k++) {
        auto &s = r->shards[k];
        cz_ctx *ctx = s.ctx;
        if (cudaSetDevice(s.device) != cudaSuccess) return run_fail(r, CZ_ERR_CUDA, "cudaSetDevice failed");
        int rc = cz_world_synchronize(s.world);                       // also surfaces a sticky device-side error of the asynchronous steps
        if (rc) return run_fail(r, rc, "shard " + std::to_string(k) + ": " + ctx->err);
        uint64_t cks = 0;
        double en = 0;
        if ((rc = cz_world_checksum_energy(s.world, &cks, &en))) return run_fail(r, rc, ctx->err);
        unsigned long long h[ST_N];
        if (cudaMemcpy(h, s.world->stats, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemset(s.world->stats, 0, sizeof(unsigned long long) * ST_STATUS) != cudaSuccess) return run_fail(r, CZ_ERR_CUDA, "reading the shard counters failed");
        float ms = 0;
        if (s.timing) cudaEventElapsedTime(&ms, s.ev0, s.ev1);
        part[k] = Part{cks, en, {(long long)s.count * s.steps, (long long)h[ST_CONTACTS], (long long)h[ST_POS], (long long)h[ST_VEL]}, ms};
        s.timing = false;
        s.steps = 0;
    }
    out->n_shards = n;
    if (!r->comms.empty()) {
        // ONE grouped all-reduce over the single-process communicator: every shard contributes its partial sums from a
        // buffer on its own device and receives the totals
        cznccl::Api &nc = cznccl::api();
        for (int k = 0; k < n; k++) {
            auto &s = r->shards[k];
            unsigned long long buf[8] = {0};
            buf[0] = part[k].checksum;
            std::memcpy(&buf[1], &part[k].energy, 8);
            for (int j = 0; j < 4; j++) buf[2 + j] = (unsigned long long)part[k].c[j];
            std::memcpy(&buf[6], &part[k].ms, 4);
            if (cudaSetDevice(s.device) != cudaSuccess || cudaMemcpyAsync(s.dRed, buf, sizeof(buf), cudaMemcpyHostToDevice, s.ctx->stream) != cudaSuccess)
                return run_fail(r, CZ_ERR_CUDA, "staging the reduce buffers failed");
        }
        int e = nc.GroupStart();
        for (int k = 0; k < n && e == 0; k++) {
            auto &s = r->shards[k];
            e = nc.AllReduce(s.dRed + 0, s.dRed + 0, 1, cznccl::Uint64, cznccl::Sum, r->comms[k], s.ctx->stream);
            if (!e) e = nc.AllReduce(s.dRed + 1, s.dRed + 1, 1, cznccl::Float64, cznccl::Sum, r->comms[k], s.ctx->stream);
            if (!e) e = nc.AllReduce(s.dRed + 2, s.dRed + 2, 4, cznccl::Int64, cznccl::Sum, r->comms[k], s.ctx->stream);
            if (!e) e = nc.AllReduce(s.dRed + 6, s.dRed + 6, 1, cznccl::Float32, cznccl::Max, r->comms[k], s.ctx->stream);
        }
        const int e2 = nc.GroupEnd();
        if (e || e2) return run_fail(r, CZ_ERR_CUDA, std::string("ncclAllReduce: ") + (nc.GetErrorString ? nc.GetErrorString(e ? e : e2) : "error"));
        unsigned long long buf[8];
        auto &s0 = r->shards[0];
        if (cudaSetDevice(s0.device) != cudaSuccess || cudaMemcpyAsync(buf, s0.dRed, sizeof(buf), cudaMemcpyDeviceToHost, s0.ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(s0.ctx->stream) != cudaSuccess) return run_fail(r, CZ_ERR_CUDA, "reading the reduced totals failed");
        for (int k = 1; k < n; k++) { cudaSetDevice(r->shards[k].device); cudaStreamSynchronize(r->shards[k].ctx->stream); }
        out->checksum = buf[0];
        std::memcpy(&out->energy, &buf[1], 8);
        out->world_steps = (int64_t)buf[2]; out->contacts = (int64_t)buf[3]; out->pos_iterations = (int64_t)buf[4]; out->vel_iterations = (int64_t)buf[5];
        std::memcpy(&out->max_device_ms, &buf[6], 4);
        out->used_nccl = 1;
    } else {
        for (int k = 0; k < n; k++) {
            out->checksum += part[k].checksum;
            out->energy += part[k].energy;
            out->world_steps += part[k].c[0]; out->contacts += part[k].c[1]; out->pos_iterations += part[k].c[2]; out->vel_iterations += part[k].c[3];
            out->max_device_ms = std::max(out->max_device_ms, part[k].ms);
        }
    }
    return CZ_OK;
}
#undef RUN_EACH

// ---- microbench / diagnostics --------------------------------------------------------------
int cz_bench_integrate(cz_ctx *ctx, int64_t n, uint64_t seed, int32_t warmup, int32_t steps, cz_real dt, float *avg_ms, uint64_t *checksum) {
    if (!ctx || n <= 0 || steps <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_bench_integrate: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    Batch b;
    int rc = batch_alloc(ctx, b, n);
    if (rc) { batch_free(b); return rc; }
    k_init_free_bodies<<<nblk(n, 256), 256, 0, ctx->stream>>>(b.st, seed, dt);
    real bias = (real)std::pow(0.5, (double)dt);
    // grid: a multiple of the SM count; 256-thread CTAs, 2+ resident per SM
    int grid = (int)std::min<long long>(nblk(n, 256), (long long)ctx->sm_count * 16);
    for (int i = 0; i < warmup; i++) k_integrate<false><<<grid, 256, 0, ctx->stream>>>(b.st, dt, bias, 0);
    cudaEventRecord(ctx->ev0, ctx->stream);
    for (int i = 0; i < steps; i++) k_integrate<false><<<grid, 256, 0, ctx->stream>>>(b.st, dt, bias, 0);
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev1);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { batch_free(b); return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); }
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    if (avg_ms) *avg_ms = ms / steps;
    if (checksum) {
        // checksum over the first 4096 bodies (keeps the check cheap)
        int W = (int)std::min<int64_t>(n, 4096);
        unsigned long long *dh;
        double *de;
        cudaMalloc(&dh, sizeof(unsigned long long) * W);
        cudaMalloc(&de, sizeof(double) * W);
        k_checksum_energy<<<nblk(W, 128), 128, 0, ctx->stream>>>(b.st, W, 1, dh, de);
        std::vector<unsigned long long> hh(W);
        cudaMemcpyAsync(hh.data(), dh, sizeof(unsigned long long) * W, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        unsigned long long s = 0;
        for (auto v : hh) s += v;
        *checksum = s;
        cudaFree(dh);
        cudaFree(de);
    }
    batch_free(b);
    return CZ_OK;
}

// FP64 pipe rate: the denominator of the fused step's roofline (that kernel is bound by FP64 issue / dependency latency,
// not by HBM).  Every thread runs 8 independent chains of alternating multiplies and adds (the product is built with
// -fmad=false, so its arithmetic is separate DMUL / DADD too); full occupancy, no memory traffic.
__global__ void __launch_bounds__(256) k_fp64_rate(double *out, int iters, double m, double a) {
    double x0 = threadIdx.x * 1e-9 + 1.0, x1 = x0 + 0.125, x2 = x0 + 0.25, x3 = x0 + 0.375, x4 = x0 + 0.5, x5 = x0 + 0.625, x6 = x0 + 0.75, x7 = x0 + 0.875;
    for (int i = 0; i < iters; i++) {
        x0 = x0 * m; x1 = x1 * m; x2 = x2 * m; x3 = x3 * m; x4 = x4 * m; x5 = x5 * m; x6 = x6 * m; x7 = x7 * m;
        x0 = x0 + a; x1 = x1 + a; x2 = x2 + a; x3 = x3 + a; x4 = x4 + a; x5 = x5 + a; x6 = x6 + a; x7 = x7 + a;
    }
    const double r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (r == 123.456) out[0] = r;   // keeps the chains alive
}
int cz_bench_fp64_rate(cz_ctx *ctx, double *ops_per_s) {
    if (!ctx || !ops_per_s) return fail(ctx, CZ_ERR_INVALID, "cz_bench_fp64_rate: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    double *d = nullptr;
    CK(ctx, cudaMalloc(&d, sizeof(double)));
    const int grid = ctx->sm_count * 8, iters = 1 << 14;
    k_fp64_rate<<<grid, 256, 0, ctx->stream>>>(d, 256, 0.999999999, 1e-9);   // warm-up
    cudaEventRecord(ctx->ev0, ctx->stream);
    k_fp64_rate<<<grid, 256, 0, ctx->stream>>>(d, iters, 0.999999999, 1e-9);
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaError_t e = cudaEventSynchronize(ctx->ev1);
    if (e == cudaSuccess) e = cudaGetLastError();
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    cudaFree(d);
    if (e != cudaSuccess || !(ms > 0)) return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
    *ops_per_s = (double)grid * 256.0 * 16.0 * (double)iters / ((double)ms * 1e-3);   // thread-level FP64 instructions per second
    return CZ_OK;
}

// ---- K2 entry points: host-buffer broadphase, microbench, sort test hooks ------------------------
__global__ void k_bp_load_host(const real *centers, const real *radii, long long n, czbp::Bounds *bounds, long long *box) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300}, rmax = 0;
    if (i < n) {
        czbp::Bounds b{centers[i * 3], centers[i * 3 + 1], centers[i * 3 + 2], radii[i]};
        bounds[i] = b;
        for (int k = 0; k < 3; k++) mn[k] = mx[k] = (double)centers[i * 3 + k];
        rmax = (double)b.r;
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int k = 0; k < 3; k++) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) { atomicMin(&box[k], czbp::ord(mn[k])); atomicMax(&box[3 + k], czbp::ord(mx[k])); }
        atomicMax(&box[6], czbp::ord(rmax));
    }
}
// n unit-radius spheres uniformly scattered in a cube sized for the requested volume fill
__global__ void k_bp_scatter_spheres(long long n, unsigned long long seed, double side, czbp::Bounds *bounds, long long *box) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    if (i < n) {
        unsigned long long st = seed * 0x100000001B3ull + (unsigned long long)i * 0x9E3779B97F4A7C15ull;
        czbp::Bounds b;
        b.x = (real)(sm64_next(st) * side); b.y = (real)(sm64_next(st) * side); b.z = (real)(sm64_next(st) * side); b.r = R_(1.0);
        bounds[i] = b;
        mn[0] = mx[0] = (double)b.x; mn[1] = mx[1] = (double)b.y; mn[2] = mx[2] = (double)b.z;
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; k++) { atomicMin(&box[k], czbp::ord(mn[k])); atomicMax(&box[3 + k], czbp::ord(mx[k])); }
        atomicMax(&box[6], czbp::ord(1.0));
    }
}

int cz_broadphase_pairs(cz_ctx *ctx, int64_t n, const cz_real *centers, const cz_real *radii, int64_t capacity, int32_t *pairs, int64_t *n_pairs) {
    if (!ctx || n <= 0 || !centers || !radii || !pairs || !n_pairs || capacity <= 0) return fail(ctx, CZ_ERR_INVALID, "cz_broadphase_pairs: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    czbp::Broadphase bp;
    cudaError_t e = czbp::bp_alloc(bp, n, (unsigned long long)capacity, 0, std::max<long long>(1ll << 22, 16 * n));
    real *dc = nullptr, *dr = nullptr;
    int rc = CZ_OK;
    do {
        if (e != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        if (cudaMalloc(&dc, sizeof(real) * 3 * n) != cudaSuccess || cudaMalloc(&dr, sizeof(real) * n) != cudaSuccess) { rc = fail(ctx, CZ_ERR_NOMEM, "cudaMalloc"); break; }
        cudaMemcpyAsync(dc, centers, sizeof(real) * 3 * n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(dr, radii, sizeof(real) * n, cudaMemcpyHostToDevice, ctx->stream);
        czbp::bp_reset_box(bp, ctx->stream);
        k_bp_load_host<<<nblk(n, 256), 256, 0, ctx->stream>>>(dc, dr, n, bp.bounds, bp.box);
        long long launches = 0;
        if ((e = czbp::bp_candidates(bp, ctx->stream, &launches)) != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        unsigned long long cnt[2];
        cudaMemcpyAsync(cnt, bp.counters, sizeof(cnt), cudaMemcpyDeviceToHost, ctx->stream);
        if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        *n_pairs = (int64_t)cnt[0];
        if (cnt[0] > (unsigned long long)capacity) { rc = fail(ctx, CZ_ERR_CAPACITY, "pair capacity exceeded"); break; }
        cudaMemcpy(pairs, bp.pairs, sizeof(uint2) * cnt[0], cudaMemcpyDeviceToHost);
    } while (0);
    if (dc) cudaFree(dc);
    if (dr) cudaFree(dr);
    czbp::bp_free(bp);
    return rc;
}

int cz_bench_broadphase(cz_ctx *ctx, int64_t n, uint64_t seed, double fill, int32_t warmup, int32_t steps, float *avg_ms, int64_t *n_pairs, float *sort_ms) {
    if (!ctx || n <= 0 || steps <= 0 || !(fill > 0)) return fail(ctx, CZ_ERR_INVALID, "cz_bench_broadphase: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    czbp::Broadphase bp;
    cudaError_t e = czbp::bp_alloc(bp, n, (unsigned long long)n * 8ull + 1024ull, 0, std::max<long long>(1ll << 22, 16 * n));
    if (e != cudaSuccess) { czbp::bp_free(bp); return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); }
    bp.trace = getenv("CUBEZ_BP_TRACE") != nullptr;
    const double side = cbrt((double)n * (4.0 / 3.0) * 3.14159265358979323846 / fill);
    float total = 0, totalSort = 0;
    cudaEvent_t s0, s1;
    cudaEventCreate(&s0); cudaEventCreate(&s1);
    int rc = CZ_OK;
    for (int it = 0; it < warmup + steps && rc == CZ_OK; it++) {
        // (re)generate the bounds outside the timed region: the timed frame is keys -> sort -> cells -> sweep
        czbp::bp_reset_box(bp, ctx->stream);
        k_bp_scatter_spheres<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, seed + it, side, bp.bounds, bp.box);
        cudaEventRecord(ctx->ev0, ctx->stream);
        long long launches = 0;
        if ((e = czbp::bp_candidates(bp, ctx->stream, &launches)) != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        cudaEventRecord(ctx->ev1, ctx->stream);
        if ((e = cudaEventSynchronize(ctx->ev1)) != cudaSuccess) { rc = fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); break; }
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        // the radix sort alone, on the same keys
        czbp::k_bp_keys<<<nblk(n, 256), 256, 0, ctx->stream>>>(bp.bounds, n, bp.grid, bp.sortCells.keys[0], bp.sortCells.vals[0]);
        cudaEventRecord(s0, ctx->stream);
        { int kb = 8; const long long cells = (long long)bp.grid.nx * bp.grid.ny * bp.grid.nz; while (kb < 32 && (1ll << kb) <= cells) kb += 8; czs::radix_sort(bp.sortCells, n, kb, ctx->stream, nullptr); }
        cudaEventRecord(s1, ctx->stream);
        cudaEventSynchronize(s1);
        float sms = 0;
        cudaEventElapsedTime(&sms, s0, s1);
        if (it >= warmup) { total += ms; totalSort += sms; }
    }
    unsigned long long cnt[2] = {0, 0};
    cudaMemcpy(cnt, bp.counters, sizeof(cnt), cudaMemcpyDeviceToHost);
    if (avg_ms) *avg_ms = total / steps;
    if (sort_ms) *sort_ms = totalSort / steps;
    if (n_pairs) *n_pairs = (int64_t)cnt[0];
    cudaEventDestroy(s0); cudaEventDestroy(s1);
    czbp::bp_free(bp);
    return rc;
}

}  // extern "C"

template <typename K> static int sort_pairs_host(cz_ctx *ctx, int64_t n, K *keys, uint32_t *vals, int bits) {
    if (!ctx || n < 0 || !keys || !vals) return fail(ctx, CZ_ERR_INVALID, "cz_sort_pairs: bad argument");
    if (n == 0) return CZ_OK;
    CK(ctx, cudaSetDevice(ctx->device));
    czs::RadixBuffers<K> buf{};
    cudaError_t e = czs::radix_alloc(buf, n);
    if (e != cudaSuccess) { czs::radix_free(buf); return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e)); }
    cudaMemcpyAsync(buf.keys[0], keys, sizeof(K) * n, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(buf.vals[0], vals, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, ctx->stream);
    int cur = czs::radix_sort(buf, n, bits, ctx->stream, nullptr);
    cudaMemcpyAsync(keys, buf.keys[cur], sizeof(K) * n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(vals, buf.vals[cur], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    czs::radix_free(buf);
    if (e != cudaSuccess) return fail(ctx, CZ_ERR_CUDA, cudaGetErrorString(e));
    return CZ_OK;
}
extern "C" {

int cz_sort_pairs_u32(cz_ctx *ctx, int64_t n, uint32_t *keys, uint32_t *vals) { return sort_pairs_host<unsigned>(ctx, n, keys, vals, 32); }
int cz_sort_pairs_u64(cz_ctx *ctx, int64_t n, uint64_t *keys, uint32_t *vals, int32_t bits) {
    return sort_pairs_host<unsigned long long>(ctx, n, (unsigned long long *)keys, vals, bits <= 0 ? 64 : (bits + 7) / 8 * 8);
}

}  // extern "C"

__global__ void k_math_op(int op, const real *in, real *out) {
    using namespace czm;
    auto v3 = [&](int o) { return mk3(in[o], in[o + 1], in[o + 2]); };
    auto q4 = [&](int o) { Q4 q; for (int k = 0; k < 4; k++) q.c[k] = in[o + k]; return q; };
    auto m3 = [&](int o) { M3 m; for (int k = 0; k < 9; k++) m.c[k] = in[o + k]; return m; };
    auto m34 = [&](int o) { M34 m; for (int k = 0; k < 12; k++) m.c[k] = in[o + k]; return m; };
    auto put3 = [&](const V3 &v) { out[0] = v.c[0]; out[1] = v.c[1]; out[2] = v.c[2]; };
    switch (op) {
    case CZ_OP_VEC_ADD: { V3 a = v3(0); v_add(a, v3(3)); put3(a); break; }
    case CZ_OP_VEC_ADD_SCALED: { V3 a = v3(0); v_add_scaled(a, v3(3), in[6]); put3(a); break; }
    case CZ_OP_VEC_COMPONENT_PRODUCT: { V3 a = v3(0); v_component_product(a, v3(3)); put3(a); break; }
    case CZ_OP_VEC_CROSS: put3(v_cross(v3(0), v3(3))); break;
    case CZ_OP_VEC_DOT: out[0] = v_dot(v3(0), v3(3)); break;
    case CZ_OP_VEC_MAGNITUDE: out[0] = v_mag(v3(0)); break;
    case CZ_OP_VEC_SQUARE_MAGNITUDE: out[0] = v_sqmag(v3(0)); break;
    case CZ_OP_VEC_MUL_WITH: { V3 a = v3(0); v_mul(a, in[3]); put3(a); break; }
    case CZ_OP_VEC_NORMALIZE: { V3 a = v3(0); v_normalize(a); put3(a); break; }
    case CZ_OP_VEC_SUB: { V3 a = v3(0); v_sub(a, v3(3)); put3(a); break; }
    case CZ_OP_QUAT_MUL: { Q4 a = q4(0); q_mul(a, q4(4)); for (int k = 0; k < 4; k++) out[k] = a.c[k]; break; }
    case CZ_OP_QUAT_LEN: out[0] = q_len(q4(0)); break;
    case CZ_OP_QUAT_NORMALIZE: { Q4 a = q4(0); q_normalize(a); for (int k = 0; k < 4; k++) out[k] = a.c[k]; break; }
    case CZ_OP_QUAT_ROTATE: put3(q_rotate(q4(0), v3(4))); break;
    case CZ_OP_QUAT_ADD_SCALED_VECTOR: { Q4 a = q4(0); q_add_scaled_vector(a, v3(4), in[7]); for (int k = 0; k < 4; k++) out[k] = a.c[k]; break; }
    case CZ_OP_M3_MUL_M3: { M3 r = m3_mul_m(m3(0), m3(9)); for (int k = 0; k < 9; k++) out[k] = r.c[k]; break; }
    case CZ_OP_M3_INVERT: { M3 r = m3_invert(m3(0)); for (int k = 0; k < 9; k++) out[k] = r.c[k]; break; }
    case CZ_OP_M3_MUL_V: put3(m3_mul_v(m3(0), v3(9))); break;
    case CZ_OP_M3_TRANSFORM_TRANSPOSE: put3(m3_transform_transpose(m3(0), v3(9))); break;
    case CZ_OP_M3_DETERMINANT: out[0] = m3_det(m3(0)); break;
    case CZ_OP_M34_MUL_M34: { M34 r = m34_mul_m34(m34(0), m34(12)); for (int k = 0; k < 12; k++) out[k] = r.c[k]; break; }
    case CZ_OP_M34_MUL_V: put3(m34_mul_v(m34(0), v3(12))); break;
    case CZ_OP_M34_TRANSFORM_INVERSE: put3(m34_transform_inverse(m34(0), v3(12))); break;
    case CZ_OP_M34_SET_AS_TRANSFORM: { M34 r; m34_set_as_transform(r, v3(0), q4(3)); for (int k = 0; k < 12; k++) out[k] = r.c[k]; break; }
    case CZ_OP_REAL_EQUAL: out[0] = real_equal(in[0], in[1]) ? R_(1) : R_(0); break;
    case CZ_OP_TRANSFORM_INERTIA: { M3 w; transform_inertia_tensor(w, m3(0), m34(9)); for (int k = 0; k < 9; k++) out[k] = w.c[k]; break; }
    default: out[0] = R_(0);
    }
}

extern "C" {

int cz_math_op(cz_ctx *ctx, int32_t op, const cz_real *in, cz_real *out) {
    if (!ctx || !in || !out) return CZ_ERR_INVALID;
    CK(ctx, cudaSetDevice(ctx->device));
    real *d = nullptr;
    CK(ctx, cudaMalloc(&d, sizeof(real) * 48));
    CK(ctx, cudaMemcpy(d, in, sizeof(real) * 24, cudaMemcpyHostToDevice));
    k_math_op<<<1, 1, 0, ctx->stream>>>(op, d, d + 24);
    CKL(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaMemcpy(out, d + 24, sizeof(real) * 12, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return CZ_OK;
}

}  // extern "C"
