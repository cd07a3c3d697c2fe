// cz_broadphase.cuh — K2: sort-based broadphase replacing the reference's O(n^2) pair scan
// (examples/cubedrop.go:54-63 "yes this is O(n^2) and not good practice").
//
// Pipeline per frame for one large world of n colliders:
//   k_bp_bounds     centre (collider transform column 3) + bounding radius -> bounds[i] (4 reals),
//                   and the axis-aligned box of all centres (block reduce + ordered atomics)
//   uniform grid, cell edge >= 2*Rmax*(1+margin); key = linear cell id.  Two sorts by cell key:
//   (default) ONE-PASS radix sort with radix = number of cells (counting sort):
//     k_bp_count    key of body i, rank of i inside its cell (L2 atomic on the cell counter)
//     scan          exclusive scan of the counters -> first sorted position of every cell
//     k_bp_resolve  sorted position of body i = start[cell] + rank (the random table reads, kept apart
//                   from the scatter below, which would evict the table from L2)
//     k_bp_place    body i -> sorted[position]: one 32-byte sector per body, one 256-bit store (float
//                   centre relative to the grid origin, inflated radius, original index, row key and
//                   the x-cell interval that can hold a partner)
//     k_bp_sweep    every sorted body walks the forward half of its neighbourhood as five contiguous
//                   ranges of the sorted array (a row of cells = consecutive keys) in ONE flattened
//                   loop, four partners per trip, float sphere test, each unordered candidate once
//   (CUBEZ_BP_SORT=radix) the LSD 8-bit radix sort of cz_sort.cuh:
//     k_bp_keys -> radix sort -> k_bp_gather -> k_bp_cells -> k_bp_pairs (f64 test)
//   k_bp_narrow     both ordered checks (i,j) and (j,i) of every candidate through czn::check_pair
//   k_bp_planes     every collider against every plane (planes bypass the grid)
//   radix sort      contacts by canonical key (check id * 8 + vertex)  == the reference's append order
//   k_bp_emit       gather into the world's as-generated contact arrays
// False positives are free, a dropped pair would not be: the test is inclusive and inflated by
// BP_MARGIN, the same margin czn::bounding_reject uses.
#pragma once
#include "cz_kernels.cuh"
#include "cz_sort.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace czbp {
using namespace czm;
using namespace czk;

struct Bounds { real x, y, z, r; };

// order-preserving map double -> int64 for atomicMin/Max
__device__ __forceinline__ long long ord(double v) {
    long long b = __double_as_longlong(v);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffll);
}
__host__ __device__ __forceinline__ double unord(long long b) {
    long long r = b >= 0 ? b : (b ^ 0x7fffffffffffffffll);
    double d;
    memcpy(&d, &r, sizeof(d));
    return d;
}

// box[0..2] = min, box[3..5] = max (ordered ints), box[6] = max radius (ordered)
__global__ void k_bp_bounds(BodyStore s, long long n, long long step, Bounds *bounds, long long *box) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300}, rmax = 0;
    if (i < n) {
        Bounds b;
        const int shape = s.shape[i];
        const bool active = shape != CZ_SHAPE_NONE && step >= (long long)s.active_from[i];
        real2 x89 = s.ld(czb::C_X89, i), x1011 = s.ld(czb::C_X1011, i);
        real2 h01 = s.ld(czb::C_H01, i), h2r = s.ld(czb::C_H2R, i);
        b.x = x89.y; b.y = x1011.x; b.z = x1011.y;
        b.r = shape == CZ_SHAPE_SPHERE ? h2r.y : rsqrt_(h01.x * h01.x + h01.y * h01.y + h2r.x * h2r.x);
        if (!active) b.r = R_(-1);
        bounds[i] = b;
        if (active) {
            mn[0] = mx[0] = (double)b.x; mn[1] = mx[1] = (double)b.y; mn[2] = mx[2] = (double)b.z;
            rmax = (double)b.r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(&box[k], ord(mn[k]));
            atomicMax(&box[3 + k], ord(mx[k]));
        }
        atomicMax(&box[6], ord(rmax));
    }
}

struct Grid {
    double ox, oy, oz, inv;   // origin and 1/cell (y and z; x too on the radix path)
    double invx;              // 1/cell along x: rows of cells are contiguous in the sorted array, so the counting path
                              // makes x-cells as long as the table budget asks and looks up the exact x-range per body
    double e1;                // float rounding bound of a grid-relative coordinate: extent * 2^-24
    double rmaxInfl;          // >= every Entry::r
    int nx, ny, nz;
};

__device__ __forceinline__ void cell_of(const Grid &g, const Bounds &b, int &cx, int &cy, int &cz) {
    cx = min(max((int)floor(((double)b.x - g.ox) * g.invx), 0), g.nx - 1);
    cy = min(max((int)floor(((double)b.y - g.oy) * g.inv), 0), g.ny - 1);
    cz = min(max((int)floor(((double)b.z - g.oz) * g.inv), 0), g.nz - 1);
}

#define BP_INACTIVE 0xffffffffu
__global__ void k_bp_keys(const Bounds *bounds, long long n, Grid g, unsigned *keys, unsigned *vals) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Bounds b = bounds[i];
    unsigned key = (unsigned)(g.nx * g.ny * g.nz);   // inactive colliders: one past the last cell, sorts to the end
    if (b.r >= R_(0)) {
        int cx, cy, cz;
        cell_of(g, b, cx, cy, cz);
        key = (unsigned)((cz * g.ny + cy) * g.nx + cx);
    }
    keys[i] = key;
    vals[i] = (unsigned)i;
}

__global__ void k_bp_gather(const Bounds *bounds, const unsigned *vals, long long n, Bounds *sorted) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) sorted[p] = bounds[vals[p]];
}

__global__ void k_bp_cells(const unsigned *keys, long long n, unsigned nCells, uint2 *cellRange, unsigned *occ) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const unsigned k = keys[p];
    if (k >= nCells) return;
    if (p == 0 || keys[p - 1] != k) { cellRange[k].x = (unsigned)p; atomicOr(&occ[k >> 5], 1u << (k & 31)); }
    if (p == n - 1 || keys[p + 1] != k) cellRange[k].y = (unsigned)(p + 1);
}

#define BP_MARGIN 1.005   // on the distance (1.01 on its square), as czn::bounding_reject

// candidate pairs (original collider indices; each unordered pair once).  Hits are rare (~0.2 per
// body at 5 % fill), so they are staged per CTA in shared memory (shared-memory atomic for the slot)
// and flushed with ONE global atomic per CTA and a coalesced copy; a per-hit warp ballot made every
// lane iterate to the longest neighbour range of its warp and the kernel was issue-bound.
#define BP_STAGE 1024
__global__ void __launch_bounds__(256) k_bp_pairs(const Bounds *sorted, const unsigned *keys, const unsigned *vals, long long n, Grid g,
                                                  const uint2 *cellRange, const unsigned *occ, uint2 *pairs, unsigned long long *nPairs,
                                                  unsigned long long capacity) {
    __shared__ uint2 stage[BP_STAGE];
    __shared__ unsigned nStage;
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) nStage = 0;
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < n && keys[p] < (unsigned)(g.nx * g.ny * g.nz);
    if (valid) {
        const Bounds me = sorted[p];
        int cx, cy, cz;
        cell_of(g, me, cx, cy, cz);
        const unsigned myVal = vals[p];
        // Each unordered pair is emitted once, from the body that comes first in sorted (key) order,
        // so only the "forward" half of the 27-cell neighbourhood is visited: the own row from the own
        // cell on, and the four rows with a larger key ((dz,dy) = (0,+1), (+1,-1), (+1,0), (+1,+1)).
#pragma unroll 1
        for (int row = 0; row < 5; row++) {
            const int dz = row >= 2 ? 1 : 0, dy = row == 0 ? 0 : (row == 1 ? 1 : row - 3);
            const int y = cy + dy, z = cz + dz;
            if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
            // the cells of a row are consecutive keys -> one contiguous range of the sorted array
            unsigned q0 = 0, q1 = 0;
            bool first = true;
            for (int dx = row == 0 ? 0 : -1; dx <= 1; dx++) {
                const int x = cx + dx;
                if (x < 0 || x >= g.nx) continue;
                // 1 bit per cell (L2-resident even when the 8-byte range table is not): most cells are empty
                const unsigned c = (unsigned)((z * g.ny + y) * g.nx + x);
                if (!((occ[c >> 5] >> (c & 31)) & 1u)) continue;
                const uint2 r = cellRange[c];
                if (first) { q0 = r.x; first = false; }
                q1 = r.y;
            }
            if (row == 0 && q0 <= (unsigned)p) q0 = (unsigned)p + 1;   // own cell: only bodies after me
            for (unsigned q = q0; q < q1; q++) {
                const Bounds ob = sorted[q];
                const double ddx = (double)ob.x - (double)me.x, ddy = (double)ob.y - (double)me.y, ddz = (double)ob.z - (double)me.z;
                const double rr = ((double)ob.r + (double)me.r) * BP_MARGIN;
                if (ddx * ddx + ddy * ddy + ddz * ddz <= rr * rr + 1e-9) {
                    const uint2 pr = make_uint2(myVal, vals[q]);
                    const unsigned slot = atomicAdd(&nStage, 1u);
                    if (slot < BP_STAGE) stage[slot] = pr;
                    else {   // staging full: straight to global
                        const unsigned long long gslot = atomicAdd(nPairs, 1ull);
                        if (gslot < capacity) pairs[gslot] = pr;
                    }
                }
            }
        }
    }
    __syncthreads();
    const unsigned cnt = min(nStage, (unsigned)BP_STAGE);
    if (threadIdx.x == 0 && cnt) base = atomicAdd(nPairs, (unsigned long long)cnt);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x)
        if (base + i < capacity) pairs[base + i] = stage[i];
}


// ------------------------------------------------------------------------------------------------
// One-pass (counting) sort by cell + flattened sweep — the default candidate generator
// ------------------------------------------------------------------------------------------------
// One 32-byte sector per sorted body: scattered writes fill whole sectors (no read-modify-write in
// L2) and the sweep's test needs only the first 16 bytes.
struct __align__(16) Entry {
    float x, y, z, r;          // centre relative to the grid origin; radius * BP_MARGIN + float slack, rounded up
    unsigned idx;              // original collider index
    unsigned rowOwn;           // key of cell 0 of the body's row of x-cells: (cz*ny + cy)*nx
    unsigned xlf;              // first x-cell that can hold a partner | flags << 29 (1: row y+1 exists, 2: row y-1, 4: slab z+1)
    unsigned xh;               // last x-cell that can hold a partner
};
static_assert(sizeof(Entry) == 32, "Entry must be one sector");

// L2 residency plan of the counting path: the cell table (4 B per cell, hit at random by one atomic
// and one load per body) is tagged evict_last; everything that streams through once (bounds, ranks,
// the scattered sorted entries) is tagged evict_first, so the streams do not push the table out.
__device__ __forceinline__ unsigned long long l2_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned atom_add_keep(unsigned *a, unsigned v, unsigned long long pol) {
    unsigned old;
    asm volatile("atom.global.add.L2::cache_hint.u32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(v), "l"(pol) : "memory");
    return old;
}
__device__ __forceinline__ unsigned ld_keep(const unsigned *a, unsigned long long pol) {
    unsigned v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
// bounds are read once per pass: streaming loads keep the cell table resident in L2
__device__ __forceinline__ Bounds ld_bounds_stream(const Bounds *p) {
    Bounds b;
#ifdef CUBEZ_REAL_FLOAT
    const float4 v = __ldcs(reinterpret_cast<const float4 *>(p));
    b.x = v.x; b.y = v.y; b.z = v.z; b.r = v.w;
#else
    const double2 v0 = __ldcs(reinterpret_cast<const double2 *>(p)), v1 = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
    b.x = v0.x; b.y = v0.y; b.z = v1.x; b.r = v1.y;
#endif
    return b;
}

__device__ __forceinline__ bool cell_key(const Grid &g, const Bounds &b, int &cx, int &cy, int &cz, unsigned &key) {
    if (!(b.r >= R_(0))) return false;   // inactive (r = -1)
    cell_of(g, b, cx, cy, cz);
    key = (unsigned)((cz * g.ny + cy) * g.nx + cx);
    return true;
}

__global__ void __launch_bounds__(256) k_bp_count(const Bounds *__restrict__ bounds, long long n, Grid g, unsigned *__restrict__ cellCount,
                                                  uint2 *__restrict__ keyRank) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Bounds b = ld_bounds_stream(bounds + i);
    int cx, cy, cz;
    unsigned key = BP_INACTIVE, rank = 0;
    if (cell_key(g, b, cx, cy, cz, key)) rank = atom_add_keep(&cellCount[key], 1u, l2_evict_last());
    __stcs(keyRank + i, make_uint2(key, rank));
}

// sorted position of body i = first position of its cell + its rank.  A kernel of its own: the only
// random access is the 4-byte table read, and the table stays in L2 as long as no scattered stores
// run beside it (measured: next to the scatter of k_bp_place every table read missed, +1 GB of DRAM reads).
__global__ void __launch_bounds__(256) k_bp_resolve(const uint2 *__restrict__ keyRank, long long n, const unsigned *__restrict__ cellStart,
                                                    unsigned *__restrict__ pos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 kr = __ldcs(keyRank + i);
    __stcs(pos + i, kr.x == BP_INACTIVE ? BP_INACTIVE : ld_keep(cellStart + kr.x, l2_evict_last()) + kr.y);
}

// slack: float rounding of the relative centre is <= extent * 2^-24 per coordinate (Grid::e1); each
// radius carries 4*e1, so the float test accepts every pair the exact inflated test accepts.
__global__ void __launch_bounds__(256) k_bp_place(const Bounds *__restrict__ bounds, long long n, Grid g, const unsigned *__restrict__ pos,
                                                  Entry *__restrict__ sorted) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned p = __ldcs(pos + i);
    if (p == BP_INACTIVE) return;
    const Bounds b = ld_bounds_stream(bounds + i);
    int cx, cy, cz;
    cell_of(g, b, cx, cy, cz);
    const float x = (float)((double)b.x - g.ox), y = (float)((double)b.y - g.oy), z = (float)((double)b.z - g.oz);
    const float r = __double2float_ru(((double)b.r * BP_MARGIN + 4.0 * g.e1) * 1.000002);   // 2e-6: the float roundings of the test itself
    // x-cells that can hold a partner: every centre within my radius + the largest radius (+ rounding slack)
    const double reach = ((double)r + g.rmaxInfl) * 1.000002 + 8.0 * g.e1;
    const int xl = min(max((int)floor(((double)x - reach) * g.invx), 0), cx);
    const int xh = max(min((int)floor(((double)x + reach) * g.invx), g.nx - 1), cx);
    const unsigned flags = (cy + 1 < g.ny ? 1u : 0u) | (cy > 0 ? 2u : 0u) | (cz + 1 < g.nz ? 4u : 0u);
    // ONE 256-bit store: the scattered write fills its sector, so L2 never fetches it first
    const unsigned long long w0 = (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
    const unsigned long long w1 = (unsigned long long)__float_as_uint(z) | ((unsigned long long)__float_as_uint(r) << 32);
    const unsigned long long w2 = (unsigned long long)(unsigned)i | ((unsigned long long)(unsigned)((cz * g.ny + cy) * g.nx) << 32);
    const unsigned long long w3 = (unsigned long long)((unsigned)xl | (flags << 29)) | ((unsigned long long)(unsigned)xh << 32);
    asm volatile("st.global.L2::cache_hint.v4.b64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(sorted + p), "l"(w0), "l"(w1), "l"(w2), "l"(w3), "l"(l2_evict_first()) : "memory");
}

__global__ void __launch_bounds__(256, 4) k_bp_sweep(const Entry *__restrict__ sorted, long long n, Grid g, const unsigned *__restrict__ cellStart,
                                                  uint2 *pairs, unsigned long long *nPairs, unsigned long long capacity) {
    __shared__ uint2 stage[BP_STAGE];
    __shared__ unsigned nStage;
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) nStage = 0;
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned cells = (unsigned)(g.nx * g.ny * g.nz);
    const unsigned nActive = cellStart[cells];
    if (p < (long long)nActive) {
        const float4 me = *reinterpret_cast<const float4 *>(sorted + p);
        const uint4 mi = *(reinterpret_cast<const uint4 *>(sorted + p) + 1);   // idx, rowOwn, xl|flags, xh
        const unsigned rowOwn = mi.y, xl = mi.z & 0x1fffffffu, fl = mi.z >> 29, xh1 = mi.w + 1u;
        // five contiguous ranges of the sorted array: the own row from the body after me to the end
        // of cell xh, and the rows (dz,dy) = (0,+1), (+1,-1), (+1,0), (+1,+1) over cells xl..xh
        unsigned a0, e0, a1 = 0, e1 = 0, a2 = 0, e2 = 0, a3 = 0, e3 = 0, a4 = 0, e4 = 0;
        a0 = (unsigned)p + 1u;
        e0 = cellStart[rowOwn + xh1];
        if (fl & 1u) { const unsigned r = rowOwn + g.nx; a1 = cellStart[r + xl]; e1 = cellStart[r + xh1]; }
        if (fl & 4u) {
            const unsigned rz = rowOwn + (unsigned)(g.nx * g.ny);
            a3 = cellStart[rz + xl]; e3 = cellStart[rz + xh1];
            if (fl & 2u) { const unsigned r = rz - g.nx; a2 = cellStart[r + xl]; e2 = cellStart[r + xh1]; }
            if (fl & 1u) { const unsigned r = rz + g.nx; a4 = cellStart[r + xl]; e4 = cellStart[r + xh1]; }
        }
        // one flattened loop over the five ranges: a warp iterates to the longest TOTAL of its lanes,
        // not to the sum of the per-row maxima.  The flat counter k maps to a sorted position through a
        // branch-free select chain; four partners per trip, their loads issued together (the test is a
        // dozen instructions: one load in flight per trip left the loop waiting on L1/L2 latency).
        const unsigned c1 = e0 - a0, c2 = c1 + (e1 - a1), c3 = c2 + (e2 - a2), c4 = c3 + (e3 - a3), total = c4 + (e4 - a4);
        const unsigned o0 = a0, o1 = a1 - c1, o2 = a2 - c2, o3 = a3 - c3, o4 = a4 - c4;
        for (unsigned k0 = 0; k0 < total; k0 += 4) {
            unsigned q[4];
            float4 ob[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const unsigned k = k0 + j;
                unsigned off = k < c1 ? o0 : o1;
                off = k < c2 ? off : o2;
                off = k < c3 ? off : o3;
                off = k < c4 ? off : o4;
                q[j] = k < total ? off + k : (unsigned)p;   // past the end: re-read myself (masked below)
                ob[j] = *reinterpret_cast<const float4 *>(sorted + q[j]);
            }
            unsigned hits = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float dx = ob[j].x - me.x, dy = ob[j].y - me.y, dz = ob[j].z - me.z;
                const float rr = ob[j].w + me.w;
                const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
                if (d2 <= rr * rr && k0 + j < total) hits |= 1u << j;
            }
            while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                const unsigned qq = j == 0 ? q[0] : (j == 1 ? q[1] : (j == 2 ? q[2] : q[3]));
                const uint2 pr = make_uint2(mi.x, sorted[qq].idx);
                const unsigned slot = atomicAdd(&nStage, 1u);
                if (slot < BP_STAGE) stage[slot] = pr;
                else {   // staging full: straight to global
                    const unsigned long long gslot = atomicAdd(nPairs, 1ull);
                    if (gslot < capacity) pairs[gslot] = pr;
                }
            }
        }
    }
    __syncthreads();
    const unsigned cnt = min(nStage, (unsigned)BP_STAGE);
    if (threadIdx.x == 0 && cnt) base = atomicAdd(nPairs, (unsigned long long)cnt);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x)
        if (base + i < capacity) pairs[base + i] = stage[i];
}

// ---- narrowphase on the candidates; contacts are produced unordered with their canonical key ----
struct KeyedContacts {
    unsigned long long *keys;   // check id * 8 + vertex
    unsigned *vals;             // slot in the payload arrays
    real *payload;              // [slot][8]: point3 normal3 pen
    int2 *ids;                  // [slot] body indices
    unsigned long long *count;
    unsigned long long capacity;
};
__device__ __forceinline__ void push_contact(const KeyedContacts &kc, unsigned long long key, const GenContact &c) {
    const unsigned long long slot = atomicAdd(kc.count, 1ull);
    if (slot >= kc.capacity) return;
    kc.keys[slot] = key;
    kc.vals[slot] = (unsigned)slot;
    real *r = kc.payload + slot * 8;
#pragma unroll
    for (int k = 0; k < 3; k++) { r[k] = c.point.c[k]; r[3 + k] = c.normal.c[k]; }
    r[6] = c.pen;
    kc.ids[slot] = make_int2(c.b0, c.b1);
}

__global__ void __launch_bounds__(128) k_bp_narrow(WorldParams p, const uint2 *pairs, unsigned long long nPairs, KeyedContacts kc) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nPairs * 2) return;
    const uint2 pr = pairs[t >> 1];
    const int a = (t & 1) ? (int)pr.y : (int)pr.x, b = (t & 1) ? (int)pr.x : (int)pr.y;   // both ordered checks
    ColliderView one = load_collider(p.st, a, a), two = load_collider(p.st, b, b);
    V3 v1 = zero3(), v2 = zero3();
    if (one.shape != two.shape) { v1 = czb::ld_velocity(p.st, a); v2 = czb::ld_velocity(p.st, b); }
    GenContact gc;
    if (czn::check_pair(one, two, v1, v2, gc)) {
        const unsigned long long check = (unsigned long long)a * (unsigned long long)(p.P + p.B) + (unsigned long long)(p.P + b);
        push_contact(kc, check * 8ull, gc);
    }
}

__global__ void __launch_bounds__(128) k_bp_planes(WorldParams p, KeyedContacts kc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.B * p.P) return;
    const int i = (int)(t / p.P), pl = (int)(t % p.P);
    if (!body_active(p, i)) return;
    ColliderView c = load_collider(p.st, i, i);
    const unsigned long long check = (unsigned long long)i * (unsigned long long)(p.P + p.B) + (unsigned long long)pl;
    if (c.shape == CZ_SHAPE_SPHERE) {
        GenContact gc;
        if (czn::sphere_halfspace(c, p.planes[pl], gc)) push_contact(kc, check * 8ull, gc);
    } else if (c.shape == CZ_SHAPE_CUBE) {
        const unsigned mask = czn::cube_halfspace_mask(c, p.planes[pl]);
#pragma unroll 1
        for (int v = 0; v < 8; v++)
            if (mask & (1u << v)) {
                GenContact gc;
                czn::cube_halfspace_contact(c, p.planes[pl], v, gc);
                push_contact(kc, check * 8ull + (unsigned long long)v, gc);
            }
    }
}

// sorted contacts -> the world's as-generated arrays (world 0)
__global__ void k_bp_emit(WorldParams p, const unsigned long long *sortedKeys, const unsigned *sortedVals, const real *payload, const int2 *idsArr,
                          const unsigned long long *count) {
    using namespace czr;
    const unsigned long long n = *count;
    const unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) {
        p.nContacts[0] = (int)(n > 0x7fffffffull ? 0x7fffffff : n);
        if (n > (unsigned long long)p.Cc) raise_status(p.stats, CZ_ERR_CAPACITY);
        atomicAdd(&p.stats[ST_CONTACTS], n);
        atomicMax(&p.stats[ST_MAXC], n);
    }
    if (c >= n || c >= (unsigned long long)p.Cc) return;
    const real *r = payload + (size_t)sortedVals[c] * 8;
    const long long gs = (long long)p.W * p.Cc;
#pragma unroll
    for (int k = 0; k < 3; k++) { p.gen[(G_POINT + k) * gs + c] = r[k]; p.gen[(G_NORMAL + k) * gs + c] = r[3 + k]; }
    p.gen[G_PEN * gs + c] = r[6];
    real fric = R_(0.9), rest = R_(0.1);
    if (p.matFric) {   // the canonical key names the check: (a, b) = (check / (P+B), check % (P+B))
        const unsigned long long check = sortedKeys[c] >> 3, per = (unsigned long long)(p.P + p.B);
        const int a = (int)(check / per), slot = (int)(check % per);
        check_material(p, 0, a, slot < p.P ? -(slot + 1) : slot - p.P, fric, rest);
    }
    p.gen[G_FRIC * gs + c] = fric;
    p.gen[G_REST * gs + c] = rest;
    const int2 ids = idsArr[sortedVals[c]];
    p.gb0[c] = ids.x;
    p.gb1[c] = ids.y;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Broadphase {
    long long n = 0;
    Bounds *bounds = nullptr, *sorted = nullptr;
    long long *box = nullptr;            // device [8]
    long long *h_box = nullptr;          // pinned [8]
    czs::RadixBuffers<unsigned> sortCells{};
    uint2 *cellRange = nullptr;          // radix path: [first,last) per cell; counting path: reused as cellStart (unsigned[2*cap])
    unsigned *occ = nullptr;             // occupancy bit per cell
    long long cellCapacity = 0;
    Entry *entries = nullptr;            // counting path: one sector per sorted body
    uint2 *keyRank = nullptr;            // counting path: cell key and rank of body i inside its cell
    unsigned *pos = nullptr;             // counting path: sorted position of body i
    unsigned *scanScratch = nullptr;
    bool useRadix = false;               // CUBEZ_BP_SORT=radix
    uint2 *pairs = nullptr;
    unsigned long long pairCapacity = 0;
    unsigned long long *counters = nullptr;      // device [2]: nPairs, nContacts
    unsigned long long *h_counters = nullptr;    // pinned [2]
    czs::RadixBuffers<unsigned long long> sortContacts{};
    real *payload = nullptr;
    int2 *ids = nullptr;
    unsigned long long contactCapacity = 0;
    Grid grid{};
    unsigned long long lastPairs = 0, lastContacts = 0;
    bool trace = false;                  // per-stage CUDA-event timing (diagnostics)
    cudaEvent_t tev[8] = {};
    float stageMs[8] = {};               // keys, sort, memset, gather, cells, pairs
};

static inline cudaError_t bp_alloc(Broadphase &bp, long long n, unsigned long long pairCap, unsigned long long contactCap, long long maxCells) {
    bp.n = n;
    cudaError_t e;
#define BPCK(x) if ((e = (x)) != cudaSuccess) return e
    BPCK(cudaMalloc(&bp.bounds, sizeof(Bounds) * n));
    BPCK(cudaMalloc(&bp.sorted, sizeof(Bounds) * n));
    BPCK(cudaMalloc(&bp.box, sizeof(long long) * 8));
    BPCK(cudaHostAlloc(&bp.h_box, sizeof(long long) * 8, cudaHostAllocDefault));
    BPCK(czs::radix_alloc(bp.sortCells, n));
    bp.cellCapacity = maxCells;
    BPCK(cudaMalloc(&bp.cellRange, sizeof(uint2) * maxCells));
    BPCK(cudaMalloc(&bp.occ, sizeof(unsigned) * (maxCells / 32 + 1)));
    BPCK(cudaMalloc(&bp.entries, sizeof(Entry) * n));
    BPCK(cudaMalloc(&bp.keyRank, sizeof(uint2) * n));
    BPCK(cudaMalloc(&bp.pos, sizeof(unsigned) * n));
    BPCK(cudaMalloc(&bp.scanScratch, sizeof(unsigned) * (size_t)czs::scan_scratch_elems(maxCells + 1)));
    { const char *m = getenv("CUBEZ_BP_SORT"); bp.useRadix = m && m[0] == 'r'; }
    bp.pairCapacity = pairCap;
    BPCK(cudaMalloc(&bp.pairs, sizeof(uint2) * pairCap));
    BPCK(cudaMalloc(&bp.counters, sizeof(unsigned long long) * 2));
    BPCK(cudaHostAlloc(&bp.h_counters, sizeof(unsigned long long) * 2, cudaHostAllocDefault));
    bp.contactCapacity = contactCap;
    if (contactCap) {
        BPCK(czs::radix_alloc(bp.sortContacts, (long long)contactCap));
        BPCK(cudaMalloc(&bp.payload, sizeof(real) * 8 * contactCap));
        BPCK(cudaMalloc(&bp.ids, sizeof(int2) * contactCap));
    }
#undef BPCK
    return cudaSuccess;
}
static inline void bp_free(Broadphase &bp) {
    if (bp.bounds) cudaFree(bp.bounds);
    if (bp.sorted) cudaFree(bp.sorted);
    if (bp.box) cudaFree(bp.box);
    if (bp.h_box) cudaFreeHost(bp.h_box);
    czs::radix_free(bp.sortCells);
    if (bp.cellRange) cudaFree(bp.cellRange);
    if (bp.occ) cudaFree(bp.occ);
    if (bp.entries) cudaFree(bp.entries);
    if (bp.keyRank) cudaFree(bp.keyRank);
    if (bp.pos) cudaFree(bp.pos);
    if (bp.scanScratch) cudaFree(bp.scanScratch);
    if (bp.pairs) cudaFree(bp.pairs);
    if (bp.counters) cudaFree(bp.counters);
    if (bp.h_counters) cudaFreeHost(bp.h_counters);
    czs::radix_free(bp.sortContacts);
    if (bp.payload) cudaFree(bp.payload);
    if (bp.ids) cudaFree(bp.ids);
    bp = Broadphase();
}

static inline cudaError_t bp_reset_box(Broadphase &bp, cudaStream_t st) {
    long long init[8] = {0x7fffffffffffffffll, 0x7fffffffffffffffll, 0x7fffffffffffffffll, (long long)0x8000000000000000ull,
                         (long long)0x8000000000000000ull, (long long)0x8000000000000000ull, 0, 0};
    memcpy(bp.h_box, init, sizeof(init));
    return cudaMemcpyAsync(bp.box, bp.h_box, sizeof(init), cudaMemcpyHostToDevice, st);
}

// Candidate generation from bounds that are already in bp.bounds (box must be reduced too).
// Returns cudaError; *launches accumulates.
static inline cudaError_t bp_candidates(Broadphase &bp, cudaStream_t st, long long *launches) {
    const long long n = bp.n;
    cudaError_t e;
    // grid from the box (one small D2H: the cell table size depends on it)
    if ((e = cudaMemcpyAsync(bp.h_box, bp.box, sizeof(long long) * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) { mn[k] = unord(bp.h_box[k]); mx[k] = unord(bp.h_box[3 + k]); }
    double rmax = unord(bp.h_box[6]);
    if (!(rmax > 0)) rmax = 1.0;
    if (!(mx[0] >= mn[0])) { for (int k = 0; k < 3; k++) { mn[k] = 0; mx[k] = 0; } }
    const double extent = std::max(mx[0] - mn[0], std::max(mx[1] - mn[1], mx[2] - mn[2]));
    const double e1 = extent * (1.0 / 16777216.0);
    Grid g;
    g.e1 = e1;
    double cell;
    auto dims = [&](double c, Grid &h) {
        h.nx = (int)floor((mx[0] - mn[0]) / c) + 1; h.ny = (int)floor((mx[1] - mn[1]) / c) + 1; h.nz = (int)floor((mx[2] - mn[2]) / c) + 1;
        return (double)h.nx * h.ny * h.nz;
    };
    if (bp.useRadix) {
        // cell edge = 1.3 x the minimum (2*Rmax): at a few % volume fill the minimum-size grid is ~10 cells
        // per body, and its table (memset + scattered writes + 14 lookups per body) costs more than the
        // extra distance tests of a slightly coarser grid (measured at 16 Mi spheres, 5 % fill; see profiles/)
        double scale = 1.3;
        if (const char *ev = getenv("CUBEZ_BP_CELL_SCALE")) { double sc = atof(ev); if (sc >= 1.0 && sc <= 8.0) scale = sc; }
        cell = 2.0 * rmax * BP_MARGIN * 1.0001 * scale;
        while (dims(cell, g) > (double)bp.cellCapacity) cell *= 1.26;   // coarser cells: more candidates, never fewer
        // one radix pass less when a slightly coarser grid brings the key width under a byte boundary
        for (int tries = 0; tries < 4; tries++) {
            const double c = (double)g.nx * g.ny * g.nz + 1.0;
            int kb = 1;
            while ((double)(1ull << kb) < c) kb++;
            const int over = kb % 8;   // bits above the last full byte
            if (over == 0 || over > 2) break;
            cell *= 1.26;
            dims(cell, g);
        }
    } else {
        // Counting sort.  y and z cells take the minimum edge (3x3 rows of cells around a body; the edge
        // covers the inflated float test of k_bp_sweep: 2*(Rmax*margin + 4*e1) plus the distance slack).
        // A row of x-cells is one contiguous range of the sorted array and k_bp_sweep looks up the exact
        // x-interval per body, so the x edge is free: it is set by the table budget (~cellsPerBody cells per
        // collider; the table is memset, hit by one atomic per body, scanned and streamed by the sweep).
        // Tests per body ~ 4.5 * (2R + ex) * eyz^2 * density: minimal eyz is optimal for any budget.
        double cpb = 1.0;
        if (const char *ev = getenv("CUBEZ_BP_CELLS_PER_BODY")) { double v = atof(ev); if (v >= 0.01 && v <= 64.0) cpb = v; }
        const double minEdge = 2.0 * (rmax * BP_MARGIN + 4.0 * e1) * 1.0001 + 4.0 * e1;
        const double target = std::min((double)bp.cellCapacity - 1.0, std::max(4096.0, cpb * (double)n));
        cell = minEdge;
        auto rows = [&](double c) { g.ny = (int)floor((mx[1] - mn[1]) / c) + 1; g.nz = (int)floor((mx[2] - mn[2]) / c) + 1; return (double)g.ny * g.nz; };
        if (rows(cell) > target) {
            const double ly = mx[1] - mn[1], lz = mx[2] - mn[2];
            if (ly > cell && lz > cell) cell = std::max(cell, sqrt(ly * lz / target));
            while (rows(cell) > target) cell *= 1.02;
        }
        const double lx = mx[0] - mn[0];
        double nxd = floor(target / ((double)g.ny * g.nz));
        nxd = std::min(nxd, floor(lx / (minEdge * 0.25)) + 1.0);   // finer than a quarter of the reach buys nothing
        g.nx = (int)std::max(1.0, std::min(nxd, 268435455.0));   // Entry::xlf keeps 29 bits for an x-cell
        const double ex = lx > 0 ? lx / g.nx : 1.0;
        g.invx = 1.0 / ex;
        g.rmaxInfl = (double)(float)((rmax * BP_MARGIN + 4.0 * e1) * 1.000002) * 1.000001;
    }
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2]; g.inv = 1.0 / cell;
    if (bp.useRadix) { g.invx = g.inv; g.rmaxInfl = rmax * BP_MARGIN; }
    bp.grid = g;
    const long long cells = (long long)g.nx * g.ny * g.nz;
    const unsigned nb = (unsigned)((n + 255) / 256);
    if (bp.trace) { for (int k = 0; k < 8; k++) if (!bp.tev[k]) cudaEventCreate(&bp.tev[k]); cudaEventRecord(bp.tev[0], st); }
    if (!bp.useRadix) {
        unsigned *cellStart = reinterpret_cast<unsigned *>(bp.cellRange);
        if ((e = cudaMemsetAsync(cellStart, 0, sizeof(unsigned) * (cells + 1), st)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(bp.counters, 0, sizeof(unsigned long long) * 2, st)) != cudaSuccess) return e;
        if (bp.trace) cudaEventRecord(bp.tev[1], st);
        k_bp_count<<<nb, 256, 0, st>>>(bp.bounds, n, g, cellStart, bp.keyRank);
        if (bp.trace) cudaEventRecord(bp.tev[2], st);
        int l = czs::exclusive_scan_u32(cellStart, cells + 1, bp.scanScratch, st);
        if (bp.trace) cudaEventRecord(bp.tev[3], st);
        k_bp_resolve<<<nb, 256, 0, st>>>(bp.keyRank, n, cellStart, bp.pos);
        k_bp_place<<<nb, 256, 0, st>>>(bp.bounds, n, g, bp.pos, bp.entries);
        if (bp.trace) cudaEventRecord(bp.tev[4], st);
        k_bp_sweep<<<nb, 256, 0, st>>>(bp.entries, n, g, cellStart, bp.pairs, bp.counters, bp.pairCapacity);
        if (launches) *launches += 4 + l;
        if (bp.trace) {
            cudaEventRecord(bp.tev[5], st);
            cudaEventSynchronize(bp.tev[5]);
            for (int k = 0; k < 5; k++) cudaEventElapsedTime(&bp.stageMs[k], bp.tev[k], bp.tev[k + 1]);
            fprintf(stderr, "[bp] cells %lld (%dx%dx%d) | memset %.3f count %.3f scan %.3f place %.3f sweep %.3f ms\n", cells, g.nx, g.ny, g.nz,
                    bp.stageMs[0], bp.stageMs[1], bp.stageMs[2], bp.stageMs[3], bp.stageMs[4]);
        }
        return cudaGetLastError();
    }
    k_bp_keys<<<nb, 256, 0, st>>>(bp.bounds, n, g, bp.sortCells.keys[0], bp.sortCells.vals[0]);
    if (bp.trace) cudaEventRecord(bp.tev[1], st);
    int bits = 8;
    while (bits < 32 && (1ll << bits) <= cells) bits += 8;   // keys <= cells (the inactive key)
    int cur = czs::radix_sort(bp.sortCells, n, bits, st, launches);
    if (bp.trace) cudaEventRecord(bp.tev[2], st);
    // only the occupancy bits are cleared: range entries are read for occupied cells only
    if ((e = cudaMemsetAsync(bp.occ, 0, sizeof(unsigned) * (cells / 32 + 1), st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(bp.counters, 0, sizeof(unsigned long long) * 2, st)) != cudaSuccess) return e;
    if (bp.trace) cudaEventRecord(bp.tev[3], st);
    k_bp_gather<<<nb, 256, 0, st>>>(bp.bounds, bp.sortCells.vals[cur], n, bp.sorted);
    if (bp.trace) cudaEventRecord(bp.tev[4], st);
    k_bp_cells<<<nb, 256, 0, st>>>(bp.sortCells.keys[cur], n, (unsigned)cells, bp.cellRange, bp.occ);
    if (bp.trace) cudaEventRecord(bp.tev[5], st);
    k_bp_pairs<<<nb, 256, 0, st>>>(bp.sorted, bp.sortCells.keys[cur], bp.sortCells.vals[cur], n, g, bp.cellRange, bp.occ, bp.pairs, bp.counters, bp.pairCapacity);
    if (launches) *launches += 4;
    if (bp.trace) {
        cudaEventRecord(bp.tev[6], st);
        cudaEventSynchronize(bp.tev[6]);
        for (int k = 0; k < 6; k++) cudaEventElapsedTime(&bp.stageMs[k], bp.tev[k], bp.tev[k + 1]);
        fprintf(stderr, "[bp] cells %lld (%dx%dx%d) bits %d | keys %.3f sort %.3f memset %.3f gather %.3f cells %.3f pairs %.3f ms\n", cells, g.nx, g.ny, g.nz,
                bits, bp.stageMs[0], bp.stageMs[1], bp.stageMs[2], bp.stageMs[3], bp.stageMs[4], bp.stageMs[5]);
    }
    return cudaGetLastError();
}

}  // namespace czbp
