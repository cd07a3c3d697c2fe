// cz_sort.cuh — hand-written device-wide exclusive scan and LSD radix sort (key/value pairs).
// Used by the sort-based broadphase (K2): cell keys -> sorted bodies, and contact keys ->
// canonical (reference append-order) contact list.  No CUB/Thrust.
//
// Radix sort: 8-bit digits, per pass  (1) k_rs_hist: per-CTA digit histogram (digit-major, so one
// exclusive scan gives every CTA its global base per digit)  (2) scan  (3) k_rs_scatter: stable
// in-CTA multi-split by warp match/ballot ranking, tile staged in shared memory in digit order so
// the global writes are coalesced runs per bucket.  Traffic per pass and element: read key (hist) +
// read key,value + write key,value.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace czs {

// ------------------------------------------------------------------------------------------------
// exclusive scan of uint32, n up to 2^31: per-CTA tiles of 4096, recursive on the tile totals
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned warp_incl_scan(unsigned v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// in-place exclusive scan of each tile; tile totals to `totals` (may be NULL when one tile)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles_u32(unsigned *data, long long n, unsigned *totals) {
    __shared__ unsigned warpTot[33];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    unsigned v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = base + k < n ? data[base + k] : 0u; sum += v[k]; }
    unsigned incl = warp_incl_scan(sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 31) warpTot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned w = warpTot[lane];
        unsigned wi = warp_incl_scan(w);
        warpTot[lane] = wi - w;
        if (lane == 31) warpTot[32] = wi;
    }
    __syncthreads();
    unsigned run = warpTot[warp] + incl - sum;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = warpTot[32];
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add_u32(unsigned *data, long long n, const unsigned *tileOffsets) {
    const unsigned off = tileOffsets[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) data[base + k] += off;
}

// scratch: at least scan_scratch_elems(n) unsigned
static inline long long scan_scratch_elems(long long n) {
    long long total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += n;
    }
    return total + 1;
}
// returns number of kernel launches
static inline int exclusive_scan_u32(unsigned *data, long long n, unsigned *scratch, cudaStream_t st) {
    if (n <= 0) return 0;
    const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        k_scan_tiles_u32<<<1, SCAN_THREADS, 0, st>>>(data, n, nullptr);
        return 1;
    }
    k_scan_tiles_u32<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(data, n, scratch);
    int launches = 1 + exclusive_scan_u32(scratch, tiles, scratch + tiles, st);
    k_scan_add_u32<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(data, n, scratch);
    return launches + 1;
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort of (key, uint32 value) pairs; K = uint32_t or uint64_t
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 2048 elements per CTA

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const K *keys, long long n, int shift, unsigned *hist, unsigned nblocks) {
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; k++) {
        long long i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];   // digit-major
}

template <typename K>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const K *keysIn, const unsigned *valsIn, K *keysOut, unsigned *valsOut, long long n,
                                                           int shift, const unsigned *histScanned, unsigned nblocks) {
    __shared__ unsigned warpCount[RS_THREADS / 32][256];   // 8 KB: per-warp digit counts, then per-warp bases
    __shared__ unsigned binStart[257];
    __shared__ K skey[RS_TILE];
    __shared__ unsigned sval[RS_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < 256 * (RS_THREADS / 32); d += RS_THREADS) (&warpCount[0][0])[d] = 0;
    __syncthreads();
    // warp w owns the contiguous segment [w*256, (w+1)*256) of the tile; round r covers 32 elements
    const long long base = (long long)blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    K key[RS_ITEMS];
    unsigned val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const long long i = base + r * 32 + lane;
        const bool valid = i < n;
        key[r] = valid ? keysIn[i] : (K)~(K)0;
        val[r] = valid ? valsIn[i] : 0u;
        const unsigned d = valid ? ((unsigned)(key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        unsigned old = 0;
        if (valid && lane == leader) { old = warpCount[warp][d]; warpCount[warp][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps, and the tile total
    {
        const int d = threadIdx.x;
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; w++) { unsigned c = warpCount[w][d]; warpCount[w][d] = run; run += c; }
        // exclusive scan of the 256 totals
        unsigned incl = warp_incl_scan(run);
        __shared__ unsigned wt[9];
        if (lane == 31) wt[warp] = incl;
        __syncthreads();
        if (threadIdx.x == 0) { unsigned a = 0; for (int w = 0; w < 8; w++) { unsigned t = wt[w]; wt[w] = a; a += t; } wt[8] = a; }
        __syncthreads();
        binStart[d] = wt[warp] + incl - run;
        if (d == 255) binStart[256] = wt[8];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const long long i = base + r * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)(key[r] >> shift) & 255u;
            const unsigned pos = binStart[d] + warpCount[warp][d] + rank[r];
            skey[pos] = key[r];
            sval[pos] = val[r];
        }
    }
    __syncthreads();
    const unsigned count = binStart[256];
    for (unsigned idx = threadIdx.x; idx < count; idx += RS_THREADS) {
        const K k = skey[idx];
        const unsigned d = (unsigned)(k >> shift) & 255u;
        const size_t g = (size_t)histScanned[(size_t)d * nblocks + blockIdx.x] + (idx - binStart[d]);
        keysOut[g] = k;
        valsOut[g] = sval[idx];
    }
}

template <typename K> struct RadixBuffers {
    K *keys[2];          // ping-pong
    unsigned *vals[2];
    unsigned *hist;      // 256 * nblocks
    unsigned *scanScratch;
    long long capacity;
};

static inline long long rs_blocks(long long n) { return (n + RS_TILE - 1) / RS_TILE; }

// Sorts the first n pairs of buf.keys[0]/vals[0] by the low `bits` bits of the key.  Returns the
// index (0/1) of the buffers holding the result; *launches accumulates kernel launches.
template <typename K>
static inline int radix_sort(RadixBuffers<K> &buf, long long n, int bits, cudaStream_t st, long long *launches) {
    int cur = 0;
    if (n <= 1) return cur;
    const unsigned nb = (unsigned)rs_blocks(n);
    for (int shift = 0; shift < bits; shift += 8) {
        k_rs_hist<K><<<nb, RS_THREADS, 0, st>>>(buf.keys[cur], n, shift, buf.hist, nb);
        int l = exclusive_scan_u32(buf.hist, 256ll * nb, buf.scanScratch, st);
        k_rs_scatter<K><<<nb, RS_THREADS, 0, st>>>(buf.keys[cur], buf.vals[cur], buf.keys[cur ^ 1], buf.vals[cur ^ 1], n, shift, buf.hist, nb);
        if (launches) *launches += 2 + l;
        cur ^= 1;
    }
    return cur;
}

template <typename K> static inline cudaError_t radix_alloc(RadixBuffers<K> &buf, long long capacity) {
    buf.capacity = capacity;
    const long long nb = rs_blocks(capacity);
    cudaError_t e;
    for (int k = 0; k < 2; k++) {
        if ((e = cudaMalloc(&buf.keys[k], sizeof(K) * (size_t)capacity)) != cudaSuccess) return e;
        if ((e = cudaMalloc(&buf.vals[k], sizeof(unsigned) * (size_t)capacity)) != cudaSuccess) return e;
    }
    if ((e = cudaMalloc(&buf.hist, sizeof(unsigned) * 256 * (size_t)nb)) != cudaSuccess) return e;
    return cudaMalloc(&buf.scanScratch, sizeof(unsigned) * (size_t)scan_scratch_elems(256 * nb));
}
template <typename K> static inline void radix_free(RadixBuffers<K> &buf) {
    for (int k = 0; k < 2; k++) { if (buf.keys[k]) cudaFree(buf.keys[k]); if (buf.vals[k]) cudaFree(buf.vals[k]); }
    if (buf.hist) cudaFree(buf.hist);
    if (buf.scanScratch) cudaFree(buf.scanScratch);
    buf = RadixBuffers<K>{};
}

}  // namespace czs
