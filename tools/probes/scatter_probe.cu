// Hardware probe: what do random 32-byte scattered writes / random 4-byte table reads cost on B200?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/scatter_probe tools/probes/scatter_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
// bijective scramble on 2^24: multiply by odd constant mod 2^24 then xor-shift
__device__ __forceinline__ unsigned perm24(unsigned i) { i = (i * 0x9E3779B1u) & 0xffffffu; i ^= i >> 12; i = (i * 0x85EBCA6Bu) & 0xffffffu; i ^= i >> 11; return i & 0xffffffu; }

__global__ void k_write256(unsigned long long *out, unsigned n, int mode) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned p = mode ? perm24(i) : i;
    unsigned long long a = i, b = p, c = 3, d = 4;
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * (size_t)p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__global__ void k_write2x128(uint4 *out, unsigned n, int mode) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned p = mode ? perm24(i) : i;
    out[2 * (size_t)p] = make_uint4(i, p, 1, 2);
    out[2 * (size_t)p + 1] = make_uint4(i, p, 3, 4);
}
__global__ void k_write128(uint4 *out, unsigned n, int mode) {   // 16-byte entries
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned p = mode ? perm24(i) : i;
    out[(size_t)p] = make_uint4(i, p, 1, 2);
}
__global__ void k_read4(const unsigned *tab, unsigned tabN, unsigned *out, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = tab[hash(i) % tabN];
}
__global__ void k_read256(const unsigned long long *in, unsigned *out, unsigned n, int mode) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned p = mode ? perm24(i) : i;
    unsigned long long a, b, c, d;
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(in + 4 * (size_t)p));
    out[i] = (unsigned)(a + b + c + d);
}
__global__ void k_atom(unsigned *tab, unsigned tabN, unsigned *out, unsigned n) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = atomicAdd(&tab[hash(i) % tabN], 1u);
}
// place-like kernel: flags 1 = stream-read 32 B/body (.cs), 2 = random table read (evict_last), 4 = random 256-bit store,
// 8 = the store carries an evict_first hint, 16 = table read is a plain load
__global__ void k_place_like(const double2 *in, const unsigned *tab, unsigned tabN, unsigned long long *out, unsigned *sink, unsigned n, int flags) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long a = i, b = 2, c = 3, d = 4;
    if (flags & 1) { double2 v0 = __ldcs(in + 2 * (size_t)i), v1 = __ldcs(in + 2 * (size_t)i + 1); a = __double_as_longlong(v0.x + v1.y); b = __double_as_longlong(v0.y + v1.x); }
    unsigned p = perm24(i);
    if (flags & 2) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        unsigned v;
        const unsigned *ad = tab + hash(i) % tabN;
        if (flags & 16) v = *ad;
        else asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(ad), "l"(pol));
        p ^= (v & 1u);
        c = v;
    }
    if (flags & 4) {
        if (flags & 8) {
            unsigned long long pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            asm volatile("st.global.L2::cache_hint.v4.b64 [%0], {%1,%2,%3,%4}, %5;" ::"l"(out + 4 * (size_t)p), "l"(a), "l"(b), "l"(c), "l"(d), "l"(pol) : "memory");
        } else {
            asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(out + 4 * (size_t)p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
        }
    } else if ((a ^ b ^ c) == 0x123456789ull) sink[0] = 1;
}
int main() {
    const unsigned n = 1u << 24;
    void *buf, *tab, *out;
    cudaMalloc(&buf, (size_t)n * 32); cudaMalloc(&tab, 1u << 30); cudaMalloc(&out, (size_t)n * 4);
    cudaMemset(buf, 0, (size_t)n * 32); cudaMemset(tab, 0, 1u << 30);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](const char *name, auto f, double bytes) {
        float best = 1e9;
        for (int r = 0; r < 4; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
        printf("%-34s %.3f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
    };
    const unsigned nb = n / 256;
    time("write256 sequential", [&] { k_write256<<<nb, 256>>>((unsigned long long *)buf, n, 0); }, n * 32.0);
    time("write256 random", [&] { k_write256<<<nb, 256>>>((unsigned long long *)buf, n, 1); }, n * 32.0);
    time("write2x128 random", [&] { k_write2x128<<<nb, 256>>>((uint4 *)buf, n, 1); }, n * 32.0);
    time("write128 random (16B entries)", [&] { k_write128<<<nb, 256>>>((uint4 *)buf, n, 1); }, n * 16.0);
    time("read256 sequential", [&] { k_read256<<<nb, 256>>>((unsigned long long *)buf, (unsigned *)out, n, 0); }, n * 36.0);
    time("read256 random", [&] { k_read256<<<nb, 256>>>((unsigned long long *)buf, (unsigned *)out, n, 1); }, n * 36.0);
    for (unsigned mb : {8u, 16u, 32u, 64u, 128u, 256u}) {
        char nm[64];
        snprintf(nm, sizeof nm, "read4 random, table %u MB", mb);
        time(nm, [&] { k_read4<<<nb, 256>>>((unsigned *)tab, mb << 18, (unsigned *)out, n); }, n * 8.0);
        snprintf(nm, sizeof nm, "atomicAdd random, table %u MB", mb);
        time(nm, [&] { k_atom<<<nb, 256>>>((unsigned *)tab, mb << 18, (unsigned *)out, n); }, n * 8.0);
    }
    void *in; cudaMalloc(&in, (size_t)n * 32); cudaMemset(in, 0, (size_t)n * 32);
    for (int flags : {1, 2, 4, 12, 5, 13, 6, 14, 7, 15, 31, 3}) {
        char nm[64];
        snprintf(nm, sizeof nm, "place-like flags=%d (64 MB table)", flags);
        time(nm, [&] { k_place_like<<<nb, 256>>>((const double2 *)in, (const unsigned *)tab, 16u << 20, (unsigned long long *)buf, (unsigned *)out, n, flags); }, n * 32.0);
    }
    return 0;
}
