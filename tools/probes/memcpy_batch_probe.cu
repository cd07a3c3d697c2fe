// Per-copy cost of the cz_world_step_host transfer pattern: 6 chunks x 9 fields per direction as
//   (a) cudaMemcpyAsync per field and chunk, (b) one cudaMemcpyBatchAsync per chunk and direction,
//   (c) kernels reading / writing the pinned host arrays directly (zero-copy), both directions at once.
// nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/memcpy_batch_probe tools/probes/memcpy_batch_probe.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); return 1; } } while (0)
static const long long NB = 65536LL * 8;
static const int IN_W[9] = {24, 32, 24, 24, 24, 72, 8, 1, 1};       // bytes per body
static const int OUT_W[9] = {24, 32, 24, 24, 8, 24, 96, 72, 1};

__global__ void k_copy16(const uint4 *__restrict__ src, uint4 *__restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
    char *hi[9], *di[9], *ho[9], *dout[9];
    double bi = 0, bo = 0;
    for (int f = 0; f < 9; f++) {
        CK(cudaHostAlloc((void **)&hi[f], NB * IN_W[f], cudaHostAllocDefault)); CK(cudaMalloc((void **)&di[f], NB * IN_W[f]));
        CK(cudaHostAlloc((void **)&ho[f], NB * OUT_W[f], cudaHostAllocDefault)); CK(cudaMalloc((void **)&dout[f], NB * OUT_W[f]));
        memset(hi[f], 1, NB * IN_W[f]); memset(ho[f], 1, NB * OUT_W[f]);
        bi += (double)NB * IN_W[f]; bo += (double)NB * OUT_W[f];
    }
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    const int chunks = 6;
    auto timeit = [&](const char *name, auto fn, double bytes) {
        fn(); cudaDeviceSynchronize();
        const int R = 10;
        auto t0 = std::chrono::steady_clock::now();
        for (int r = 0; r < R; r++) fn();
        auto t1 = std::chrono::steady_clock::now();
        cudaDeviceSynchronize();
        auto t2 = std::chrono::steady_clock::now();
        const double el = std::chrono::duration<double>(t2 - t0).count() / R, enq = std::chrono::duration<double>(t1 - t0).count() / R;
        printf("%-44s %6.2f ms per frame, %6.1f GB/s, host enqueue %.2f ms  (%s)\n", name, el * 1e3, bytes / el / 1e9, enq * 1e3, cudaGetErrorString(cudaGetLastError()));
    };
    for (int dir = 0; dir < 3; dir++) {
        const bool up = dir != 1, down = dir != 0;
        const double bytes = (up ? bi : 0) + (down ? bo : 0);
        printf("--- up=%d down=%d\n", up, down);
        timeit("cudaMemcpyAsync per field and chunk", [&] {
            for (int c = 0; c < chunks; c++) {
                const long long a = NB * c / chunks / 16 * 16, b = c + 1 == chunks ? NB : NB * (c + 1) / chunks / 16 * 16;
                if (up) for (int f = 0; f < 9; f++) cudaMemcpyAsync(di[f] + a * IN_W[f], hi[f] + a * IN_W[f], (b - a) * IN_W[f], cudaMemcpyHostToDevice, s1);
                if (down) for (int f = 0; f < 9; f++) cudaMemcpyAsync(ho[f] + a * OUT_W[f], dout[f] + a * OUT_W[f], (b - a) * OUT_W[f], cudaMemcpyDeviceToHost, s2);
            }
        }, bytes);
        timeit("cudaMemcpyBatchAsync per chunk", [&] {
            for (int c = 0; c < chunks; c++) {
                const long long a = NB * c / chunks / 16 * 16, b = c + 1 == chunks ? NB : NB * (c + 1) / chunks / 16 * 16;
                void *dsts[9], *srcs[9]; size_t sizes[9]; size_t fail = 0, idx0 = 0;
                cudaMemcpyAttributes at{}; at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
                if (up) {
                    for (int f = 0; f < 9; f++) { dsts[f] = di[f] + a * IN_W[f]; srcs[f] = hi[f] + a * IN_W[f]; sizes[f] = (b - a) * IN_W[f]; }
                    cudaMemcpyBatchAsync(dsts, srcs, sizes, 9, &at, &idx0, 1, &fail, s1);
                }
                if (down) {
                    for (int f = 0; f < 9; f++) { dsts[f] = ho[f] + a * OUT_W[f]; srcs[f] = dout[f] + a * OUT_W[f]; sizes[f] = (b - a) * OUT_W[f]; }
                    cudaMemcpyBatchAsync(dsts, srcs, sizes, 9, &at, &idx0, 1, &fail, s2);
                }
            }
        }, bytes);
        for (int grid : {32, 148, 592}) {
            char nm[64]; snprintf(nm, sizeof nm, "zero-copy kernels per field and chunk, grid %d", grid);
            timeit(nm, [&] {
                for (int c = 0; c < chunks; c++) {
                    const long long a = NB * c / chunks / 16 * 16, b = c + 1 == chunks ? NB : NB * (c + 1) / chunks / 16 * 16;
                    if (up) for (int f = 0; f < 7; f++) k_copy16<<<grid, 256, 0, s1>>>((const uint4 *)(hi[f] + a * IN_W[f]), (uint4 *)(di[f] + a * IN_W[f]), (b - a) * IN_W[f] / 16);
                    if (down) for (int f = 0; f < 8; f++) k_copy16<<<grid, 256, 0, s2>>>((const uint4 *)(dout[f] + a * OUT_W[f]), (uint4 *)(ho[f] + a * OUT_W[f]), (b - a) * OUT_W[f] / 16);
                }
            }, bytes);
        }
        timeit("one cudaMemcpyAsync per field", [&] {
            if (up) for (int f = 0; f < 9; f++) cudaMemcpyAsync(di[f], hi[f], NB * IN_W[f], cudaMemcpyHostToDevice, s1);
            if (down) for (int f = 0; f < 9; f++) cudaMemcpyAsync(ho[f], dout[f], NB * OUT_W[f], cudaMemcpyDeviceToHost, s2);
        }, bytes);
    }
    return 0;
}
