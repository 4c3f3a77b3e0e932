// Tuning harness (not part of the product library): times force_kernel variants on random bodies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I nbody-deep-sim_b200/csrc -o tools/tune_force tools/tune_force.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "force.cuh"

using namespace nb;

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("%s: %s\n", #x, cudaGetErrorString(e));                         \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

template <int kPairs, int kWarps, int kMinB, int kTileJ, int kUnroll, int kFold>
void run(const char* name, const float4* bodies, int n_j, float* acc, int sms) {
    auto k = force_kernel<kPairs, kWarps, kMinB, kTileJ, false, kUnroll, kFold>;
    const size_t smem = TileRing<kTileJ, kStages, kWarps>::smem_bytes();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kWarps * 32, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    const int tile_i = kWarps * 32 * kPairs * 2;
    const int waves = 2;
    const int i_tiles = sms * occ * waves;
    const int n_i = i_tiles * tile_i;  // exact waves; i indices wrap onto the j array through i_begin = 0 (n_i <= n_j needed)
    if (n_i > n_j) {
        printf("%-28s skipped (n_i %d > n_j %d)\n", name, n_i, n_j);
        return;
    }
    ForceParams p{};
    p.bodies = bodies, p.j_begin = 0, p.j_end = n_j, p.i_begin = 0, p.i_count = n_i;
    p.eps2 = 1e-4f, p.g = 1.f, p.splits_total = 1, p.mode = MODE_ACCEL, p.acc = acc;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k<<<dim3(i_tiles, 1), kWarps * 32, smem>>>(p);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        k<<<dim3(i_tiles, 1), kWarps * 32, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double rate = double(n_i) * n_j / (best * 1e-3);
    printf("%-28s regs %3d occ %d n_i %7d  %8.3f ms  %.4e int/s  %.1f%% of 74.45TF\n", name, fa.numRegs, occ, n_i, best, rate,
           rate * 20 / 74.45e12 * 100);
}

int main(int argc, char** argv) {
    const int n_j = argc > 1 ? atoi(argv[1]) : 1 << 20;
    int sms;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<float4> h(n_j);
    srand(1);
    for (auto& b : h) b = make_float4(rand() / float(RAND_MAX), rand() / float(RAND_MAX), rand() / float(RAND_MAX), 1.f / n_j);
    float4* d;
    float* acc;
    CK(cudaMalloc(&d, sizeof(float4) * n_j));
    CK(cudaMalloc(&acc, sizeof(float) * 3 * n_j));
    CK(cudaMemcpy(d, h.data(), sizeof(float4) * n_j, cudaMemcpyHostToDevice));
#define RUN(P, W, B, T, U, F) run<P, W, B, T, U, F>("<" #P "," #W "," #B "," #T ",u" #U ",f" #F ">", d, n_j, acc, sms)
    RUN(2, 16, 1, 1024, 32, 32);
    RUN(2, 8, 2, 1024, 32, 32);
    RUN(2, 8, 3, 1024, 32, 32);
    RUN(2, 24, 1, 1024, 32, 32);
    RUN(2, 12, 2, 512, 32, 32);
    RUN(2, 16, 1, 1024, 8, 32);
    RUN(2, 16, 1, 1024, 4, 32);
    RUN(1, 8, 2, 512, 32, 32);
    RUN(1, 8, 4, 512, 32, 32);
    RUN(1, 16, 2, 512, 32, 32);
    return 0;
}
