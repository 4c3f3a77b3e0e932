// Experiment (not part of the product library): force kernel variants for round 2, timed against the shipped kernel
// and checked against its output.
//   kDual     : two accumulator sets (even / odd j) per i-pair, folded together -> half the FFMA2 dependency chain
//   kXY       : accumulate (x,y) per body with the weight as a scalar-broadcast operand (2-cycle FFMA2) and z per
//               pair (3-cycle FFMA2): 7 cycles per (i-pair, j) guaranteed instead of 7..9 depending on operand reuse,
//               at the price of four register moves.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "force.cuh"

using namespace nb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int kWarps, int kTileJ, bool kDual, bool kXY>
__global__ void __launch_bounds__(kWarps * 32, 1) variant_kernel(const float4* bodies, int n_j, int n_i, float eps2s, float* acc_out) {
    constexpr int kPairs = 2, kCT = kWarps * 32, kI = 4, kTileI = kCT * kI, kFold = 32;
    using Ring = TileRing<kTileJ, kStages, kWarps>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    Ring ring;
    ring.attach(smem_raw, bodies, n_j);
    const int ntiles = ring.num_tiles();
    if (tid == 0) ring.init_barriers();
    __syncthreads();
    if (tid == 0) for (int t = 0; t < min(kLookahead, ntiles); ++t) ring.issue(t);
    const int tile_base = blockIdx.x * kTileI;
    float4 me[kI]; int li[kI];
    float2 nx[kPairs], ny[kPairs], nz[kPairs];
#pragma unroll
    for (int k = 0; k < kI; ++k) { li[k] = tile_base + k * kCT + tid; me[k] = bodies[min(li[k], n_i - 1)]; }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        nx[q] = make_float2(-me[2*q].x, -me[2*q+1].x); ny[q] = make_float2(-me[2*q].y, -me[2*q+1].y); nz[q] = make_float2(-me[2*q].z, -me[2*q+1].z);
    }
    double tot[kI][3];
#pragma unroll
    for (int k = 0; k < kI; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;
    const float2 eps2 = make_float2(eps2s, eps2s);
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + kLookahead < ntiles) ring.issue(t + kLookahead);
        const int count = ring.tile_count(t);
        const float4* __restrict__ tj = ring.tile(t);
        ring.wait(t);
        for (int jb = 0; jb + kFold <= count; jb += kFold) {
            float2 ax[2][kPairs], ay[2][kPairs], az[2][kPairs];   // [set][pair]; kXY: ax = (x,y) of body 0, ay = (x,y) of body 1
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int q = 0; q < kPairs; ++q) ax[s][q] = ay[s][q] = az[s][q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kFold; ++u) {
                const int s = kDual ? (u & 1) : 0;
                const float4 b = tj[jb + u];
                const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
                for (int q = 0; q < kPairs; ++q) {
                    const float2 dx = __fadd2_rn(bx, nx[q]), dy = __fadd2_rn(by, ny[q]), dz = __fadd2_rn(bz, nz[q]);
                    float2 r2 = __ffma2_rn(dz, dz, eps2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                    const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                    const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), __fmul2_rn(ri, bm));
                    if (kXY) {
                        ax[s][q] = __ffma2_rn(make_float2(w.x, w.x), make_float2(dx.x, dy.x), ax[s][q]);
                        ay[s][q] = __ffma2_rn(make_float2(w.y, w.y), make_float2(dx.y, dy.y), ay[s][q]);
                        az[s][q] = __ffma2_rn(w, dz, az[s][q]);
                    } else {
                        ax[s][q] = __ffma2_rn(w, dx, ax[s][q]); ay[s][q] = __ffma2_rn(w, dy, ay[s][q]); az[s][q] = __ffma2_rn(w, dz, az[s][q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                float2 X = ax[0][q], Y = ay[0][q], Z = az[0][q];
                if (kDual) { X = __fadd2_rn(X, ax[1][q]); Y = __fadd2_rn(Y, ay[1][q]); Z = __fadd2_rn(Z, az[1][q]); }
                if (kXY) {  // X = (x,y) of body 2q, Y = (x,y) of body 2q+1
                    tot[2*q][0] += double(X.x); tot[2*q][1] += double(X.y); tot[2*q+1][0] += double(Y.x); tot[2*q+1][1] += double(Y.y);
                } else {
                    tot[2*q][0] += double(X.x); tot[2*q+1][0] += double(X.y); tot[2*q][1] += double(Y.x); tot[2*q+1][1] += double(Y.y);
                }
                tot[2*q][2] += double(Z.x); tot[2*q+1][2] += double(Z.y);
            }
        }
        ring.release(t);
    }
#pragma unroll
    for (int k = 0; k < kI; ++k)
        if (li[k] < n_i) { acc_out[3*li[k]] = float(tot[k][0]); acc_out[3*li[k]+1] = float(tot[k][1]); acc_out[3*li[k]+2] = float(tot[k][2]); }
}

template <int kWarps, int kTileJ, bool kDual, bool kXY>
void run(const char* name, const float4* d, int n_j, int n_i, float* out, const std::vector<float>& ref) {
    auto k = variant_kernel<kWarps, kTileJ, kDual, kXY>;
    const size_t smem = TileRing<kTileJ, kStages, kWarps>::smem_bytes();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    const int grid = n_i / (kWarps * 32 * 4);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<grid, kWarps * 32, smem>>>(d, n_j, n_i, 1e-4f, out); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(e0)); k<<<grid, kWarps * 32, smem>>>(d, n_j, n_i, 1e-4f, out); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = fminf(best, ms); }
    std::vector<float> h(size_t(n_i) * 3); CK(cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0; for (int i = 0; i < 3000; ++i) { double e = fabs(h[i] - ref[i]) / (fabs(ref[i]) + 1e-30); if (e > worst) worst = e; }
    const double rate = double(n_i) * n_j / (best * 1e-3);
    printf("%-34s regs %3d  %8.3f ms  %.4e int/s  %.1f%% of 74.45TF   max rel diff vs shipped kernel %.1e\n", name, fa.numRegs, best, rate, rate * 20 / 74.45e12 * 100, worst);
}

int main() {
    const int n_j = 1 << 20; int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int n_i = sms * 2 * 2048;
    std::vector<float4> h(n_j); srand(1);
    for (auto& b : h) b = make_float4(rand() / float(RAND_MAX), rand() / float(RAND_MAX), rand() / float(RAND_MAX), 1.f / n_j);
    float4* d; float* out; CK(cudaMalloc(&d, sizeof(float4) * n_j)); CK(cudaMalloc(&out, sizeof(float) * 3 * n_j));
    CK(cudaMemcpy(d, h.data(), sizeof(float4) * n_j, cudaMemcpyHostToDevice));
    // reference output: the shipped kernel
    { auto k = force_kernel<2, 16, 1, 1024, false, 32, 32>; const size_t smem = TileRing<1024, kStages, 16>::smem_bytes();
      CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      ForceParams p{}; p.bodies = d; p.j_end = n_j; p.i_count = n_i; p.eps2 = 1e-4f; p.g = 1.f; p.splits_total = 1; p.mode = MODE_ACCEL; p.acc = out;
      k<<<dim3(n_i / 2048, 1), 512, smem>>>(p); CK(cudaDeviceSynchronize()); }
    std::vector<float> ref(3000); CK(cudaMemcpy(ref.data(), out, 3000 * 4, cudaMemcpyDeviceToHost));
    run<16, 1024, false, false>("shipped loop (re-stated)", d, n_j, n_i, out, ref);
    run<16, 1024, true, false>("dual accumulator sets", d, n_j, n_i, out, ref);
    run<16, 1024, false, true>("xy-packed accumulate", d, n_j, n_i, out, ref);
    run<16, 1024, true, true>("dual + xy-packed", d, n_j, n_i, out, ref);
    return 0;
}
