// Operand-form microbenchmarks for the packed FP32 pipe on sm_100a (not part of the product library).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int CH = 8;
// MODE 0: FFMA2 acc[i] = x[i]*y[i] + acc[i]       (3 distinct pairs, nothing reusable between neighbours)
// MODE 1: FFMA2 acc[i] = x[i]*x[i] + acc[i]       (2 distinct pairs)
// MODE 2: FFMA2 acc[i] = x[i]*s    + acc[i]       (2 pairs + scalar broadcast)
// MODE 3: scalar FFMA, 3 distinct regs, 2 per slot
// MODE 4: FADD2 acc[i] = s + acc[i]  (scalar + pair)
// MODE 5: FMUL2 acc[i] = acc[i]*x[i]
// MODE 6: FFMA2 acc[i] = w*x[i] + acc[i]          (w shared by 3 consecutive, like the force loop's ax/ay/az)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, const float2* in, float s) {
    float2 acc[CH], x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { acc[i] = in[threadIdx.x + i]; x[i] = in[threadIdx.x + 32 + i]; y[i] = in[threadIdx.x + 64 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);
                if (MODE == 1) acc[i] = __ffma2_rn(x[i], x[i], acc[i]);
                if (MODE == 2) acc[i] = __ffma2_rn(x[i], make_float2(s, s), acc[i]);
                if (MODE == 3) { acc[i].x = __fmaf_rn(x[i].x, y[i].x, acc[i].x); acc[i].y = __fmaf_rn(x[i].y, y[i].y, acc[i].y); }
                if (MODE == 4) acc[i] = __fadd2_rn(make_float2(s, s), acc[i]);
                if (MODE == 5) acc[i] = __fmul2_rn(acc[i], x[i]);
                if (MODE == 6) acc[i] = __ffma2_rn(y[i / 3], x[i], acc[i]);
            }
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += acc[i].x + acc[i].y;
    if (r == 123.456f) out[0] = r;
}
template <int MODE>
void run(const char* name, float* out, const float2* in, int sms) {
    const int blocks = sms * 8, iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<blocks, 256>>>(out, iters, in, 1e-9f); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<MODE><<<blocks, 256>>>(out, iters, in, 1e-9f); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double slots = double(blocks) * 8 /*warps*/ * iters * 8.0 * CH;  // packed-slot warp-instructions
    const double cyc = best * 1e-3 * 1.965e9 * sms * 4;                      // SMSP-cycles available
    printf("%-52s %8.3f ms  %.3f cycles per packed slot (2 lanes-ops)\n", name, best, cyc / slots);
}
int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out; float2* in; CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&in, 8192)); CK(cudaMemset(in, 0, 8192));
    run<0>("0: FFMA2 x[i]*y[i]+acc[i]  (3 distinct pairs)", out, in, sms);
    run<1>("1: FFMA2 x[i]*x[i]+acc[i]  (2 distinct pairs)", out, in, sms);
    run<2>("2: FFMA2 x[i]*s+acc[i]     (2 pairs + scalar)", out, in, sms);
    run<3>("3: 2x scalar FFMA, 3 distinct regs", out, in, sms);
    run<4>("4: FADD2 s+acc[i]          (scalar + pair)", out, in, sms);
    run<5>("5: FMUL2 acc[i]*x[i]       (2 pairs)", out, in, sms);
    run<6>("6: FFMA2 w*x[i]+acc[i]     (w shared by 3)", out, in, sms);
    return 0;
}
