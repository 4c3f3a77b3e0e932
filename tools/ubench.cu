// Pipe-interaction microbenchmarks (not part of the product library): how do MUFU / LDS / operand forms
// affect the packed-FP32 issue rate on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int CH = 12;  // independent packed chains per thread

// MODE 0: 12 FFMA2 (all-register operands)
// MODE 1: 12 FFMA2 with one scalar-broadcast operand
// MODE 2: 3 FADD2 + 3 FMUL2 + 6 FFMA2
// MODE 3: MODE 0 + 2 MUFU per 12 packed
// MODE 4: MODE 0 + 1 MUFU per 12 packed
// MODE 5: MODE 0 + 4 MUFU per 12 packed
// MODE 6: MODE 0 + 1 LDS.128 per 24 packed
// MODE 7: 24 scalar FFMA + 2 MUFU
// MODE 8: 24 scalar FFMA
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, const float4* gsm) {
    __shared__ float4 sm[256];
    sm[threadIdx.x] = gsm[threadIdx.x];
    __syncthreads();
    float2 acc[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    float m[4] = {1.5f + threadIdx.x, 2.5f, 3.5f, 4.5f};
    float2 a2 = make_float2(a, a + 1e-7f), b2 = make_float2(b, b * 1.0001f);
    float4 lds = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5 || MODE == 6) {
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = __ffma2_rn(acc[i], a2, b2);
            }
            if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < CH; ++i) acc[i] = __ffma2_rn(acc[i], make_float2(a, a), b2);
            }
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < CH; i += 4) {
                    acc[i] = __fadd2_rn(acc[i], a2);
                    acc[i + 1] = __fmul2_rn(acc[i + 1], a2);
                    acc[i + 2] = __ffma2_rn(acc[i + 2], a2, b2);
                    acc[i + 3] = __ffma2_rn(acc[i + 3], a2, b2);
                }
            }
            if (MODE == 3 || MODE == 7) { m[0] = rsq(m[0]); m[1] = rsq(m[1]); }
            if (MODE == 4) { m[u & 3] = rsq(m[u & 3]); }
            if (MODE == 5) { m[0] = rsq(m[0]); m[1] = rsq(m[1]); m[2] = rsq(m[2]); m[3] = rsq(m[3]); }
            if (MODE == 6 && (u & 1)) { float4 t = sm[(it + u) & 255]; lds.x += t.x; }
            if (MODE == 7 || MODE == 8) {
#pragma unroll
                for (int i = 0; i < CH; ++i) { acc[i].x = __fmaf_rn(acc[i].x, a, b); acc[i].y = __fmaf_rn(acc[i].y, a, b); }
            }
        }
    }
    float s = lds.x + m[0] + m[1] + m[2] + m[3];
#pragma unroll
    for (int i = 0; i < CH; ++i) s += acc[i].x + acc[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int MODE>
void run(const char* name, float* out, const float4* gsm, int sms) {
    const int blocks = sms * 8, iters = 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<MODE><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f, gsm);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<MODE><<<blocks, 256>>>(out, iters, 1.0000001f, 1e-9f, gsm); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    // packed-op (or scalar-pair) slots per thread: iters * 8 * 12; each = 2 lanes
    const double lane_ops = double(blocks) * 256 * iters * 8.0 * CH * 2;
    printf("%-44s %8.3f ms  %.2f T lane-ops/s  (%.1f%% of 37.22 T = 148*128*1.965G)\n", name, best, lane_ops / (best * 1e-3) / 1e12,
           lane_ops / (best * 1e-3) / 37.22e12 * 100);
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out; float4* gsm; CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&gsm, 4096)); CK(cudaMemset(gsm, 0, 4096));
    run<0>("0: 12 FFMA2 reg operands", out, gsm, sms);
    run<1>("1: 12 FFMA2 scalar-broadcast operand", out, gsm, sms);
    run<2>("2: 3 FADD2 + 3 FMUL2 + 6 FFMA2", out, gsm, sms);
    run<3>("3: 12 FFMA2 + 2 MUFU", out, gsm, sms);
    run<4>("4: 12 FFMA2 + 1 MUFU", out, gsm, sms);
    run<5>("5: 12 FFMA2 + 4 MUFU", out, gsm, sms);
    run<6>("6: 24 FFMA2 + 1 LDS.128", out, gsm, sms);
    run<7>("7: 24 FFMA + 2 MUFU", out, gsm, sms);
    run<8>("8: 24 FFMA", out, gsm, sms);
    return 0;
}
