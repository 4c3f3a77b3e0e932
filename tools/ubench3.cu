// Feasibility microbenchmark (not part of the product library): a warp-systolic, Newton's-third-law inner loop.
// Each lane owns 2 i-pairs (4 bodies) and one travelling j-bundle (x,y,z,m + 3 packed reaction accumulators).
// Per step: 2 chains of 16 packed ops + 2 MUFU each (force on i AND reaction on j), then the bundle moves one lane.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float rsq(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int PAIRS, bool SHUFFLE, bool SCALAR_REACTION>
__global__ void __launch_bounds__(256) k(float* out, int iters, const float4* in) {
    const int lane = threadIdx.x & 31;
    float2 nx[PAIRS], ny[PAIRS], nz[PAIRS], mi[PAIRS], ax[PAIRS], ay[PAIRS], az[PAIRS];
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) {
        float4 a = in[threadIdx.x * 2 * PAIRS + 2 * q], b = in[threadIdx.x * 2 * PAIRS + 2 * q + 1];
        nx[q] = make_float2(-a.x, -b.x), ny[q] = make_float2(-a.y, -b.y), nz[q] = make_float2(-a.z, -b.z);
        mi[q] = make_float2(a.w, b.w);
        ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
    }
    float4 bj = in[4096 + threadIdx.x];
    float2 rx = make_float2(0.f, 0.f), ry = rx, rz = rx;  // packed reaction accumulators of the travelling j
    float sx = 0.f, sy = 0.f, sz = 0.f;                    // scalar variant
    const float2 eps2 = make_float2(1e-4f, 1e-4f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int s = 0; s < 32; ++s) {
#pragma unroll
            for (int q = 0; q < PAIRS; ++q) {
                const float2 dx = __fadd2_rn(make_float2(bj.x, bj.x), nx[q]);
                const float2 dy = __fadd2_rn(make_float2(bj.y, bj.y), ny[q]);
                const float2 dz = __fadd2_rn(make_float2(bj.z, bj.z), nz[q]);
                float2 r2 = __ffma2_rn(dz, dz, eps2);
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dx, dx, r2);
                const float2 ri = make_float2(rsq(r2.x), rsq(r2.y));
                const float2 ri3 = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                const float2 wi = __fmul2_rn(ri3, make_float2(bj.w, bj.w));
                const float2 wj = __fmul2_rn(ri3, mi[q]);
                ax[q] = __ffma2_rn(wi, dx, ax[q]);
                ay[q] = __ffma2_rn(wi, dy, ay[q]);
                az[q] = __ffma2_rn(wi, dz, az[q]);
                if (SCALAR_REACTION) {
                    sx = __fmaf_rn(wj.x, dx.x, sx); sx = __fmaf_rn(wj.y, dx.y, sx);
                    sy = __fmaf_rn(wj.x, dy.x, sy); sy = __fmaf_rn(wj.y, dy.y, sy);
                    sz = __fmaf_rn(wj.x, dz.x, sz); sz = __fmaf_rn(wj.y, dz.y, sz);
                } else {
                    rx = __ffma2_rn(wj, dx, rx);
                    ry = __ffma2_rn(wj, dy, ry);
                    rz = __ffma2_rn(wj, dz, rz);
                }
            }
            if (SHUFFLE) {
                const int src = (lane + 1) & 31;
                bj.x = __shfl_sync(0xffffffffu, bj.x, src); bj.y = __shfl_sync(0xffffffffu, bj.y, src);
                bj.z = __shfl_sync(0xffffffffu, bj.z, src); bj.w = __shfl_sync(0xffffffffu, bj.w, src);
                if (SCALAR_REACTION) {
                    sx = __shfl_sync(0xffffffffu, sx, src); sy = __shfl_sync(0xffffffffu, sy, src); sz = __shfl_sync(0xffffffffu, sz, src);
                } else {
                    rx.x = __shfl_sync(0xffffffffu, rx.x, src); rx.y = __shfl_sync(0xffffffffu, rx.y, src);
                    ry.x = __shfl_sync(0xffffffffu, ry.x, src); ry.y = __shfl_sync(0xffffffffu, ry.y, src);
                    rz.x = __shfl_sync(0xffffffffu, rz.x, src); rz.y = __shfl_sync(0xffffffffu, rz.y, src);
                }
            }
        }
    }
    float r = rx.x + rx.y + ry.x + ry.y + rz.x + rz.y + sx + sy + sz + bj.x;
#pragma unroll
    for (int q = 0; q < PAIRS; ++q) r += ax[q].x + ax[q].y + ay[q].x + ay[q].y + az[q].x + az[q].y;
    if (r == 123.456f) out[0] = r;
}

template <int PAIRS, bool SHUFFLE, bool SCALAR>
void run(const char* name, float* out, const float4* in, int sms, int ctas_per_sm) {
    const int blocks = sms * ctas_per_sm, iters = 200;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<PAIRS, SHUFFLE, SCALAR><<<blocks, 256>>>(out, iters, in); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<PAIRS, SHUFFLE, SCALAR><<<blocks, 256>>>(out, iters, in); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double directed = double(blocks) * 256 * iters * 32.0 * PAIRS * 2 /*bodies per pair*/ * 2 /*directions*/;
    const double cyc_per_step = best * 1e-3 * 1.965e9 * sms * 4 / (double(blocks) * 8 * iters * 32.0);
    printf("%-46s ctas/sm %d  %8.3f ms  %.3e directed int/s  (%.1f%% of 3.7225e12 = 20-flop roofline)  %.1f SMSP-cycles/step\n", name,
           ctas_per_sm, best, directed / (best * 1e-3), directed / (best * 1e-3) / 3.7225e12 * 100, cyc_per_step);
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out; float4* in; CK(cudaMalloc(&out, 256)); CK(cudaMalloc(&in, 16 * 8192)); 
    float4* h = (float4*)malloc(16 * 8192); srand(1);
    for (int i = 0; i < 8192; ++i) h[i] = make_float4(rand() / float(RAND_MAX), rand() / float(RAND_MAX), rand() / float(RAND_MAX), 1e-3f);
    CK(cudaMemcpy(in, h, 16 * 8192, cudaMemcpyHostToDevice));
    for (int c : {2, 3, 4}) {
        run<2, true, false>("2 pairs, shuffle, packed reaction", out, in, sms, c);
        run<2, false, false>("2 pairs, NO shuffle, packed reaction", out, in, sms, c);
        run<2, true, true>("2 pairs, shuffle, scalar reaction", out, in, sms, c);
        run<3, true, false>("3 pairs, shuffle, packed reaction", out, in, sms, c);
        run<1, true, false>("1 pair, shuffle, packed reaction", out, in, sms, c);
    }
    return 0;
}
