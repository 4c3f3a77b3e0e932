// Round-2 experiment (not part of the product library): mass-folded force loop.
//
// j-records carry pre-scaled coordinates x' = x*q with q = m^(-1/2) and e = q*q*eps2, so that
//   d' = x' - q*x_i (FFMA2),  r2' = d'.d' + e (3 FFMA2),  ri = rsqrt(r2'),  w = ri*ri*ri (2 FMUL2),  a += w*d' (3 FFMA2)
// gives m * d / (r2+eps2)^(3/2) with 11 packed FMA-pipe instructions per (i-pair, j) instead of 12.
// Layouts (kLayout):
//   0  shipped loop (float4 x,y,z,m), for reference
//   1  8 floats per j: x',x',y',y',z',z',q,e   (pairs arrive duplicated: no register moves)
//   2  float4 (x',y',z',q) + float e in a second array
//   3  8 floats per j: x',y',z',z',q,e,-,-  with the (x,y) of each i-body packed in one register pair: the (x,y)
//      accumulates take the weight as a scalar-broadcast operand (2-cycle FFMA2), r2 is a scalar FFMA chain
// Output is checked against an FP64 host sum on a sample of i-bodies.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "async_copy.cuh"

using namespace nb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int kStagesX = 4, kLook = 2;

template <int kRec /*floats per j record*/, int kTileJ, int kWarps>
struct Ring {
    float* tiles; uint64_t* full; uint64_t* empty; const float* src; int count;
    static constexpr size_t smem_bytes() { return size_t(kStagesX) * kTileJ * kRec * 4 + 2 * kStagesX * 8; }
    __device__ void attach(unsigned char* smem, const float* s, int c) {
        tiles = reinterpret_cast<float*>(smem);
        full = reinterpret_cast<uint64_t*>(smem + size_t(kStagesX) * kTileJ * kRec * 4);
        empty = full + kStagesX; src = s; count = c;
    }
    __device__ void init() { for (int s = 0; s < kStagesX; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); } mbar_fence_init(); }
    __device__ int num_tiles() const { return (count + kTileJ - 1) / kTileJ; }
    __device__ int tile_count(int t) const { return min(kTileJ, count - t * kTileJ); }
    __device__ const float* tile(int t) const { return tiles + size_t(t % kStagesX) * kTileJ * kRec; }
    __device__ void issue(int t) {
        const int s = t % kStagesX;
        if (t >= kStagesX) mbar_wait(&empty[s], ((t / kStagesX) & 1) ^ 1);
        const uint32_t bytes = uint32_t(tile_count(t)) * kRec * 4;
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_copy_g2s(tiles + size_t(s) * kTileJ * kRec, src + size_t(t) * kTileJ * kRec, bytes, &full[s]);
    }
    __device__ void wait(int t) { mbar_wait(&full[t % kStagesX], (t / kStagesX) & 1); }
    __device__ void release(int t) { __syncwarp(); if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[t % kStagesX]); }
};

template <int kLayout, int kPairs, int kWarps, int kMinB, int kTileJ, int kFold>
__global__ void __launch_bounds__(kWarps * 32, kMinB)
fold_kernel(const float* __restrict__ jrec, const float* __restrict__ jrec2, const float4* __restrict__ ibodies, int n_j, int n_i,
            float eps2s, float* __restrict__ acc_out) {
    constexpr int kRec = (kLayout == 1 || kLayout == 3) ? 8 : 4;
    constexpr int kCT = kWarps * 32, kI = 2 * kPairs, kTileI = kCT * kI;
    using R = Ring<kRec, kTileJ, kWarps>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    R ring;
    ring.attach(smem_raw, jrec, n_j);
    const int ntiles = ring.num_tiles();
    if (tid == 0) ring.init();
    __syncthreads();
    if (tid == 0) for (int t = 0; t < min(kLook, ntiles); ++t) ring.issue(t);
    const int tile_base = blockIdx.x * kTileI;
    float4 me[kI]; int li[kI];
    float2 nx[kPairs], ny[kPairs], nz[kPairs];
#pragma unroll
    for (int k = 0; k < kI; ++k) { li[k] = tile_base + k * kCT + tid; me[k] = ibodies[min(li[k], n_i - 1)]; }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        nx[q] = make_float2(-me[2*q].x, -me[2*q+1].x); ny[q] = make_float2(-me[2*q].y, -me[2*q+1].y); nz[q] = make_float2(-me[2*q].z, -me[2*q+1].z);
        if (kLayout == 3) { nx[q] = make_float2(-me[2*q].x, -me[2*q].y); ny[q] = make_float2(-me[2*q+1].x, -me[2*q+1].y); }
    }
    double tot[kI][3];
#pragma unroll
    for (int k = 0; k < kI; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;
    const float2 eps2 = make_float2(eps2s, eps2s);
    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + kLook < ntiles) ring.issue(t + kLook);
        const int count = ring.tile_count(t);
        const float* __restrict__ tj = ring.tile(t);
        ring.wait(t);
        for (int jb = 0; jb + kFold <= count; jb += kFold) {
            float2 ax[kPairs], ay[kPairs], az[kPairs];
#pragma unroll
            for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kFold; ++u) {
                if (kLayout == 0) {
                    const float4 b = reinterpret_cast<const float4*>(tj)[jb + u];
                    const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) {
                        const float2 dx = __fadd2_rn(bx, nx[q]), dy = __fadd2_rn(by, ny[q]), dz = __fadd2_rn(bz, nz[q]);
                        float2 r2 = __ffma2_rn(dz, dz, eps2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                        const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                        const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), __fmul2_rn(ri, bm));
                        ax[q] = __ffma2_rn(w, dx, ax[q]); ay[q] = __ffma2_rn(w, dy, ay[q]); az[q] = __ffma2_rn(w, dz, az[q]);
                    }
                } else if (kLayout == 3) {
                    const float4 b0 = reinterpret_cast<const float4*>(tj)[2 * (jb + u)];
                    const float2 b1 = reinterpret_cast<const float2*>(tj)[4 * (jb + u) + 2];
                    const float2 bxy = make_float2(b0.x, b0.y), bzz = make_float2(b0.z, b0.w), qq = make_float2(b1.x, b1.x);
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) {
                        // nx[q] = (-x0,-y0), ny[q] = (-x1,-y1), nz[q] = (-z0,-z1); ax = (ax0,ay0), ay = (ax1,ay1), az = (az0,az1)
                        const float2 d0 = __ffma2_rn(qq, nx[q], bxy), d1 = __ffma2_rn(qq, ny[q], bxy), dz = __ffma2_rn(qq, nz[q], bzz);
                        float r0 = __fmaf_rn(dz.x, dz.x, b1.y); r0 = __fmaf_rn(d0.y, d0.y, r0); r0 = __fmaf_rn(d0.x, d0.x, r0);
                        float r1 = __fmaf_rn(dz.y, dz.y, b1.y); r1 = __fmaf_rn(d1.y, d1.y, r1); r1 = __fmaf_rn(d1.x, d1.x, r1);
                        const float2 ri = make_float2(rsqrt_approx(r0), rsqrt_approx(r1));
                        const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                        ax[q] = __ffma2_rn(make_float2(w.x, w.x), d0, ax[q]);
                        ay[q] = __ffma2_rn(make_float2(w.y, w.y), d1, ay[q]);
                        az[q] = __ffma2_rn(w, dz, az[q]);
                    }
                } else {
                    float2 xx, yy, zz, qq, ee;
                    if (kLayout == 1) {
                        const float4 b0 = reinterpret_cast<const float4*>(tj)[2 * (jb + u)];
                        const float4 b1 = reinterpret_cast<const float4*>(tj)[2 * (jb + u) + 1];
                        xx = make_float2(b0.x, b0.y); yy = make_float2(b0.z, b0.w); zz = make_float2(b1.x, b1.y);
                        qq = make_float2(b1.z, b1.z); ee = make_float2(b1.w, b1.w);
                    } else {
                        const float4 b = reinterpret_cast<const float4*>(tj)[jb + u];
                        xx = make_float2(b.x, b.x); yy = make_float2(b.y, b.y); zz = make_float2(b.z, b.z); qq = make_float2(b.w, b.w);
                        const float e = __fmul_rn(__fmul_rn(b.w, b.w), eps2s);
                        ee = make_float2(e, e);
                    }
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) {
                        const float2 dx = __ffma2_rn(qq, nx[q], xx), dy = __ffma2_rn(qq, ny[q], yy), dz = __ffma2_rn(qq, nz[q], zz);
                        float2 r2 = __ffma2_rn(dz, dz, ee); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                        const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                        const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                        ax[q] = __ffma2_rn(w, dx, ax[q]); ay[q] = __ffma2_rn(w, dy, ay[q]); az[q] = __ffma2_rn(w, dz, az[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                if (kLayout == 3) {
                    tot[2*q][0] += double(ax[q].x); tot[2*q][1] += double(ax[q].y); tot[2*q+1][0] += double(ay[q].x); tot[2*q+1][1] += double(ay[q].y);
                } else {
                    tot[2*q][0] += double(ax[q].x); tot[2*q+1][0] += double(ax[q].y); tot[2*q][1] += double(ay[q].x); tot[2*q+1][1] += double(ay[q].y);
                }
                tot[2*q][2] += double(az[q].x); tot[2*q+1][2] += double(az[q].y);
            }
        }
        ring.release(t);
    }
#pragma unroll
    for (int k = 0; k < kI; ++k)
        if (li[k] < n_i) { acc_out[3*li[k]] = float(tot[k][0]); acc_out[3*li[k]+1] = float(tot[k][1]); acc_out[3*li[k]+2] = float(tot[k][2]); }
}

struct Data { float *rec4, *rec8, *recs4, *rec8b, *e; float4* bodies; float* acc; int n_j; std::vector<float4> h; };

template <int kLayout, int kPairs, int kWarps, int kMinB, int kTileJ, int kFold>
void run(const char* name, Data& d, int sms, float eps2) {
    auto k = fold_kernel<kLayout, kPairs, kWarps, kMinB, kTileJ, kFold>;
    constexpr int kRec = (kLayout == 1 || kLayout == 3) ? 8 : 4;
    const size_t smem = Ring<kRec, kTileJ, kWarps>::smem_bytes();
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kWarps * 32, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    const int tile_i = kWarps * 32 * kPairs * 2, i_tiles = sms * occ * 2, n_i = i_tiles * tile_i;
    if (n_i > d.n_j) { printf("%-34s skipped\n", name); return; }
    const float* rec = kLayout == 0 ? d.rec4 : kLayout == 1 ? d.rec8 : kLayout == 3 ? d.rec8b : d.recs4;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k<<<i_tiles, kWarps * 32, smem>>>(rec, d.e, d.bodies, d.n_j, n_i, eps2, d.acc); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0)); k<<<i_tiles, kWarps * 32, smem>>>(rec, d.e, d.bodies, d.n_j, n_i, eps2, d.acc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    // accuracy on 64 sampled i-bodies vs FP64
    std::vector<float> acc(3 * size_t(n_i)); CK(cudaMemcpy(acc.data(), d.acc, acc.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int s = 0; s < 64; ++s) {
        const int i = int((long long)s * 7919 % n_i);
        double a[3] = {0, 0, 0};
        for (int j = 0; j < d.n_j; ++j) {
            const double dx = double(d.h[j].x) - d.h[i].x, dy = double(d.h[j].y) - d.h[i].y, dz = double(d.h[j].z) - d.h[i].z;
            const double r2 = dx * dx + dy * dy + dz * dz + double(eps2); const double w = d.h[j].w / (r2 * sqrt(r2));
            a[0] += w * dx; a[1] += w * dy; a[2] += w * dz;
        }
        const double num = sqrt(pow(acc[3*i] - a[0], 2) + pow(acc[3*i+1] - a[1], 2) + pow(acc[3*i+2] - a[2], 2));
        const double den = sqrt(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]);
        if (num / den > worst) worst = num / den;
    }
    const double rate = double(n_i) * d.n_j / (best * 1e-3);
    printf("%-34s regs %3d occ %d  %8.3f ms  %.4e int/s  %.1f%% of 74.45TF  max rel err %.2e\n", name, fa.numRegs, occ, best, rate,
           rate * 20 / 74.45e12 * 100, worst);
}

int main(int argc, char** argv) {
    const int n_j = argc > 1 ? atoi(argv[1]) : 1 << 20;
    const float eps2 = 0.0025f;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    Data d; d.n_j = n_j; d.h.resize(n_j);
    srand(1);
    // disk-like cloud: exponential radii, masses spanning 4 decades, one heavy body away from the origin
    for (auto& b : d.h) {
        const float r = -3.f * logf(1.f - 0.999f * rand() / float(RAND_MAX)), th = 6.2831853f * rand() / float(RAND_MAX);
        const float mexp = -4.f * rand() / float(RAND_MAX);
        b = make_float4(r * cosf(th) + 10.f, r * sinf(th), 0.3f * (rand() / float(RAND_MAX) - 0.5f), powf(10.f, mexp) / n_j);
    }
    d.h[0] = make_float4(10.f, 0.f, 0.f, 0.01f);
    std::vector<float> r8(size_t(n_j) * 8), r8b(size_t(n_j) * 8), s4(size_t(n_j) * 4), ev(n_j);
    for (int j = 0; j < n_j; ++j) {
        const float q = 1.f / sqrtf(d.h[j].w);
        const float x = d.h[j].x * q, y = d.h[j].y * q, z = d.h[j].z * q, ee = q * q * eps2;
        float* p = &r8[size_t(j) * 8]; p[0] = p[1] = x; p[2] = p[3] = y; p[4] = p[5] = z; p[6] = q; p[7] = ee;
        float* pb = &r8b[size_t(j) * 8]; pb[0] = x; pb[1] = y; pb[2] = pb[3] = z; pb[4] = q; pb[5] = ee; pb[6] = pb[7] = 0.f;
        float* s = &s4[size_t(j) * 4]; s[0] = x; s[1] = y; s[2] = z; s[3] = q; ev[j] = ee;
    }
    CK(cudaMalloc(&d.bodies, 16 * size_t(n_j))); CK(cudaMalloc(&d.rec8, 32 * size_t(n_j))); CK(cudaMalloc(&d.recs4, 16 * size_t(n_j))); CK(cudaMalloc(&d.rec8b, 32 * size_t(n_j)));
    CK(cudaMalloc(&d.e, 4 * size_t(n_j))); CK(cudaMalloc(&d.acc, 12 * size_t(n_j)));
    d.rec4 = reinterpret_cast<float*>(d.bodies);
    CK(cudaMemcpy(d.bodies, d.h.data(), 16 * size_t(n_j), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.rec8, r8.data(), 32 * size_t(n_j), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.rec8b, r8b.data(), 32 * size_t(n_j), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.recs4, s4.data(), 16 * size_t(n_j), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d.e, ev.data(), 4 * size_t(n_j), cudaMemcpyHostToDevice));
#define RUN(L, P, W, B, T, F) run<L, P, W, B, T, F>("<L" #L ",p" #P ",w" #W ",b" #B ",t" #T ",f" #F ">", d, sms, eps2)
    RUN(0, 2, 16, 1, 1024, 32);
    RUN(1, 2, 16, 1, 1024, 32);
    RUN(1, 2, 16, 1, 512, 32);
    RUN(3, 2, 16, 1, 1024, 32);
    RUN(3, 2, 16, 1, 512, 32);
    RUN(3, 2, 8, 2, 512, 32);
    RUN(2, 2, 16, 1, 1024, 32);
    RUN(1, 2, 16, 1, 1024, 64);
    RUN(0, 2, 16, 1, 1024, 64);
    RUN(1, 3, 8, 1, 512, 32);
    RUN(1, 2, 8, 2, 512, 32);
    RUN(1, 1, 8, 2, 512, 32);
    RUN(1, 1, 8, 4, 512, 32);
    RUN(0, 1, 8, 2, 512, 32);
    return 0;
}
