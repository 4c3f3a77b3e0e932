// Round-2 experiment (not part of the product library): Newton's-third-law ("pair") force kernel prototype.
//
// Every unordered pair of bodies is evaluated once; the force on i accumulates in registers, the reaction on j travels
// around the warp with its j-body (systolic rotation: lane L meets j-body (L+s)&31 at step s, the three reaction sums
// move one lane per step by shuffle, the j coordinates come from shared memory). After 32 steps each lane holds the
// reaction of one j-body summed over the warp's 32*kI i-bodies and adds it to an FP64 buffer in shared memory; warps
// walk the 32-body blocks of a j-tile in a rotated order with one named barrier per round, so no two warps touch the
// same block at once and no atomics are needed inside the CTA. Per j-tile the buffer is flushed to FP64 global
// accumulators with RED.ADD.F64; per work item the i-sums are flushed the same way.
//
// Work items (I-tile, run of J-tiles) are pulled from a global counter by persistent CTAs. J-tiles above the I-tile are
// symmetric items; the J-tiles inside the I-tile's own range (the diagonal) are directed items (no reaction).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "async_copy.cuh"

using namespace nb;
#define CK(x) do { cudaError_t err_ = (x); if (err_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(err_)); exit(1);} } while (0)

struct Item { int i_tile, j_tile_begin, j_tile_end, sym; };

constexpr int kSt = 4, kLa = 2;

template <int kTileJ, int kWarps>
struct JRing {
    float4* tiles; uint64_t* full; uint64_t* empty;
    __host__ __device__ static constexpr size_t bytes() { return size_t(kSt) * kTileJ * 16 + 2 * kSt * 8; }
    __device__ void attach(unsigned char* smem) {
        tiles = reinterpret_cast<float4*>(smem);
        full = reinterpret_cast<uint64_t*>(smem + size_t(kSt) * kTileJ * 16); empty = full + kSt;
    }
    __device__ void init() { for (int s = 0; s < kSt; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); } mbar_fence_init(); }
    // `seq` is a CTA-lifetime running tile counter (the ring never restarts between work items)
    __device__ void issue(unsigned seq, const float4* src, int count) {
        const int s = seq % kSt;
        if (seq >= kSt) mbar_wait(&empty[s], ((seq / kSt) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], uint32_t(count) * 16);
        bulk_copy_g2s(tiles + size_t(s) * kTileJ, src, uint32_t(count) * 16, &full[s]);
    }
    __device__ const float4* tile(unsigned seq) const { return tiles + size_t(seq % kSt) * kTileJ; }
    __device__ void wait(unsigned seq) { mbar_wait(&full[seq % kSt], (seq / kSt) & 1); }
    __device__ void release(unsigned seq) { __syncwarp(); if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[seq % kSt]); }
};

template <int kPairs, int kWarps, int kMinB>
__global__ void __launch_bounds__(kWarps * 32, kMinB)
pair_kernel(const float4* __restrict__ bodies, int n, float eps2s, const Item* __restrict__ items, int n_items,
            unsigned* __restrict__ counter, double* __restrict__ acc64 /* [n][3] */) {
    constexpr int kCT = kWarps * 32, kI = 2 * kPairs, kTileI = kCT * kI, kTileJ = kWarps * 32;
    using Ring = JRing<kTileJ, kWarps>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* react = reinterpret_cast<double*>(smem_raw + Ring::bytes());  // [2][kTileJ][3]
    __shared__ int s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ring ring; ring.attach(smem_raw);
    if (tid == 0) ring.init();
    for (int k = tid; k < 2 * kTileJ * 3; k += kCT) react[k] = 0.0;
    __syncthreads();
    const float2 eps2 = make_float2(eps2s, eps2s);
    unsigned seq_issue = 0, seq_use = 0;  // producer / consumer running tile counters (identical sequences)
    unsigned rbuf = 0;

    for (;;) {
        if (tid == 0) s_item = int(atomicAdd(counter, 1u));
        __syncthreads();
        const int it = s_item;
        __syncthreads();
        if (it >= n_items) break;
        const Item item = items[it];
        const int ntiles = item.j_tile_end - item.j_tile_begin;
        if (tid == 0)
            for (int t = 0; t < min(kLa, ntiles); ++t) {
                const int j0 = (item.j_tile_begin + t) * kTileJ;
                ring.issue(seq_issue++, bodies + j0, min(kTileJ, n - j0));
            }
        float4 me[kI];
        float2 nx[kPairs], ny[kPairs], nz[kPairs], mi[kPairs];
        int gi[kI];
#pragma unroll
        for (int k = 0; k < kI; ++k) {  // component-wise scalar loads: see csrc/pair.cuh (no per-use MOV re-packing)
            gi[k] = item.i_tile * kTileI + k * kCT + tid;
            const float* src = reinterpret_cast<const float*>(bodies + min(gi[k], n - 1));
            me[k] = make_float4(ldg_f32(src), ldg_f32(src + 1), ldg_f32(src + 2), ldg_f32(src + 3));
            if (gi[k] >= n) me[k].w = 0.f;
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            nx[q] = make_float2(-me[2*q].x, -me[2*q+1].x); ny[q] = make_float2(-me[2*q].y, -me[2*q+1].y);
            nz[q] = make_float2(-me[2*q].z, -me[2*q+1].z); mi[q] = make_float2(me[2*q].w, me[2*q+1].w);
        }
        double tot[kI][3];
#pragma unroll
        for (int k = 0; k < kI; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;

        for (int t = 0; t < ntiles; ++t) {
            if (tid == 0 && t + kLa < ntiles) {
                const int j0 = (item.j_tile_begin + t + kLa) * kTileJ;
                ring.issue(seq_issue++, bodies + j0, min(kTileJ, n - j0));
            }
            const int jt0 = (item.j_tile_begin + t) * kTileJ;
            const int count = min(kTileJ, n - jt0);
            const float4* __restrict__ tj = ring.tile(seq_use);
            ring.wait(seq_use);
            if (item.sym) {
                double* rb = react + size_t(rbuf) * kTileJ * 3;
                for (int r = 0; r < kWarps; ++r) {
                    const int jblk = (warp + r) % kWarps;
                    const float4* __restrict__ blk = tj + jblk * 32;
                    if (jblk * 32 < count) {
                        float2 ax[kPairs], ay[kPairs], az[kPairs];
#pragma unroll
                        for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
                        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                        for (int s = 0; s < 32; ++s) {
                            const int jl = (lane + s) & 31;
                            const float4 b = blk[jl];  // prototype: n is a multiple of 32, no ragged tail inside a block
                            const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
                            for (int q = 0; q < kPairs; ++q) {
                                const float2 dx = __fadd2_rn(bx, nx[q]), dy = __fadd2_rn(by, ny[q]), dz = __fadd2_rn(bz, nz[q]);
                                float2 r2 = __ffma2_rn(dz, dz, eps2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                                const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                                const float2 ri3 = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                                const float2 wi = __fmul2_rn(ri3, bm), wj = __fmul2_rn(ri3, mi[q]);
                                ax[q] = __ffma2_rn(wi, dx, ax[q]); ay[q] = __ffma2_rn(wi, dy, ay[q]); az[q] = __ffma2_rn(wi, dz, az[q]);
                                sx = __fmaf_rn(-wj.x, dx.x, sx); sx = __fmaf_rn(-wj.y, dx.y, sx);
                                sy = __fmaf_rn(-wj.x, dy.x, sy); sy = __fmaf_rn(-wj.y, dy.y, sy);
                                sz = __fmaf_rn(-wj.x, dz.x, sz); sz = __fmaf_rn(-wj.y, dz.y, sz);
                            }
                            const int src = (lane + 1) & 31;
                            sx = __shfl_sync(0xffffffffu, sx, src); sy = __shfl_sync(0xffffffffu, sy, src); sz = __shfl_sync(0xffffffffu, sz, src);
                        }
                        // lane L now holds the reaction on j-body jblk*32 + L from this warp's i-bodies
                        double* rj = rb + size_t(jblk * 32 + lane) * 3;
                        rj[0] += double(sx); rj[1] += double(sy); rj[2] += double(sz);
#pragma unroll
                        for (int q = 0; q < kPairs; ++q) {
                            tot[2*q][0] += double(ax[q].x); tot[2*q+1][0] += double(ax[q].y); tot[2*q][1] += double(ay[q].x);
                            tot[2*q+1][1] += double(ay[q].y); tot[2*q][2] += double(az[q].x); tot[2*q+1][2] += double(az[q].y);
                        }
                    }
                    compute_barrier<kCT>();  // rounds in lockstep: no two warps on one j-block
                }
                // flush this tile's reactions; the other buffer serves the next tile
                for (int k = tid; k < count * 3; k += kCT) {
                    const double v = rb[k];
                    rb[k] = 0.0;
                    atomicAdd(&acc64[size_t(jt0) * 3 + k], v);
                }
                rbuf ^= 1;
            } else {
                for (int jb = 0; jb < count; jb += 32) {
                    float2 ax[kPairs], ay[kPairs], az[kPairs];
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
                    const int lim = min(32, count - jb);
                    if (lim == 32) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) {
                            const float4 b = tj[jb + u];
                            const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
                            for (int q = 0; q < kPairs; ++q) {
                                const float2 dx = __fadd2_rn(bx, nx[q]), dy = __fadd2_rn(by, ny[q]), dz = __fadd2_rn(bz, nz[q]);
                                float2 r2 = __ffma2_rn(dz, dz, eps2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                                const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                                const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), __fmul2_rn(ri, bm));
                                ax[q] = __ffma2_rn(w, dx, ax[q]); ay[q] = __ffma2_rn(w, dy, ay[q]); az[q] = __ffma2_rn(w, dz, az[q]);
                            }
                        }
                    } else {
                        for (int u = 0; u < lim; ++u) {
                            const float4 b = tj[jb + u];
                            const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z), bm = make_float2(b.w, b.w);
#pragma unroll
                            for (int q = 0; q < kPairs; ++q) {
                                const float2 dx = __fadd2_rn(bx, nx[q]), dy = __fadd2_rn(by, ny[q]), dz = __fadd2_rn(bz, nz[q]);
                                float2 r2 = __ffma2_rn(dz, dz, eps2); r2 = __ffma2_rn(dy, dy, r2); r2 = __ffma2_rn(dx, dx, r2);
                                const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                                const float2 w = __fmul2_rn(__fmul2_rn(ri, ri), __fmul2_rn(ri, bm));
                                ax[q] = __ffma2_rn(w, dx, ax[q]); ay[q] = __ffma2_rn(w, dy, ay[q]); az[q] = __ffma2_rn(w, dz, az[q]);
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) {
                        tot[2*q][0] += double(ax[q].x); tot[2*q+1][0] += double(ax[q].y); tot[2*q][1] += double(ay[q].x);
                        tot[2*q+1][1] += double(ay[q].y); tot[2*q][2] += double(az[q].x); tot[2*q+1][2] += double(az[q].y);
                    }
                }
            }
            ring.release(seq_use);
            ++seq_use;
        }
#pragma unroll
        for (int k = 0; k < kI; ++k)
            if (gi[k] < n) {
                atomicAdd(&acc64[size_t(gi[k]) * 3 + 0], tot[k][0]); atomicAdd(&acc64[size_t(gi[k]) * 3 + 1], tot[k][1]);
                atomicAdd(&acc64[size_t(gi[k]) * 3 + 2], tot[k][2]);
            }
    }
}

template <int kPairs, int kWarps, int kMinB>
void run(const char* name, const float4* d_bodies, const std::vector<float4>& h, int n, float eps2, int sms, int chunk_tiles) {
    constexpr int kTileI = kWarps * 32 * 2 * kPairs, kTileJ = kWarps * 32;
    auto k = pair_kernel<kPairs, kWarps, kMinB>;
    const size_t smem = JRing<kTileJ, kWarps>::bytes() + size_t(2) * kTileJ * 3 * 8;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    int occ = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, kWarps * 32, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    // work items, largest first
    std::vector<Item> items;
    const int i_tiles = (n + kTileI - 1) / kTileI, j_tiles = (n + kTileJ - 1) / kTileJ, per = kTileI / kTileJ;
    for (int a = 0; a < i_tiles; ++a) {
        const int d0 = a * per, d1 = std::min(j_tiles, (a + 1) * per);
        for (int b = d1; b < j_tiles; b += chunk_tiles) items.push_back({a, b, std::min(j_tiles, b + chunk_tiles), 1});
        items.push_back({a, d0, d1, 0});
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& x, const Item& y) {
        return (x.j_tile_end - x.j_tile_begin) * (x.sym ? 16 : 10) > (y.j_tile_end - y.j_tile_begin) * (y.sym ? 16 : 10); });
    Item* d_items; unsigned* d_counter; double* d_acc;
    CK(cudaMalloc(&d_items, items.size() * sizeof(Item))); CK(cudaMalloc(&d_counter, 4)); CK(cudaMalloc(&d_acc, size_t(n) * 24));
    CK(cudaMemcpy(d_items, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaMemset(d_counter, 0, 4)); CK(cudaMemset(d_acc, 0, size_t(n) * 24));
        CK(cudaEventRecord(e0));
        k<<<sms * occ, kWarps * 32, smem>>>(d_bodies, n, eps2, d_items, int(items.size()), d_counter, d_acc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0 && ms < best) best = ms;
    }
    std::vector<double> acc(size_t(n) * 3); CK(cudaMemcpy(acc.data(), d_acc, acc.size() * 8, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (int s = 0; s < 48; ++s) {
        const int i = s < 4 ? (s & 1 ? n - 1 - s : s) : int((long long)s * 104729 % n);
        double a[3] = {0, 0, 0};
        for (int j = 0; j < n; ++j) {
            if (j == i) continue;
            const double dx = double(h[j].x) - h[i].x, dy = double(h[j].y) - h[i].y, dz = double(h[j].z) - h[i].z;
            const double r2 = dx * dx + dy * dy + dz * dz + double(eps2); const double w = h[j].w / (r2 * sqrt(r2));
            a[0] += w * dx; a[1] += w * dy; a[2] += w * dz;
        }
        const double num = sqrt(pow(acc[3*i] - a[0], 2) + pow(acc[3*i+1] - a[1], 2) + pow(acc[3*i+2] - a[2], 2));
        const double den = sqrt(a[0]*a[0] + a[1]*a[1] + a[2]*a[2]);
        if (num / den > worst) worst = num / den;
    }
    const double rate = double(n) * n / (best * 1e-3);
    printf("%-28s regs %3d occ %d items %6zu chunk %2d  %8.3f ms  %.4e int/s  %.1f%% of 74.45TF  max rel err %.2e\n", name, fa.numRegs, occ,
           items.size(), chunk_tiles, best, rate, rate * 20 / 74.45e12 * 100, worst);
    cudaFree(d_items); cudaFree(d_counter); cudaFree(d_acc);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1 << 20;
    const float eps2 = 0.0025f;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    std::vector<float4> h(n);
    srand(1);
    for (auto& b : h) {
        const float r = -3.f * logf(1.f - 0.999f * rand() / float(RAND_MAX)), th = 6.2831853f * rand() / float(RAND_MAX);
        const float mexp = -4.f * rand() / float(RAND_MAX);
        b = make_float4(r * cosf(th), r * sinf(th), 0.3f * (rand() / float(RAND_MAX) - 0.5f), powf(10.f, mexp) / n);
    }
    h[0] = make_float4(0.f, 0.f, 0.f, 0.01f);
    float4* d; CK(cudaMalloc(&d, 16 * size_t(n))); CK(cudaMemcpy(d, h.data(), 16 * size_t(n), cudaMemcpyHostToDevice));
#define RUN(P, W, B, C) run<P, W, B>("<p" #P ",w" #W ",b" #B ">", d, h, n, eps2, sms, C)
    const int c1 = n >= (1 << 20) ? 64 : n >= 262144 ? 20 : n >= 65536 ? 5 : 2;
    RUN(3, 4, 2, c1);
    RUN(3, 4, 2, c1 / 2 > 0 ? c1 / 2 : 1);
    RUN(3, 4, 2, c1 * 2);
    RUN(2, 4, 3, c1);
    RUN(2, 4, 3, c1 * 2);
    RUN(2, 4, 4, c1);
    RUN(3, 2, 4, c1 * 2);
    RUN(3, 2, 5, c1 * 2);
    RUN(4, 4, 2, c1);
    RUN(4, 2, 4, c1 * 2);
    return 0;
}
