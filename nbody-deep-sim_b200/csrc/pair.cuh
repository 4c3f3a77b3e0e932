// Newton's-third-law ("pair") force kernel for large N: every unordered pair of bodies is evaluated ONCE.
//
// Replaces (reference, read-only): src/galaxify/simulation.py:71-89 (compute_accelerations), like force.cuh; the
// integrator of :153-187 runs in finish_kernel below with the same rounding-faithful epilogue (epilogue_body).
//
// Why: the directed loop of force.cuh issues 12 packed FMA-pipe instructions per two interactions and is bound by
// register-file operand bandwidth (48 operand words per 24 issue cycles; profiles/r2_tune_force3.log shows that
// cutting instructions without cutting operand words does not help). Evaluating m_i m_j d / (r^2+eps^2)^(3/2) once per
// unordered pair and using it for both bodies costs 16 packed-equivalent instructions per FOUR directed interactions.
//
// How a CTA works on one item (an I-tile of kTileI bodies against a run of J-tiles of kTileJ = 32*kWarps bodies):
//   * i-bodies live in registers (2*kPairs per thread, packed two by two as in force.cuh); J-tiles stream through the
//     same TMA bulk-copy ring (cp.async.bulk + mbarrier), which never restarts between items.
//   * symmetric items: warp w walks the 32-body blocks of the tile in the rotated order (w + round) % kWarps, one named
//     barrier per round, so no two warps are ever on the same block. Inside a round the warp is a systolic ring: at step
//     s lane L meets j-body (L+s)&31 (read from shared memory), accumulates the force on its own i-bodies, and adds the
//     reaction to three scalars that move one lane per step by shuffle; after the 32 steps the lane adds them (plain
//     FP64 read-add-write, no atomics: lanes hold distinct j-bodies) to the tile's reaction buffer in shared memory at
//     the j-body they belong to. Per tile the buffer is flushed to the FP64 global accumulators with
//     RED.ADD.F64; per item the i-sums (FP32 runs of 32 folded into FP64 registers, as force.cuh) go the same way.
//   * directed items (the J-tiles inside the I-tile's own index range, i.e. the diagonal): the force.cuh loop, no
//     reaction; the self term is d = 0 times a finite weight, exactly zero (callers route a tiny softening to
//     force.cuh's index-masked variant instead).
// Items are pulled from a global counter by persistent CTAs (kMinBlocks per SM), largest first, so there is no wave
// quantisation: the tail is at most one item.
//
// Accumulation order across items is the order the RED.ADD.F64 operations land, which is not fixed: results are
// reproducible to FP64 rounding of the sums, i.e. the FP32 accelerations can differ in the last bit in about one value
// per 1e9. Callers that need bitwise run-to-run reproducibility use force.cuh (all sizes below kPairMinBodies do).
#pragma once
#include "async_copy.cuh"
#include "force.cuh"

namespace nb {

struct PairItem {
    int i_begin, i_count;  // global index of the first i-body, number of valid i-bodies (<= kTileI)
    int j_begin, j_count;  // global index of the first j-body of the run, bodies in the run (any length)
    int sym;               // 1: every j is a distinct body "after" every i in the pair order -> evaluate once, both ways
    int pad0, pad1, pad2;
};

// A rectangle or a triangle of the interaction matrix, in global body indices.
struct PairBlock {
    int i_lo, i_hi;  // i range
    int j_lo, j_hi;  // j range; for a triangle it equals the i range
    int triangle;    // 1: unordered pairs within one range (diagonal tiles directed, tiles above symmetric)
};
constexpr int kMaxPairBlocks = 16;

struct PairPlanParams {
    PairBlock blocks[kMaxPairBlocks];
    int n_blocks;
    int tile_i, tile_j, chunk_tiles;
    PairItem* items;  // capacity max_items
    int max_items;
    int* n_items;  // out
};

// Single-CTA planner: writes the item list of the given blocks, full-size symmetric chunks first and the short
// remainders / diagonal items last, so that the dynamic schedule ends on small items.
__global__ void __launch_bounds__(256) pair_plan_kernel(const PairPlanParams p) {
    __shared__ int s_base;
    __shared__ int s_cnt[256];
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int chunk = p.chunk_tiles * p.tile_j;
    for (int pass = 0; pass < 2; ++pass) {
        for (int b = 0; b < p.n_blocks; ++b) {
            const PairBlock blk = p.blocks[b];
            const int i_tiles = (blk.i_hi - blk.i_lo + p.tile_i - 1) / p.tile_i;
            for (int a0 = 0; a0 < i_tiles; a0 += blockDim.x) {
                const int a = a0 + threadIdx.x;
                int n_mine = 0, i_begin = 0, i_count = 0, s_lo = 0, s_hi = 0, full = 0;
                if (a < i_tiles) {
                    i_begin = blk.i_lo + a * p.tile_i;
                    i_count = min(p.tile_i, blk.i_hi - i_begin);
                    s_lo = blk.triangle ? min(i_begin + p.tile_i, blk.j_hi) : blk.j_lo;  // symmetric part of the row
                    s_hi = blk.j_hi;
                    full = (s_hi - s_lo) / chunk;
                    const int rest = (s_hi - s_lo) - full * chunk;
                    n_mine = pass == 0 ? full : (rest > 0 ? 1 : 0) + (blk.triangle ? 1 : 0);
                }
                // exclusive scan over the CTA by a serial pass of thread 0 (a few hundred tiles at most)
                s_cnt[threadIdx.x] = n_mine;
                __syncthreads();
                if (threadIdx.x == 0) {
                    int run = s_base;
                    for (int t = 0; t < int(blockDim.x); ++t) {
                        const int c = s_cnt[t];
                        s_cnt[t] = run;
                        run += c;
                    }
                    s_base = run;
                }
                __syncthreads();
                int at = s_cnt[threadIdx.x];
                __syncthreads();
                if (a < i_tiles) {
                    auto put = [&](int j_begin, int j_count, int sym) {
                        if (at < p.max_items) p.items[at] = PairItem{i_begin, i_count, j_begin, j_count, sym, 0, 0, 0};
                        ++at;
                    };
                    if (pass == 0) {
                        for (int c = 0; c < full; ++c) put(s_lo + c * chunk, chunk, 1);
                    } else {
                        const int rest = (s_hi - s_lo) - full * chunk;
                        if (rest > 0) put(s_lo + full * chunk, rest, 1);
                        if (blk.triangle) put(i_begin, min(p.tile_i, blk.j_hi - i_begin), 0);
                    }
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.n_items = min(s_base, p.max_items);
}

struct PairParams {
    const float4* bodies;  // (x,y,z,m) of every body the items refer to
    float eps2;
    const PairItem* items;
    const int* n_items;
    unsigned* counter;  // zero before the launch (finish_kernel resets it)
    double* acc64;      // [n_total][3] un-scaled FP64 sums, accumulated with RED.ADD.F64
};

constexpr int kPairStages = 4;
constexpr int kPairLookahead = 2;
// Steps between folds of the travelling reaction sums into FP64. 32 = once per round: FP32 runs of 32 steps x 4
// i-bodies = 128 terms. Measured on the config4 merger (tools/diag_pair_accuracy.py, profiles/r2_pair_fold_steps.log):
// folding every 8 / 16 steps (32- / 64-term runs) costs 7.5% / 5.5% of the kernel and does not move the error of the
// worst-conditioned bodies (1.2e-7 / 1.4e-7 / 1.5e-7 x condition number at 8 / 16 / 32, against 1.3e-7 for force.cuh's
// 32-term runs): at that level the rounding of the individual terms dominates, not the summation.
constexpr int kReactFoldSteps = 32;

template <int kWarps>
struct PairRing {
    static constexpr int kTileJ = kWarps * 32;
    float4* tiles;
    uint64_t* full;
    uint64_t* empty;
    __host__ __device__ static constexpr size_t ring_bytes() { return size_t(kPairStages) * kTileJ * sizeof(float4) + 2 * kPairStages * sizeof(uint64_t); }
    __host__ __device__ static constexpr size_t smem_bytes() { return ring_bytes() + size_t(2) * kTileJ * 3 * sizeof(double); }
    __device__ __forceinline__ void attach(unsigned char* smem) {
        tiles = reinterpret_cast<float4*>(smem);
        full = reinterpret_cast<uint64_t*>(smem + size_t(kPairStages) * kTileJ * sizeof(float4));
        empty = full + kPairStages;
    }
    __device__ __forceinline__ void init_barriers() {
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        mbar_fence_init();
    }
    // `seq` counts tiles over the CTA's lifetime: the ring keeps its phase across items.
    __device__ __forceinline__ void issue(unsigned seq, const float4* src, int count) {
        const int s = seq % kPairStages;
        if (seq >= kPairStages) mbar_wait(&empty[s], ((seq / kPairStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], uint32_t(count) * sizeof(float4));
        bulk_copy_g2s(tiles + size_t(s) * kTileJ, src, uint32_t(count) * sizeof(float4), &full[s]);
    }
    __device__ __forceinline__ float4* tile(unsigned seq) const { return tiles + size_t(seq % kPairStages) * kTileJ; }
    __device__ __forceinline__ void wait(unsigned seq) { mbar_wait(&full[seq % kPairStages], (seq / kPairStages) & 1); }
    __device__ __forceinline__ void release(unsigned seq) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[seq % kPairStages]);
    }
};

// FP64 read-add-write of one shared-memory word, as an asm statement WITHOUT a memory clobber: the compiler keeps
// software-pipelining the tile loads of later steps across it (a C++ store to the buffer would order them, because it
// cannot prove that the buffer and the tile ring do not alias; measured: 7% of the kernel). Ordering against the other
// accesses to the buffer (zeroing, flush) comes from the named barriers between them, which do clobber memory.
__device__ __forceinline__ void smem_add_f64(uint32_t addr, double v) {
    asm volatile(
        "{\n"
        ".reg .f64 t;\n"
        "ld.shared.f64 t, [%0];\n"
        "add.f64 t, t, %1;\n"
        "st.shared.f64 [%0], t;\n"
        "}\n" ::"r"(addr),
        "d"(v));
}

// Orders generic-proxy writes to shared memory before later async-proxy (TMA) writes to the same bytes.
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int kPairs, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) pair_kernel(const PairParams p) {
    constexpr int kCT = kWarps * 32, kI = 2 * kPairs, kTileJ = kWarps * 32;
    using Ring = PairRing<kWarps>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* react = reinterpret_cast<double*>(smem_raw + Ring::ring_bytes());  // [2][kTileJ][3]
    __shared__ int s_item;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    Ring ring;
    ring.attach(smem_raw);
    if (tid == 0) ring.init_barriers();
    for (int k = tid; k < 2 * kTileJ * 3; k += kCT) react[k] = 0.0;
    __syncthreads();

    const int n_items = *p.n_items;
    unsigned seq_issue = 0, seq_use = 0;  // producer / consumer running tile numbers (same sequence)
    unsigned rbuf = 0;

    for (;;) {
        if (tid == 0) s_item = int(atomicAdd(p.counter, 1u));
        __syncthreads();
        const int it = s_item;
        __syncthreads();
        if (it >= n_items) break;
        const PairItem item = p.items[it];
        const int ntiles = (item.j_count + kTileJ - 1) / kTileJ;
        if (tid == 0)
            for (int t = 0; t < min(kPairLookahead, ntiles); ++t)
                ring.issue(seq_issue++, p.bodies + item.j_begin + t * kTileJ, min(kTileJ, item.j_count - t * kTileJ));

        // this thread's i-bodies; slots past the end of the tile are massless copies of the first body
        f32x2 nx[kPairs], ny[kPairs], nz[kPairs], mi[kPairs];
        bool valid[kI];
        {
            // Component-wise scalar loads on purpose: after a float4 load ptxas keeps the quad registers as the home of
            // the coordinates and re-packs (x_i0, x_i1) with two MOVs at EVERY use inside the systolic loop (9 MOVs
            // per step, profiles/r2_sass_pair_kernel.txt); loading each component into its own register lets the two
            // halves of a pair be allocated next to each other once.
            float cx[kI], cy[kI], cz[kI], cm[kI];
#pragma unroll
            for (int k = 0; k < kI; ++k) {
                const int li = k * kCT + tid;
                valid[k] = li < item.i_count;
                const float* src = reinterpret_cast<const float*>(p.bodies + item.i_begin + (valid[k] ? li : 0));
                cx[k] = ldg_f32(src), cy[k] = ldg_f32(src + 1), cz[k] = ldg_f32(src + 2), cm[k] = ldg_f32(src + 3);
                if (!valid[k]) cm[k] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                nx[q] = pack2(-cx[2 * q], -cx[2 * q + 1]);
                ny[q] = pack2(-cy[2 * q], -cy[2 * q + 1]);
                nz[q] = pack2(-cz[2 * q], -cz[2 * q + 1]);
                mi[q] = pack2(cm[2 * q], cm[2 * q + 1]);
            }
        }
        double tot[kI][3];
#pragma unroll
        for (int k = 0; k < kI; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;
        f32x2 ax[kPairs], ay[kPairs], az[kPairs];
        auto fold = [&]() {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                tot[2 * q][0] += double(lo2(ax[q])), tot[2 * q][1] += double(lo2(ay[q])), tot[2 * q][2] += double(lo2(az[q]));
                tot[2 * q + 1][0] += double(hi2(ax[q])), tot[2 * q + 1][1] += double(hi2(ay[q]));
                tot[2 * q + 1][2] += double(hi2(az[q]));
            }
        };
        const f32x2 eps2 = pack2(p.eps2, p.eps2);

        for (int t = 0; t < ntiles; ++t) {
            if (tid == 0 && t + kPairLookahead < ntiles)
                ring.issue(seq_issue++, p.bodies + item.j_begin + (t + kPairLookahead) * kTileJ,
                           min(kTileJ, item.j_count - (t + kPairLookahead) * kTileJ));
            const int jt0 = item.j_begin + t * kTileJ;
            const int count = min(kTileJ, item.j_count - t * kTileJ);
            float4* tile_w = ring.tile(seq_use);
            const float4* __restrict__ tj = tile_w;
            ring.wait(seq_use);

            if (item.sym) {
                if (count & 31) {  // ragged last block: fill it up with massless bodies far away (they add exactly zero)
                    if (warp == 0) {
                        const int idx = (count & ~31) + lane;
                        if (idx >= count) tile_w[idx] = make_float4(1e18f, 1e18f, 1e18f, 0.f);
                        fence_proxy_async_smem();
                    }
                    compute_barrier<kCT>();
                }
                double* rb = react + size_t(rbuf) * kTileJ * 3;
                const uint32_t rb_s = smem_u32(rb);
                // one copy of the (fully unrolled, ~35 KB) round body: unrolling the rounds as well put 190 KB of code in
                // the kernel, more than the instruction cache holds, and cost 4.5%
#pragma unroll 1
                for (int r = 0; r < kWarps; ++r) {
                    const int jblk = (warp + r) % kWarps;
                    if (jblk * 32 < count) {
                        const float4* __restrict__ blk = tj + jblk * 32;
#pragma unroll
                        for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = 0ull;
                        float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
                        for (int s = 0; s < 32; ++s) {
                            const float4 b = blk[(lane + s) & 31];
                            const f32x2 bx = pack2(b.x, b.x), by = pack2(b.y, b.y), bz = pack2(b.z, b.z), bm = pack2(b.w, b.w);
#pragma unroll
                            for (int q = 0; q < kPairs; ++q) {
                                const f32x2 dx = add2(bx, nx[q]);
                                const f32x2 dy = add2(by, ny[q]);
                                const f32x2 dz = add2(bz, nz[q]);
                                f32x2 r2 = fma2(dz, dz, eps2);
                                r2 = fma2(dy, dy, r2);
                                r2 = fma2(dx, dx, r2);
                                const f32x2 ri = pack2(rsqrt_approx(lo2(r2)), rsqrt_approx(hi2(r2)));
                                const f32x2 ri3 = mul2(mul2(ri, ri), ri);
                                const f32x2 wi = mul2(ri3, bm);     // weight of j's pull on the two i-bodies
                                const f32x2 wj = mul2(ri3, mi[q]);  // weight of their pull on j
                                ax[q] = fma2(wi, dx, ax[q]);
                                ay[q] = fma2(wi, dy, ay[q]);
                                az[q] = fma2(wi, dz, az[q]);
                                const float wj0 = lo2(wj), wj1 = hi2(wj);
                                // grouped by weight: three consecutive FFMAs share one multiplicand (operand reuse)
                                sx = __fmaf_rn(-wj0, lo2(dx), sx), sy = __fmaf_rn(-wj0, lo2(dy), sy), sz = __fmaf_rn(-wj0, lo2(dz), sz);
                                sx = __fmaf_rn(-wj1, hi2(dx), sx), sy = __fmaf_rn(-wj1, hi2(dy), sy), sz = __fmaf_rn(-wj1, hi2(dz), sz);
                            }
                            if ((s & (kReactFoldSteps - 1)) == kReactFoldSteps - 1) {
                                // every kReactFoldSteps steps the travelling FP32 sums (kReactFoldSteps * 2*kPairs
                                // terms) are folded into the tile's FP64 buffer at the j-body they belong to; lanes
                                // hold distinct j-bodies, so there is no conflict
                                const uint32_t rj = rb_s + uint32_t(jblk * 32 + ((lane + s) & 31)) * 24u;
                                smem_add_f64(rj, double(sx)), smem_add_f64(rj + 8, double(sy)), smem_add_f64(rj + 16, double(sz));
                                sx = sy = sz = 0.f;
                            } else {
                                // the j-body this lane meets next is the one lane+1 just met: fetch its reaction sums
                                const int src = (lane + 1) & 31;
                                sx = __shfl_sync(0xffffffffu, sx, src);
                                sy = __shfl_sync(0xffffffffu, sy, src);
                                sz = __shfl_sync(0xffffffffu, sz, src);
                            }
                        }
                        fold();
                    }
                    compute_barrier<kCT>();  // rounds in lockstep: no two warps on one j-block
                }
                // flush the tile's reactions; the other buffer serves the next tile
                for (int k = tid; k < count * 3; k += kCT) {
                    const double v = rb[k];
                    rb[k] = 0.0;
                    atomicAdd(&p.acc64[size_t(jt0) * 3 + k], v);
                }
                if (count & 31)  // reactions on the filler bodies
                    for (int k = count * 3 + tid; k < ((count + 31) & ~31) * 3; k += kCT) rb[k] = 0.0;
                rbuf ^= 1;
            } else {
                for (int jb = 0; jb < count; jb += 32) {
#pragma unroll
                    for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = 0ull;
                    auto interact = [&](int jj) {
                        const float4 b = tj[jj];
                        const f32x2 bx = pack2(b.x, b.x), by = pack2(b.y, b.y), bz = pack2(b.z, b.z), bm = pack2(b.w, b.w);
#pragma unroll
                        for (int q = 0; q < kPairs; ++q) {
                            const f32x2 dx = add2(bx, nx[q]);
                            const f32x2 dy = add2(by, ny[q]);
                            const f32x2 dz = add2(bz, nz[q]);
                            f32x2 r2 = fma2(dz, dz, eps2);
                            r2 = fma2(dy, dy, r2);
                            r2 = fma2(dx, dx, r2);
                            const f32x2 ri = pack2(rsqrt_approx(lo2(r2)), rsqrt_approx(hi2(r2)));
                            const f32x2 w = mul2(mul2(ri, ri), mul2(ri, bm));
                            ax[q] = fma2(w, dx, ax[q]);
                            ay[q] = fma2(w, dy, ay[q]);
                            az[q] = fma2(w, dz, az[q]);
                        }
                    };
                    if (jb + 32 <= count) {
#pragma unroll
                        for (int u = 0; u < 32; ++u) interact(jb + u);
                    } else {
                        for (int jj = jb; jj < count; ++jj) interact(jj);
                    }
                    fold();
                }
            }
            ring.release(seq_use);
            ++seq_use;
        }
#pragma unroll
        for (int k = 0; k < kI; ++k)
            if (valid[k]) {
                double* dst = p.acc64 + size_t(item.i_begin + k * kCT + tid) * 3;
                atomicAdd(dst + 0, tot[k][0]);
                atomicAdd(dst + 1, tot[k][1]);
                atomicAdd(dst + 2, tot[k][2]);
            }
    }
}

// The integrator epilogue of the pair path: acceleration = fl32(G) * fl32(FP64 sum), then the same rounding-faithful
// update as force.cuh's epilogue_body (simulation.py:153-187). Also clears what the next pair launch accumulates
// into and rewinds the item counters.
struct FinishParams {
    ForceParams f;           // bodies / bodies_next / i_begin / i_count / g / mode / dt / state and record pointers
    double* acc_own;         // [i_count][3] sums of this rank's bodies (local index); cleared after reading
    double* zero_extra;      // optional second array to clear (the full accumulator of a sharded rank), may be null
    long long zero_count;    // doubles in zero_extra
    unsigned* counters;      // item counters to rewind, may be null
    int n_counters;
};

__global__ void __launch_bounds__(256) finish_kernel(const FinishParams p) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = gid; i < p.f.i_count; i += stride) {
        double* a = p.acc_own + i * 3;
        const double sx = a[0], sy = a[1], sz = a[2];
        a[0] = a[1] = a[2] = 0.0;
        const float4 me = p.f.bodies[p.f.i_begin + i];
        epilogue_body(p.f, int(i), me, float(sx), float(sy), float(sz));
    }
    if (p.zero_extra)
        for (long long k = gid; k < p.zero_count; k += stride) p.zero_extra[k] = 0.0;
    if (gid < p.n_counters) p.counters[gid] = 0u;
}

}  // namespace nb
