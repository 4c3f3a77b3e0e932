// Total potential and kinetic energy, as the reference defines them.
//
// Replaces (reference, read-only): src/galaxify/simulation.py:91-115 (BaseSimulator.compute_energies):
//   k = sum_i 0.5 * m_i * |v_i|^2
//   u = sum_{i<j} -G m_i m_j / (|r_i - r_j| + eps)        <- softening enters as |r| + eps, NOT sqrt(r^2+eps^2)
//
// The pair sum is evaluated as -G * sum_i m_i * phi_i with phi_i = sum_{j > i} m_j / (|r_ij| + eps) (the upper
// triangle, as simulation.py:113), on the same TMA-fed j-tile ring as the force kernel; j tiles that lie entirely
// below a CTA's i-bodies are skipped, so only half of the N^2 pairs are evaluated.
// 1/(|r| + eps) needs |r| = sqrt(r^2) AND a reciprocal: two MUFU operations per interaction where the force needs one,
// which is why this kernel is kept apart from the force kernel. With both on the XU pipe (MUFU.SQRT + MUFU.RCP) the
// kernel sat at 92% of the XU pipe's throughput (round 1: 255 ms at N = 1M). Now |r| = r^2 * rsqrt(r^2) (one MUFU),
// and of the two bodies packed in a register pair one takes MUFU.RCP and the other a Newton reciprocal on the FMA pipe
// (rcp_newton), which balances the two pipes: per pair of interactions 3 MUFU (24 XU cycles) and 9 packed + 6 scalar
// FMA-pipe instructions (24 cycles). Pairs with j <= i are masked by index (the reference masks the diagonal with +inf and
// keeps triu(1), simulation.py:107-113); distinct coincident bodies contribute m_i m_j / eps exactly as there.
// Sums over runs of 32 bodies are FP32, everything across runs / tiles / threads / CTAs is FP64, and every cross-CTA sum is taken in a
// fixed order, so the result is deterministic and closer to the exact value than the reference's FP32 reduction.
#pragma once
#include "async_copy.cuh"

namespace nb {

constexpr int kEnergyStages = 4;
constexpr int kEnergyLookahead = 2;
constexpr int kEnergyFold = 32;

struct EnergyParams {
    const float4* bodies;  // (x,y,z,m), all n_total bodies
    int n_total;
    int i_begin, i_count;
    float eps;
    double* cta_partial;  // [gridDim.y][gridDim.x] : sum over the CTA's i-bodies of m_i * phi_i (its j split)
};

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 1/x for a positive normal x without the MUFU unit: the integer seed (magic constant minus the bit pattern, within
// 12.5% of 1/x) refined by two cubically convergent steps s <- s + s*(e + e*e), e = 1 - x*s: 0.125 -> 2e-3 -> 8e-9
// relative, i.e. correct to FP32 rounding. Six FMA-pipe operations and one integer subtraction.
__device__ __forceinline__ float rcp_newton(float x) {
    float s = __int_as_float(0x7EF311C7 - __float_as_int(x));
    float e = __fmaf_rn(-x, s, 1.f);
    s = __fmaf_rn(s, __fmaf_rn(e, e, e), s);
    e = __fmaf_rn(-x, s, 1.f);
    s = __fmaf_rn(s, __fmaf_rn(e, e, e), s);
    return s;
}

template <int kBlock>
__device__ __forceinline__ double block_sum(double v, double* scratch /* kBlock/32 doubles */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double total = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kBlock / 32; ++w) total += scratch[w];
    return total;  // valid on thread 0
}

struct MaskOn {
    static constexpr bool value = true;
};
struct MaskOff {
    static constexpr bool value = false;
};

template <int kPairs, int kWarps, int kMinBlocks, int kTileJ>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) potential_kernel(const EnergyParams p) {
    constexpr int kCT = kWarps * 32;
    constexpr int kI = 2 * kPairs;
    constexpr int kTileI = kCT * kI;
    using Ring = TileRing<kTileJ, kEnergyStages, kWarps>;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double s_red[kWarps];

    const int tid = threadIdx.x;
    const int nj = p.n_total;
    const int per = (nj + gridDim.y - 1) / gridDim.y;
    const int j0 = min(int(blockIdx.y) * per, nj);
    const int j1 = min(int(blockIdx.y + 1) * per, nj);

    // upper triangle: the smallest global i of this CTA bounds the j range from below (rounded down to a tile so
    // that TMA sources stay 16-byte aligned and tiles line up between CTAs)
    const int tile_base = blockIdx.x * kTileI;
    const int i_min = p.i_begin + min(tile_base, p.i_count - 1);
    const int j_lo = max(j0, (i_min / kTileJ) * kTileJ);
    Ring ring;
    ring.attach(smem_raw, p.bodies + min(j_lo, j1), max(j1 - j_lo, 0));
    const int ntiles = ring.num_tiles();
    if (tid == 0) ring.init_barriers();
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < min(kEnergyLookahead, ntiles); ++t) ring.issue(t);

    float4 me[kI];
    int gi[kI];
    bool valid[kI];
    float2 nx[kPairs], ny[kPairs], nz[kPairs];
#pragma unroll
    for (int k = 0; k < kI; ++k) {
        const int li = tile_base + k * kCT + tid;
        valid[k] = li < p.i_count;
        gi[k] = p.i_begin + min(li, p.i_count - 1);
        me[k] = p.bodies[gi[k]];
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        nx[q] = make_float2(-me[2 * q].x, -me[2 * q + 1].x);
        ny[q] = make_float2(-me[2 * q].y, -me[2 * q + 1].y);
        nz[q] = make_float2(-me[2 * q].z, -me[2 * q + 1].z);
    }
    double phi[kI];
#pragma unroll
    for (int k = 0; k < kI; ++k) phi[k] = 0.0;
    const float2 eps = make_float2(p.eps, p.eps);
    const float2 tiny = make_float2(1e-30f, 1e-30f);

    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + kEnergyLookahead < ntiles) ring.issue(t + kEnergyLookahead);
        const int jt = j_lo + ring.tile_offset(t);
        const int count = ring.tile_count(t);
        const float4* __restrict__ tj = ring.tile(t);
        float2 acc[kPairs];
        ring.wait(t);
        // Only tiles that reach down to this CTA's own i-bodies need the j > i mask; for every tile above them (all but
        // a handful) the mask-free variant saves three integer compares and selects per (i-pair, j), which matters:
        // with the reciprocal split over two pipes the loop is bound by instruction issue, not by a pipe.
        const bool masked = jt <= p.i_begin + tile_base + kTileI - 1;
        auto interact = [&](int jj, auto use_mask) {
            const float4 b = tj[jj];
            const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z);
            const float2 bm = make_float2(b.w, b.w);
            const int jg = jt + jj;
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                const float2 dx = __fadd2_rn(bx, nx[q]);
                const float2 dy = __fadd2_rn(by, ny[q]);
                const float2 dz = __fadd2_rn(bz, nz[q]);
                float2 r2 = __ffma2_rn(dx, dx, tiny);  // + 1e-30: |r| = r2 * rsqrt(r2) stays finite for coincident bodies
                r2 = __ffma2_rn(dy, dy, r2);
                r2 = __ffma2_rn(dz, dz, r2);
                // |r| = r2 * rsqrt(r2): one MUFU per body; then 1/(|r| + eps) by MUFU.RCP for one body of the pair and
                // by two cubic Newton steps on the FMA pipe for the other, so that the XU and FMA pipes carry equal
                // loads (3 MUFU and 24 FMA-pipe cycles per pair of interactions, against 4 MUFU = 32 XU cycles before)
                const float2 ry = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                const float2 d = __fadd2_rn(__fmul2_rn(r2, ry), eps);
                float2 inv = make_float2(rcp_approx(d.x), rcp_newton(d.y));
                if (decltype(use_mask)::value) {
                    if (jg <= gi[2 * q]) inv.x = 0.f;
                    if (jg <= gi[2 * q + 1]) inv.y = 0.f;
                }
                acc[q] = __ffma2_rn(bm, inv, acc[q]);
            }
        };
        // FP32 runs of kEnergyFold bodies, folded into the FP64 row sums (same scheme as force.cuh)
        for (int jb = 0; jb < count; jb += kEnergyFold) {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) acc[q] = make_float2(0.f, 0.f);
            if (jb + kEnergyFold <= count && !masked) {
#pragma unroll 8
                for (int u = 0; u < kEnergyFold; ++u) interact(jb + u, MaskOff{});
            } else if (jb + kEnergyFold <= count) {
#pragma unroll 4
                for (int u = 0; u < kEnergyFold; ++u) interact(jb + u, MaskOn{});
            } else {
                for (int jj = jb; jj < count; ++jj) interact(jj, MaskOn{});
            }
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                phi[2 * q] += double(acc[q].x);
                phi[2 * q + 1] += double(acc[q].y);
            }
        }
        ring.release(t);
    }
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < kI; ++k)
        if (valid[k]) mine += double(me[k].w) * phi[k];
    const double total = block_sum<kCT>(mine, s_red);
    if (tid == 0) p.cta_partial[size_t(blockIdx.y) * gridDim.x + blockIdx.x] = total;
}

// Single-CTA finish: u = -G * sum(cta_partial) and k = sum 0.5 m v^2, both in fixed order.
struct EnergyFinishParams {
    const double* cta_partial;
    int n_partials;
    const float4* bodies;  // for the masses, global index
    const float* vel;      // (i_count,3), local index
    int i_begin, i_count;
    float g;
    double* out_uk;  // (u, k); with accumulate != 0 the values are added (sharded ranks sum afterwards)
};

__global__ void __launch_bounds__(1024) energy_finish_kernel(const EnergyFinishParams p) {
    __shared__ double s_red[32];
    double u = 0.0;
    for (int i = threadIdx.x; i < p.n_partials; i += 1024) u += p.cta_partial[i];
    const double u_tot = block_sum<1024>(u, s_red);
    __syncthreads();
    double k = 0.0;
    for (int i = threadIdx.x; i < p.i_count; i += 1024) {
        const float vx = p.vel[3 * i], vy = p.vel[3 * i + 1], vz = p.vel[3 * i + 2];
        // per-body term in FP32 exactly as simulation.py:100: 0.5 * m * (vx^2 + vy^2 + vz^2)
        const float v2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        k += double(__fmul_rn(__fmul_rn(0.5f, p.bodies[p.i_begin + i].w), v2));
    }
    const double k_tot = block_sum<1024>(k, s_red);
    if (threadIdx.x == 0) {
        p.out_uk[0] = -double(p.g) * u_tot;
        p.out_uk[1] = k_tot;
    }
}

// Energies of every recorded state of a (batched) trajectory buffer, one CTA per (slot, system): the state's bodies
// go to shared memory, thread t takes rows i = t, t+T, ... of the upper triangle. Same arithmetic as
// potential_kernel / energy_finish_kernel (simulation.py:91-115). Used for systems that fit one CTA (n <= 2048),
// where the whole run is one persistent kernel and the energies are evaluated afterwards, all states in parallel.
struct TrajEnergyParams {
    const float* traj;  // [slot][3][n_systems][n][3]
    const float* mass;  // [n_systems][n]
    int n_systems, n;
    float g, eps;
    double* out;  // [slot][n_systems][2]
};

constexpr int kTrajEnergyThreads = 128;

__global__ void __launch_bounds__(kTrajEnergyThreads) traj_energy_kernel(const TrajEnergyParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* bodies = reinterpret_cast<float4*>(smem_raw);
    __shared__ double s_red[kTrajEnergyThreads / 32];
    const int slot = blockIdx.y, sys = blockIdx.x, tid = threadIdx.x;
    const size_t plane = size_t(p.n_systems) * p.n * 3;
    const float* pos = p.traj + size_t(slot) * 3 * plane + size_t(sys) * p.n * 3;
    const float* vel = pos + plane;
    const float* mass = p.mass + size_t(sys) * p.n;
    for (int j = tid; j < p.n; j += kTrajEnergyThreads)
        bodies[j] = make_float4(pos[3 * j], pos[3 * j + 1], pos[3 * j + 2], mass[j]);
    __syncthreads();
    double u = 0.0, k = 0.0;
    for (int i = tid; i < p.n; i += kTrajEnergyThreads) {
        const float4 me = bodies[i];
        double phi = 0.0;
        for (int jb = i + 1; jb < p.n; jb += kEnergyFold) {
            float run = 0.f;
            const int jend = min(p.n, jb + kEnergyFold);
            for (int j = jb; j < jend; ++j) {
                const float4 b = bodies[j];
                const float dx = b.x - me.x, dy = b.y - me.y, dz = b.z - me.z;
                const float r2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, dx * dx));
                run = __fmaf_rn(b.w, rcp_approx(sqrt_approx(r2) + p.eps), run);
            }
            phi += double(run);
        }
        u += double(me.w) * phi;
        const float vx = vel[3 * i], vy = vel[3 * i + 1], vz = vel[3 * i + 2];
        const float v2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        k += double(__fmul_rn(__fmul_rn(0.5f, me.w), v2));
    }
    const double u_tot = block_sum<kTrajEnergyThreads>(u, s_red);
    __syncthreads();
    const double k_tot = block_sum<kTrajEnergyThreads>(k, s_red);
    if (tid == 0) {
        double* o = p.out + (size_t(slot) * p.n_systems + sys) * 2;
        o[0] = -double(p.g) * u_tot;
        o[1] = k_tot;
    }
}

}  // namespace nb
