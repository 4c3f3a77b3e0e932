// Many independent small systems, all steps inside one persistent kernel.
//
// No reference counterpart: src/s01-dataset-generation.py:130-214 builds and runs scenes one at a time. Each
// system evolves exactly as LeapFrogSimulator.step / EulerSimulator.step (src/galaxify/simulation.py:153-187)
// would evolve it alone (same separately-rounded multiply/add updates), with the force of simulation.py:71-89.
//
// One thread-block cluster (1, 2 or 4 CTAs of 32..128 threads) per system. The (x,y,z,m) bodies live in shared memory
// (double-buffered, one cluster barrier per step), every thread keeps the position, velocity and acceleration of its
// 2*kPairs bodies in registers across steps, and the j loop reads shared memory with broadcast LDS.128 and runs the
// same packed-FP32 inner loop as the large-N force kernel. Global memory is touched only to load the initial state
// and to record trajectory slots.
#pragma once
#include <cooperative_groups.h>

#include "async_copy.cuh"
#include "force.cuh"

namespace nb {

constexpr int kBatchedMaxThreads = 128;
constexpr int kBatchedMaxPairs = 2;
constexpr int kBatchedMaxCluster = 4;

struct BatchedParams {
    int n;  // bodies per system
    int mode;
    int steps;
    int record_every;
    float g, eps2, dt, half_dt;
    const float* mass;  // (S,n)
    float* pos;         // (S,n,3) in/out
    float* vel;
    float* acc;
    float* traj;  // [slot][3][S][n][3] or null
    int n_systems;
};

constexpr int kBatchedFold = 32;  // FP32 accumulation run length, as in force.cuh

template <int kPairs, bool kExactDiag>
__device__ __forceinline__ void batched_force(const float4* __restrict__ bodies, int n, float eps2s,
                                              const float (&x)[2 * kPairs][3], const int (&idx)[2 * kPairs],
                                              float (&sum)[2 * kPairs][3]) {
    float2 nx[kPairs], ny[kPairs], nz[kPairs], ax[kPairs], ay[kPairs], az[kPairs];
    double tot[2 * kPairs][3];
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        nx[q] = make_float2(-x[2 * q][0], -x[2 * q + 1][0]);
        ny[q] = make_float2(-x[2 * q][1], -x[2 * q + 1][1]);
        nz[q] = make_float2(-x[2 * q][2], -x[2 * q + 1][2]);
    }
#pragma unroll
    for (int k = 0; k < 2 * kPairs; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;
    const float2 eps2 = make_float2(eps2s, eps2s);

    auto interact = [&](int j) {
        const float4 b = bodies[j];
        const float2 bx = make_float2(b.x, b.x), by = make_float2(b.y, b.y), bz = make_float2(b.z, b.z);
        const float2 bm = make_float2(b.w, b.w);
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const float2 dx = __fadd2_rn(bx, nx[q]);
            const float2 dy = __fadd2_rn(by, ny[q]);
            const float2 dz = __fadd2_rn(bz, nz[q]);
            float2 r2 = __ffma2_rn(dz, dz, eps2);
            r2 = __ffma2_rn(dy, dy, r2);
            r2 = __ffma2_rn(dx, dx, r2);
            float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
            if (kExactDiag) {
                if (j == idx[2 * q]) ri.x = 0.f;
                if (j == idx[2 * q + 1]) ri.y = 0.f;
            }
            const float2 ri2 = __fmul2_rn(ri, ri);
            const float2 mri = __fmul2_rn(ri, bm);
            const float2 w = __fmul2_rn(ri2, mri);
            ax[q] = __ffma2_rn(w, dx, ax[q]);
            ay[q] = __ffma2_rn(w, dy, ay[q]);
            az[q] = __ffma2_rn(w, dz, az[q]);
        }
    };

    for (int jb = 0; jb < n; jb += kBatchedFold) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
        if (jb + kBatchedFold <= n) {
#pragma unroll
            for (int u = 0; u < kBatchedFold; ++u) interact(jb + u);
        } else {
            for (int j = jb; j < n; ++j) interact(j);
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            tot[2 * q][0] += double(ax[q].x), tot[2 * q][1] += double(ay[q].x), tot[2 * q][2] += double(az[q].x);
            tot[2 * q + 1][0] += double(ax[q].y), tot[2 * q + 1][1] += double(ay[q].y), tot[2 * q + 1][2] += double(az[q].y);
        }
    }
#pragma unroll
    for (int k = 0; k < 2 * kPairs; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) sum[k][c] = float(tot[k][c]);
}

// One thread-block CLUSTER per system: the system's i-bodies are split evenly over the cluster's CTAs (so that the
// schedulable unit is a fraction of a system and 512 systems spread evenly over 148 SMs), every CTA keeps a full copy
// of the (x,y,z,m) array in its own shared memory, and after the drift each CTA stores its updated bodies into every
// copy through distributed shared memory; one cluster barrier per step replaces the __syncthreads.
template <int kPairs, int kBatchedThreads, bool kExactDiag>
__global__ void __launch_bounds__(kBatchedThreads) batched_kernel(const BatchedParams p) {
    namespace cg = cooperative_groups;
    constexpr int kI = 2 * kPairs;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* buf0 = reinterpret_cast<float4*>(smem_raw);
    float4* buf1 = buf0 + p.n;

    cg::cluster_group cluster = cg::this_cluster();
    const int csize = int(cluster.num_blocks());
    const int crank = int(cluster.block_rank());
    const int sys = blockIdx.x / csize;
    const int tid = threadIdx.x;
    const int per = (p.n + csize - 1) / csize;  // i-bodies of one CTA
    const int i_lo = crank * per;
    const int i_hi = min(p.n, i_lo + per);

    const size_t base3 = size_t(sys) * p.n * 3;
    const float* mass = p.mass + size_t(sys) * p.n;
    float* gpos = p.pos + base3;
    float* gvel = p.vel + base3;
    float* gacc = p.acc + base3;

    // every CTA loads the whole system into its own copy
    for (int j = tid; j < p.n; j += kBatchedThreads)
        buf0[j] = make_float4(gpos[3 * j], gpos[3 * j + 1], gpos[3 * j + 2], mass[j]);

    int idx[kI];
    bool valid[kI];
    float x[kI][3], v[kI][3], a[kI][3], m[kI], sum[kI][3];
#pragma unroll
    for (int k = 0; k < kI; ++k) {
        const int i = i_lo + k * kBatchedThreads + tid;
        valid[k] = i < i_hi;
        idx[k] = min(i, p.n - 1);
        m[k] = mass[idx[k]];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            x[k][c] = gpos[3 * idx[k] + c];
            v[k][c] = gvel[3 * idx[k] + c];
            a[k][c] = (p.mode == MODE_LEAPFROG) ? gacc[3 * idx[k] + c] : 0.f;
        }
    }
    // Every CTA of the cluster must be running before any of them writes into a peer's shared memory (publish()),
    // and a peer's initial load of buf0 must have finished before it is read: a cluster-wide barrier does both.
    if (csize > 1)
        cluster.sync();
    else
        __syncthreads();

    float4* cur = buf0;
    float4* nxt = buf1;

    if (p.mode == MODE_ACCEL) {
        batched_force<kPairs, kExactDiag>(cur, p.n, p.eps2, x, idx, sum);
#pragma unroll
        for (int k = 0; k < kI; ++k)
            if (valid[k])
                store3(gacc, idx[k], __fmul_rn(p.g, sum[k][0]), __fmul_rn(p.g, sum[k][1]), __fmul_rn(p.g, sum[k][2]));
        return;
    }

    // publishes this thread's drifted bodies into every CTA's copy of `dst`
    auto publish = [&](float4* dst) {
#pragma unroll
        for (int k = 0; k < kI; ++k) {
            if (!valid[k]) continue;
            const float4 b = make_float4(x[k][0], x[k][1], x[k][2], m[k]);
            for (int r = 0; r < csize; ++r) cluster.map_shared_rank(dst, r)[idx[k]] = b;
        }
    };

    const size_t plane = size_t(p.n_systems) * p.n * 3;
    for (int s = 0; s < p.steps; ++s) {
        if (p.mode == MODE_LEAPFROG) {
            // half-kick + drift (simulation.py:164-166), publish the drifted bodies, then force + closing half-kick
#pragma unroll
            for (int k = 0; k < kI; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[k][c] = __fadd_rn(v[k][c], __fmul_rn(p.half_dt, a[k][c]));
                    x[k][c] = __fadd_rn(x[k][c], __fmul_rn(p.dt, v[k][c]));
                }
            publish(nxt);
            cluster.sync();
            batched_force<kPairs, kExactDiag>(nxt, p.n, p.eps2, x, idx, sum);
#pragma unroll
            for (int k = 0; k < kI; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[k][c] = __fmul_rn(p.g, sum[k][c]);
                    v[k][c] = __fadd_rn(v[k][c], __fmul_rn(p.half_dt, a[k][c]));
                }
        } else {
            // force at the current positions, then v += dt*a ; x += dt*v (simulation.py:183-187)
            batched_force<kPairs, kExactDiag>(cur, p.n, p.eps2, x, idx, sum);
#pragma unroll
            for (int k = 0; k < kI; ++k)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[k][c] = __fmul_rn(p.g, sum[k][c]);
                    v[k][c] = __fadd_rn(v[k][c], __fmul_rn(p.dt, a[k][c]));
                    x[k][c] = __fadd_rn(x[k][c], __fmul_rn(p.dt, v[k][c]));
                }
            publish(nxt);
            cluster.sync();
        }
        float4* tmp = cur;
        cur = nxt;
        nxt = tmp;

        if (p.traj && (s + 1) % p.record_every == 0) {
            float* slot = p.traj + size_t((s + 1) / p.record_every - 1) * 3 * plane + base3;
#pragma unroll
            for (int k = 0; k < kI; ++k)
                if (valid[k]) {
                    store3(slot, idx[k], x[k][0], x[k][1], x[k][2]);
                    store3(slot + plane, idx[k], v[k][0], v[k][1], v[k][2]);
                    store3(slot + 2 * plane, idx[k], a[k][0], a[k][1], a[k][2]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < kI; ++k)
        if (valid[k]) {
            store3(gpos, idx[k], x[k][0], x[k][1], x[k][2]);
            store3(gvel, idx[k], v[k][0], v[k][1], v[k][2]);
            store3(gacc, idx[k], a[k][0], a[k][1], a[k][2]);
        }
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace nb
