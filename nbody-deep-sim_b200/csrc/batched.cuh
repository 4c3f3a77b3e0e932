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

// The state of a thread's bodies is held as packed pairs (two bodies per 64-bit register, one register per component)
// for the whole kernel: the j loop then uses them as they are. With scalar state the pairs were re-packed by two MOVs at
// every use inside the loop (2.8 MOVs per j on top of 12 FMA-pipe instructions, profiles/r2_sass_batched.txt).
template <int kPairs, bool kExactDiag>
__device__ __forceinline__ void batched_force(const float4* __restrict__ bodies, int n, float eps2s,
                                              const f32x2 (&x)[kPairs][3], const int (&idx)[2 * kPairs],
                                              f32x2 (&sum)[kPairs][3]) {
    f32x2 ax[kPairs], ay[kPairs], az[kPairs];
    double tot[2 * kPairs][3];
#pragma unroll
    for (int k = 0; k < 2 * kPairs; ++k) tot[k][0] = tot[k][1] = tot[k][2] = 0.0;
    const f32x2 eps2 = pack2(eps2s, eps2s);

    auto interact = [&](int j) {
        const float4 b = bodies[j];
        const f32x2 bx = pack2(b.x, b.x), by = pack2(b.y, b.y), bz = pack2(b.z, b.z), bm = pack2(b.w, b.w);
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const f32x2 dx = sub2(bx, x[q][0]);
            const f32x2 dy = sub2(by, x[q][1]);
            const f32x2 dz = sub2(bz, x[q][2]);
            f32x2 r2 = fma2(dz, dz, eps2);
            r2 = fma2(dy, dy, r2);
            r2 = fma2(dx, dx, r2);
            float r0 = rsqrt_approx(lo2(r2)), r1 = rsqrt_approx(hi2(r2));
            if (kExactDiag) {
                if (j == idx[2 * q]) r0 = 0.f;
                if (j == idx[2 * q + 1]) r1 = 0.f;
            }
            const f32x2 ri = pack2(r0, r1);
            const f32x2 w = mul2(mul2(ri, ri), mul2(ri, bm));
            ax[q] = fma2(w, dx, ax[q]);
            ay[q] = fma2(w, dy, ay[q]);
            az[q] = fma2(w, dz, az[q]);
        }
    };

    for (int jb = 0; jb < n; jb += kBatchedFold) {
#pragma unroll
        for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = 0ull;
        if (jb + kBatchedFold <= n) {
#pragma unroll
            for (int u = 0; u < kBatchedFold; ++u) interact(jb + u);
        } else {
            for (int j = jb; j < n; ++j) interact(j);
        }
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            tot[2 * q][0] += double(lo2(ax[q])), tot[2 * q][1] += double(lo2(ay[q])), tot[2 * q][2] += double(lo2(az[q]));
            tot[2 * q + 1][0] += double(hi2(ax[q])), tot[2 * q + 1][1] += double(hi2(ay[q]));
            tot[2 * q + 1][2] += double(hi2(az[q]));
        }
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) sum[q][c] = pack2(float(tot[2 * q][c]), float(tot[2 * q + 1][c]));
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
    float m[kI];
    f32x2 x[kPairs][3], v[kPairs][3], a[kPairs][3], sum[kPairs][3];  // (body 2q, body 2q+1) per component
#pragma unroll
    for (int k = 0; k < kI; ++k) {
        const int i = i_lo + k * kBatchedThreads + tid;
        valid[k] = i < i_hi;
        idx[k] = min(i, p.n - 1);
        m[k] = mass[idx[k]];
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int i0 = 3 * idx[2 * q] + c, i1 = 3 * idx[2 * q + 1] + c;
            x[q][c] = pack2(gpos[i0], gpos[i1]);
            v[q][c] = pack2(gvel[i0], gvel[i1]);
            a[q][c] = (p.mode == MODE_LEAPFROG) ? pack2(gacc[i0], gacc[i1]) : 0ull;
        }
    // Every CTA of the cluster must be running before any of them writes into a peer's shared memory (publish()),
    // and a peer's initial load of buf0 must have finished before it is read: a cluster-wide barrier does both.
    if (csize > 1)
        cluster.sync();
    else
        __syncthreads();

    float4* cur = buf0;
    float4* nxt = buf1;
    auto comp = [&](const f32x2 (&arr)[kPairs][3], int k, int c) { return (k & 1) ? hi2(arr[k >> 1][c]) : lo2(arr[k >> 1][c]); };
    auto store_state = [&](float* base, const f32x2 (&arr)[kPairs][3]) {
#pragma unroll
        for (int k = 0; k < kI; ++k)
            if (valid[k]) store3(base, idx[k], comp(arr, k, 0), comp(arr, k, 1), comp(arr, k, 2));
    };
    const f32x2 g2 = pack2(p.g, p.g), dt2 = pack2(p.dt, p.dt), h2 = pack2(p.half_dt, p.half_dt);

    if (p.mode == MODE_ACCEL) {
        batched_force<kPairs, kExactDiag>(cur, p.n, p.eps2, x, idx, sum);
#pragma unroll
        for (int q = 0; q < kPairs; ++q)
#pragma unroll
            for (int c = 0; c < 3; ++c) a[q][c] = mul2(g2, sum[q][c]);
        store_state(gacc, a);
        return;
    }

    // publishes this thread's drifted bodies into every CTA's copy of `dst`
    auto publish = [&](float4* dst) {
#pragma unroll
        for (int k = 0; k < kI; ++k) {
            if (!valid[k]) continue;
            const float4 b = make_float4(comp(x, k, 0), comp(x, k, 1), comp(x, k, 2), m[k]);
            for (int r = 0; r < csize; ++r) cluster.map_shared_rank(dst, r)[idx[k]] = b;
        }
    };

    // Each update is a rounded multiply followed by a rounded add (mul.rn / add.rn per lane), never an FMA, as torch
    // evaluates `v += c * a` (simulation.py:164-170, 183-187).
    const size_t plane = size_t(p.n_systems) * p.n * 3;
    for (int s = 0; s < p.steps; ++s) {
        if (p.mode == MODE_LEAPFROG) {
            // half-kick + drift (simulation.py:164-166), publish the drifted bodies, then force + closing half-kick
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    v[q][c] = add2(v[q][c], mul2(h2, a[q][c]));
                    x[q][c] = add2(x[q][c], mul2(dt2, v[q][c]));
                }
            publish(nxt);
            cluster.sync();
            batched_force<kPairs, kExactDiag>(nxt, p.n, p.eps2, x, idx, sum);
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[q][c] = mul2(g2, sum[q][c]);
                    v[q][c] = add2(v[q][c], mul2(h2, a[q][c]));
                }
        } else {
            // force at the current positions, then v += dt*a ; x += dt*v (simulation.py:183-187)
            batched_force<kPairs, kExactDiag>(cur, p.n, p.eps2, x, idx, sum);
#pragma unroll
            for (int q = 0; q < kPairs; ++q)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a[q][c] = mul2(g2, sum[q][c]);
                    v[q][c] = add2(v[q][c], mul2(dt2, a[q][c]));
                    x[q][c] = add2(x[q][c], mul2(dt2, v[q][c]));
                }
            publish(nxt);
            cluster.sync();
        }
        float4* tmp = cur;
        cur = nxt;
        nxt = tmp;

        if (p.traj && (s + 1) % p.record_every == 0) {
            float* slot = p.traj + size_t((s + 1) / p.record_every - 1) * 3 * plane + base3;
            store_state(slot, x);
            store_state(slot + plane, v);
            store_state(slot + 2 * plane, a);
        }
    }
    store_state(gpos, x);
    store_state(gvel, v);
    store_state(gacc, a);
    cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

}  // namespace nb
