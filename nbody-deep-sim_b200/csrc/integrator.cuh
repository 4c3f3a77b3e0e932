// Leapfrog kick/drift with a caller-supplied force, and the momentum diagnostics.
//
// Replaces (reference, read-only): trainer.py:217-226 (Trainer.step) and gnn.py:223-232 (GraphModel.step), the
// kick-drift-kick update the surrogate-model rollout wraps around `model.predict`:
//     vel_ = vel + 0.5*dt*acc ; pos_ = pos + dt*vel_ ; acc_ = predict(pos_, ...) ; vel_ += 0.5*dt*acc_
// i.e. LeapFrogSimulator.step (src/galaxify/simulation.py:164-170) with the force slot left to the caller. Every
// update is fl32(a + fl32(c*b)): a separately rounded multiply and add, never an FMA, as torch evaluates it.
//
// Both kernels are HBM-bound elementwise passes over the (n,3) state, 128-bit vectorised when the arrays allow it:
//   kick_drift: 3 reads + 2 writes = 20 B per scalar; kick: 2 reads + 1 write = 12 B per scalar.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb {

struct KickDriftParams {
    long long count;  // scalars = 3 * n
    float dt, half_dt;
    const float* pos;
    const float* vel;
    const float* acc;
    float* pos_out;  // may alias pos
    float* vel_out;  // may alias vel
};

__device__ __forceinline__ float kick1(float v, float c, float a) { return __fadd_rn(v, __fmul_rn(c, a)); }

template <bool kVec4>
__global__ void __launch_bounds__(256) kick_drift_kernel(const KickDriftParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (kVec4) {
        const long long n4 = p.count / 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 x = reinterpret_cast<const float4*>(p.pos)[i];
            const float4 v = reinterpret_cast<const float4*>(p.vel)[i];
            const float4 a = reinterpret_cast<const float4*>(p.acc)[i];
            float4 nv, nx;
            nv.x = kick1(v.x, p.half_dt, a.x), nv.y = kick1(v.y, p.half_dt, a.y);
            nv.z = kick1(v.z, p.half_dt, a.z), nv.w = kick1(v.w, p.half_dt, a.w);
            nx.x = kick1(x.x, p.dt, nv.x), nx.y = kick1(x.y, p.dt, nv.y);
            nx.z = kick1(x.z, p.dt, nv.z), nx.w = kick1(x.w, p.dt, nv.w);
            reinterpret_cast<float4*>(p.vel_out)[i] = nv;
            reinterpret_cast<float4*>(p.pos_out)[i] = nx;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += stride) {
            const float nv = kick1(p.vel[i], p.half_dt, p.acc[i]);
            p.vel_out[i] = nv;
            p.pos_out[i] = kick1(p.pos[i], p.dt, nv);
        }
    }
}

struct KickParams {
    long long count;
    float half_dt;
    const float* vel;
    const float* acc;
    float* vel_out;  // may alias vel
};

template <bool kVec4>
__global__ void __launch_bounds__(256) kick_kernel(const KickParams p) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (kVec4) {
        const long long n4 = p.count / 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const float4 v = reinterpret_cast<const float4*>(p.vel)[i];
            const float4 a = reinterpret_cast<const float4*>(p.acc)[i];
            float4 nv;
            nv.x = kick1(v.x, p.half_dt, a.x), nv.y = kick1(v.y, p.half_dt, a.y);
            nv.z = kick1(v.z, p.half_dt, a.z), nv.w = kick1(v.w, p.half_dt, a.w);
            reinterpret_cast<float4*>(p.vel_out)[i] = nv;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.count; i += stride)
            p.vel_out[i] = kick1(p.vel[i], p.half_dt, p.acc[i]);
    }
}

// Total linear momentum sum m v and angular momentum sum m (x cross v) of a state, FP64 sums of FP64 products of the
// FP32 inputs, fixed order (deterministic). The reference has no momentum function; SURVEY.md §8a/§8f1 name the
// momentum DRIFT as a parity metric next to the energy drift, and at N >= 32k it has to be evaluated on the device.
struct MomentumParams {
    int n;
    const float* pos;   // (n,3) or null (then the angular part is zero)
    const float* vel;   // (n,3)
    const float* mass;  // (n)
    double* out;        // 6 doubles: px, py, pz, lx, ly, lz
};

__global__ void __launch_bounds__(1024) momentum_kernel(const MomentumParams p) {
    __shared__ double s_red[32][6];
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < p.n; i += 1024) {
        const double m = p.mass[i];
        const double vx = p.vel[3 * i], vy = p.vel[3 * i + 1], vz = p.vel[3 * i + 2];
        acc[0] += m * vx, acc[1] += m * vy, acc[2] += m * vz;
        if (p.pos) {
            const double x = p.pos[3 * i], y = p.pos[3 * i + 1], z = p.pos[3 * i + 2];
            acc[3] += m * (y * vz - z * vy), acc[4] += m * (z * vx - x * vz), acc[5] += m * (x * vy - y * vx);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += s_red[w][threadIdx.x];
        p.out[threadIdx.x] = t;
    }
}

}  // namespace nb
