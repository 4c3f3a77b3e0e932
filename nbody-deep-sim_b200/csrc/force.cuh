// All-pairs softened gravitational acceleration with the integrator fused into the epilogue.
//
// Replaces (reference, read-only): src/galaxify/simulation.py:71-89 (compute_accelerations),
// :153-170 (LeapFrogSimulator.step), :173-187 (EulerSimulator.step).
//
// Shape of the kernel
//   grid  = (i_tiles, j_splits); block = kWarps warps, all computing; lane 0 of warp 0 also feeds the TMA ring.
//   Each compute thread owns 2*kPairs i-bodies in registers, packed two-by-two into f32x2 lanes so that the
//   whole inner loop issues FADD2/FFMA2/FMUL2 (sm_100 packed FP32) with the j-body as a scalar-broadcast
//   operand: 12 packed instructions + 2 MUFU.RSQ per (i-pair, j) = 6 FMA-pipe issues + 1 MUFU per interaction.
//   j-bodies (x,y,z,m float4) stream through a kStages-deep shared-memory ring filled kLookahead tiles ahead with
//   1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx); warps release stages with mbarrier arrives, so
//   there is no CTA-wide barrier in the tile loop and warps may drift up to kStages-kLookahead tiles apart.
//   Accumulation: FP32 runs of kFold (32) j-bodies, each run folded into an FP64 total (F2F + DADD, off the FMA
//   pipe). A long plain-FP32 run loses the small terms that follow a large one; for bodies whose forces cancel to a
//   few 1e-3 of their sum (close to a galaxy's centre) 1024-term runs gave 1e-4 relative error where the reference's
//   cascade summation gives 1e-6 — 32-term runs bring it to the reference's level (tools/diag_shard_accuracy.py).
//   With j_splits > 1 every CTA writes its FP64 partial to scratch and the last CTA to arrive for an i-tile reduces
//   the splits in fixed order (deterministic) and runs the epilogue.
//
// Epilogue (per i-body, rounding exactly as the reference: multiply and add rounded separately, no FMA)
//   a = fl32(G) * sum
//   ACCEL    : store a
//   LEAPFROG : v = vhalf + h*a (state k complete) ; if do_next: vhalf = v + h*a ; x' = x + dt*vhalf
//   EULER    : v = v + dt*a ; x' = x + dt*v
#pragma once
#include "async_copy.cuh"

namespace nb {

constexpr int kStages = 4;
constexpr int kLookahead = 2;  // tiles in flight ahead of the one being consumed

enum Mode : int { MODE_ACCEL = 0, MODE_LEAPFROG = 1, MODE_EULER = 2 };

struct ForceParams {
    const float4* bodies;  // current (x,y,z,m) of ALL bodies: j source, and i source at [i_begin, i_begin+i_count)
    float4* bodies_next;   // where drifted bodies are written (global index), may be null
    int j_begin, j_end;    // j range this launch covers ...
    int j2_begin, j2_end;  // ... plus an optional second range (empty when j2_begin >= j2_end)
    int i_begin, i_count;  // i range this launch covers (global index of first, count)
    float eps2, g;
    // split-j reduction scratch
    double* partial;     // [splits_total][3][partial_stride] FP64 partial sums
    int partial_stride;  // >= i_count
    int split_offset;    // slot of this launch's split 0 (lets several launches share one reduction)
    int splits_total;    // arrivals per i-tile that trigger the epilogue
    unsigned* counters;  // [i_tiles], zero before the first launch; the finishing CTA resets its slot
    // epilogue
    int mode;
    int do_next;
    float dt, half_dt;
    // state arrays, (i_count,3) floats, indexed by local i
    float* pos;
    float* vel;
    float* acc;
    float* vhalf;
    // optional trajectory slots, same layout, may be null
    float* rec_pos;
    float* rec_vel;
    float* rec_acc;
};

__device__ __forceinline__ void store3(float* base, int i, float x, float y, float z) {
    if (base) {
        base[3 * i + 0] = x;
        base[3 * i + 1] = y;
        base[3 * i + 2] = z;
    }
}

// One i-body's integrator update. `sum` is the un-scaled j-sum; `me` is the body's current (x,y,z,m).
__device__ __forceinline__ void epilogue_body(const ForceParams& p, int li, float4 me, float sx, float sy, float sz) {
    const float ax = __fmul_rn(p.g, sx), ay = __fmul_rn(p.g, sy), az = __fmul_rn(p.g, sz);
    store3(p.acc, li, ax, ay, az);
    store3(p.rec_acc, li, ax, ay, az);
    if (p.mode == MODE_ACCEL) return;

    float vx, vy, vz;
    if (p.mode == MODE_LEAPFROG) {
        const float h = p.half_dt;
        vx = __fadd_rn(p.vhalf[3 * li + 0], __fmul_rn(h, ax));
        vy = __fadd_rn(p.vhalf[3 * li + 1], __fmul_rn(h, ay));
        vz = __fadd_rn(p.vhalf[3 * li + 2], __fmul_rn(h, az));
        store3(p.vel, li, vx, vy, vz);
        store3(p.rec_vel, li, vx, vy, vz);
        if (!p.do_next) return;
        vx = __fadd_rn(vx, __fmul_rn(h, ax));
        vy = __fadd_rn(vy, __fmul_rn(h, ay));
        vz = __fadd_rn(vz, __fmul_rn(h, az));
        store3(p.vhalf, li, vx, vy, vz);
    } else {  // MODE_EULER
        vx = __fadd_rn(p.vel[3 * li + 0], __fmul_rn(p.dt, ax));
        vy = __fadd_rn(p.vel[3 * li + 1], __fmul_rn(p.dt, ay));
        vz = __fadd_rn(p.vel[3 * li + 2], __fmul_rn(p.dt, az));
        store3(p.vel, li, vx, vy, vz);
        store3(p.rec_vel, li, vx, vy, vz);
    }
    const float nx = __fadd_rn(me.x, __fmul_rn(p.dt, vx));
    const float ny = __fadd_rn(me.y, __fmul_rn(p.dt, vy));
    const float nz = __fadd_rn(me.z, __fmul_rn(p.dt, vz));
    store3(p.pos, li, nx, ny, nz);
    store3(p.rec_pos, li, nx, ny, nz);
    if (p.bodies_next) p.bodies_next[p.i_begin + li] = make_float4(nx, ny, nz, me.w);
}

// kExactDiag: softening^2 underflows FP32 (or is 0), so the self term would be 0*inf. Mask it by index, which is
// what fill_diagonal_(0) does in the reference (simulation.py:85); two distinct coincident bodies still give NaN there.
template <int kPairs, int kWarps, int kMinBlocks, int kTileJ, bool kExactDiag, int kUnroll = 4, int kFold = 32>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks) force_kernel(const ForceParams p) {
    constexpr int kCT = kWarps * 32;  // threads
    constexpr int kI = 2 * kPairs;    // i-bodies per thread
    constexpr int kTileI = kCT * kI;
    using Ring = TileRing<kTileJ, kStages, kWarps>;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_is_last;

    const int tid = threadIdx.x;

    // j ranges of this split: an even share of each of the launch's (one or two) ranges
    auto share = [&](int begin, int end, int& lo, int& hi) {
        const int len = max(end - begin, 0);
        const int per = (len + gridDim.y - 1) / gridDim.y;
        lo = begin + min(int(blockIdx.y) * per, len);
        hi = begin + min(int(blockIdx.y + 1) * per, len);
    };
    int j0, j1, k0, k1;
    share(p.j_begin, p.j_end, j0, j1);
    share(p.j2_begin, p.j2_end, k0, k1);

    Ring ring;
    ring.attach(smem_raw, p.bodies + j0, j1 - j0, p.bodies + k0, k1 - k0);
    const int ntiles = ring.num_tiles();
    if (tid == 0) ring.init_barriers();
    __syncthreads();
    if (tid == 0)
        for (int t = 0; t < min(kLookahead, ntiles); ++t) ring.issue(t);

    const int tile_base = blockIdx.x * kTileI;
    float4 me[kI];
    int li[kI];
    float2 nx[kPairs], ny[kPairs], nz[kPairs];  // negated i positions, packed
#pragma unroll
    for (int k = 0; k < kI; ++k) {
        li[k] = tile_base + k * kCT + tid;
        me[k] = p.bodies[p.i_begin + min(li[k], p.i_count - 1)];
    }
#pragma unroll
    for (int q = 0; q < kPairs; ++q) {
        nx[q] = make_float2(-me[2 * q].x, -me[2 * q + 1].x);
        ny[q] = make_float2(-me[2 * q].y, -me[2 * q + 1].y);
        nz[q] = make_float2(-me[2 * q].z, -me[2 * q + 1].z);
    }
    float2 ax[kPairs], ay[kPairs], az[kPairs];
    double tot[kI][3];
#pragma unroll
    for (int k = 0; k < kI; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) tot[k][c] = 0.0;

    const float2 eps2 = make_float2(p.eps2, p.eps2);

    for (int t = 0; t < ntiles; ++t) {
        if (tid == 0 && t + kLookahead < ntiles) ring.issue(t + kLookahead);
        const int jt = (ring.segment(t) == 0 ? j0 : k0) + ring.tile_offset(t);  // global index of the tile's first body
        const int count = ring.tile_count(t);
        const float4* __restrict__ tj = ring.tile(t);
        ring.wait(t);

        // one (i-pair, j) interaction, packed over the pair
        auto interact = [&](int jj) {
            const float4 b = tj[jj];
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
                    const int jg = jt + jj;
                    if (jg == p.i_begin + li[2 * q]) ri.x = 0.f;
                    if (jg == p.i_begin + li[2 * q + 1]) ri.y = 0.f;
                }
                const float2 ri2 = __fmul2_rn(ri, ri);
                const float2 mri = __fmul2_rn(ri, bm);
                const float2 w = __fmul2_rn(ri2, mri);
                ax[q] = __ffma2_rn(w, dx, ax[q]);
                ay[q] = __ffma2_rn(w, dy, ay[q]);
                az[q] = __ffma2_rn(w, dz, az[q]);
            }
        };

        for (int jb = 0; jb < count; jb += kFold) {
#pragma unroll
            for (int q = 0; q < kPairs; ++q) ax[q] = ay[q] = az[q] = make_float2(0.f, 0.f);
            if (jb + kFold <= count) {
#pragma unroll kUnroll
                for (int u = 0; u < kFold; ++u) interact(jb + u);
            } else {
                for (int jj = jb; jj < count; ++jj) interact(jj);
            }
            // fold the run into the FP64 totals
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                tot[2 * q][0] += double(ax[q].x), tot[2 * q][1] += double(ay[q].x), tot[2 * q][2] += double(az[q].x);
                tot[2 * q + 1][0] += double(ax[q].y), tot[2 * q + 1][1] += double(ay[q].y);
                tot[2 * q + 1][2] += double(az[q].y);
            }
        }
        ring.release(t);
    }

    if (p.splits_total == 1) {
#pragma unroll
        for (int k = 0; k < kI; ++k)
            if (li[k] < p.i_count) epilogue_body(p, li[k], me[k], float(tot[k][0]), float(tot[k][1]), float(tot[k][2]));
        return;
    }

    // ---- split-j: publish the partial, last arriver reduces in slot order ----
    double* mine = p.partial + size_t(p.split_offset + blockIdx.y) * 3 * p.partial_stride;
#pragma unroll
    for (int k = 0; k < kI; ++k)
        if (li[k] < p.i_count) {
#pragma unroll
            for (int c = 0; c < 3; ++c) mine[size_t(c) * p.partial_stride + li[k]] = tot[k][c];
        }
    __threadfence();
    compute_barrier<kCT>();
    if (tid == 0) {
        const unsigned prev = atomicAdd(&p.counters[blockIdx.x], 1u);
        s_is_last = (prev == unsigned(p.splits_total - 1));
    }
    compute_barrier<kCT>();
    if (!s_is_last) return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < kI; ++k) {
        if (li[k] >= p.i_count) continue;
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (int s = 0; s < p.splits_total; ++s) {
            const double* src = p.partial + size_t(s) * 3 * p.partial_stride + li[k];
            sx += __ldcg(src);
            sy += __ldcg(src + p.partial_stride);
            sz += __ldcg(src + 2 * size_t(p.partial_stride));
        }
        epilogue_body(p, li[k], me[k], float(sx), float(sy), float(sz));
    }
    if (tid == 0) p.counters[blockIdx.x] = 0u;
}

// O(N) pre-pass of a stepping call: builds the (x,y,z,m) body array from the reference-layout state and, for
// leapfrog, performs the opening half-kick + drift of the first step (simulation.py:164-166).
struct PrepParams {
    int n;          // local bodies
    int i_begin;    // global index of local body 0 in `bodies`
    int mode;
    float dt, half_dt;
    const float* mass;
    float* pos;
    const float* vel;
    const float* acc;
    float* vhalf;
    float4* bodies;
    float* rec_pos;
};

__global__ void prep_kernel(const PrepParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float x = p.pos[3 * i], y = p.pos[3 * i + 1], z = p.pos[3 * i + 2];
    if (p.mode == MODE_LEAPFROG) {
        const float vx = __fadd_rn(p.vel[3 * i + 0], __fmul_rn(p.half_dt, p.acc[3 * i + 0]));
        const float vy = __fadd_rn(p.vel[3 * i + 1], __fmul_rn(p.half_dt, p.acc[3 * i + 1]));
        const float vz = __fadd_rn(p.vel[3 * i + 2], __fmul_rn(p.half_dt, p.acc[3 * i + 2]));
        store3(p.vhalf, i, vx, vy, vz);
        x = __fadd_rn(x, __fmul_rn(p.dt, vx));
        y = __fadd_rn(y, __fmul_rn(p.dt, vy));
        z = __fadd_rn(z, __fmul_rn(p.dt, vz));
        store3(p.pos, i, x, y, z);
        store3(p.rec_pos, i, x, y, z);
    }
    p.bodies[p.i_begin + i] = make_float4(x, y, z, p.mass[i]);
}

}  // namespace nb
