// mbarrier + TMA bulk-copy primitives for sm_100a (inline PTX).
// Used by the j-body tile pipeline of the force and energy kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Makes mbarrier.init visible to the async proxy (the TMA unit) before the first bulk copy.
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// 1-D TMA bulk copy global -> shared; completion is signalled as `bytes` of transaction count on `bar`.
// `bytes` must be a multiple of 16 and both addresses 16-byte aligned. SASS: UBLKCP.
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Shared-memory ring of j-body tiles fed by TMA bulk copies.
//
// The j range is one or two contiguous segments of the body array (two when a launch covers "everything but the
// rank's own slice"); tiles are numbered through both. One lane of the CTA is the producer: issue(t) arms
// full[t % kStages] with the tile's byte count and starts the copy; it first waits on empty[] until every warp has
// released the stage's previous occupant. Consumers call wait(t) before reading tile t and release(t) (one arrive
// per warp) when done. Tiles are kTileJ float4 bodies.
template <int kTileJ, int kStages, int kWarps>
struct TileRing {
    float4* tiles;
    uint64_t* full;
    uint64_t* empty;
    const float4* src[2];  // first body of each segment
    int count[2];          // bodies in each segment
    int tiles0;            // tiles of segment 0

    static constexpr size_t smem_bytes() {
        return size_t(kStages) * kTileJ * sizeof(float4) + 2 * kStages * sizeof(uint64_t);
    }
    __device__ __forceinline__ void attach(unsigned char* smem, const float4* src0, int count0,
                                           const float4* src1 = nullptr, int count1 = 0) {
        tiles = reinterpret_cast<float4*>(smem);
        full = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * kTileJ * sizeof(float4));
        empty = full + kStages;
        src[0] = src0, src[1] = src1;
        count[0] = max(count0, 0), count[1] = max(count1, 0);
        tiles0 = (count[0] + kTileJ - 1) / kTileJ;
    }
    // Call from one thread, then __syncthreads().
    __device__ __forceinline__ void init_barriers() {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWarps);
        }
        mbar_fence_init();
    }
    __device__ __forceinline__ int num_tiles() const { return tiles0 + (count[1] + kTileJ - 1) / kTileJ; }
    __device__ __forceinline__ int segment(int t) const { return t < tiles0 ? 0 : 1; }
    // Offset of tile t's first body within its segment.
    __device__ __forceinline__ int tile_offset(int t) const { return (t < tiles0 ? t : t - tiles0) * kTileJ; }
    __device__ __forceinline__ int tile_count(int t) const { return min(kTileJ, count[segment(t)] - tile_offset(t)); }
    __device__ __forceinline__ const float4* tile(int t) const { return tiles + size_t(t % kStages) * kTileJ; }
    __device__ __forceinline__ void issue(int t) {
        const int s = t % kStages;
        if (t >= kStages) mbar_wait(&empty[s], ((t / kStages) & 1) ^ 1);
        const uint32_t bytes = uint32_t(tile_count(t)) * sizeof(float4);
        mbar_arrive_expect_tx(&full[s], bytes);
        bulk_copy_g2s(tiles + size_t(s) * kTileJ, src[segment(t)] + tile_offset(t), bytes, &full[s]);
    }
    __device__ __forceinline__ void wait(int t) { mbar_wait(&full[t % kStages], (t / kStages) & 1); }
    // All lanes of a warp call this after their last read of tile t.
    __device__ __forceinline__ void release(int t) {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[t % kStages]);
    }
};

// Named barrier over `kCount` threads (id 1; id 0 is __syncthreads).
template <int kCount>
__device__ __forceinline__ void compute_barrier() {
    asm volatile("bar.sync 1, %0;" ::"n"(kCount) : "memory");
}

// One MUFU.RSQ. Denormal inputs flush to zero (-> +inf); callers keep r2 >= FLT_MIN or mask the term.
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One 32-bit read-only global load that the compiler may not merge with its neighbours into a vector load.
__device__ __forceinline__ float ldg_f32(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Packed FP32 pairs as opaque 64-bit registers. The CUDA float2 intrinsics (__ffma2_rn, ...) split a pair into two
// 32-bit values and re-pack them (mov.b64) at every use, and ptxas does not always keep loop-invariant pairs in an
// aligned register pair: the pair kernel's systolic loop carried 9 MOVs per step for that. A value of this type IS the
// register pair, so nothing has to be re-packed.
using f32x2 = unsigned long long;

__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) {
    float lo;
    asm("{ .reg .f32 t; mov.b64 {%0, t}, %1; }" : "=f"(lo) : "l"(v));
    return lo;
}
__device__ __forceinline__ float hi2(f32x2 v) {
    float hi;
    asm("{ .reg .f32 t; mov.b64 {t, %0}, %1; }" : "=f"(hi) : "l"(v));
    return hi;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

}  // namespace nb
