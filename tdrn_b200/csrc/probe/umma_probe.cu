// umma_probe.cu -- development probe (not part of the public ABI): which shared-memory bytes does a
// SWIZZLE_128B K-major UMMA A-descriptor read when its start address is not 1024-byte aligned and/or its
// stride-byte-offset (SBO) is not a multiple of 1024?  The halo-tile convolution (conv_tc.cu) addresses the
// nine taps of a 3x3 kernel as row-shifted views of ONE shared-memory tile and depends on the answer.
//
// Smem rows (128 B = 64 bf16) are filled in the layout TMA SWIZZLE_128B produces (16-byte chunk c of absolute
// row R stored at chunk position c ^ (R & 7)); value = R (mode 0) or k (mode 1).  B = 64x64 identity, so
// D[m][n] = A[m][n] and the output reveals the (row, column) every A element was fetched from.
#include "../halo_common.cuh"

namespace tdrn {
namespace tc {

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(float *out, int r0, int sbo_bytes, int base_offset, int mode, int rows)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                        // rows x 128 B
    uint8_t *sB = base + ((rows * 128 + 1023) & ~1023);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < rows * 64; e += 128) {
        const int R = e >> 6, k = e & 63;
        const float v = mode == 0 ? (float)R : (float)k;
        *(__nv_bfloat16 *)(sA + R * 128 + (((k >> 3) ^ (R & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
    }
    for (int e = tid; e < 64 * 64; e += 128) {
        const int n = e >> 6, k = e & 63;
        *(__nv_bfloat16 *)(sB + n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(n == k ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t a_addr = smem_u32(sA) + (uint32_t)r0 * 128u;
        uint64_t ad = 0;
        ad |= (uint64_t)((a_addr & 0x3FFFFu) >> 4);
        ad |= (uint64_t)1 << 16;
        ad |= (uint64_t)((uint32_t)sbo_bytes >> 4) << 32;
        ad |= (uint64_t)1 << 46;
        ad |= (uint64_t)(base_offset & 7) << 49;
        ad |= (uint64_t)2 << 61;
        const uint64_t bd = umma_desc_sw128(smem_u32(sB));
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, k != 0);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < 64; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);
        for (int j = 0; j < 16; ++j) out[tid * 64 + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

}  // namespace tc
}  // namespace tdrn

using namespace tdrn;
using namespace tdrn::tc;

// out: device [128][64] fp32.  Synchronous.  Development aid (scripts/umma_probe.py), not declared in the header.
extern "C" int tdrn_debug_umma_probe(float *out, int r0, int sbo_bytes, int base_offset, int mode, int rows)
{
    TDRN_REQUIRE(out && rows > 0 && rows <= 512, "umma probe: bad argument");
    const int smem = ((rows * 128 + 1023) & ~1023) + 64 * 128 + 1024;
    TDRN_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_probe_kernel<<<1, 128, smem>>>(out, r0, sbo_bytes, base_offset, mode, rows);
    TDRN_LAUNCH_CHECK();
    TDRN_CUDA(cudaDeviceSynchronize());
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// MMA issue-rate probe: every CTA issues `iters` groups of 4 x (M=128, N=n, K=16) tcgen05.mma on smem-resident
// operands (no TMA, no epilogue), alternating between `nacc` TMEM accumulators, and reports clock64 cycles.
// Answers "what can the tensor pipe sustain for narrow N tiles" (the bound of conv_halo_tc.cu's main loop).
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
namespace tc {

__global__ void __launch_bounds__(128, 1) umma_rate_kernel(long long *cycles, int n, int iters, int nacc, int a_shift_rows)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                        // 24 KB
    uint8_t *sB = base + 24 * 1024;            // 256 x 128 B
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (24 * 1024 + 32 * 1024) / 4; e += 128) ((uint32_t *)base)[e] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, n);
        const uint64_t bd = umma_desc_sw128(smem_u32(sB));
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a_addr = smem_u32(sA) + (uint32_t)((i % 9) * a_shift_rows) * 128u;
            const uint64_t ad = umma_desc_sw128(a_addr);
            const uint32_t d = tmem_base + (uint32_t)((i % nacc) * n);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tdrn

extern "C" int tdrn_debug_umma_rate(long long *cycles_dev, int grid, int n, int iters, int nacc, int a_shift_rows)
{
    TDRN_REQUIRE(cycles_dev && grid > 0 && n >= 16 && n <= 256 && n % 16 == 0 && nacc >= 1 && nacc * n <= 512, "umma rate: bad argument");
    const int smem = 24 * 1024 + 32 * 1024 + 1024;
    TDRN_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_rate_kernel<<<grid, 128, smem>>>(cycles_dev, n, iters, nacc, a_shift_rows);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Second rate probe (development): is the 73-cycle cost of narrow-N MMAs the shared-memory read of the A operand?
// (Superseded: the 71-73 cycles of this probe and of umma_rate_kernel are their own issue loops -- the run-time `i % nacc`
// per group of four MMAs -- see umma_rate_bg below, which measures 48 / 64 / 128 cycles at N = 64 / 128 / 256.)
//   mode 0  SS  : A and B from shared memory (the case above, repeated as the reference point of the same run)
//   mode 1  TS  : A from tensor memory ([taddr] operand), B from shared memory
//   mode 2  CP  : tcgen05.cp.128x256b only (one K = 16 slice of a 128-row A tile, shared -> tensor memory)
//   mode 3  CP+TS: every MMA preceded by the tcgen05.cp of its A slice into one of two tensor-memory slots
//   mode 4  SS, cta_group::2: M = 256 over an SM pair (cluster of two CTAs, the leader issues), N = n
//   mode 5  SS, A and B descriptors recomputed once per group of 4 MMAs (what a k-block loop does)
//   mode 6  SS, A and B descriptors recomputed before EVERY MMA (what the nine-tap halo loop does per tap, taken further)
// Values are not checked (operands are constant fills); only the issue rate is read.
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
namespace tc {

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t s_desc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) umma_rate2_kernel(long long *cycles, int n, int iters, int nacc)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                        // 24 KB, as in umma_rate_kernel
    uint8_t *sB = base + 24 * 1024;            // 256 x 128 B (+ 4 KB: modes 5/6 shift the B start by up to 3 KB)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (24 * 1024 + 36 * 1024) / 4; e += 128) ((uint32_t *)base)[e] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, n);
        const uint64_t ad = umma_desc_sw128(smem_u32(sA));
        const uint64_t bd = umma_desc_sw128(smem_u32(sB));
        const uint32_t a_tm = tmem_base + 448u;           // A slots: 8 columns (16 bf16) per K = 16 slice
        if (MODE == 1) {                                   // give the TS MMAs defined A values
#pragma unroll
            for (int k = 0; k < 4; ++k) tmem_cp_128x256b(a_tm + (uint32_t)(k * 8), ad + (uint64_t)(k * 2));
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem_base + (uint32_t)((i % nacc) * n);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (MODE == 0) umma_bf16(d, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u);
                else if (MODE == 5 || MODE == 6) {
                    const uint32_t j = MODE == 5 ? (uint32_t)i : (uint32_t)(i * 4 + k);
                    const uint64_t ax = umma_desc_sw128(smem_u32(sA) + (j & 7u) * 128u);      // cheap, not hoistable
                    const uint64_t bx = umma_desc_sw128(smem_u32(sB) + ((j >> 1) & 3u) * 1024u);
                    umma_bf16(d, ax + (uint64_t)(k * 2), bx + (uint64_t)(k * 2), idesc, 1u);
                }
                else if (MODE == 1) umma_bf16_ts(d, a_tm + (uint32_t)(k * 8), bd + (uint64_t)(k * 2), idesc, 1u);
                else if (MODE == 2) tmem_cp_128x256b(a_tm + (uint32_t)(k * 8), ad + (uint64_t)(k * 2));
                else {                                     // slot alternates so that cp k+1 may run under MMA k
                    const uint32_t slot = a_tm + (uint32_t)(((i * 4 + k) & 1) * 32 + k * 8);
                    tmem_cp_128x256b(slot, ad + (uint64_t)(k * 2));
                    umma_bf16_ts(d, slot, bd + (uint64_t)(k * 2), idesc, 1u);
                }
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// cta_group::2: both CTAs of the pair allocate, the leader (cluster rank 0) issues M = 256 MMAs whose A descriptor
// names the 128 rows each CTA holds at the same shared-memory offset and whose B descriptor names N/2 rows per CTA;
// the commit is multicast to the barrier of both CTAs so that the peer stays resident until the MMAs have drained.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate_cg2_kernel(long long *cycles, int n, int iters, int nacc)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;
    uint8_t *sB = base + 16 * 1024;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int e = tid; e < (16 * 1024 + 32 * 1024) / 4; e += 128) ((uint32_t *)base)[e] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        long long t0 = 0;
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_bf16(256, n);
            const uint64_t ad = umma_desc_sw128(smem_u32(sA));
            const uint64_t bd = umma_desc_sw128(smem_u32(sB));
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem_base + (uint32_t)((i % nacc) * n);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    asm volatile("{\n\t.reg .pred p;\n\t"
                                 "setp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(ad + (uint64_t)(k * 2)), "l"(bd + (uint64_t)(k * 2)), "r"(idesc), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        }
        mbar_wait(&bar, 0);
        if (rank == 0) cycles[blockIdx.x >> 1] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace tc
}  // namespace tdrn

// ---------------------------------------------------------------------------------------------------------
// Third rate probe (development): the MMA rate with loop-invariant descriptors, next to OTHER shared-memory traffic, and
// with the halo kernels' shifted A views.  Production N = 128 layers run at ~100 cycles per MMA even with their epilogue compiled out; the tensor core alone reads A 4 KB + B 4 KB per
// instruction = 115 of the 128 B/clk of shared-memory bandwidth, so the TMA writes of the next halo box (and any staged
// epilogue) may be what stretches them.  cta_group::2 halves the B rows each SM reads (6 KB per instruction = 86 B/clk).
// Warps 1..bg_warps stream 16-byte-per-lane loads (bg_kind 0) or stores (1) over an 8 KB region while thread 0 issues
// the MMAs; `gap` idle cycles between batches of 8 throttle them.  Reports MMA cycles and background bytes per CTA.
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
namespace tc {

template <int CG>
__device__ __forceinline__ void umma_rate_bg_body(long long *cycles, unsigned long long *bg_bytes, int n, int iters, int bg_warps,
                                                  int bg_kind, int gap, int a_shift_rows, int a_sbo_bytes)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int stop;
    __shared__ unsigned long long bg_total;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = base;                        // 24 KB (128 rows at up to 1280 B per 8-row group + a start shift)
    uint8_t *sB = base + 24 * 1024;            // 32 KB (256 rows)
    uint8_t *sG = base + 56 * 1024;            // 8 KB of background traffic
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    for (int e = tid; e < (64 * 1024) / 4; e += 128) ((uint32_t *)base)[e] = 0x3c003c00u;
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; bg_total = 0ull; }
    fence_proxy_async_smem();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(&tmem_base_s, 512);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        long long t0 = 0;
        if (rank == 0) {
            const uint32_t idesc = umma_idesc_bf16(CG == 2 ? 256 : 128, n);
            // the halo-tile kernels' A view: start shifted by whole 128-byte rows, 8-row groups a_sbo_bytes apart
            const uint64_t ad = umma_desc_sw128_sbo(smem_u32(sA) + (uint32_t)a_shift_rows * 128u, (uint32_t)a_sbo_bytes);
            const uint64_t bd = umma_desc_sw128(smem_u32(sB));
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {      // no per-iteration arithmetic: descriptors stay in uniform registers
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (CG == 2)
                        asm volatile("{\n\t.reg .pred p;\n\t"
                                     "setp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(tmem_base), "l"(ad + (uint64_t)(k * 2)), "l"(bd + (uint64_t)(k * 2)), "r"(idesc), "r"(1u) : "memory");
                    else
                        umma_bf16(tmem_base, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, 1u);
                }
            }
            if (CG == 2)
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
            else
                umma_commit(&bar);
        }
        mbar_wait(&bar, 0);
        if (rank == 0) cycles[CG == 2 ? blockIdx.x >> 1 : blockIdx.x] = clock64() - t0;
        *(volatile int *)&stop = 1;
    } else if (warp >= 1 && warp <= bg_warps) {
        // background traffic: 8 independent 16-byte accesses per lane per batch (4 KB per warp-batch), conflict-free
        const uint32_t a0 = smem_u32(sG) + (uint32_t)lane * 16u;
        unsigned long long batches = 0;
        uint32_t x = 0;
        while (*(volatile int *)&stop == 0) {
            if (bg_kind == 0) {
                uint32_t r[32];
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(r[4 * u]), "=r"(r[4 * u + 1]), "=r"(r[4 * u + 2]), "=r"(r[4 * u + 3]) : "r"(a0 + (uint32_t)u * 512u) : "memory");
#pragma unroll
                for (int u = 0; u < 32; ++u) x ^= r[u];
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a0 + (uint32_t)u * 512u), "r"(x) : "memory");
                x += 1u;
            }
            ++batches;
            if (gap > 0) { const long long t = clock64(); while (clock64() - t < gap) { } }
        }
        if (x == 0x12345678u) sG[0] = 1;           // keep the loads alive
        if (lane == 0) atomicAdd(&bg_total, batches * 4096ull);
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) bg_bytes[blockIdx.x] = bg_total;
    if (CG == 2) cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else tmem_dealloc(tmem_base, 512);
    }
}

__global__ void __launch_bounds__(128, 1) umma_rate_bg1_kernel(long long *cycles, unsigned long long *bg_bytes, int n, int iters,
                                                               int bg_warps, int bg_kind, int gap, int a_shift_rows, int a_sbo_bytes)
{
    umma_rate_bg_body<1>(cycles, bg_bytes, n, iters, bg_warps, bg_kind, gap, a_shift_rows, a_sbo_bytes);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_rate_bg2_kernel(long long *cycles, unsigned long long *bg_bytes,
                                                                                          int n, int iters, int bg_warps, int bg_kind, int gap,
                                                                                          int a_shift_rows, int a_sbo_bytes)
{
    umma_rate_bg_body<2>(cycles, bg_bytes, n, iters, bg_warps, bg_kind, gap, a_shift_rows, a_sbo_bytes);
}

}  // namespace tc
}  // namespace tdrn

// cycles_dev: one entry per issuing CTA (grid, or grid / 2 with cta_group 2); bg_bytes_dev: one entry per CTA.
// a_shift_rows / a_sbo_bytes: the A descriptor starts that many 128-byte rows into the tile and steps a_sbo_bytes between
// 8-row groups (0 / 1024 = the aligned case; the halo kernels use shifts of r * 10 + s rows and 1280 or 2304 bytes).
extern "C" int tdrn_debug_umma_rate_bg(long long *cycles_dev, unsigned long long *bg_bytes_dev, int grid, int n, int iters, int cta_group,
                                       int bg_warps, int bg_kind, int gap, int a_shift_rows, int a_sbo_bytes)
{
    TDRN_REQUIRE(a_shift_rows >= 0 && a_sbo_bytes >= 1024 && a_sbo_bytes % 16 == 0 &&
                 a_shift_rows * 128 + 15 * a_sbo_bytes + 1024 <= 24 * 1024, "umma rate bg: A view leaves the 24 KB tile");
    TDRN_REQUIRE(cycles_dev && bg_bytes_dev && grid > 0 && n >= 16 && n <= 256 && n % 16 == 0 && (cta_group == 1 || cta_group == 2) &&
                 bg_warps >= 0 && bg_warps <= 3 && (bg_kind == 0 || bg_kind == 1) && gap >= 0, "umma rate bg: bad argument");
    const int smem = 64 * 1024 + 1024;
    if (cta_group == 2) {
        TDRN_REQUIRE(grid % 2 == 0, "umma rate bg: cta_group::2 needs an even grid");
        TDRN_CUDA(cudaFuncSetAttribute(tdrn::tc::umma_rate_bg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tdrn::tc::umma_rate_bg2_kernel<<<grid, 128, smem>>>(cycles_dev, bg_bytes_dev, n, iters, bg_warps, bg_kind, gap, a_shift_rows, a_sbo_bytes);
    } else {
        TDRN_CUDA(cudaFuncSetAttribute(tdrn::tc::umma_rate_bg1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tdrn::tc::umma_rate_bg1_kernel<<<grid, 128, smem>>>(cycles_dev, bg_bytes_dev, n, iters, bg_warps, bg_kind, gap, a_shift_rows, a_sbo_bytes);
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Fourth rate probe (development): the issue pattern of conv_halo_kernel without TMA and without an epilogue.
// Per "tile": nine taps x four K = 16 MMAs; tap (r, s) reads the A tile through the halo view (start (r * 10 + s) rows in,
// 8-row groups 1280 bytes apart) or, with `aligned`, through the nearest 1024-byte-aligned view with SBO 1024 (different
// data, same volume); every tap has its own n x 128-byte weight block, as the resident-weight kernel has.
// handshake = 1 adds the production kernel's TMEM double-buffer protocol (tcgen05.commit per tile to t_full, a consumer warp
// that waits for it and arrives on t_empty, the issuer waiting for t_empty before it re-uses an accumulator).
// Answers: does the pattern itself reach 48 / 64 cycles per MMA, and if not, is it the views or the protocol?
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
namespace tc {

__global__ void __launch_bounds__(128, 1) umma_rate_halo_kernel(long long *cycles, int n, int tiles, int aligned, int handshake)
{
    extern __shared__ uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t t_full[2], t_empty[2], done_bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t *base = (uint8_t *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
    const uint32_t w_bytes = 9u * (uint32_t)n * 128u;
    uint8_t *sW = base;
    uint8_t *sA = base + w_bytes;                       // two halo tiles, HL_A_STRIDE apart
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (uint32_t e = tid; e < (w_bytes + 2u * HL_A_STRIDE) / 4u; e += 128) ((uint32_t *)base)[e] = 0x3c003c00u;
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < 2; ++b) { mbar_init(&t_full[b], 1); mbar_init(&t_empty[b], 1); }
        mbar_init(&done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, n);
        const uint32_t sW_u = smem_u32(sW), sA_u = smem_u32(sA);
        const uint32_t kb_bytes = (uint32_t)n * 128u;
        const long long t0 = clock64();
        for (int tile = 0; tile < tiles; ++tile) {
            const uint32_t buf = (uint32_t)tile & 1u;
            if (handshake) { mbar_wait(&t_empty[buf], (((uint32_t)tile >> 1) & 1u) ^ 1u); tc_fence_after(); }
            const uint32_t d_tmem = tmem_base + buf * (uint32_t)n;
            const uint32_t a0 = sA_u + buf * (uint32_t)HL_A_STRIDE;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int tr = tap / 3, ts = tap - tr * 3;
                const uint32_t row = (uint32_t)(tr * HL_PW + ts);
                const uint64_t adesc = aligned ? umma_desc_sw128_sbo(a0 + (row & ~7u) * 128u, 1024u)
                                               : umma_desc_sw128_sbo(a0 + row * 128u, HL_PW * 128u);
                const uint64_t bdesc = umma_desc_sw128(sW_u + (uint32_t)tap * kb_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (tap | k) != 0);
            }
            if (handshake) umma_commit(&t_full[buf]);
        }
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    } else if (warp == 1 && handshake) {
        // stands in for the epilogue: wait for the accumulator, hand it back at once
        for (int tile = 0; tile < tiles; ++tile) {
            const uint32_t buf = (uint32_t)tile & 1u;
            mbar_wait(&t_full[buf], ((uint32_t)tile >> 1) & 1u);
            tc_fence_after();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[buf]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace tdrn

extern "C" int tdrn_debug_umma_rate_halo(long long *cycles_dev, int grid, int n, int tiles, int aligned, int handshake)
{
    TDRN_REQUIRE(cycles_dev && grid > 0 && (n == 64 || n == 128) && tiles > 0, "umma rate halo: bad argument");
    const int smem = 9 * n * 128 + 2 * tdrn::tc::HL_A_STRIDE + 1024;
    TDRN_CUDA(cudaFuncSetAttribute(tdrn::tc::umma_rate_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tdrn::tc::umma_rate_halo_kernel<<<grid, 128, smem>>>(cycles_dev, n, tiles, aligned ? 1 : 0, handshake ? 1 : 0);
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

extern "C" int tdrn_debug_umma_rate2(long long *cycles_dev, int grid, int n, int iters, int nacc, int mode)
{
    TDRN_REQUIRE(cycles_dev && grid > 0 && n >= 16 && n <= 256 && n % 16 == 0 && nacc >= 1 && nacc * n <= 448 && mode >= 0 && mode <= 6,
                 "umma rate2: bad argument");
    const int smem = 24 * 1024 + 36 * 1024 + 1024;
    if (mode == 4) {
        TDRN_REQUIRE(grid % 2 == 0, "umma rate2: cta_group::2 needs an even grid");
        TDRN_CUDA(cudaFuncSetAttribute(tdrn::tc::umma_rate_cg2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tdrn::tc::umma_rate_cg2_kernel<<<grid, 128, smem>>>(cycles_dev, n, iters, nacc);
    } else {
#define TDRN_RATE2(M)                                                                                                        \
    case M:                                                                                                                  \
        TDRN_CUDA(cudaFuncSetAttribute(tdrn::tc::umma_rate2_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        tdrn::tc::umma_rate2_kernel<M><<<grid, 128, smem>>>(cycles_dev, n, iters, nacc);                                  \
        break;
        switch (mode) { TDRN_RATE2(0) TDRN_RATE2(1) TDRN_RATE2(2) TDRN_RATE2(3) TDRN_RATE2(5) TDRN_RATE2(6) }
#undef TDRN_RATE2
    }
    TDRN_LAUNCH_CHECK();
    return TDRN_OK;
}

// ---------------------------------------------------------------------------------------------------------
// TMA fp32 box probe (development): load one (pw, ph, 3, 1) box of an NCHW fp32 image at (cx, cy) and copy it out.
// ---------------------------------------------------------------------------------------------------------
namespace tdrn {
namespace tc {
__global__ void tma_f32_probe_kernel(const __grid_constant__ CUtensorMap tm, float *out, int n, int cx, int cy, int b, int rank)
{
    __shared__ __align__(1024) float buf[4096];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
        mbar_expect_tx(&bar, (uint32_t)n * 4u);
        if (rank == 4) tma_load_4d(buf, &tm, &bar, cx, cy, 0, b);
        else asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                          ::"r"(smem_u32(buf)), "l"((uint64_t)&tm), "r"(smem_u32(&bar)), "r"(cx), "r"(cy), "r"(3 * b) : "memory");
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
}  // namespace tc
}  // namespace tdrn

extern "C" int tdrn_debug_tma_f32(const float *x, int B, int H, int W, int pw, int ph, int cx, int cy, int b, int rank, int l2promo, float *out)
{
    EncodeTiledFn enc = get_encode_tiled();
    TDRN_REQUIRE(enc, "no encoder");
    CUtensorMap tm;
    CUresult r;
    const CUtensorMapL2promotion promo = l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (l2promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (rank == 4) {
        const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
        const cuuint32_t bdim[4] = {(cuuint32_t)pw, (cuuint32_t)ph, 3, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)3 * B};
        const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
        const cuuint32_t bdim[3] = {(cuuint32_t)pw, (cuuint32_t)ph, 3};
        const cuuint32_t estr[3] = {1, 1, 1};
        r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(x), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    TDRN_REQUIRE(r == CUDA_SUCCESS, "encode failed %d", (int)r);
    tma_f32_probe_kernel<<<1, 128>>>(tm, out, pw * ph * 3, cx, cy, b, rank);
    TDRN_LAUNCH_CHECK();
    TDRN_CUDA(cudaDeviceSynchronize());
    return TDRN_OK;
}
