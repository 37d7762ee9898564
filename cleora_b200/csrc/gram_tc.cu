// K2b on the 5th-generation tensor cores: the centred Gram matrix  sum_r (x_r - mean)(x_r - mean)^T  computed
// EXACTLY in integers (pycleora/__init__.py:138-142 accumulates it in f64; the PCA basis amplifies covariance
// error by 1/eigengap, so floating-point tensor kinds with f32 accumulators are not admissible).
//
// Method (error-free byte-plane split; no floating-point rounding anywhere before the final f64 combination):
//   q[r,j]  = rint(x[r,j] * 2^e) - m[j]                  int32, |q| < 2^31;  e, m[j] = rint(mean[j] * 2^e) from
//                                                        quant_params_kernel (2^e <= 2^30 / (max|x| + max|mean|))
//   q       = b0 + 2^8 b1 + 2^16 b2 + 2^24 b3            b0..b2 unsigned bytes, b3 signed (two's complement planes)
//   G_s[i,j] = sum_{k+l=s} sum_r b_k[r,i] * b_l[r,j]     tcgen05.mma kind::i8, int32 accumulators in TMEM, one
//                                                        accumulator per power-of-two weight group s = 0..6
//   Q[i,j] = sum_r q_i q_j = sum_s 2^(8s) G_s[i,j]       exact; int64 global accumulators, drained every 6144 rows
//   result  = ( Q - S delta^T - delta S^T + n delta delta^T ) / 4^e     S = exact int64 column sums of q,
//                                                        delta_j = mean_j 2^e - m_j; combined in f64
// The only approximation is the input quantisation 2^-e (<= 2^-30 of max|x|, i.e. well below one f32 ulp of a
// typical element).  Integer accumulation is order-independent, so the result is bit-identical for any slicing,
// any atomic order and any GPU count.
//
// One CTA = a 128 x 64 output tile (row block i, column stripe j at or right of block i -- every G_s is symmetric, the
// combine kernel mirrors the rest) for a slice of rows: TMEM holds the 7 weight-group accumulators of 128 x 64 int32 =
// 448 of its 512 columns.  Per 32-row stage: 4 loader warps stage the rows (cp.async, 4 stages deep), 8 converter warps
// quantise them into four byte planes in a canonical no-swizzle UMMA layout (MN-major for d >= 256, K-major for d = 128)
// -- since round 2 only the 128 + 64 columns the tile reads, in a compact staging whose size does not depend on d, which
// is what admits d = 384 and 512 (template parameter NEEDED; the round-1 form stages all d columns) --, one lane issues
// the 16 tcgen05.mma (both operands are views of the planes: A = the 128 columns of block
// i in plane k, B = the stripe's 64 columns in plane l), and 4 drain warps move the accumulators to int64 global memory
// every 192 stages.
// Round-2 measurements (profiles/r2d_*): the shared-memory port bounds this kernel, not the tensor pipe (which issues a
// 128x64x32 MMA every 48 clk with both operands in shared memory, profiles/r2c_mma_probe.txt): per 32-row stage the
// cp.async ring writes 32 KB, the converters read it and write 32 KB of planes, and the 16 MMAs read 96 KB of operands
// -- ~1900 wavefronts of a 128 B/clk port against 773 clk of MMA.  Letting the converters read global memory directly
// (no ring) was tried and is slower (2.3 vs 1.8 ms at C2): the loads then occupy the same L1/shared-memory pipeline with
// worse sector efficiency.  The step that would help is the A operand in TMEM (tcgen05.cp once per plane, 32 clk per
// MMA): see DESIGN.md section 9.
#include "device.cuh"
#include "../../include/cleora_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <string>

namespace cleora {
namespace g8 {

constexpr int ROWS = 32;           // rows (K) per stage = one MMA k-step
constexpr int MAX_STAGES = 6;        // barrier slots; the actual raw / plane stage counts are launch parameters
constexpr int STRIPE = 64;         // output columns per CTA (UMMA N); the CTA owns ONE 128-row block i x this stripe
constexpr int GROUPS = 7;          // weight groups s = k + l
constexpr int DRAIN_STAGES = 192;  // 192*32 = 6144 rows: 4 * 255^2 * 6144 < 2^31
constexpr int CONV_THREADS = 256;  // 8 converter warps (16 were tried with the K-major converter: +7 % only)
constexpr int LOAD_THREADS = 128;  // 4 loader warps (cp.async, fully coalesced)
constexpr int THREADS = CONV_THREADS + 128 + 32 + LOAD_THREADS;   // + 4 drain warps + MMA warp + loaders
constexpr int PS = 2;              // plane stages; compile-time, like RAW_STAGES: a runtime modulo in the per-stage
                                   // bookkeeping cost every role ~10 % (measured 2.10 -> 1.88 ms)
constexpr int RAW_STAGES = 4;
constexpr int NGC = 12;            // compact staging (NEEDED): 16-column groups per tile = 8 of the row block + 4 of the stripe

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waiting must be cheap: a warp that polls flat out steals issue slots from the producer warps on its scheduler
// (measured: 6.7k warp-instructions per stage per SM, 80 % of them polling).  After the first failed probe the
// waiter backs off with nanosleep (`sleep_ns`: ~32 for pipeline hand-offs, ~1000 for the rare accumulator drains).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, uint32_t sleep_ns = 32) {
    uint32_t done, spins = 0;
    long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(sleep_ns);
        if ((++spins & 0xFFF) == 0) {                       // watchdog: a protocol bug must not hang the GPU
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000LL) __trap();
        }
    }
}
// 16-byte LDGSTS (cp.async): measured far faster than UBLKCP (cp.async.bulk) for this access pattern -- the 1-D bulk
// engine delivered only ~8 B/clk per SM here, which capped the kernel at 4.1 ms regardless of pipeline depth.
__device__ __forceinline__ void cp_async_cg16(void *dst, const void *src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier gets one (pre-counted) arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// One lane of a converged warp, chosen by the hardware.  `if (lane == 0)` makes ptxas treat the operands of the
// tcgen05.mma (descriptors, TMEM address: uniform-register operands) as possibly divergent and wrap EVERY MMA in an
// ELECT / R2UR.BROADCAST / BRA.U.ANY loop -- ~11 extra instructions and >100 clk per MMA, which was the real limit of
// both tensor-core kernels in round 1 (profiles/r2b_mma_probe_issue_bound.txt); under elect.sync it knows exactly one
// lane is active and moves the values to uniform registers directly.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "@p mov.u32 %0, 1;\n\t}"
        : "+r"(pred));
    return pred != 0;
}

// SWIZZLE_NONE descriptor; for MN-major operands LBO = byte distance between 8-row K groups, SBO = byte distance
// between 16-byte MN groups (cute make_umma_desc<Major::MN>, INTERLEAVE case).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// kind::i8 instruction descriptor: D = s32 (2), A/B format 0 = u8 / 1 = s8, both operands MN-major (bits 15/16 set)
// or both K-major, M = 128, N = STRIPE.
__device__ __forceinline__ uint32_t make_idesc_i8(int a_signed, int b_signed, bool mn_major) {
    return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) |
           ((uint32_t)(STRIPE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct QuantParams {
    float scale;        // 2^e
    int32_t e;
    int32_t pad[2];
};

}  // namespace g8

// max |x| over the matrix (per-block partial maxima, then one block) -- feeds the fixed-point scale.
__global__ void __launch_bounds__(256) absmax_stage1(const float *__restrict__ x, int64_t count, float *__restrict__ partial) {
    float m = 0.f;
    const float4 *xv = reinterpret_cast<const float4 *>(x);
    const int64_t n4 = count / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(xv + i);
        m = fmaxf(fmaxf(fmaxf(m, fabsf(v.x)), fmaxf(fabsf(v.y), fabsf(v.z))), fabsf(v.w));
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
        partial[blockIdx.x] = m;
    }
}

// scale = 2^e with (max|x| + max|mean|) * 2^e < 2^30;  m[j] = rint(mean[j] * 2^e);  zero the accumulators.
__global__ void quant_params_kernel(const float *__restrict__ absmax_partial, int nb, const double *__restrict__ mean,
                                    int d, g8::QuantParams *__restrict__ qp, int32_t *__restrict__ m_int) {
    __shared__ float s_scale;
    if (threadIdx.x == 0) {
        float mx = 0.f;
        for (int i = 0; i < nb; ++i) mx = fmaxf(mx, absmax_partial[i]);
        double mm = 0.0;
        for (int j = 0; j < d; ++j) mm = fmax(mm, fabs(mean[j]));
        const double bound = (double)mx + mm;
        int e = 30;
        if (bound > 0.0) {
            int ex;
            frexp(bound, &ex);                 // bound = f * 2^ex, f in [0.5, 1)  =>  bound < 2^ex
            e = 30 - ex;
        }
        e = max(-60, min(e, 60));
        qp->e = e;
        qp->scale = (float)ldexp(1.0, e);
        s_scale = qp->scale;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += blockDim.x) m_int[j] = (int32_t)llrint(mean[j] * (double)s_scale);
}

// KMAJOR selects the shared-memory layout of the byte planes (both are canonical no-swizzle UMMA layouts):
//   false: MN-major -- core matrix = 8 rows (K) x 16 columns (MN); a converter thread owns 16 columns of one row;
//   true : K-major  -- core matrix = 8 columns (MN) x 16 rows (K); a converter thread owns one column of 16 rows.
// Measured (n = 1M, d = 256 / n = 300k, d = 128): MN-major 1.88 / 0.30 ms, K-major 2.29 / 0.24 ms -- the launcher
// picks MN-major for d = 256 and K-major for d = 128 (where the MN-major mapping leaves half the converters idle).
template <bool KMAJOR, bool NEEDED>
__global__ void __launch_bounds__(g8::THREADS, 1)
gram_i8_kernel(const float *__restrict__ x, int64_t n, int d, const g8::QuantParams *__restrict__ qp,
               const int32_t *__restrict__ m_int, long long *__restrict__ G /* [7][d][d] */,
               long long *__restrict__ colsum /* [d] */, int64_t rows_per_slice) {
    constexpr bool needed_only = NEEDED;
    using namespace g8;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // NEEDED: only the columns this tile reads are loaded, converted and staged, in a COMPACT layout -- group slots 0..7 =
    // the 128 columns of row block mb, slots 8..11 = the 64 columns of stripe js when it lies outside the block (inside,
    // the B operand is a view of slots 0..7).  Shared memory then no longer grows with d (24 KB per plane stage and per
    // raw stage for every d), which is what admits d = 384 and 512.
    const int plane_bytes = NEEDED ? ROWS * NGC * 16 : ROWS * d;      // one byte plane of one stage
    const int stage_bytes = 4 * plane_bytes;
    // Raw rows are staged densely (row stride d*4, every cp.async piece 16-byte aligned inside its 128-byte line) with
    // an XOR swizzle: piece p of stage row r lives at piece p ^ (r & 7).  Eight converter lanes that read the same
    // piece index of eight consecutive rows then hit eight different 16-byte slots (conflict-free LDS.128), and a
    // loader quarter-warp still fills exactly one 128-byte line.  (Round 1 skewed the rows by 16 bytes instead, which
    // kept the LDS conflict-free but made most LDGSTS quarter-warps straddle two lines: 102 M wavefronts for 48 M ideal.)
    const int raw_stride = NEEDED ? NGC * 64 : d * 4;
    const int raw_bytes = ROWS * raw_stride;
    unsigned char *sP = smem_raw;                                     // [STAGES] byte planes (MMA operands)
    unsigned char *sR = sP + PS * stage_bytes;                        // [RAW_STAGES] raw f32 rows (cp.async ring)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sR + RAW_STAGES * raw_bytes);
    uint64_t *raw_full = bars;                   // [RAW_STAGES] count LOAD_THREADS (cp.async noinc arrivals)
    uint64_t *raw_empty = bars + MAX_STAGES;     // [RAW_STAGES] count CONV_THREADS
    uint64_t *full = bars + 2 * MAX_STAGES;      // [PS] count CONV_THREADS: planes ready
    uint64_t *empty = bars + 3 * MAX_STAGES;     // [PS] count 1 (tcgen05.commit): planes consumed
    uint64_t *acc_full = bars + 4 * MAX_STAGES;  // count 1
    uint64_t *acc_empty = acc_full + 1;          // count 128 (drain threads)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Tile (mb, js) = output row block i (128 columns of x) x output column stripe j (64 columns of x).  The Gram
    // matrix is symmetric group by group (the (k,l) sum is), so only stripes at or right of the diagonal block are
    // computed; the combine kernel mirrors the rest.  Tiles are numbered row block by row block.
    const int n_mb = d / 128, n_js = d / STRIPE;
    int mb = 0, js = 0;
    for (int t = blockIdx.x; mb < n_mb; ++mb) {
        const int first = mb * (128 / STRIPE), cnt = n_js - first;
        if (t < cnt) { js = first + t; break; }
        t -= cnt;
    }
    // needed_only (MN-major planes): a tile reads only the columns of its row block and of its stripe -- 128 columns when
    // the stripe lies inside the block, 192 otherwise -- so only those are loaded and converted (at d = 256: 42 % fewer
    // conversions, plane bytes and cp.async pieces over the six tiles of a row slice); the exact column sums are then
    // taken by the first tile of each block row for that block's columns.
    const bool owns_colsum = needed_only ? (js == mb * (128 / STRIPE)) : blockIdx.x == 0;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice, r1 = min(n, r0 + rows_per_slice);
    const int n_stages = (int)((r1 - r0 + ROWS - 1) / ROWS);
    const int n_cg = d / 16;                                          // 16-byte column groups per row
    constexpr int CONV_WARPS = CONV_THREADS / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < MAX_STAGES; ++s) {
            mbar_init(&raw_full[s], LOAD_THREADS); mbar_init(&raw_empty[s], CONV_THREADS);
            mbar_init(&full[s], CONV_THREADS); mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CONV_WARPS + 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < CONV_WARPS) {
      if constexpr (KMAJOR) {
        // ------------------------------------------------------------ converters, K-major planes
        // item = (column j, 16-row K chunk kc); thread t owns items t, t + 256, ...  (d = 256: column t, both chunks;
        // d = 128: column t % 128, chunk t / 128).  A warp reads 32 consecutive floats of one raw row (conflict-free
        // LDS.32) and stores 32 x 16 bytes = four adjacent core matrices per plane (conflict-free STS.128).
        // Plane byte offset of (j, r) = (j / 8) * 256 + (r / 16) * 128 + (j % 8) * 16 + r % 16.
        const float scale = qp->scale;
        const int total_items = d * (ROWS / 16);                        // (column, 16-row chunk) pairs per stage
        const int n_items = (total_items + CONV_THREADS - 1) / CONV_THREADS;
        long long csum = 0;                                             // all of a thread's items share one column
        const int j = threadIdx.x % d;
        const int mi = __ldg(m_int + j);
        for (int st = 0; st < n_stages; ++st) {
            const int rs = st % RAW_STAGES, s = st % PS;
            mbar_wait(&raw_full[rs], (st / RAW_STAGES) & 1);
            mbar_wait(&empty[s], ((st / PS) & 1) ^ 1);
            const unsigned char *raw = sR + rs * raw_bytes;
            unsigned char *base = sP + s * stage_bytes;
            const int64_t row0 = r0 + (int64_t)st * ROWS;
            for (int it = 0; it < n_items; ++it) {
                if ((int)threadIdx.x + it * CONV_THREADS >= total_items) break;   // spare threads only keep the barriers' counts
                const int kc = (threadIdx.x + it * CONV_THREADS) / d;
                int qv[16];
#pragma unroll
                for (int rr = 0; rr < 16; ++rr) {
                    const int r = kc * 16 + rr;
                    const float v = *reinterpret_cast<const float *>(raw + r * raw_stride + (((j >> 2) ^ (r & 7)) << 4) + (j & 3) * 4);
                    qv[rr] = (row0 + r < r1) ? __float2int_rn(v * scale) - mi : 0;
                }
                uint4 pl[4];
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) {
                    const uint32_t a = (uint32_t)qv[4 * wq], b2 = (uint32_t)qv[4 * wq + 1];
                    const uint32_t c2 = (uint32_t)qv[4 * wq + 2], d2 = (uint32_t)qv[4 * wq + 3];
                    const uint32_t t0 = __byte_perm(a, b2, 0x5140), t1 = __byte_perm(a, b2, 0x7362);
                    const uint32_t t2 = __byte_perm(c2, d2, 0x5140), t3 = __byte_perm(c2, d2, 0x7362);
                    (&pl[0].x)[wq] = __byte_perm(t0, t2, 0x5410);
                    (&pl[1].x)[wq] = __byte_perm(t0, t2, 0x7632);
                    (&pl[2].x)[wq] = __byte_perm(t1, t3, 0x5410);
                    (&pl[3].x)[wq] = __byte_perm(t1, t3, 0x7632);
                }
                const uint32_t off = (uint32_t)((j >> 3) * 256 + kc * 128 + (j & 7) * 16);
#pragma unroll
                for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4 *>(base + p * plane_bytes + off) = pl[p];
                if (owns_colsum) {
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) csum += qv[rr];
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
            mbar_arrive(&raw_empty[rs]);
        }
        if (owns_colsum && csum != 0) atomicAdd(reinterpret_cast<unsigned long long *>(colsum + j), (unsigned long long)csum);
      } else if constexpr (NEEDED) {
        // ------------------------------------------------------------ converters, MN-major planes, needed columns only
        // item = (group slot, stage row rr); items 0..255 = the 8 groups of row block mb (slots 0..7), items 256..383 = the
        // 4 groups of stripe js when it lies outside the block (slots 8..11).  Thread t owns item t and, in warps 0-3 of
        // such a tile, item 256 + t.  A quarter-warp = 8 consecutive rows of one group: conflict-free LDS.128 / STS.128.
        // cg_* = absolute 16-column group (centres, column sums), slot_* = its place in the compact staging.
        const int rr = threadIdx.x & 31, gi = threadIdx.x >> 5;               // gi = 0..7
        const int blk0 = mb * 8, str0 = js * (STRIPE / 16);
        const bool stripe_outside = str0 < blk0 || str0 >= blk0 + 8;
        const int cg_a = blk0 + gi, slot_a = gi;
        const bool has_b = stripe_outside && gi < STRIPE / 16;
        const int cg_b = has_b ? str0 + gi : cg_a, slot_b = has_b ? 8 + gi : gi;
        const float scale = qp->scale;
        int4 mia[4];                                     // (the second item's centres are re-read from L1 each stage: registers)
#pragma unroll
        for (int q = 0; q < 4; ++q) mia[q] = __ldg(reinterpret_cast<const int4 *>(m_int + cg_a * 16) + q);
        long long csum[16];                              // exact column sums of q over this thread's rows (block columns)
#pragma unroll
        for (int c = 0; c < 16; ++c) csum[c] = 0;
        for (int st = 0; st < n_stages; ++st) {
            const int rs = st % RAW_STAGES, s = st % PS;
            mbar_wait(&raw_full[rs], (st / RAW_STAGES) & 1);
            mbar_wait(&empty[s], ((st / PS) & 1) ^ 1);
            const unsigned char *raw = sR + rs * raw_bytes;
            unsigned char *base = sP + s * stage_bytes;
            const int64_t row = r0 + (int64_t)st * ROWS + rr;
            const unsigned char *xrow = raw + rr * raw_stride;
#pragma unroll 1
            for (int item = 0; item < 2; ++item) {
                if (item == 1 && !has_b) break;
                const int slot = item ? slot_b : slot_a;
                int qv[16];
                if (row < r1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4 *>(xrow + (((slot * 4 + q) ^ (rr & 7)) << 4));
                        const int4 m = item ? __ldg(reinterpret_cast<const int4 *>(m_int + cg_b * 16) + q) : mia[q];
                        qv[4 * q + 0] = __float2int_rn(v.x * scale) - m.x;
                        qv[4 * q + 1] = __float2int_rn(v.y * scale) - m.y;
                        qv[4 * q + 2] = __float2int_rn(v.z * scale) - m.z;
                        qv[4 * q + 3] = __float2int_rn(v.w * scale) - m.w;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) qv[c] = 0;
                }
                const uint32_t off = (uint32_t)((rr & 7) * 16 + (rr >> 3) * 128 + slot * (ROWS / 8) * 128);
                uint4 pl[4];
#pragma unroll
                for (int wq = 0; wq < 4; ++wq) {
                    const uint32_t a = (uint32_t)qv[4 * wq], b2 = (uint32_t)qv[4 * wq + 1];
                    const uint32_t c2 = (uint32_t)qv[4 * wq + 2], d2 = (uint32_t)qv[4 * wq + 3];
                    const uint32_t t0 = __byte_perm(a, b2, 0x5140), t1 = __byte_perm(a, b2, 0x7362);
                    const uint32_t t2 = __byte_perm(c2, d2, 0x5140), t3 = __byte_perm(c2, d2, 0x7362);
                    (&pl[0].x)[wq] = __byte_perm(t0, t2, 0x5410);
                    (&pl[1].x)[wq] = __byte_perm(t0, t2, 0x7632);
                    (&pl[2].x)[wq] = __byte_perm(t1, t3, 0x5410);
                    (&pl[3].x)[wq] = __byte_perm(t1, t3, 0x7632);
                }
#pragma unroll
                for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4 *>(base + p * plane_bytes + off) = pl[p];
                if (owns_colsum && item == 0) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) csum[c] += qv[c];
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
            mbar_arrive(&raw_empty[rs]);
        }
        if (owns_colsum) {                              // integer atomics: exact and order-independent
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (csum[c] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(colsum + cg_a * 16 + c), (unsigned long long)csum[c]);
        }
      } else {
        // ------------------------------------------------------------ converters: raw f32 (smem) -> 4 byte planes
        // thread -> (row lane rl = tid % 8, column group cg = (tid / 8) % 16, half h = tid / 128); it converts rows
        // rl + 8*i, i in {2h, 2h+1}.  A quarter-warp = 8 consecutive rows of one column group = one 128-byte core matrix
        // on the plane side (conflict-free STS.128); on the raw side the 16-byte row skew puts the 8 rows in 8 distinct
        // 16-byte bank slots (conflict-free LDS.128).
        const int rl = threadIdx.x & 7, cs = (threadIdx.x >> 3) & 15, half = threadIdx.x >> 7;
        const bool has_cg = cs < n_cg;
        const int cg = has_cg ? cs : 0;
        const float scale = qp->scale;
        int4 mi[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) mi[q] = __ldg(reinterpret_cast<const int4 *>(m_int + cg * 16) + q);
        long long csum[16];                              // exact column sums of q over this thread's rows
#pragma unroll
        for (int c = 0; c < 16; ++c) csum[c] = 0;
        for (int st = 0; st < n_stages; ++st) {
            const int rs = st % RAW_STAGES, s = st % PS;
            mbar_wait(&raw_full[rs], (st / RAW_STAGES) & 1);   // the loaders' cp.async for this stage have landed
            mbar_wait(&empty[s], ((st / PS) & 1) ^ 1);   // the MMAs that read these planes last time have retired
            const unsigned char *raw = sR + rs * raw_bytes;
            unsigned char *base = sP + s * stage_bytes;
            if (has_cg) {
#pragma unroll
                for (int ii = 0; ii < 2; ++ii) {
                    const int rr = rl + 8 * (2 * half + ii);
                    const int64_t row = r0 + (int64_t)st * ROWS + rr;
                    int qv[16];
                    if (row < r1) {
                        const unsigned char *xrow = raw + rr * raw_stride;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 v = *reinterpret_cast<const float4 *>(xrow + (((cg * 4 + q) ^ (rr & 7)) << 4));
                            qv[4 * q + 0] = __float2int_rn(v.x * scale) - mi[q].x;
                            qv[4 * q + 1] = __float2int_rn(v.y * scale) - mi[q].y;
                            qv[4 * q + 2] = __float2int_rn(v.z * scale) - mi[q].z;
                            qv[4 * q + 3] = __float2int_rn(v.w * scale) - mi[q].w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 16; ++c) qv[c] = 0;
                    }
                    const uint32_t off = (uint32_t)((rr & 7) * 16 + (rr >> 3) * 128 + cg * (ROWS / 8) * 128);
                    // 4x4 byte transposes: word w of plane p = byte p of qv[4w .. 4w+3]  (8 PRMT per 4 values)
                    uint4 pl[4];
#pragma unroll
                    for (int wq = 0; wq < 4; ++wq) {
                        const uint32_t a = (uint32_t)qv[4 * wq], b2 = (uint32_t)qv[4 * wq + 1];
                        const uint32_t c2 = (uint32_t)qv[4 * wq + 2], d2 = (uint32_t)qv[4 * wq + 3];
                        const uint32_t t0 = __byte_perm(a, b2, 0x5140), t1 = __byte_perm(a, b2, 0x7362);   // a0 b0 a1 b1 | a2 b2 a3 b3
                        const uint32_t t2 = __byte_perm(c2, d2, 0x5140), t3 = __byte_perm(c2, d2, 0x7362);
                        (&pl[0].x)[wq] = __byte_perm(t0, t2, 0x5410);      // a0 b0 c0 d0
                        (&pl[1].x)[wq] = __byte_perm(t0, t2, 0x7632);      // a1 b1 c1 d1
                        (&pl[2].x)[wq] = __byte_perm(t1, t3, 0x5410);
                        (&pl[3].x)[wq] = __byte_perm(t1, t3, 0x7632);
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p) *reinterpret_cast<uint4 *>(base + p * plane_bytes + off) = pl[p];
                    if (owns_colsum) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) csum[c] += qv[c];
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
            mbar_arrive(&raw_empty[rs]);
        }
        if (owns_colsum && has_cg) {                    // integer atomics: exact and order-independent
#pragma unroll
            for (int c = 0; c < 16; ++c)
                if (csum[c] != 0) atomicAdd(reinterpret_cast<unsigned long long *>(colsum + cg * 16 + c), (unsigned long long)csum[c]);
        }
      }
    } else if (warp < CONV_WARPS + 4) {
        // ------------------------------------------------------------ drain: TMEM int32 -> global int64 (atomic)
        const int q4 = warp - CONV_WARPS;
        uint32_t v[32];
        int drains = 0;
        for (int st0 = 0; st0 < n_stages; st0 += DRAIN_STAGES, ++drains) {
            mbar_wait(acc_full, drains & 1, 2000);
            tc_fence_after();
            for (int s = 0; s < GROUPS; ++s)
                for (int h = 0; h < STRIPE / 32; ++h) {
                    const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * STRIPE + h * 32);
                    tmem_ld32(taddr, v);
                    const int i = mb * 128 + q4 * 32 + lane;
                    long long *dst = G + ((int64_t)s * d + i) * d + js * STRIPE + h * 32;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const long long val = (long long)(int32_t)v[c];
                        if (val != 0) atomicAdd(reinterpret_cast<unsigned long long *>(dst + c), (unsigned long long)val);
                    }
                }
            tc_fence_before();
            mbar_arrive(acc_empty);
        }
    } else if (warp == CONV_WARPS + 4) {
        // ------------------------------------------------------------ MMA issuer
        int drains = 0;
        // MN-major: LBO = step between 8-row K groups (128 B), SBO = step between 16-column MN groups (512 B);
        // K-major: LBO = step between 16-row K chunks (128 B), SBO = step between 8-column MN groups (256 B).
        const uint32_t lbo = 128, sbo = KMAJOR ? (ROWS / 16) * 128 : (ROWS / 8) * 128;
        uint32_t a_off = KMAJOR ? (uint32_t)(mb * 16) * sbo : (uint32_t)(mb * 8) * sbo;              // first MN group of block i
        uint32_t b_off = KMAJOR ? (uint32_t)(js * (STRIPE / 8)) * sbo : (uint32_t)(js * (STRIPE / 16)) * sbo;
        if (NEEDED) {                                                 // compact staging: block at slot 0, stripe inside it or at slot 8
            const int blk0 = mb * 8, str0 = js * (STRIPE / 16);
            a_off = 0;
            b_off = (uint32_t)((str0 >= blk0 && str0 < blk0 + 8) ? str0 - blk0 : 8) * sbo;
        }
        for (int st = 0; st < n_stages; ++st) {
            const int s = st % PS;
            const bool first = (st % DRAIN_STAGES) == 0;              // first stage after a drain: overwrite
            if (first && st > 0) {                                    // wait until the previous accumulators are drained
                mbar_wait(acc_empty, (drains - 1) & 1, 200);
                tc_fence_after();
            }
            mbar_wait(&full[s], (st / PS) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t pbase = smem_u32(sP + s * stage_bytes);
                uint32_t used = 0;                                    // bit g: accumulator already written this stage
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = make_desc(pbase + k * plane_bytes + a_off, lbo, sbo);
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const uint64_t bd = make_desc(pbase + l * plane_bytes + b_off, lbo, sbo);
                        const int g = k + l;
                        const uint32_t acc = (first && !(used & (1u << g))) ? 0u : 1u;
                        used |= 1u << g;
                        mma_i8(tmem_base + (uint32_t)(g * STRIPE), ad, bd, make_idesc_i8(k == 3, l == 3, !KMAJOR), acc);
                    }
                }
                mma_commit(&empty[s]);
                if (((st + 1) % DRAIN_STAGES) == 0 || st + 1 == n_stages) mma_commit(acc_full);
            }
            __syncwarp();
            if (((st + 1) % DRAIN_STAGES) == 0 || st + 1 == n_stages) ++drains;
        }
    } else if (warp >= CONV_WARPS + 5) {
        // ------------------------------------------------------------ loaders: coalesced 16-byte cp.async, 4 stages deep.
        // (Separate warps on purpose: the converters' fence.proxy.async would otherwise wait for their own in-flight
        //  prefetches and serialise the ring.)
        // thread -> fixed 16-byte column piece pc and a fixed row phase; it walks down the stage's rows with constant
        // strides (no per-piece division: the first version spent ~80 instructions per piece on index arithmetic)
        const int lt = threadIdx.x - (CONV_WARPS + 5) * 32;            // 0..127
        if constexpr (NEEDED) {
            // compact rows: 32 (stripe inside the block) or 48 pieces of 16 bytes = the block's 512 contiguous bytes and,
            // behind them, the stripe's 256.  Piece i of the stage (row-major, i = rr * ppr + pc) goes to thread i % 128;
            // (rr, pc) advance incrementally from item to item (no division in the loop).
            const int blk0 = mb * 8, str0 = js * (STRIPE / 16);
            const bool outside = str0 < blk0 || str0 >= blk0 + 8;
            const int ppr = outside ? NGC * 4 : 32;
            const int n_items = ROWS * ppr / LOAD_THREADS;              // 8 or 12 per thread and stage
            const int step_rr = LOAD_THREADS / ppr, step_pc = LOAD_THREADS % ppr;
            // every item's (stage offset, source offset, row) is the same in every stage: computed once, kept in
            // registers (a first version did this arithmetic per piece and was slower than loading whole rows)
            constexpr int MAX_ITEMS = ROWS * NGC * 4 / LOAD_THREADS;    // 12
            int dst_pk[MAX_ITEMS], src_off[MAX_ITEMS];                  // dst_pk = byte offset in the stage | row << 16
            {
                int rr = lt / ppr, pc = lt % ppr;
#pragma unroll
                for (int it = 0; it < MAX_ITEMS; ++it) {
                    const int slot = pc >> 2;
                    const int col = (slot < 8 ? mb * 128 + slot * 16 : js * STRIPE + (slot - 8) * 16) + (pc & 3) * 4;
                    dst_pk[it] = (rr * raw_stride + ((pc ^ (rr & 7)) << 4)) | (rr << 16);
                    src_off[it] = rr * d + col;
                    rr += step_rr; pc += step_pc;
                    if (pc >= ppr) { pc -= ppr; ++rr; }
                }
            }
            for (int st = 0; st < n_stages; ++st) {
                const int rs = st % RAW_STAGES;
                mbar_wait(&raw_empty[rs], ((st / RAW_STAGES) & 1) ^ 1);
                const int64_t row0 = r0 + (int64_t)st * ROWS;
                const int rows_left = (int)min(r1 - row0, (int64_t)ROWS);   // >= 1
                unsigned char *stage = sR + rs * raw_bytes;
                const float *xs = x + row0 * (int64_t)d;
#pragma unroll
                for (int it = 0; it < MAX_ITEMS; ++it) {
                    if (it < n_items) {                                  // (uniform: 8 items when the stripe is inside the block)
                        const bool in = (dst_pk[it] >> 16) < rows_left;
                        cp_async_cg16(stage + (dst_pk[it] & 0xFFFF), in ? xs + src_off[it] : x, in ? 16 : 0);
                    }
                }
                cp_async_arrive_noinc(&raw_full[rs]);
            }
            cp_async_wait_all();
        } else {
        const int ppr = d / 4;                                          // 16-byte pieces per row: 32 (d=128) or 64 (d=256)
        const int pc = lt % ppr, rr0 = lt / ppr, rstep = LOAD_THREADS / ppr;   // rows rr0, rr0+rstep, ...
        for (int st = 0; st < n_stages; ++st) {
            const int rs = st % RAW_STAGES;
            mbar_wait(&raw_empty[rs], ((st / RAW_STAGES) & 1) ^ 1);
            const int64_t row0 = r0 + (int64_t)st * ROWS;
            unsigned char *dst = sR + rs * raw_bytes + rr0 * raw_stride;
            const float *src = x + (row0 + rr0) * (int64_t)d + pc * 4;
            const int64_t rows_left = r1 - row0;                        // >= 1
#pragma unroll 4
            for (int rr = rr0; rr < ROWS; rr += rstep) {
                const bool in = rr < rows_left;
                cp_async_cg16(dst + ((pc ^ (rr & 7)) << 4), in ? src : x, in ? 16 : 0);
                dst += rstep * raw_stride;
                src += (int64_t)rstep * d;
            }
            cp_async_arrive_noinc(&raw_full[rs]);
        }
        cp_async_wait_all();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == CONV_WARPS + 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// With Q = sum_s 2^(8s) G_s (exact sum_r q_i q_j), S = column sums of q (exact) and delta_j = mean_j 2^e - m_j (the
// part of the requested centre the integer centring did not remove, |delta| <~ 1):
//   sum_r (x_i - mean_i)(x_j - mean_j) = ( Q_ij - S_i delta_j - delta_i S_j + n delta_i delta_j ) * 4^-e
// `mean` may be any centre (the global mean when the rows are one rank's shard).
__global__ void gram_i8_combine_kernel(const long long *__restrict__ G, const long long *__restrict__ colsum,
                                       const int32_t *__restrict__ m_int, const double *__restrict__ mean, int d,
                                       int64_t n, const g8::QuantParams *__restrict__ qp, double *__restrict__ cov) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)d * d) return;
    const int i = (int)(idx / d), j = (int)(idx - (int64_t)i * d);
    // tiles left of the diagonal block were not computed: G_s[i][j] = G_s[j][i]
    const int64_t src = (j / 128 < i / 128) ? (int64_t)j * d + i : idx;
    double acc = 0.0;
    for (int s = 0; s < g8::GROUPS; ++s) acc += ldexp((double)G[(int64_t)s * d * d + src], 8 * s);   // small to large
    const double di = ldexp(mean[i], qp->e) - (double)m_int[i], dj = ldexp(mean[j], qp->e) - (double)m_int[j];
    acc += -(double)colsum[i] * dj - di * (double)colsum[j] + (double)n * di * dj;
    cov[idx] = ldexp(acc, -2 * qp->e);
}

std::atomic<int> g_gram_needed_only{[] {
    const char *e = getenv("CLEORA_B200_GRAM_COLS");
    return (e && std::string(e) == "all") ? 0 : 1;
}()};

// d = 128, 256: both stagings; d = 384, 512: only with the compact staging of the needed columns (option gram_needed_cols)
bool gram_i8_supported(int64_t n, int64_t d) {
    if (n < 4096) return false;
    if (d == 128 || d == 256) return true;
    return (d == 384 || d == 512) && g_gram_needed_only.load() != 0;
}

void launch_centered_gram_i8(const float *x, int64_t n, int64_t d, const double *mean, double *cov, cudaStream_t st,
                             const AbsmaxPartials *known_absmax) {
    using namespace g8;
    // scratch: absmax partials | QuantParams | m_int[d] | colsum[d] | G[7][d][d]
    int nb = 148 * 4;
    const size_t off_qp = 1024 * sizeof(float);
    const size_t off_m = off_qp + sizeof(QuantParams);
    const size_t off_cs = off_m + sizeof(int32_t) * 512;
    const size_t off_G = off_cs + sizeof(long long) * 512;
    const size_t total = off_G + sizeof(long long) * GROUPS * d * d;
    unsigned char *ws = (unsigned char *)workspace().gram_partials.get(total);
    const float *absmax = (const float *)ws;
    QuantParams *qp = (QuantParams *)(ws + off_qp);
    int32_t *m_int = (int32_t *)(ws + off_m);
    long long *colsum = (long long *)(ws + off_cs);
    long long *G = (long long *)(ws + off_G);
    CUDA_TRY(cudaMemsetAsync(ws + off_cs, 0, total - off_cs, st));
    if (known_absmax && known_absmax->count > 0) {                   // the column-sum pass over the same matrix made them
        absmax = known_absmax->p;
        nb = known_absmax->count;
    } else {
        absmax_stage1<<<nb, 256, 0, st>>>(x, n * d, (float *)ws);
        LAUNCH_CHECK();
    }
    quant_params_kernel<<<1, 256, 0, st>>>(absmax, nb, mean, (int)d, qp, m_int);
    LAUNCH_CHECK();
    int stripes = 0;                                                  // tiles: (128-row block i) x (64-column stripe j >= block i)
    for (int mb = 0; mb < (int)(d / 128); ++mb) stripes += (int)(d / STRIPE) - mb * (128 / STRIPE);
    int64_t slices = std::max<int64_t>(1, std::min<int64_t>(148 / stripes, (n + 4 * ROWS - 1) / (4 * ROWS)));   // one wave
    const int64_t rows_per_slice = ((n + slices - 1) / slices + ROWS - 1) / ROWS * ROWS;
    slices = (n + rows_per_slice - 1) / rows_per_slice;
    // plane layout: CLEORA_B200_GRAM_LAYOUT=mn|k overrides the per-shape default (see the kernel's comment)
    static const int forced = [] {
        const char *e = getenv("CLEORA_B200_GRAM_LAYOUT");
        const std::string v = e ? e : "";
        return v == "mn" ? 0 : v == "k" ? 1 : -1;
    }();
    const bool kmajor = forced >= 0 ? forced == 1 : d < 256;
    // convert only the columns a tile reads (MN-major path, two row blocks): option "gram_needed_cols"
    const bool needed_only = !kmajor && d >= 256 && g_gram_needed_only.load() != 0;
    if (d > 256 && !needed_only) throw CudaFail{"integer Gram: d > 256 needs the compact staging (gram_needed_cols = 1, MN-major planes)"};
    const size_t smem = (needed_only ? (size_t)(PS + RAW_STAGES) * 4 * ROWS * NGC * 16
                                     : (size_t)PS * 4 * ROWS * d + (size_t)RAW_STAGES * ROWS * (d * 4)) +
                        (4 * MAX_STAGES + 4) * sizeof(uint64_t) + 16;
    auto kernel = kmajor ? gram_i8_kernel<true, false> : (needed_only ? gram_i8_kernel<false, true> : gram_i8_kernel<false, false>);
    const int threads = THREADS;
    // per device, not per process: set on every launch (a host-side table write)
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid((unsigned)stripes, (unsigned)slices);
    kernel<<<grid, threads, smem, st>>>(x, n, (int)d, qp, m_int, G, colsum, rows_per_slice);
    LAUNCH_CHECK();
    gram_i8_combine_kernel<<<(unsigned)((d * d + 255) / 256), 256, 0, st>>>(G, colsum, m_int, mean, (int)d, n, qp, cov);
    LAUNCH_CHECK();
}

}  // namespace cleora
