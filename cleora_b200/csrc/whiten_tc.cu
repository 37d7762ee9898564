// K3 on the 5th-generation tensor cores:  out[n, N] = rownorm?( (x - mean_f32) @ T )   (pycleora/__init__.py:157-163)
//
// tcgen05.mma kind::tf32 with the accumulator in TMEM, "3xTF32" error compensation so the result keeps fp32-class
// accuracy (each f32 operand is split a = a_hi + a_lo with a_hi = the top 19 bits; D = A_lo*B_hi + A_hi*B_lo +
// A_hi*B_hi accumulated in fp32): the reference multiplies in f32 (np.dot), plain TF32 (10-bit mantissa) would
// miss its 1e-5 bar.
//
// Shape: one CTA = 128 rows of x (UMMA M = 128), all N output columns (one UMMA N = N <= 256 accumulator; 256 < N <= 512
// runs the two column halves of a row tile as two passes into the two TMEM buffers), K = d in chunks of 32.
// Warp roles (352 threads, persistent over row tiles):
//   warps 0-3  A producers: centre, split hi/lo, 16-byte stores into the A tiles.  Default (ASW = 1, round 2): eight lanes
//              own the eight 16-byte chunks of one 128-byte row of the stage -- coalesced 128-bit loads, 4 whole lines per
//              warp instruction -- and the tiles are in the SWIZZLE_128B K-major layout (chunk c of row r at chunk
//              c ^ (r % 8): conflict-free STS.128).  ASW = 0 (round 1): thread = row, canonical K-major no-swizzle layout
//              (8-row x 16-byte core matrices, LBO 128 B, SBO 1024 B) -- every warp load touched 32 different lines;
//   warps 4-7  epilogue: tcgen05.ld of the warp's 32 TMEM lanes (thread = row), optional row L2 norm (one division per
//              row, a multiplication per element), then every 32 x 32
//              block goes through a padded shared-memory tile so that rows leave as 128-byte pieces, four rows per
//              store instruction -- coalesced for HBM and, in the fused-collective modes (PeerOut), for NVLink: the
//              same rows into the peers' copies (all-gather) or each 32-column slice into its owner's column-sharded
//              matrix (all-to-all), 512 contiguous bytes per instruction;
//   warp  8    TMEM allocation + MMA issuer (one elected lane): 3 x 4 tcgen05.mma per K chunk, tcgen05.commit to free
//              smem stages and to publish the accumulator;
//   warps 9-10 B loaders: the transform, pre-split and pre-tiled in global memory by prep_transform_kernel, is copied
//              chunk by chunk with 16-byte cp.async (LDGSTS) arriving on an mbarrier (cp.async.bulk was tried first
//              and sustained only ~15 B/clk per SM here).
// TMEM holds two accumulator buffers (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "device.cuh"
#include "../../include/cleora_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <type_traits>

namespace cleora {

namespace tc {

constexpr int BM = 128;            // rows per tile (UMMA M)
// K per stage (floats) and pipeline depth are template parameters of the kernel: (32, 2) = round 1's shape, (16, 4) = the
// same bytes in flight cut into twice as many stages.  With two stages the refill of a stage (L2 latency + 64 KB of
// transform) can only start when the MMAs that read it retire and must finish within ONE stage time; with four half-size
// stages it has three stage times.
constexpr int BK_MAX = 32;
constexpr int NMAX = 256;          // UMMA N limit
constexpr int B_LOAD_THREADS = 64;  // 2 loader warps for the transform
constexpr int THREADS = 288 + B_LOAD_THREADS;
constexpr int EPI_LD = 36;          // floats per row of an epilogue staging tile (32 + 4: conflict-free 128-bit rows)
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;   // one 32 x 32 tile per epilogue warp

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
// 16-byte LDGSTS for the B operand: the 1-D bulk engine (cp.async.bulk / UBLKCP) sustained only ~15 B/clk per SM on
// these 64 KB chunks and was the stage-time limiter (measured); LDGSTS from two warps is several times faster.
__device__ __forceinline__ void cp_async_cg16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
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

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell): start address, leading (K-direction) and
// stride (M/N-direction 8-row group) byte offsets, all in 16-byte units.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::tf32: D = f32, A = B = tf32, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
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

__device__ __forceinline__ float tf32_hi(float a) { return __uint_as_float(__float_as_uint(a) & 0xFFFFE000u); }

}  // namespace tc

// Pre-split and pre-tile the transform for the tensor-core kernel.  For every K chunk c (32 rows of T) the image
// is exactly what the UMMA descriptor expects for a K-major B operand [N x BK]: element (n, k) of the chunk at
// byte (n%8)*16 + (k%4)*4 + (k/4)*128 + (n/8)*(BK*32).  Bt = [chunk][hi|lo][N*BK floats].
__global__ void prep_transform_kernel(const float *__restrict__ T, int d, int dout, int BK, float *__restrict__ Bt) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)d * dout) return;
    const int k = (int)(idx / dout), n = (int)(idx - (int64_t)k * dout);
    const int c = k / BK, kk = k % BK;
    const float v = T[idx];
    const float hi = tc::tf32_hi(v);
    const float lo = v - hi;
    const int64_t chunk_floats = (int64_t)dout * BK;
    const int64_t off = (int64_t)(n % 8) * 4 + (kk % 4) + (int64_t)(kk / 4) * 32 + (int64_t)(n / 8) * (BK * 8);
    float *base = Bt + (int64_t)c * 2 * chunk_floats;
    base[off] = hi;
    base[chunk_floats + off] = lo;
}

// NORM: 0 none, 2 row L2 (x / max(norm, 1e-10)) fused into the epilogue.
// SCALED: 0 -> a = x - mean;  1 -> a = x - rowscale[r] * mean  (x = A*Y with A not yet applied to the centring:
//         A (Y - 1 mean^T) = A Y - (A 1) mean^T, rowscale = A 1; used by the pipelined loop, see abi.cu).
// ASW: 0 -> A tiles in the no-swizzle canonical layout, produced by one thread per row (each warp load instruction
//           touches 32 different 128-byte lines);
//      1 -> A tiles in the SWIZZLE_128B K-major layout (row r of a stage = 128 contiguous bytes, 16-byte chunk c stored at
//           chunk c ^ (r % 8); what a 2-D TMA load with CU_TENSOR_MAP_SWIZZLE_128B would deposit): eight lanes own the
//           eight chunks of one row, so a warp load instruction covers 4 whole lines (coalesced) and a quarter-warp's
//           STS.128 fills the 8 distinct chunk slots of its row (conflict-free).  BK = 32 only (one swizzle atom).
template <int NORM, int SCALED, int BK, int STAGES, int ASW>
__global__ void __launch_bounds__(tc::THREADS, 1)
whiten_apply_tc_kernel(const float *__restrict__ x, int64_t n, int d, const float *__restrict__ mean,
                       const float *__restrict__ rowscale, const float *__restrict__ Bt, int NT, int upper,
                       float *__restrict__ out, PeerOut peers) {
    using namespace tc;
    constexpr int A_BYTES = BM * BK * 4;                             // one of hi / lo: 16 KB (BK = 32) or 8 KB (BK = 16)
    constexpr int SBO = BK * 32;                                     // bytes between 8-row groups of an operand tile
    constexpr int Q = BK / 4;                                        // float4 per row and stage
    static_assert(!ASW || BK == 32, "the swizzled A layout is one 128-byte atom per row");
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    // 1024-byte alignment: the SWIZZLE_128B pattern is a function of the shared-memory address bits [4, 10)
    unsigned char *smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    const int halves = NT > NMAX ? 2 : 1;                            // column halves of a row tile (two accumulator passes)
    const int N = NT / halves;                                       // UMMA N
    const int b_bytes = N * BK * 4;                                  // one of hi / lo, one half
    unsigned char *sA = smem_raw;                                    // STAGES x (A_hi, A_lo)
    unsigned char *sB = sA + STAGES * 2 * A_BYTES;                   // STAGES x (B_hi, B_lo)
    float *sE = reinterpret_cast<float *>(sB + STAGES * 2 * b_bytes);   // epilogue staging tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(sE) + EPI_BYTES);
    uint64_t *full_a = bars;                 // [STAGES] count 128
    uint64_t *full_b = bars + STAGES;        // [STAGES] count 1 + tx
    uint64_t *empty = bars + 2 * STAGES;     // [STAGES] count 1 (tcgen05.commit)
    uint64_t *acc_full = bars + 3 * STAGES;  // [2] count 1 (tcgen05.commit)
    uint64_t *acc_empty = acc_full + 2;      // [2] count 128
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = d / BK;
    const int64_t n_tiles = (n + BM - 1) / BM;
    // upper != 0: T is upper triangular (the Cholesky whitening transform): rows [32c, 32c+32) of T are zero left of
    // column 32c, so K chunk c only touches the accumulator columns from n0(h, c) on -- the MMAs shrink to N - n0 columns
    // (44 % fewer tensor clocks and operand bytes at d = 256) and steps with n0 >= N are skipped by every role.
    auto first_col = [&](int h, int c) { return upper ? min(max(c * BK - h * N, 0), N) : 0; };
    int steps_per_tile = 0;
    for (int h = 0; h < halves; ++h)
        for (int c = 0; c < n_chunks; ++c) steps_per_tile += first_col(h, c) < N ? 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_a[s], 128); mbar_init(&full_b[s], B_LOAD_THREADS); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {   // TMEM: 512 columns = two 256-column accumulator buffers
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4 && ASW) {
        // ------------------------------------------------------------------ A producers, swizzled tiles (8 lanes = one row)
        // lane -> (rsub = lane / 8, chunk q = lane % 8); the thread owns chunk q of rows 32*warp + 4*j + rsub, j = 0..7.
        const int rsub = lane >> 3, q = lane & 7;
        const int per_tile = halves * n_chunks;
        const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t total = my_tiles * steps_per_tile;
        auto advance = [&](int64_t &t, int &p) {       // next executed (half, chunk) step
            do {
                if (++p == per_tile) { p = 0; t += gridDim.x; }
            } while (first_col(p / n_chunks, p % n_chunks) >= N);
        };
        float4 cur[8], nxt[8];
        auto fetch = [&](int64_t tile, int c, float4 (&v)[8]) {
            const int64_t row0 = tile * BM + warp * 32 + rsub;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t row = row0 + 4 * j;
                v[j] = row < n ? __ldg(reinterpret_cast<const float4 *>(x + row * (int64_t)d + c * BK) + q)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        int64_t tile = blockIdx.x;
        int pos = 0;
        if (total > 0) fetch(tile, 0, cur);
        float rs[8];
        int64_t rs_tile = -1;
        uint32_t it = 0;
        for (int64_t i = 0; i < total; ++i, ++it) {
            const int c = pos % n_chunks;
            int64_t ntile = tile;
            int npos = pos;
            if (i + 1 < total) { advance(ntile, npos); fetch(ntile, npos % n_chunks, nxt); }
            const int64_t row0 = tile * BM + warp * 32 + rsub;
            if (SCALED && rs_tile != tile) {           // A*1 of this thread's 8 rows, once per tile
#pragma unroll
                for (int j = 0; j < 8; ++j) rs[j] = row0 + 4 * j < n ? __ldg(rowscale + row0 + 4 * j) : 0.f;
                rs_tile = tile;
            }
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const float4 m0 = __ldg(reinterpret_cast<const float4 *>(mean + c * BK) + q);
            mbar_wait(&empty[s], ph ^ 1);              // stage free (first round passes immediately)
            unsigned char *hi = sA + s * 2 * A_BYTES, *lo = hi + A_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = warp * 32 + 4 * j + rsub;                     // row of the tile
                const bool in = row0 + 4 * j < n;
                float4 m = m0;
                if (SCALED) { m.x = __fmul_rn(rs[j], m.x); m.y = __fmul_rn(rs[j], m.y); m.z = __fmul_rn(rs[j], m.z); m.w = __fmul_rn(rs[j], m.w); }
                float4 a, hh, l;
                a.x = in ? __fsub_rn(cur[j].x, m.x) : 0.f; a.y = in ? __fsub_rn(cur[j].y, m.y) : 0.f;
                a.z = in ? __fsub_rn(cur[j].z, m.z) : 0.f; a.w = in ? __fsub_rn(cur[j].w, m.w) : 0.f;
                hh.x = tf32_hi(a.x); hh.y = tf32_hi(a.y); hh.z = tf32_hi(a.z); hh.w = tf32_hi(a.w);
                l.x = a.x - hh.x; l.y = a.y - hh.y; l.z = a.z - hh.z; l.w = a.w - hh.w;
                const int off = (r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) << 4);
                *reinterpret_cast<float4 *>(hi + off) = hh;
                *reinterpret_cast<float4 *>(lo + off) = l;
            }
            fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&full_a[s]);
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = nxt[j];
            tile = ntile;
            pos = npos;
        }
    } else if (warp < 4) {
        // ------------------------------------------------------------------ A producers (thread = row of the tile)
        // The global loads of step i+1 are issued before step i is converted: a producer that loads, waits and
        // converts one chunk at a time spends most of each step in the load latency (ncu, round 2: the first use of the
        // loaded data was the kernel's top stall and the tensor pipe idled 64 % of the time).
        const int r = threadIdx.x;                     // 0..127
        const int per_tile = halves * n_chunks;
        const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        const int64_t total = my_tiles * steps_per_tile;
        auto advance = [&](int64_t &t, int &p) {       // next executed (half, chunk) step
            do {
                if (++p == per_tile) { p = 0; t += gridDim.x; }
            } while (first_col(p / n_chunks, p % n_chunks) >= N);
        };
        float4 cur[Q], nxt[Q];
        auto fetch = [&](int64_t tile, int c, float4 (&v)[Q]) {
            const int64_t row = tile * BM + r;
            const bool in = row < n;
            const float4 *xr = reinterpret_cast<const float4 *>(x + (in ? row : 0) * (int64_t)d);
#pragma unroll
            for (int q = 0; q < Q; ++q) v[q] = in ? __ldg(xr + c * Q + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        int64_t tile = blockIdx.x;
        int pos = 0;                                    // position inside the tile's (half, chunk) sequence
        if (total > 0) fetch(tile, 0, cur);
        uint32_t it = 0;
        for (int64_t i = 0; i < total; ++i, ++it) {
            const int c = pos % n_chunks;
            int64_t ntile = tile;
            int npos = pos;
            if (i + 1 < total) { advance(ntile, npos); fetch(ntile, npos % n_chunks, nxt); }
            const int64_t row = tile * BM + r;
            const bool in = row < n;
            const float rs = (SCALED && in) ? __ldg(rowscale + row) : 1.f;
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const float4 *mp = reinterpret_cast<const float4 *>(mean + c * BK);
            mbar_wait(&empty[s], ph ^ 1);              // stage free (first round passes immediately)
            unsigned char *hi = sA + s * 2 * A_BYTES, *lo = hi + A_BYTES;
            const int off = (r & 7) * 16 + (r >> 3) * SBO;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float4 m = __ldg(mp + q);
                if (SCALED) { m.x = __fmul_rn(rs, m.x); m.y = __fmul_rn(rs, m.y); m.z = __fmul_rn(rs, m.z); m.w = __fmul_rn(rs, m.w); }
                float4 a, hh, l;
                a.x = in ? __fsub_rn(cur[q].x, m.x) : 0.f; a.y = in ? __fsub_rn(cur[q].y, m.y) : 0.f;
                a.z = in ? __fsub_rn(cur[q].z, m.z) : 0.f; a.w = in ? __fsub_rn(cur[q].w, m.w) : 0.f;
                hh.x = tf32_hi(a.x); hh.y = tf32_hi(a.y); hh.z = tf32_hi(a.z); hh.w = tf32_hi(a.w);
                l.x = a.x - hh.x; l.y = a.y - hh.y; l.z = a.z - hh.z; l.w = a.w - hh.w;
                *reinterpret_cast<float4 *>(hi + off + q * 128) = hh;
                *reinterpret_cast<float4 *>(lo + off + q * 128) = l;
            }
            fence_proxy_async();                       // generic-proxy smem writes -> visible to the tensor core
            mbar_arrive(&full_a[s]);
#pragma unroll
            for (int q = 0; q < Q; ++q) cur[q] = nxt[q];
            tile = ntile;
            pos = npos;
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------------ epilogue (thread = row, own TMEM lane)
        const int q4 = warp - 4;                       // TMEM lane quarter of this warp
        const int r = q4 * 32 + lane;
        float *tile_s = sE + q4 * 32 * EPI_LD;         // this warp's staging tile
        const int sub = lane >> 3, c4 = lane & 7;      // store phase: lane -> (row within a group of 4, float4 column)
        uint32_t t = 0;                                // accumulator passes consumed so far
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t row = tile * BM + r, row_w0 = tile * BM + q4 * 32;     // this thread's row; first row of the warp
            uint32_t taddr[2];
            for (int h = 0; h < halves; ++h) {
                const uint32_t tt = t + h;
                mbar_wait(&acc_full[tt & 1], (tt >> 1) & 1, 200);
                taddr[h] = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)((tt & 1) * NMAX);
            }
            tc_fence_after();
            float scale = 1.f;
            uint32_t v[32];
            if (NORM == CLEORA_NORM_L2_NUMPY) {
                float s4[4] = {0.f, 0.f, 0.f, 0.f};         // four interleaved partial sums: a 4x shorter dependency chain
                for (int h = 0; h < halves; ++h)
                    for (int c0 = 0; c0 < N; c0 += 32) {
                        tmem_ld32(taddr[h] + c0, v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) { const float f = __uint_as_float(v[j]); s4[j & 3] = fmaf(f, f, s4[j & 3]); }
                    }
                const float ss = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                // One IEEE division per row, then a multiplication per element.  (Round 2, profiles/r2n: 256 inline
                // __fdiv_rn per row -- MUFU.RCP + Newton steps + a range check with a slow-path call each, ~28 SASS
                // instructions, serialised by the checks -- made this epilogue the bound of the whole kernel: 1.01 ms
                // for every main-loop variant.  The product differs from x / norm by at most 1 ulp, far below the
                // 3xTF32 error of the GEMM that produced x, and K3's rows have no bit-exact counterpart elsewhere.)
                scale = __fdiv_rn(1.0f, fmaxf(sqrtf(ss), 1e-10f));
            }
            for (int h = 0; h < halves; ++h)
                for (int c0 = 0; c0 < N; c0 += 32) {
                    tmem_ld32(taddr[h] + c0, v);
                    __syncwarp();                              // the previous block has been read out of the tile
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4 o;
                        o.x = __uint_as_float(v[4 * j + 0]); o.y = __uint_as_float(v[4 * j + 1]);
                        o.z = __uint_as_float(v[4 * j + 2]); o.w = __uint_as_float(v[4 * j + 3]);
                        if (NORM == CLEORA_NORM_L2_NUMPY) {
                            o.x = __fmul_rn(o.x, scale); o.y = __fmul_rn(o.y, scale);
                            o.z = __fmul_rn(o.z, scale); o.w = __fmul_rn(o.w, scale);
                        }
                        *reinterpret_cast<float4 *>(tile_s + lane * EPI_LD + 4 * j) = o;
                    }
                    __syncwarp();
                    const int col0 = h * N + c0;               // first output column of this 32 x 32 block
                    // SLICES mode: this lane's float4 column belongs to the rank that owns its column slice
                    float *slice_ptr = nullptr;
                    if (peers.mode == PEER_SLICES) {
                        const int col = col0 + 4 * c4, owner = col / peers.slice_cols;
#pragma unroll
                        for (int p = 0; p < 8; ++p)
                            if (p == owner) slice_ptr = peers.extra[p] + (col - owner * peers.slice_cols);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {              // 4 rows x 128 bytes per instruction
                        const int rr = 4 * i + sub;
                        const int64_t grow = row_w0 + rr;
                        if (grow >= n) continue;
                        const float4 o = *reinterpret_cast<const float4 *>(tile_s + rr * EPI_LD + 4 * c4);
                        reinterpret_cast<float4 *>(out + grow * (int64_t)NT + col0)[c4] = o;
                        if (peers.mode == PEER_SLICES) {
                            *reinterpret_cast<float4 *>(slice_ptr + (peers.row_base + grow) * (int64_t)peers.slice_cols) = o;
                        } else {
#pragma unroll
                            for (int p = 0; p < 7; ++p)        // fused all-gather into the peers' copies
                                if (p < peers.n_extra) reinterpret_cast<float4 *>(peers.extra[p] + grow * (int64_t)NT + col0)[c4] = o;
                        }
                    }
                }
            tc_fence_before();
            for (int h = 0; h < halves; ++h) mbar_arrive(&acc_empty[(t + h) & 1]);   // buffers may be overwritten
            t += halves;
            (void)row;
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = make_idesc_tf32(BM, N);
        uint32_t it = 0, t = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int h = 0; h < halves; ++h, ++t) {
                const int buf = t & 1;
                mbar_wait(&acc_empty[buf], ((t >> 1) & 1) ^ 1, 100);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NMAX);
                int last_c = 0;
                for (int c = 0; c < n_chunks; ++c) if (first_col(h, c) < N) last_c = c;
                for (int c = 0; c <= last_c; ++c, ++it) {                  // (steps right of last_c touch no column)
                    const int n0 = first_col(h, c);
                    const uint32_t idesc_c = upper ? make_idesc_tf32(BM, N - n0) : idesc;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full_a[s], ph);
                    mbar_wait(&full_b[s], ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_hi = smem_u32(sA + s * 2 * A_BYTES), a_lo = a_hi + A_BYTES;
                        const uint32_t b_hi = smem_u32(sB + s * 2 * b_bytes) + (uint32_t)(n0 / 8) * (uint32_t)SBO, b_lo = b_hi + b_bytes;
                        const uint32_t tmem_c = tmem_d + (uint32_t)n0;
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {                 // UMMA K = 8 tf32 = two 16-byte core columns
                            const uint32_t ko = k * 256;
                            // swizzled A: the K step moves 32 bytes inside the 128-byte rows (LBO field unused = 16 B), rows
                            // groups are 1024 bytes apart; the hardware applies the XOR from the address bits
                            const uint64_t dah = ASW ? make_desc(a_hi + k * 32, 16, 1024) | (2ull << 61) : make_desc(a_hi + ko, 128, SBO);
                            const uint64_t dal = ASW ? make_desc(a_lo + k * 32, 16, 1024) | (2ull << 61) : make_desc(a_lo + ko, 128, SBO);
                            const uint64_t dbh = make_desc(b_hi + ko, 128, SBO), dbl = make_desc(b_lo + ko, 128, SBO);
                            mma_tf32(tmem_c, dal, dbh, idesc_c, (c | k) != 0);   // small terms first
                            mma_tf32(tmem_c, dah, dbl, idesc_c, 1);
                            mma_tf32(tmem_c, dah, dbh, idesc_c, 1);
                        }
                        mma_commit(&empty[s]);                             // stage reusable once these MMAs retire
                        if (c == last_c) mma_commit(&acc_full[buf]);       // accumulator complete
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ B loaders (cp.async, 2 warps)
        const int lt = threadIdx.x - 9 * 32;                    // 0..63
        const int plane_pieces = b_bytes / 16;                  // 16-byte pieces of one half of one plane
        const int64_t plane_floats = (int64_t)NT * BK;          // one full-width plane of a chunk in the image
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int h = 0; h < halves; ++h)
            for (int c = 0; c < n_chunks; ++c) {
                const int n0 = first_col(h, c);
                if (n0 >= N) continue;                                     // nothing of this chunk lands in this half
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                ++it;
                mbar_wait(&empty[s], ph ^ 1);
                unsigned char *dst = sB + s * 2 * b_bytes;
                // rows [h*N, (h+1)*N) of the chunk: 8-row groups are SBO bytes apart in the image; rows below n0 are zero
                const unsigned char *src = reinterpret_cast<const unsigned char *>(Bt + (int64_t)c * 2 * plane_floats) + (size_t)h * b_bytes;
                for (int p = lt + (n0 / 8) * (SBO / 16); p < plane_pieces; p += B_LOAD_THREADS) {
                    cp_async_cg16(dst + p * 16, src + p * 16);                                            // hi
                    cp_async_cg16(dst + b_bytes + p * 16, src + plane_floats * 4 + p * 16);                // lo
                }
                cp_async_arrive_noinc(&full_b[s]);
            }
        }
        cp_async_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// Stage shape of K3: 32 = (BK 32, 2 stages), 16 = (BK 16, 4 stages).  cleora_set_option("k3_bk", ...) or CLEORA_B200_K3_BK.
std::atomic<int> g_k3_bk{[] {
    const char *e = getenv("CLEORA_B200_K3_BK");
    const int v = e ? atoi(e) : 0;
    return (v == 16 || v == 32) ? v : 32;
}()};

// A-tile layout of K3 (BK = 32): 0 = row-per-thread producers / no swizzle, 1 = coalesced producers / SWIZZLE_128B.
std::atomic<int> g_k3_asw{[] {
    const char *e = getenv("CLEORA_B200_K3_ASW");
    return e ? (atoi(e) != 0 ? 1 : 0) : 1;
}()};

bool whiten_apply_tc_supported(int64_t d, int64_t dout) {
    if (d % tc::BK_MAX != 0 || d < tc::BK_MAX || dout < 16) return false;
    if (dout <= tc::NMAX) return dout % 32 == 0;                    // the epilogue moves 32-column blocks
    return dout <= 2 * tc::NMAX && dout % 64 == 0;                  // two halves of <= 256 columns each
}

// Scratch for the pre-tiled transform lives in the caller's workspace (misc).
void launch_whiten_apply_tc(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T, int64_t dout,
                            float *out, int norm, const float *rowscale, cudaStream_t st, const PeerOut *peers_in,
                            bool upper_triangular) {
    using namespace tc;
    if (n == 0) return;
    PeerOut peers{};
    if (peers_in) peers = *peers_in;
    if (peers.mode == PEER_SLICES && (peers.slice_cols % 4 != 0 || peers.slice_cols * peers.n_extra != dout))
        throw CudaFail{"column slices must be multiples of 4 columns and cover the output"};
    if (peers.mode == PEER_OWNERS) throw CudaFail{"tensor-core apply: unsupported destination mode"};
    float *Bt = (float *)workspace().misc.get((size_t)2 * d * dout * sizeof(float));
    const int64_t tot = d * dout;
    const int bk = g_k3_bk.load(), stages = bk == 16 ? 4 : 2;
    prep_transform_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(T, (int)d, (int)dout, bk, Bt);
    LAUNCH_CHECK();
    const int NT = (int)dout, N = NT > NMAX ? NT / 2 : NT;
    const size_t smem = (size_t)stages * 2 * (BM * bk * 4) + (size_t)stages * 2 * N * bk * 4 + EPI_BYTES +
                        (3 * stages + 4) * sizeof(uint64_t) + 16 + 1024;      // + slack for the 1024-byte alignment
    const int64_t n_tiles = (n + BM - 1) / BM;
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, 148);
    auto launch = [&](auto kernel) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        kernel<<<grid, THREADS, smem, st>>>(x, n, (int)d, mean_f32, rowscale, Bt, NT, (upper_triangular && d == dout) ? 1 : 0, out, peers);
    };
    const bool l2 = norm == CLEORA_NORM_L2_NUMPY;
    if (norm != CLEORA_NORM_NONE && !l2) throw CudaFail{"tensor-core apply: unsupported fused normalisation"};
    auto pick = [&](auto norm_c, auto scaled_c) {
        constexpr int NORM = decltype(norm_c)::value, SCALED = decltype(scaled_c)::value;
        if (bk == 16) launch(whiten_apply_tc_kernel<NORM, SCALED, 16, 4, 0>);
        else if (g_k3_asw.load()) launch(whiten_apply_tc_kernel<NORM, SCALED, 32, 2, 1>);
        else launch(whiten_apply_tc_kernel<NORM, SCALED, 32, 2, 0>);
    };
    using std::integral_constant;
    if (rowscale) { if (l2) pick(integral_constant<int, CLEORA_NORM_L2_NUMPY>{}, integral_constant<int, 1>{}); else pick(integral_constant<int, 0>{}, integral_constant<int, 1>{}); }
    else          { if (l2) pick(integral_constant<int, CLEORA_NORM_L2_NUMPY>{}, integral_constant<int, 0>{}); else pick(integral_constant<int, 0>{}, integral_constant<int, 0>{}); }
    LAUNCH_CHECK();
}

}  // namespace cleora
