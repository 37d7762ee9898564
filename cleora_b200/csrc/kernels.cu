// sm_100a kernels of libcleora_b200.
//
//   K0 init_kernel            deterministic init                     src/lib.rs:69-81,478-488
//   K1 spmm_rows_kernel       CSR x dense SpMM + residual + row norm src/embedding.rs:52-104,121-131,
//                                                                    pycleora/__init__.py:114-115,943-950
//   K2 col_sums / gram_f64    column mean, centred covariance (f64)  pycleora/__init__.py:136-143
//   K3 whiten_apply_kernel    (X - mean) @ T in f32                  pycleora/__init__.py:157-163
//      sq_diff                rmse numerator                         src/embedding.rs:169-177, __init__.py:974-976
//
// Numerical contract of K1: per output element the products are accumulated in the row's stored (column)
// order with a separate f32 multiply and f32 add (__fmul_rn/__fadd_rn, no FMA contraction) -- the exact
// operation sequence of the reference's spmm_kernel, so the un-normalised SpMM is bit-identical to the CPU path.
// All of these are HBM-bound gather/stream kernels: no tensor cores here by design.
#include "device.cuh"
#include "../../include/cleora_b200.h"

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>

namespace cleora {

static constexpr unsigned FULL = 0xffffffffu;

// ================================================================================================ K0
__global__ void init_kernel(const uint64_t *__restrict__ hash, int64_t n, int64_t d, int64_t seed,
                            float *__restrict__ out) {
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d, j = i - r * d;
        const uint64_t num = hash[r] + (uint64_t)j + (uint64_t)seed;          // wrapping i64 adds
        const int64_t hashed = (int64_t)(num * 0x517cc1b727220a95ULL);       // FxHasher::write_i64 from zero state
        out[i] = (float)(hashed % 8388608LL) / 8388608.0f;                     // % keeps the dividend's sign
    }
}

void launch_init(const uint64_t *hash, int64_t n, int64_t d, int64_t seed, float *out, cudaStream_t st) {
    if (n * d == 0) return;
    const int threads = 256;
    const int64_t blocks = std::min<int64_t>((n * d + threads - 1) / threads, 148 * 32);
    init_kernel<<<(unsigned)blocks, threads, 0, st>>>(hash, n, d, seed, out);
    LAUNCH_CHECK();
}

// ================================================================================================ K1
__device__ __forceinline__ float group_sum(float v, int width) {
    for (int off = width >> 1; off > 0; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(FULL, v, off));
    return v;
}

// Row-norm epilogue shared by the SpMM kernels: `a` holds this lane's NV values of the row, `part` the lane's
// partial (sum of squares or of |x|) already reduced over the lane group.
__device__ __forceinline__ float norm_scale_factor(float part, int norm) {
    float nrm = (norm == CLEORA_NORM_L1_NUMPY) ? part : sqrtf(part);
    return fmaxf(nrm, 1e-10f);
}

// LPR lanes cooperate on one row (32/LPR rows per warp); each lane owns VEC float4 column groups, so
// D = 4*LPR*VEC.  Edge (col,val) pairs are fetched LPR at a time with one coalesced load per lane and broadcast
// by shuffle; X rows are read with 128-bit read-only loads, U rows in flight before the first dependent add.
// acc += sum over edges [s, e) in stored order (separate multiply and add); all 32 lanes of the warp must call it.
template <int LPR, int VEC, int U>
__device__ __forceinline__ void accumulate_edges(const uint32_t *__restrict__ col, const float *__restrict__ val,
                                                 const float4 *__restrict__ xv, int64_t s, int64_t e, int gl,
                                                 float4 (&acc)[VEC]) {
    constexpr int D4 = LPR * VEC;
    for (int64_t base = s; __any_sync(FULL, base < e); base += LPR) {
        const int64_t rem = e - base;
        const int cnt = rem > LPR ? LPR : (int)rem;            // <= 0 for groups that are already done
        uint32_t my_c = 0;
        float my_v = 0.f;
        if (gl < cnt) { my_c = __ldg(col + base + gl); my_v = __ldg(val + base + gl); }
#pragma unroll 1
        for (int k = 0; k < LPR; k += U) {
            if (!__any_sync(FULL, k < cnt)) break;
            float4 xr[U][VEC];
            float vv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t c = __shfl_sync(FULL, my_c, k + u, LPR);
                vv[u] = __shfl_sync(FULL, my_v, k + u, LPR);
                if (k + u < cnt) {
                    const float4 *p = xv + (int64_t)c * D4 + gl;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) xr[u][v] = __ldg(p + v * LPR);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (k + u < cnt) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        acc[v].x = __fadd_rn(acc[v].x, __fmul_rn(vv[u], xr[u][v].x));
                        acc[v].y = __fadd_rn(acc[v].y, __fmul_rn(vv[u], xr[u][v].y));
                        acc[v].z = __fadd_rn(acc[v].z, __fmul_rn(vv[u], xr[u][v].z));
                        acc[v].w = __fadd_rn(acc[v].w, __fmul_rn(vv[u], xr[u][v].w));
                    }
                }
            }
        }
    }
}

// Epilogue: residual mix (embedding.rs:121-129), row norm, store.  All lanes of the warp must call it.
template <int LPR, int VEC>
__device__ __forceinline__ void finish_row(float4 (&acc)[VEC], int64_t row, bool valid, int gl, float *__restrict__ out,
                                           const float *__restrict__ resid, float alpha, float rw, int norm,
                                           const PeerOut &peers) {
    constexpr int D4 = LPR * VEC;
    if (resid != nullptr && valid) {                    // dst = alpha*dst + rw*src
        const float4 *rp = reinterpret_cast<const float4 *>(resid) + row * D4 + gl;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float4 r = __ldg(rp + v * LPR);
            acc[v].x = __fadd_rn(__fmul_rn(alpha, acc[v].x), __fmul_rn(rw, r.x));
            acc[v].y = __fadd_rn(__fmul_rn(alpha, acc[v].y), __fmul_rn(rw, r.y));
            acc[v].z = __fadd_rn(__fmul_rn(alpha, acc[v].z), __fmul_rn(rw, r.z));
            acc[v].w = __fadd_rn(__fmul_rn(alpha, acc[v].w), __fmul_rn(rw, r.w));
        }
    }
    if (norm != CLEORA_NORM_NONE) {
        float part = 0.f;
        if (norm == CLEORA_NORM_L1_NUMPY) {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                part = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(part, fabsf(acc[v].x)), fabsf(acc[v].y)), fabsf(acc[v].z)), fabsf(acc[v].w));
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                part = __fadd_rn(part, __fmul_rn(acc[v].x, acc[v].x));
                part = __fadd_rn(part, __fmul_rn(acc[v].y, acc[v].y));
                part = __fadd_rn(part, __fmul_rn(acc[v].z, acc[v].z));
                part = __fadd_rn(part, __fmul_rn(acc[v].w, acc[v].w));
            }
        }
        part = group_sum(part, LPR);
        const float nrm = norm_scale_factor(part, norm);
        if (norm == CLEORA_NORM_L2_RUST) {
            const float inv = __fdiv_rn(1.0f, nrm);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc[v].x = __fmul_rn(acc[v].x, inv); acc[v].y = __fmul_rn(acc[v].y, inv);
                acc[v].z = __fmul_rn(acc[v].z, inv); acc[v].w = __fmul_rn(acc[v].w, inv);
            }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc[v].x = __fdiv_rn(acc[v].x, nrm); acc[v].y = __fdiv_rn(acc[v].y, nrm);
                acc[v].z = __fdiv_rn(acc[v].z, nrm); acc[v].w = __fdiv_rn(acc[v].w, nrm);
            }
        }
    }
    if (valid) {
        if (peers.mode == PEER_OWNERS) {                      // column-sharded SpMM: the row belongs to another rank's block
            // (row and block_rows are below 2^32: column indices are 32-bit, the matrix is square)
            const uint32_t owner = (uint32_t)row / (uint32_t)peers.block_rows;
            const int64_t local = row - (int64_t)owner * peers.block_rows;
            float4 *pp = nullptr;
#pragma unroll
            for (int p = 0; p < 8; ++p)                       // static indices keep the pointers in the constant bank
                if ((uint32_t)p == owner) pp = reinterpret_cast<float4 *>(peers.extra[p] + local * peers.ld_cols + peers.col_off) + gl;
#pragma unroll
            for (int v = 0; v < VEC; ++v) pp[v * LPR] = acc[v];
            return;
        }
        float4 *op = reinterpret_cast<float4 *>(out) + row * D4 + gl;
#pragma unroll
        for (int v = 0; v < VEC; ++v) op[v * LPR] = acc[v];
        if (peers.mode == PEER_SLICES) {                      // row-sharded -> column-sharded: every float4 to its slice owner
            const int sc4 = peers.slice_cols >> 2;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int q = gl + v * LPR, owner = q / sc4;
                float4 *pp = nullptr;
#pragma unroll
                for (int p = 0; p < 8; ++p)
                    if (p == owner) pp = reinterpret_cast<float4 *>(peers.extra[p]) + (peers.row_base + row) * sc4 + (q - owner * sc4);
                *pp = acc[v];
            }
            return;
        }
#pragma unroll
        for (int p = 0; p < 7; ++p) {                         // fused all-gather: the same row into the peers' copies
            if (p < peers.n_extra) {                          // (static indices keep the pointers in the constant bank)
                float4 *pp = reinterpret_cast<float4 *>(peers.extra[p]) + row * D4 + gl;
#pragma unroll
                for (int v = 0; v < VEC; ++v) pp[v * LPR] = acc[v];
            }
        }
    }
}

// Rows with more than `long_threshold` edges are left to the long-row kernels below.
template <int LPR, int VEC, int U>
__global__ void __launch_bounds__(256) spmm_rows_kernel(const int64_t *__restrict__ rowptr,
                                                        const uint32_t *__restrict__ col,
                                                        const float *__restrict__ val, const float *__restrict__ x,
                                                        float *__restrict__ out, const float *__restrict__ resid,
                                                        int64_t n_rows, float alpha, float rw, int norm,
                                                        int64_t long_threshold, const uint32_t *__restrict__ order,
                                                        PeerOut peers) {
    constexpr int RPW = 32 / LPR;          // rows per warp
    const int lane = threadIdx.x & 31;
    const int gl = lane & (LPR - 1);
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int64_t row = warp * RPW + lane / LPR;
    bool valid = row < n_rows;
    // Several rows share a warp when the rows are narrow: walk them in degree order (skewed graphs only) so that the
    // lane groups of a warp finish together -- on a power-law graph the warp otherwise runs as long as its longest row.
    // Measured on the products-shaped graph (ms for the full product, index order -> degree order inside 4096-row
    // windows; full rows: 17.5): 64-float slices 20.8 -> 16.5, 32-float 27.2 -> 24.1, but 16-float 40.0 -> 60.5 and
    // 8-float 94 -> 172 (profiles/r2d_, r2g_k1_slices_c3.txt) -- so only the lane-group widths 8 and 16 use it.
    if (LPR < 32 && LPR >= 8 && order != nullptr && valid) row = order[row];
    int64_t s = 0, e = 0;
    if (valid) { s = rowptr[row]; e = rowptr[row + 1]; }
    if (e - s > long_threshold) {                               // a hub row: the chunked kernels own it
        if (LPR == 32) return;                                  // warp-uniform: one row per warp
        e = s;
        valid = false;
    }
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_edges<LPR, VEC, U>(col, val, reinterpret_cast<const float4 *>(x), s, e, gl, acc);
    finish_row<LPR, VEC>(acc, row, valid, gl, out, resid, alpha, rw, norm, peers);
}

// Long rows (hubs of power-law graphs): one lane group per CHUNK of a long row writes a partial sum; one lane group per
// long row then adds the partials in chunk order and runs the usual epilogue.  Deterministic; the summation tree differs
// from the sequential reference order for these rows only (documented deviation, a few ulp).
template <int LPR, int VEC, int U>
__global__ void __launch_bounds__(256) spmm_long_partial_kernel(const int64_t *__restrict__ rowptr,
                                                                const uint32_t *__restrict__ col,
                                                                const float *__restrict__ val,
                                                                const float *__restrict__ x,
                                                                const int64_t *__restrict__ long_rows,
                                                                const int64_t *__restrict__ chunk_ptr,
                                                                const int32_t *__restrict__ chunk_owner, int64_t n_chunks,
                                                                int64_t chunk_edges, int64_t long_threshold,
                                                                float *__restrict__ partial) {
    constexpr int RPW = 32 / LPR, D4 = LPR * VEC;
    const int lane = threadIdx.x & 31, gl = lane & (LPR - 1);
    const int64_t c = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    bool valid = c < n_chunks;
    int64_t s = 0, e = 0;
    if (valid) {
        const int32_t ri = chunk_owner[c];
        const int64_t row = long_rows[ri];
        if (rowptr[row + 1] - rowptr[row] > long_threshold) {   // the schedule lists every row that is long for SOME width
            s = rowptr[row] + (c - chunk_ptr[ri]) * chunk_edges;
            e = min(rowptr[row + 1], s + chunk_edges);
        } else {
            valid = false;
        }
    }
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    accumulate_edges<LPR, VEC, U>(col, val, reinterpret_cast<const float4 *>(x), s, e, gl, acc);
    if (valid) {
        float4 *pp = reinterpret_cast<float4 *>(partial) + c * D4 + gl;
#pragma unroll
        for (int v = 0; v < VEC; ++v) pp[v * LPR] = acc[v];
    }
}

template <int LPR, int VEC>
__global__ void __launch_bounds__(256) spmm_long_finish_kernel(const int64_t *__restrict__ rowptr,
                                                               const int64_t *__restrict__ long_rows,
                                                               const int64_t *__restrict__ chunk_ptr, int64_t n_long,
                                                               int64_t long_threshold,
                                                               const float *__restrict__ partial, float *__restrict__ out,
                                                               const float *__restrict__ resid, float alpha, float rw,
                                                               int norm, PeerOut peers) {
    constexpr int RPW = 32 / LPR, D4 = LPR * VEC;
    const int lane = threadIdx.x & 31, gl = lane & (LPR - 1);
    const int64_t ri = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    bool valid = ri < n_long;
    if (valid) { const int64_t row = long_rows[ri]; valid = rowptr[row + 1] - rowptr[row] > long_threshold; }
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid)
        for (int64_t c = chunk_ptr[ri]; c < chunk_ptr[ri + 1]; ++c) {
            const float4 *pp = reinterpret_cast<const float4 *>(partial) + c * D4 + gl;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float4 p = pp[v * LPR];
                acc[v].x = __fadd_rn(acc[v].x, p.x); acc[v].y = __fadd_rn(acc[v].y, p.y);
                acc[v].z = __fadd_rn(acc[v].z, p.z); acc[v].w = __fadd_rn(acc[v].w, p.w);
            }
        }
    finish_row<LPR, VEC>(acc, valid ? long_rows[ri] : 0, valid, gl, out, resid, alpha, rw, norm, peers);
}

// Any d: one warp per row, lane owns columns lane, lane+32, ... in passes of T*32 columns.  Same accumulation
// order as above.  With more than one pass the row is written un-normalised first and rescaled afterwards.
template <int T>
__global__ void __launch_bounds__(256) spmm_generic_kernel(const int64_t *__restrict__ rowptr,
                                                           const uint32_t *__restrict__ col,
                                                           const float *__restrict__ val, const float *__restrict__ x,
                                                           float *__restrict__ out, const float *__restrict__ resid,
                                                           int64_t n_rows, int d, float alpha, float rw, int norm) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;                                     // whole warp exits together
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    float part = 0.f;
    for (int c0 = 0; c0 < d; c0 += T * 32) {
        float acc[T];
#pragma unroll
        for (int t = 0; t < T; ++t) acc[t] = 0.f;
        for (int64_t base = s; base < e; base += 32) {
            const int cnt = (e - base) > 32 ? 32 : (int)(e - base);
            uint32_t my_c = 0;
            float my_v = 0.f;
            if (lane < cnt) { my_c = __ldg(col + base + lane); my_v = __ldg(val + base + lane); }
            for (int k = 0; k < cnt; ++k) {
                const uint32_t c = __shfl_sync(FULL, my_c, k);
                const float v = __shfl_sync(FULL, my_v, k);
                const float *xr = x + (int64_t)c * d + c0 + lane;
#pragma unroll
                for (int t = 0; t < T; ++t)
                    if (c0 + t * 32 + lane < d) acc[t] = __fadd_rn(acc[t], __fmul_rn(v, __ldg(xr + t * 32)));
            }
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const int j = c0 + t * 32 + lane;
            if (j < d) {
                float a = acc[t];
                if (resid != nullptr) a = __fadd_rn(__fmul_rn(alpha, a), __fmul_rn(rw, __ldg(resid + row * d + j)));
                part = __fadd_rn(part, norm == CLEORA_NORM_L1_NUMPY ? fabsf(a) : __fmul_rn(a, a));
                out[row * d + j] = a;
            }
        }
    }
    if (norm == CLEORA_NORM_NONE) return;
    part = group_sum(part, 32);
    const float nrm = norm_scale_factor(part, norm);
    const float inv = __fdiv_rn(1.0f, nrm);
    __syncwarp();
    for (int j = lane; j < d; j += 32) {
        const float a = out[row * d + j];
        out[row * d + j] = (norm == CLEORA_NORM_L2_RUST) ? __fmul_rn(a, inv) : __fdiv_rn(a, nrm);
    }
}

template <int LPR, int VEC, int U>
static void launch_rows(const DeviceGraph &g, const float *val, const float *x, float *out, const float *resid,
                        float alpha, float rw, int norm, cudaStream_t st, const PeerOut &peers) {
    constexpr int RPW = 32 / LPR;
    const int threads = 256;
    const int64_t rows_per_block = (int64_t)(threads / 32) * RPW;
    const int64_t blocks = (g.n_rows + rows_per_block - 1) / rows_per_block;
    // A row is walked sequentially by ONE lane group, LPR edges per memory round trip: the time of the longest row is a
    // floor for the launch, and with narrow rows (few lanes per row) that floor is reached 32/LPR times sooner --
    // measured on the products-shaped graph, 25k-edge hubs: 2.5 ms per launch at 4 lanes per row against 1.3 ms for
    // all other rows together.  So the splitting threshold scales with the lane-group width (full-width rows keep the
    // configured value and with it their bit-exact sequential sums up to that degree).
    const int64_t thr = std::max<int64_t>(g.long_sched_threshold, g.long_threshold / 32 * LPR);
    const bool split = g.n_long > 0 && thr < g.max_degree;
    spmm_rows_kernel<LPR, VEC, U><<<(unsigned)blocks, threads, 0, st>>>(
        g.rowptr, g.col, val, x, out, resid, g.n_rows, alpha, rw, norm, split ? thr : INT64_MAX,
        (LPR < 32 && LPR >= 8) ? g.row_order : nullptr, peers);
    LAUNCH_CHECK();
    if (split) {
        float *partial = (float *)workspace().spmm_partials.get((size_t)g.n_long_chunks * LPR * VEC * 4 * sizeof(float));
        spmm_long_partial_kernel<LPR, VEC, U><<<(unsigned)((g.n_long_chunks + rows_per_block - 1) / rows_per_block), threads, 0, st>>>(
            g.rowptr, g.col, val, x, g.long_rows, g.long_chunk_ptr, g.long_chunk_owner, g.n_long_chunks,
            g.long_chunk_edges, thr, partial);
        LAUNCH_CHECK();
        spmm_long_finish_kernel<LPR, VEC><<<(unsigned)((g.n_long + rows_per_block - 1) / rows_per_block), threads, 0, st>>>(
            g.rowptr, g.long_rows, g.long_chunk_ptr, g.n_long, thr, partial, out, resid, alpha, rw, norm, peers);
        LAUNCH_CHECK();
    }
}

void launch_spmm(const DeviceGraph &g, const float *val, const float *x, int64_t d, float *out, const float *resid,
                 float alpha, float rw, int norm, cudaStream_t st, const PeerOut *peers_in) {
    if (g.n_rows == 0 || d == 0) return;
    PeerOut peers{};
    if (peers_in) peers = *peers_in;
    if (g.n_rows > (int64_t)0x7fffffff * 8) throw CudaFail{"too many rows for one launch"};
    if (peers.mode == PEER_OWNERS && norm != CLEORA_NORM_NONE) throw CudaFail{"row scatter produces column slices: no fused row norm"};
    switch (d) {
        case 8:    launch_rows<2, 1, 2>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 16:   launch_rows<4, 1, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 32:   launch_rows<8, 1, 8>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 64:   launch_rows<16, 1, 8>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 96:   launch_rows<8, 3, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 128:  launch_rows<32, 1, 8>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 192:  launch_rows<16, 3, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 256:  launch_rows<32, 2, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 384:  launch_rows<32, 3, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 512:  launch_rows<32, 4, 4>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        case 1024: launch_rows<32, 8, 2>(g, val, x, out, resid, alpha, rw, norm, st, peers); return;
        default: break;
    }
    if (peers.n_extra || peers.mode != PEER_REPLICATE) throw CudaFail{"peer push needs a vectorised feature dimension (8..1024 as listed)"};
    const int threads = 256;
    const int64_t blocks = (g.n_rows + 7) / 8;
    spmm_generic_kernel<8><<<(unsigned)blocks, threads, 0, st>>>(g.rowptr, g.col, val, x, out, resid, g.n_rows,
                                                                 (int)d, alpha, rw, norm);
    LAUNCH_CHECK();
}

// Row normalisation alone: warp per row, two passes over the row (second pass hits L1/L2).
__global__ void __launch_bounds__(256) normalize_kernel(const float *__restrict__ x, int64_t n, int d, int norm,
                                                        float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const float *xr = x + row * d;
    float part = 0.f;
    for (int j = lane; j < d; j += 32) {
        const float a = xr[j];
        part = __fadd_rn(part, norm == CLEORA_NORM_L1_NUMPY ? fabsf(a) : __fmul_rn(a, a));
    }
    part = group_sum(part, 32);
    const float nrm = norm_scale_factor(part, norm);
    const float inv = __fdiv_rn(1.0f, nrm);
    for (int j = lane; j < d; j += 32) {
        const float a = xr[j];
        out[row * d + j] = norm == CLEORA_NORM_NONE ? a : (norm == CLEORA_NORM_L2_RUST ? __fmul_rn(a, inv) : __fdiv_rn(a, nrm));
    }
}

// The same normalisation as K1's fused epilogue (identical lane mapping and summation tree, hence identical bits) for
// rows that were produced elsewhere -- the row-sharded half of the column-sharded multi-GPU loop -- with the
// destinations of finish_row (local copy + column slices to their owners).
template <int LPR, int VEC>
__global__ void __launch_bounds__(256) normalize_rows_kernel(const float *__restrict__ x, int64_t n_rows, int norm,
                                                             float *__restrict__ out, PeerOut peers) {
    constexpr int RPW = 32 / LPR, D4 = LPR * VEC;
    const int lane = threadIdx.x & 31, gl = lane & (LPR - 1);
    const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool valid = row < n_rows;
    float4 acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        acc[v] = valid ? __ldg(reinterpret_cast<const float4 *>(x) + row * D4 + gl + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
    finish_row<LPR, VEC>(acc, row, valid, gl, out, nullptr, 1.f, 0.f, norm, peers);
}

template <int LPR, int VEC>
static void launch_normalize_rows(const float *x, int64_t n, int norm, float *out, cudaStream_t st, const PeerOut &peers) {
    const int64_t rows_per_block = 8 * (32 / LPR);
    normalize_rows_kernel<LPR, VEC><<<(unsigned)((n + rows_per_block - 1) / rows_per_block), 256, 0, st>>>(x, n, norm, out, peers);
    LAUNCH_CHECK();
}

bool normalize_rows_supported(int64_t d) {
    switch (d) { case 8: case 16: case 32: case 64: case 96: case 128: case 192: case 256: case 384: case 512: case 1024: return true; }
    return false;
}

void launch_normalize_rows(const float *x, int64_t n, int64_t d, int norm, float *out, cudaStream_t st, const PeerOut *peers_in) {
    if (n == 0 || d == 0) return;
    PeerOut peers{};
    if (peers_in) peers = *peers_in;
    if (peers.mode == PEER_SLICES && (peers.slice_cols % 4 != 0 || (int64_t)peers.slice_cols * peers.n_extra != d))
        throw CudaFail{"column slices must be multiples of 4 columns and cover the row"};
    switch (d) {                                    // the lane mappings of launch_spmm
        case 8:    launch_normalize_rows<2, 1>(x, n, norm, out, st, peers); return;
        case 16:   launch_normalize_rows<4, 1>(x, n, norm, out, st, peers); return;
        case 32:   launch_normalize_rows<8, 1>(x, n, norm, out, st, peers); return;
        case 64:   launch_normalize_rows<16, 1>(x, n, norm, out, st, peers); return;
        case 96:   launch_normalize_rows<8, 3>(x, n, norm, out, st, peers); return;
        case 128:  launch_normalize_rows<32, 1>(x, n, norm, out, st, peers); return;
        case 192:  launch_normalize_rows<16, 3>(x, n, norm, out, st, peers); return;
        case 256:  launch_normalize_rows<32, 2>(x, n, norm, out, st, peers); return;
        case 384:  launch_normalize_rows<32, 3>(x, n, norm, out, st, peers); return;
        case 512:  launch_normalize_rows<32, 4>(x, n, norm, out, st, peers); return;
        case 1024: launch_normalize_rows<32, 8>(x, n, norm, out, st, peers); return;
        default: throw CudaFail{"row normaliser with destinations needs a vectorised feature dimension"};
    }
}

void launch_normalize(const float *x, int64_t n, int64_t d, int norm, float *out, cudaStream_t st) {
    if (n * d == 0) return;
    normalize_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(x, n, (int)d, norm, out);
    LAUNCH_CHECK();
}

// ================================================================================================ K2a column sums
// Stage 1: every warp of the grid walks rows warp, warp+W, ... and keeps f64 partial sums of its columns (lane owns
// columns lane, lane+32, ...); the 8 warps of a block then add their partials through shared memory in warp order,
// one partial row per block.  Stage 2: one thread per column adds the block partials in block order.
// Deterministic for a fixed launch shape.
// `absmax` (nullable, one float per block): max |x| over the elements this block read -- free here, and it saves the
// integer Gram kernel a pass of its own.
template <int T>
__global__ void __launch_bounds__(256) col_sums_stage1(const float *__restrict__ x, int64_t n, int d, int c0,
                                                       double *__restrict__ partial, float *__restrict__ absmax) {
    __shared__ double sh_acc[8][T * 32];
    __shared__ float sh_mx[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t W = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    double acc[T];
    float mx = 0.f;
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.0;
    for (int64_t r = w; r < n; r += W) {
        const float *xr = x + r * d + c0 + lane;
#pragma unroll
        for (int t = 0; t < T; ++t)
            if (c0 + t * 32 + lane < d) {
                const float v = __ldg(xr + t * 32);
                acc[t] += (double)v;
                mx = fmaxf(mx, fabsf(v));
            }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) sh_acc[wib][t * 32 + lane] = acc[t];
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if (lane == 0) sh_mx[wib] = mx;
    __syncthreads();
    for (int c = threadIdx.x; c < T * 32; c += blockDim.x)
        if (c0 + c < d) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q) s += sh_acc[q][c];
            partial[(int64_t)blockIdx.x * d + c0 + c] = s;
        }
    if (absmax != nullptr && threadIdx.x == 0) {
#pragma unroll
        for (int q = 1; q < 8; ++q) mx = fmaxf(mx, sh_mx[q]);
        absmax[blockIdx.x] = mx;
    }
}

// blockDim = (32 columns, 8 parts): thread (c, p) adds the partials w = p, p+8, ... of its column (coalesced over c),
// then the 8 parts are combined in fixed order -> deterministic.
__global__ void __launch_bounds__(256) col_sums_stage2(const double *__restrict__ partial, int64_t W, int d,
                                                       double *__restrict__ sums, int accumulate) {
    __shared__ double sh[8][33];
    const int c = threadIdx.x, p = threadIdx.y;
    const int j = blockIdx.x * 32 + c;
    double s = 0.0;
    if (j < d)
        for (int64_t w = p; w < W; w += 8) s += partial[w * d + j];
    sh[p][c] = s;
    __syncthreads();
    if (p == 0 && j < d) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += sh[q][c];
        sums[j] = accumulate ? sums[j] + t : t;
    }
}

void launch_col_sums(const float *x, int64_t n, int64_t d, double *sums, bool accumulate, cudaStream_t st,
                     AbsmaxPartials *absmax) {
    if (absmax) *absmax = AbsmaxPartials{};
    if (d == 0) return;
    const int threads = 256;
    int64_t blocks = std::min<int64_t>((n + 7) / 8, 148 * 4);
    if (blocks < 1) blocks = 1;
    const int64_t W = blocks;                                  // one partial row per block
    double *partial = (double *)workspace().colsum_partials.get(size_t(W) * size_t(d) * sizeof(double));
    float *mx = nullptr;
    if (absmax && d <= 8 * 32 && n > 0) {                      // one column pass covers the whole matrix
        mx = (float *)workspace().absmax_partials.get(size_t(blocks) * sizeof(float));
        absmax->p = mx;
        absmax->count = (int)blocks;
    }
    for (int c0 = 0; c0 < (int)d; c0 += 8 * 32) {
        col_sums_stage1<8><<<(unsigned)blocks, threads, 0, st>>>(x, n, (int)d, c0, partial, mx);
        LAUNCH_CHECK();
    }
    col_sums_stage2<<<(unsigned)((d + 31) / 32), dim3(32, 8), 0, st>>>(partial, W, (int)d, sums, accumulate ? 1 : 0);
    LAUNCH_CHECK();
}

// ================================================================================================ K2b centred Gram, f64
// cov = sum_r (x_r - mean)(x_r - mean)^T in IEEE f64 (the reference accumulates the covariance in f64,
// pycleora/__init__.py:138-142; the PCA eigenbasis is sensitive to ~1/eigengap, so lower precision is not an
// option).  FP64 tensor-core path: mma.sync m8n8k4 f64 (DMMA) -- tcgen05 has no f64 kind.
// One CTA = one 128x128 block (bi <= bj) of the d x d matrix for one slice of rows, 16 warps each owning a 32x32
// warp tile (4x4 DMMA fragments); in diagonal blocks the warp tiles strictly below the diagonal are skipped.
// Rows are staged 16 at a time: 128-bit global loads -> centre in f64 -> smem (row stride 132 doubles, which makes
// both the 32-byte staging stores and the m8n8k4 fragment loads bank-conflict free); the next chunk is prefetched
// into registers while the current one feeds the tensor pipe.  Partial blocks are reduced in slice order by
// gram_reduce_kernel (deterministic).
static constexpr int GB = 128;    // block edge
static constexpr int GK = 16;     // rows per staged chunk
static constexpr int GLD = 132;   // smem leading dimension in doubles (132 mod 16 == 4)
static constexpr int GT = 32;     // warp tile edge

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool VEC4>
__global__ void __launch_bounds__(512, 1) gram_f64_kernel(const float *__restrict__ x, int64_t n, int d,
                                                          const double *__restrict__ mean, double *__restrict__ partial,
                                                          int nblk, int64_t rows_per_slice) {
    __shared__ __align__(16) double As[GK][GLD];
    __shared__ __align__(16) double Bs[GK][GLD];
    __shared__ double mean_a[GB], mean_b[GB];
    int bi = 0, rest = blockIdx.x;                     // decode (bi, bj), bi <= bj, from the linear pair index
    while (rest >= nblk - bi) { rest -= nblk - bi; ++bi; }
    const int bj = bi + rest;
    const bool diag = bi == bj;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
    const int64_t r1 = min(n, r0 + rows_per_slice);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w >> 2, wn = w & 3;                 // warp tile: rows wm*32.., cols wn*32.. of the block
    const bool active = !(diag && wm > wn);
    if (tid < GB) {
        mean_a[tid] = (bi * GB + tid < d) ? mean[bi * GB + tid] : 0.0;
        mean_b[tid] = (bj * GB + tid < d) ? mean[bj * GB + tid] : 0.0;
    }
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // staging map: thread -> (row tid/32, 4 consecutive columns (tid%32)*4 ..) of the 16 x 128 chunk
    const int lr = tid >> 5, lc = (tid & 31) * 4;
    const int ca = bi * GB + lc, cb = bj * GB + lc;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    auto fetch = [&](int64_t rbase) {
        const int64_t r = rbase + lr;
        if (VEC4) {
            ra = (r < r1 && ca < d) ? __ldg(reinterpret_cast<const float4 *>(x + r * d + ca)) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (!diag) rb = (r < r1 && cb < d) ? __ldg(reinterpret_cast<const float4 *>(x + r * d + cb)) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float t[4], u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                t[q] = (r < r1 && ca + q < d) ? __ldg(x + r * d + ca + q) : 0.f;
                u[q] = (!diag && r < r1 && cb + q < d) ? __ldg(x + r * d + cb + q) : 0.f;
            }
            ra = make_float4(t[0], t[1], t[2], t[3]);
            rb = make_float4(u[0], u[1], u[2], u[3]);
        }
    };
    auto stage = [&](int64_t rbase) {
        const bool in = rbase + lr < r1;
        // out-of-range rows / columns contribute exactly zero (x - mean is forced to 0, not to -mean)
        double2 v0, v1;
        v0.x = (in && ca + 0 < d) ? (double)ra.x - mean_a[lc + 0] : 0.0;
        v0.y = (in && ca + 1 < d) ? (double)ra.y - mean_a[lc + 1] : 0.0;
        v1.x = (in && ca + 2 < d) ? (double)ra.z - mean_a[lc + 2] : 0.0;
        v1.y = (in && ca + 3 < d) ? (double)ra.w - mean_a[lc + 3] : 0.0;
        *reinterpret_cast<double2 *>(&As[lr][lc]) = v0;
        *reinterpret_cast<double2 *>(&As[lr][lc + 2]) = v1;
        if (!diag) {
            v0.x = (in && cb + 0 < d) ? (double)rb.x - mean_b[lc + 0] : 0.0;
            v0.y = (in && cb + 1 < d) ? (double)rb.y - mean_b[lc + 1] : 0.0;
            v1.x = (in && cb + 2 < d) ? (double)rb.z - mean_b[lc + 2] : 0.0;
            v1.y = (in && cb + 3 < d) ? (double)rb.w - mean_b[lc + 3] : 0.0;
            *reinterpret_cast<double2 *>(&Bs[lr][lc]) = v0;
            *reinterpret_cast<double2 *>(&Bs[lr][lc + 2]) = v1;
        }
    };

    if (r0 < r1) fetch(r0);
    for (int64_t rb0 = r0; rb0 < r1; rb0 += GK) {
        __syncthreads();                       // previous chunk fully consumed (and mean_* visible on the first trip)
        stage(rb0);
        __syncthreads();
        if (rb0 + GK < r1) fetch(rb0 + GK);    // prefetch the next chunk into registers while computing
        if (active) {
            const double(*Bp)[GLD] = diag ? As : Bs;
#pragma unroll
            for (int kk = 0; kk < GK / 4; ++kk) {
                const int kr = kk * 4 + (lane & 3);
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = As[kr][wm * GT + mi * 8 + (lane >> 2)];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = Bp[kr][wn * GT + ni * 8 + (lane >> 2)];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
    }
    if (!active) return;
    double *P = partial + (int64_t)blockIdx.y * d * d;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int gr = bi * GB + wm * GT + mi * 8 + (lane >> 2);
                const int gc = bj * GB + wn * GT + ni * 8 + (lane & 3) * 2 + j;
                if (gr < d && gc < d) P[(int64_t)gr * d + gc] = acc[mi][ni][j];
            }
}

// v3: the same Gram with the staging taken off the critical path.  Raw f32 tiles are copied global -> shared with
// cp.async (16-byte LDGSTS, 4 stages in flight, one __syncthreads per 16-row chunk); the f32 -> f64 conversion and
// the centring happen when a fragment is loaded (each thread keeps the 8 means it needs in registers).  Numerics
// are identical to gram_f64_kernel: (double)x - mean_f64, IEEE f64 DMMA accumulation in row order.
static constexpr int G3_STAGES = 4;
static constexpr int G3_LD = 136;          // floats per staged row (136 mod 32 == 8: conflict-free fragment loads)

__device__ __forceinline__ void cp_async16(void *dst, const void *src, int src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(512, 1) gram_f64_v3_kernel(const float *__restrict__ x, int64_t n, int d,
                                                             const double *__restrict__ mean, double *__restrict__ partial,
                                                             int nblk, int64_t rows_per_slice) {
    extern __shared__ __align__(16) float g3_smem[];      // [G3_STAGES][2][GK][G3_LD]
    int bi = 0, rest = blockIdx.x;
    while (rest >= nblk - bi) { rest -= nblk - bi; ++bi; }
    const int bj = bi + rest;
    const bool diag = bi == bj;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slice;
    const int64_t r1 = min(n, r0 + rows_per_slice);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w >> 2, wn = w & 3;
    const bool active = !(diag && wm > wn);

    double ma[4], mb[4];                                  // means of the 4+4 columns this lane's fragments touch
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int ca = bi * GB + wm * GT + q * 8 + (lane >> 2), cb = bj * GB + wn * GT + q * 8 + (lane >> 2);
        ma[q] = ca < d ? mean[ca] : 0.0;
        mb[q] = cb < d ? mean[cb] : 0.0;
    }
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // copy map: thread -> (row tid/32, 16-byte column chunk tid%32) of the 16 x 128 tile, for both operands
    const int lr = tid >> 5, lc = (tid & 31) * 4;
    const int ca0 = bi * GB + lc, cb0 = bj * GB + lc;
    const int64_t n_chunks = (r1 - r0 + GK - 1) / GK;
    auto issue = [&](int64_t chunk) {
        if (chunk < n_chunks) {
            float *sa = g3_smem + (size_t)(chunk % G3_STAGES) * 2 * GK * G3_LD, *sb = sa + GK * G3_LD;
            const int64_t r = r0 + chunk * GK + lr;
            const bool rin = r < r1;
            const float *ga = x + (rin ? r : 0) * (int64_t)d + (ca0 < d ? ca0 : 0);
            cp_async16(sa + lr * G3_LD + lc, ga, (rin && ca0 < d) ? 16 : 0);          // zero-fill outside
            if (!diag) {
                const float *gb = x + (rin ? r : 0) * (int64_t)d + (cb0 < d ? cb0 : 0);
                cp_async16(sb + lr * G3_LD + lc, gb, (rin && cb0 < d) ? 16 : 0);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < G3_STAGES - 1; ++s) issue(s);
    for (int64_t c = 0; c < n_chunks; ++c) {
        cp_async_wait<G3_STAGES - 2>();                  // chunk c has landed (for this thread's copies)
        __syncthreads();                                 // ... and for everyone's; chunk c-1 is fully consumed
        issue(c + G3_STAGES - 1);                        // refill the stage consumed in the previous trip
        if (active) {
            const float *sa = g3_smem + (size_t)(c % G3_STAGES) * 2 * GK * G3_LD;
            const float *sb = diag ? sa : sa + GK * G3_LD;
            const int rows_valid = (int)min((int64_t)GK, r1 - (r0 + c * GK));
#pragma unroll
            for (int kk = 0; kk < GK / 4; ++kk) {
                const int kr = kk * 4 + (lane & 3);
                const bool kin = kr < rows_valid;        // rows past the slice end contribute exactly zero
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) {
                    const float f = sa[kr * G3_LD + wm * GT + mi * 8 + (lane >> 2)];
                    a[mi] = kin ? (double)f - ma[mi] : 0.0;
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const float f = sb[kr * G3_LD + wn * GT + ni * 8 + (lane >> 2)];
                    b[ni] = kin ? (double)f - mb[ni] : 0.0;
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
        }
    }
    cp_async_wait<0>();
    if (!active) return;
    double *P = partial + (int64_t)blockIdx.y * d * d;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int gr = bi * GB + wm * GT + mi * 8 + (lane >> 2);
                const int gc = bj * GB + wn * GT + ni * 8 + (lane & 3) * 2 + j;
                if (gr < d && gc < d) P[(int64_t)gr * d + gc] = acc[mi][ni][j];
            }
}

// cov[i][j] = cov[j][i] = sum over slices, for the computed entries (32-tile of i <= 32-tile of j).
__global__ void gram_reduce_kernel(const double *__restrict__ partial, int slices, int d, double *__restrict__ cov) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)d * d) return;
    const int i = (int)(idx / d), j = (int)(idx - (int64_t)i * d);
    if (i / GT > j / GT) return;
    double s = 0.0;
    for (int k = 0; k < slices; ++k) s += partial[(int64_t)k * d * d + idx];
    cov[idx] = s;
    if (i / GT != j / GT) cov[(int64_t)j * d + i] = s;
}

bool gram_i8_supported(int64_t n, int64_t d);
void launch_centered_gram_i8(const float *x, int64_t n, int64_t d, const double *mean, double *cov, cudaStream_t st,
                             const AbsmaxPartials *absmax);

void launch_centered_gram(const float *x, int64_t n, int64_t d, const double *mean, double *cov, cudaStream_t st,
                          const AbsmaxPartials *absmax, bool ieee_f64) {
    if (d == 0) return;
    {   // exact-integer tcgen05 path for the large, common shapes; CLEORA_B200_GRAM=v3|v2 forces the FP64 DMMA kernels
        static const bool allow_i8 = [] { const char *e = getenv("CLEORA_B200_GRAM"); return !e || std::string(e) == "i8"; }();
        if (allow_i8 && !ieee_f64 && gram_i8_supported(n, d) && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            launch_centered_gram_i8(x, n, d, mean, cov, st, absmax);
            return;
        }
    }
    const int nblk = (int)((d + GB - 1) / GB);
    const int npairs = nblk * (nblk + 1) / 2;
    // one resident CTA per SM: ~148 CTAs in flight, a few waves when n is large
    int64_t slices = std::max<int64_t>(1, std::min<int64_t>((n + 8 * GK - 1) / (8 * GK), (148 * 2 + npairs - 1) / npairs));
    slices = std::min<int64_t>(slices, 65535);
    const int64_t rows_per_slice = std::max<int64_t>(GK, ((n + slices - 1) / slices + GK - 1) / GK * GK);
    slices = std::max<int64_t>(1, (n + rows_per_slice - 1) / rows_per_slice);
    double *partial = (double *)workspace().gram_partials.get(size_t(slices) * size_t(d) * size_t(d) * sizeof(double));
    dim3 grid((unsigned)npairs, (unsigned)slices);
    // (mma.sync.m16n8k16.f64 was tried: ptxas lowers it to the same DMMA.8x8x4 sequence on sm_100a -- no gain.)
    static const bool use_v3 = [] { const char *e = getenv("CLEORA_B200_GRAM"); return !(e && std::string(e) == "v2"); }();
    if (use_v3 && d % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const size_t smem = (size_t)G3_STAGES * 2 * GK * G3_LD * sizeof(float);      // 69,632 B
        CUDA_TRY(cudaFuncSetAttribute(gram_f64_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
        gram_f64_v3_kernel<<<grid, 512, smem, st>>>(x, n, (int)d, mean, partial, nblk, rows_per_slice);
    } else if (d % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
        gram_f64_kernel<true><<<grid, 512, 0, st>>>(x, n, (int)d, mean, partial, nblk, rows_per_slice);
    else
        gram_f64_kernel<false><<<grid, 512, 0, st>>>(x, n, (int)d, mean, partial, nblk, rows_per_slice);
    LAUNCH_CHECK();
    const int64_t tot = d * d;
    gram_reduce_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(partial, (int)slices, (int)d, cov);
    LAUNCH_CHECK();
}

// ================================================================================================ K3 whiten apply
// out[n, dout] = (x - mean_f32) @ T, f32 FMA accumulation in k order.  128x128x8 register-tiled SGEMM
// (8x8 per thread); the centring is applied while staging the A tile.
static constexpr int AM = 128, AN = 128, AK = 8;

__global__ void __launch_bounds__(256) whiten_apply_kernel(const float *__restrict__ x, int64_t n, int d,
                                                           const float *__restrict__ mean, const float *__restrict__ T,
                                                           int dout, float *__restrict__ out) {
    __shared__ __align__(16) float As[AK][AM + 4];
    __shared__ __align__(16) float Bs[AK][AN + 4];
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * AM;
    const int col0 = blockIdx.y * AN;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int a_r = tid >> 1, a_k = (tid & 1) * 4;      // A tile: 128 rows x 8 k, 4 consecutive k per thread
    const int b_k = tid >> 5, b_c = (tid & 31) * 4;     // B tile: 8 k x 128 cols, 4 consecutive cols per thread
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
        const int64_t r = row0 + a_r;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + a_k + q;
            ra[q] = (r < n && k < d) ? __ldg(x + r * d + k) - __ldg(mean + k) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + b_k, c = col0 + b_c + q;
            rb[q] = (k < d && c < dout) ? __ldg(T + (int64_t)k * dout + c) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < d; k0 += AK) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) { As[a_k + q][a_r] = ra[q]; Bs[b_k][b_c + q] = rb[q]; }
        __syncthreads();
        if (k0 + AK < d) fetch(k0 + AK);
#pragma unroll
        for (int k = 0; k < AK; ++k) {
            // thread owns rows {ty*4..+3, 64+ty*4..+3} x cols {tx*4..+3, 64+tx*4..+3}: 128-bit conflict-free LDS
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= n) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (c < dout) out[r * dout + c] = acc[i][j];
        }
    }
}


static bool use_tensor_core_apply() {
    static const bool on = [] { const char *e = getenv("CLEORA_B200_APPLY"); return !(e && std::string(e) == "simt"); }();
    return on;
}

void launch_whiten_apply(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T, int64_t dout,
                         float *out, cudaStream_t st) {
    if (n == 0 || dout == 0) return;
    if (use_tensor_core_apply() && whiten_apply_tc_supported(d, dout) &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(mean_f32)) & 15) == 0) {
        launch_whiten_apply_tc(x, n, d, mean_f32, T, dout, out, CLEORA_NORM_NONE, nullptr, st);    // tcgen05 / TMEM path
        return;
    }
    dim3 grid((unsigned)((n + AM - 1) / AM), (unsigned)((dout + AN - 1) / AN));
    whiten_apply_kernel<<<grid, 256, 0, st>>>(x, n, (int)d, mean_f32, T, (int)dout, out);
    LAUNCH_CHECK();
}

// ================================================================================================ rmse numerator
// f64_diff == 0: Rust semantics (f32 difference and square, src/embedding.rs:173-174); == 1: numpy semantics
// (difference and square in f64, pycleora/__init__.py:975-976).  The SUM is f64 and tree-shaped in both cases
// (the reference's serial f32 running sum cannot be reproduced in parallel).
__global__ void __launch_bounds__(256) sq_diff_stage1(const float *__restrict__ a, const float *__restrict__ b,
                                                      int64_t n, int f64_diff, double *__restrict__ partial) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (f64_diff) {
            const double dl = (double)a[i] - (double)b[i];
            acc += dl * dl;
        } else {
            const float dl = __fsub_rn(a[i], b[i]);
            acc += (double)__fmul_rn(dl, dl);
        }
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FULL, acc, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += sh[w];
        partial[blockIdx.x] = s;
    }
}
__global__ void sq_diff_stage2(const double *__restrict__ partial, int nb, double *__restrict__ result) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nb; ++i) s += partial[i];
        result[0] = s;
    }
}

void launch_sq_diff_sum(const float *a, const float *b, int64_t n, bool f64_diff, double *result, cudaStream_t st) {
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 148 * 8));
    double *partial = (double *)workspace().sqdiff_partials.get(size_t(blocks) * sizeof(double));
    sq_diff_stage1<<<blocks, 256, 0, st>>>(a, b, n, f64_diff ? 1 : 0, partial);
    LAUNCH_CHECK();
    sq_diff_stage2<<<1, 32, 0, st>>>(partial, blocks, result);
    LAUNCH_CHECK();
}

// ================================================================================================ PCA transform
// T[i, k] = V[i, d-1-k] / sqrt(max(w[d-1-k], 1e-10)) as f32 -- pycleora/__init__.py:147-156 (eigh returns ascending
// eigenvalues; the reference re-orders descending).  V is column-major (cuSOLVER): V[i + c*d].
// scaled == 0: the bare rotation V (descending order) -- normalization="spectral", pycleora/__init__.py:951-956.
__global__ void build_transform_kernel(const double *__restrict__ V, const double *__restrict__ w, int d, int dout,
                                       float *__restrict__ T, int scaled) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)d * dout) return;
    const int i = (int)(idx / dout), k = (int)(idx - (int64_t)i * dout);
    const int src = d - 1 - k;
    const double scale = scaled ? 1.0 / sqrt(fmax(w[src], 1e-10)) : 1.0;
    T[idx] = (float)(V[(int64_t)src * d + i] * scale);
}
void launch_build_transform(const double *V, const double *w, int64_t d, int64_t dout, float *T, cudaStream_t st, bool scaled) {
    const int64_t tot = d * dout;
    if (tot == 0) return;
    build_transform_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(V, w, (int)d, (int)dout, T, scaled ? 1 : 0);
    LAUNCH_CHECK();
}

// ================================================================================================ A * 1
// rowscale[r] = sum of the row's Markov values in stored order (f32) -- the vector A*1 of the pipelined loop.
__global__ void row_value_sums_kernel(const int64_t *__restrict__ rowptr, const float *__restrict__ val, int64_t n,
                                      float *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float s = 0.f;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s = __fadd_rn(s, val[k]);
    out[r] = s;
}
void launch_row_value_sums(const int64_t *rowptr, const float *val, int64_t n, float *out, cudaStream_t st) {
    if (n == 0) return;
    row_value_sums_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowptr, val, n, out);
    LAUNCH_CHECK();
}

// ================================================================================================ small helpers
__global__ void scale_f64_kernel(double *v, int64_t n, double f) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= f;
}
void launch_scale_f64(double *v, int64_t n, double factor, cudaStream_t st) {
    if (n == 0) return;
    scale_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(v, n, factor);
    LAUNCH_CHECK();
}
__global__ void f64_to_f32_kernel(const double *in, float *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
void launch_f64_to_f32(const double *in, float *out, int64_t n, cudaStream_t st) {
    if (n == 0) return;
    f64_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
    LAUNCH_CHECK();
}

}  // namespace cleora
