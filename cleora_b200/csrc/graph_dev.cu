// Device-side integer ingest: the CSR that cleora_graph_from_pairs / from_iterator("u v" lines,
// "complex::reflexive::<name>") builds, constructed on the GPU from device-resident pair arrays -- HBM-bound integer
// work (histograms, one radix sort, run-length reduce), with the reference's semantics kept bit for bit:
//   * entity index = order of first appearance in the stream u0 v0 u1 v1 ...   (src/sparse_matrix_builder.rs:59-70)
//   * per line (A = B = [u, v], value = 1/4, :170-233): u != v adds 1/2 to M[u,v], M[v,u], M[u,u], M[v,v] and 1 to
//     both row sums; u == v adds 2 to M[u,u] and 2 to row_sum[u]; duplicates merge (all sums are exact in f32)
//   * rows sorted by column, left = M/row_sum[r], sym = M/sqrt(row_sum[r] row_sum[c])                  (:292-332)
//   * entity hash = XXH64(decimal string of the id, seed 0)                                  (src/entity.rs:109-114)
// A row SHARD can be built directly: every rank passes the same pair arrays (the synthetic generators below are
// counter-based, so each rank regenerates them locally), rows are split into `world` contiguous blocks balanced by
// entry count, and rank g keeps rows [bounds[g], bounds[g+1]) with column indices remapped to the padded gathered
// layout of cleora_b200/sharded.py (owner * block + offset).  No host copy of the CSR is made (the 1.5 B-edge config
// has 24 GB of it); host accessors download on demand.
// The sort and the scans are CUB device primitives (part of the CUDA toolkit, like cuSOLVER for the d x d eigensolve);
// every kernel that carries the reference's semantics is written here.
#include "device.cuh"
#include "graph.hpp"
#include "../../include/cleora_b200.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

namespace cleora {
namespace gd {

constexpr unsigned long long NEVER = ~0ull;

template <class T>
struct Buf {                                 // RAII device array
    T *p = nullptr;
    size_t n = 0;
    Buf() = default;
    explicit Buf(size_t count) { alloc(count); }
    Buf(const Buf &) = delete;
    Buf &operator=(const Buf &) = delete;
    void alloc(size_t count) {
        release();
        n = count;
        CUDA_TRY(cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T)));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    T *take() { T *q = p; p = nullptr; n = 0; return q; }
    ~Buf() { release(); }
};

__global__ void first_pos_kernel(const uint32_t *__restrict__ u, const uint32_t *__restrict__ v, int64_t n_pairs,
                                 unsigned long long *__restrict__ first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        atomicMin(first + u[i], (unsigned long long)(2 * i));
        atomicMin(first + v[i], (unsigned long long)(2 * i + 1));
    }
}
__global__ void iota_kernel(uint32_t *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)i;
}
// sorted by first position: entry k is the k-th entity to appear; entries with NEVER (ids that do not occur) sort last
__global__ void label_kernel(const uint32_t *__restrict__ orig_sorted, const unsigned long long *__restrict__ first_sorted,
                             int64_t n_ids, uint32_t *__restrict__ label, unsigned long long *__restrict__ n_entities) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_ids) return;
    const bool used = first_sorted[k] != NEVER;
    label[orig_sorted[k]] = used ? (uint32_t)k : 0xFFFFFFFFu;
    if (used && (k + 1 == n_ids || first_sorted[k + 1] == NEVER)) *n_entities = (unsigned long long)(k + 1);
}
// ne[a] = pairs (a, b != a) that contain a;  self[a] = pairs (a, a)
__global__ void degree_kernel(const uint32_t *__restrict__ u, const uint32_t *__restrict__ v, int64_t n_pairs,
                              const uint32_t *__restrict__ label, uint32_t *__restrict__ ne, uint32_t *__restrict__ self) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t a = label[u[i]], b = label[v[i]];
        if (a == b) atomicAdd(self + a, 1u);
        else { atomicAdd(ne + a, 1u); atomicAdd(ne + b, 1u); }
    }
}
// weight of a row for the partition = its entries before merging (off-diagonal ones + the diagonal)
__global__ void weight_kernel(const uint32_t *__restrict__ ne, int64_t n, unsigned long long *__restrict__ w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = (unsigned long long)ne[i] + 1ull;
}
// bounds[g] = first row whose exclusive prefix weight reaches total * g / world (g = 1 .. world-1)
__global__ void bounds_kernel(const unsigned long long *__restrict__ prefix /* exclusive, n + 1 */, int64_t n, int world,
                              long long *__restrict__ bounds) {
    const int g = threadIdx.x;
    if (g > world) return;
    if (g == 0) { bounds[0] = 0; return; }
    if (g == world) { bounds[world] = n; return; }
    const unsigned long long total = prefix[n];
    const unsigned long long target = (unsigned long long)(((unsigned __int128)total * (unsigned)g) / (unsigned)world);
    int64_t lo = 0, hi = n;                                   // lower_bound over prefix[0..n]
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (prefix[mid] < target) lo = mid + 1; else hi = mid;
    }
    bounds[g] = lo;
}
// keys of the rows [r0, r1): (row - r0) << col_bits | col; off-diagonal entries per pair, positions from a
// warp-aggregated counter (the order is irrelevant: the keys are sorted next)
__global__ void emit_kernel(const uint32_t *__restrict__ u, const uint32_t *__restrict__ v, int64_t n_pairs,
                            const uint32_t *__restrict__ label, int64_t r0, int64_t r1, int col_bits,
                            unsigned long long *__restrict__ keys, unsigned long long *__restrict__ counter) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (n_pairs + stride - 1) / stride;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t it = 0; it < rounds; ++it, i += stride) {
        int64_t a = -1, b = -1;
        if (i < n_pairs) { a = label[u[i]]; b = label[v[i]]; }
        const bool ea = a != b && a >= r0 && a < r1, eb = a != b && b >= r0 && b < r1;
        const unsigned mine = (unsigned)ea + (unsigned)eb;
        // warp-aggregated slot allocation
        unsigned incl = mine;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, off);
            if ((threadIdx.x & 31) >= off) incl += t;
        }
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long base = 0;
        if ((threadIdx.x & 31) == 31 && total) base = atomicAdd(counter, (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned long long pos = base + incl - mine;
        if (ea) keys[pos++] = ((unsigned long long)(a - r0) << col_bits) | (unsigned long long)b;
        if (eb) keys[pos] = ((unsigned long long)(b - r0) << col_bits) | (unsigned long long)a;
    }
}
__global__ void diag_keys_kernel(int64_t r0, int64_t n_local, int col_bits, unsigned long long *__restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_local) keys[i] = ((unsigned long long)i << col_bits) | (unsigned long long)(r0 + i);
}
__device__ __forceinline__ int owner_of(const long long *bounds, int world, int64_t x) {
    int g = 0;
    while (g + 1 < world && bounds[g + 1] <= x) ++g;
    return g;
}
// Merging duplicates = run-length reduction of the sorted keys, done here in two passes over blocks of DCH keys so
// that nothing of the size of the key array has to be materialised (cub::DeviceRunLengthEncode is also limited to
// 2^31 items): pass A counts the run heads per block; after a scan of the block counts pass B writes, for every
// head, the final column index and Markov values straight into the CSR arrays and the row offsets.
constexpr int DCH_THREADS = 256, DCH_PER_THREAD = 8, DCH = DCH_THREADS * DCH_PER_THREAD;

__global__ void __launch_bounds__(DCH_THREADS) head_count_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                                unsigned long long *__restrict__ block_heads) {
    using BlockReduce = cub::BlockReduce<unsigned, DCH_THREADS>;
    __shared__ typename BlockReduce::TempStorage tmp;
    const int64_t base = (int64_t)blockIdx.x * DCH;
    unsigned heads = 0;
#pragma unroll
    for (int k = 0; k < DCH_PER_THREAD; ++k) {
        const int64_t i = base + k * DCH_THREADS + threadIdx.x;
        if (i < n) heads += (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
    }
    const unsigned total = BlockReduce(tmp).Sum(heads);
    if (threadIdx.x == 0) block_heads[blockIdx.x] = total;
}

// values and final column indices.  M[r,c] = count/2 off the diagonal, ne/2 + 2 self on it.
__global__ void __launch_bounds__(DCH_THREADS)
compact_values_kernel(const unsigned long long *__restrict__ keys, int64_t n, const unsigned long long *__restrict__ block_offset,
                      int col_bits, int64_t r0, int64_t n_local, int64_t nnz, const uint32_t *__restrict__ ne,
                      const uint32_t *__restrict__ self, const long long *__restrict__ bounds, int world, int64_t block,
                      long long *__restrict__ rowptr, uint32_t *__restrict__ col, float *__restrict__ left,
                      float *__restrict__ sym) {
    using BlockScan = cub::BlockScan<unsigned, DCH_THREADS>;
    __shared__ typename BlockScan::TempStorage tmp;
    const int64_t first = (int64_t)blockIdx.x * DCH + (int64_t)threadIdx.x * DCH_PER_THREAD;     // 8 consecutive keys per thread
    unsigned long long k[DCH_PER_THREAD];
    unsigned long long prev = 0;
    unsigned flags = 0, cnt = 0;
    if (first < n && first > 0) prev = keys[first - 1];
#pragma unroll
    for (int q = 0; q < DCH_PER_THREAD; ++q) {
        const int64_t i = first + q;
        k[q] = i < n ? keys[i] : 0ull;
        const bool head = i < n && (i == 0 || k[q] != (q ? k[q - 1] : prev));
        if (head) { flags |= 1u << q; ++cnt; }
    }
    unsigned excl;
    BlockScan(tmp).ExclusiveSum(cnt, excl);
    int64_t pos = (int64_t)block_offset[blockIdx.x] + excl;
#pragma unroll
    for (int q = 0; q < DCH_PER_THREAD; ++q) {
        if (!(flags & (1u << q))) continue;
        const int64_t i = first + q;
        const unsigned long long key = k[q];
        int64_t j = i + 1;                                      // run end: duplicates are rare, hubs gallop
        int64_t step = 1;
        while (j < n && keys[j] == key) { j += step; step <<= 1; }
        if (step > 1) {                                         // overshoot possible: binary search in (j - step/2, min(j, n)]
            int64_t lo = j - (step >> 1), hi = j < n ? j : n;   // keys[lo - 1] == key (or lo == i + 1); first index with keys != key
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (keys[mid] == key) lo = mid + 1; else hi = mid;
            }
            j = lo;
        }
        const unsigned count = (unsigned)(j - i);
        const int64_t rl = (int64_t)(key >> col_bits), r = rl + r0, c = (int64_t)(key & ((1ull << col_bits) - 1ull));
        const float rs = (float)ne[r] + 2.0f * (float)self[r], cs = (float)ne[c] + 2.0f * (float)self[c];   // exact integers
        const float m = (r == c) ? __fadd_rn(__fmul_rn(0.5f, (float)ne[r]), __fmul_rn(2.0f, (float)self[r]))
                                 : __fmul_rn(0.5f, (float)count);
        left[pos] = __fdiv_rn(m, rs);                                                     // sparse_matrix_builder.rs:328
        if (sym != nullptr) sym[pos] = __fdiv_rn(m, __fsqrt_rn(__fmul_rn(rs, cs)));       // :325-329
        int64_t cc = c;
        if (world > 1) { const int g = owner_of(bounds, world, c); cc = (int64_t)g * block + (c - bounds[g]); }
        col[pos] = (uint32_t)cc;
        const int64_t prev_row = i > 0 ? (int64_t)((q ? k[q - 1] : prev) >> col_bits) : -1;   // row of the previous key
        for (int64_t rr = prev_row + 1; rr <= rl; ++rr) rowptr[rr] = pos;
        ++pos;
    }
    if (first + DCH_PER_THREAD >= n && first < n) {            // the thread that holds the last key closes the offsets
        const int64_t last_row = (int64_t)(keys[n - 1] >> col_bits);
        for (int64_t rr = last_row + 1; rr <= n_local; ++rr) rowptr[rr] = nnz;
    }
}
__global__ void row_sum_kernel(const uint32_t *__restrict__ ne, const uint32_t *__restrict__ self, int64_t n, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)ne[i] + 2.0f * (float)self[i];
}

// XXH64 (seed 0) of a short byte string (< 32 bytes): the tail-only path of the public algorithm.
__device__ uint64_t xxh64_short(const unsigned char *p, int len) {
    const uint64_t P1 = 11400714785074694791ULL, P2 = 14029467366897019727ULL, P3 = 1609587929392839161ULL,
                   P4 = 9650029242287828579ULL, P5 = 2870177450012600261ULL;
    auto rotl = [](uint64_t x, int r) { return (x << r) | (x >> (64 - r)); };
    uint64_t h = P5 + (uint64_t)len;
    int i = 0;
    for (; i + 8 <= len; i += 8) {
        uint64_t k = 0;
        for (int b = 0; b < 8; ++b) k |= (uint64_t)p[i + b] << (8 * b);
        k *= P2; k = rotl(k, 31); k *= P1;
        h ^= k; h = rotl(h, 27) * P1 + P4;
    }
    if (i + 4 <= len) {
        uint64_t k = 0;
        for (int b = 0; b < 4; ++b) k |= (uint64_t)p[i + b] << (8 * b);
        h ^= k * P1; h = rotl(h, 23) * P2 + P3;
        i += 4;
    }
    for (; i < len; ++i) { h ^= (uint64_t)p[i] * P5; h = rotl(h, 11) * P1; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}
// hash[k] = XXH64(decimal(orig[k])); also scattered to the padded gathered layout
__global__ void hash_kernel(const uint32_t *__restrict__ orig, int64_t n, const long long *__restrict__ bounds, int world,
                            int64_t block, uint64_t *__restrict__ hash, uint64_t *__restrict__ hash_padded) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned char dec[12];
    uint32_t x = orig[k];
    int len = 0;
    do { dec[len++] = (unsigned char)('0' + x % 10); x /= 10; } while (x);
    for (int a = 0, b = len - 1; a < b; ++a, --b) { const unsigned char t = dec[a]; dec[a] = dec[b]; dec[b] = t; }
    const uint64_t h = xxh64_short(dec, len);
    hash[k] = h;
    if (hash_padded != nullptr) {
        const int g = owner_of(bounds, world, k);
        hash_padded[(int64_t)g * block + (k - bounds[g])] = h;
    }
}

// ---- synthetic pair generators (counter-based: pair i depends only on (seed, i), so every rank can regenerate the
// same stream locally).  kind 0: endpoints i.i.d. uniform over [0, n).  kind 1: Chung-Lu with weights (i + i0)^-alpha
// (continuous inverse CDF), node ids decoupled from the weight rank by a fixed pseudo-random permutation.
__device__ __forceinline__ uint64_t splitmix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }
// bijection of [0, n): 4-round Feistel over the next even bit width with cycle walking
__device__ uint32_t permute(uint32_t x, uint32_t n, int half_bits, uint64_t seed) {
    const uint32_t mask = (1u << half_bits) - 1u;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (int round = 0; round < 4; ++round) {
            const uint32_t f = (uint32_t)splitmix(seed ^ ((uint64_t)round << 32) ^ r) & mask;
            const uint32_t t = l ^ f;
            l = r; r = t;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}
__global__ void synth_pairs_kernel(int kind, uint32_t n, int64_t n_pairs, uint64_t seed, double alpha, double i0,
                                   uint32_t *__restrict__ u, uint32_t *__restrict__ v) {
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < (unsigned long long)n) ++half_bits;
    const double e1 = 1.0 - alpha, A = pow(i0, e1), B = pow((double)n + i0, e1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t ab[2];
        uint64_t ctr = 2 * (uint64_t)i;
        for (;;) {
            for (int s = 0; s < 2; ++s) {
                const double r = u01(splitmix(seed * 0x9E3779B97F4A7C15ULL + ctr + s));
                uint32_t x;
                if (kind == 0) x = min((uint32_t)(r * (double)n), n - 1);
                else {
                    const double t = pow(A + r * (B - A), 1.0 / e1) - i0;        // inverse CDF of (x + i0)^-alpha
                    x = permute(min((uint32_t)fmax(t, 0.0), n - 1), n, half_bits, seed);
                }
                ab[s] = x;
            }
            if (ab[0] != ab[1]) break;                                            // u != v, as bench.py's generators
            ctr += 2 * (uint64_t)n_pairs;                                         // redraw from a disjoint counter range
        }
        u[i] = ab[0];
        v[i] = ab[1];
    }
}

inline unsigned grid_for(int64_t n, int threads = 256) {
    return (unsigned)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
}
inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads); }
inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 64 && (max_value >> b)) ++b;
    return b;
}

}  // namespace gd

void synth_pairs_device(int kind, int64_t n_nodes, int64_t n_pairs, uint64_t seed, double alpha, uint32_t *u, uint32_t *v,
                        cudaStream_t st) {
    if (n_nodes < 2 || n_nodes > 0xFFFFFFFFll) throw BuildError{"synthetic generator needs 2 <= nodes < 2^32"};
    if (kind != 0 && kind != 1) throw BuildError{"unknown synthetic graph kind"};
    if (kind == 1 && !(alpha > 0.0 && alpha < 1.0)) throw BuildError{"Chung-Lu weight exponent must be in (0, 1)"};
    if (n_pairs == 0) return;
    gd::synth_pairs_kernel<<<gd::grid_for(n_pairs), 256, 0, st>>>(kind, (uint32_t)n_nodes, n_pairs, seed, alpha, 10.0, u, v);
    LAUNCH_CHECK();
}

// Build the (shard of the) graph.  Everything is enqueued on `st`; the few scalar read-backs synchronise it.
std::unique_ptr<Graph> build_from_pairs_device(const uint32_t *u, const uint32_t *v, int64_t n_pairs,
                                               const std::string &column_name, int rank, int world, bool want_sym,
                                               cudaStream_t st, std::vector<int64_t> *bounds_out) {
    using namespace gd;
    if (world < 1 || rank < 0 || rank >= world || world > 64) throw BuildError{"bad shard rank / world"};
    if (n_pairs < 0) throw BuildError{"negative pair count"};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    auto g = std::make_unique<Graph>();
    g->desc = Descriptor{0, 1, column_name, column_name};
    auto dg = std::make_unique<DeviceGraph>();
    dg->device = dev;

    // ---- ids and labels
    uint32_t max_id = 0;
    size_t tmp_bytes = 0;
    Buf<unsigned char> tmp;
    auto ensure_tmp = [&](size_t bytes) { if (bytes > tmp.n) tmp.alloc(bytes); };
    if (n_pairs) {
        Buf<uint32_t> d_max(2);
        for (int s = 0; s < 2; ++s) {
            const uint32_t *src = s ? v : u;
            CUDA_TRY(cub::DeviceReduce::Max(nullptr, tmp_bytes, src, d_max.p + s, n_pairs, st));
            ensure_tmp(tmp_bytes);
            CUDA_TRY(cub::DeviceReduce::Max(tmp.p, tmp_bytes, src, d_max.p + s, n_pairs, st));
        }
        uint32_t h[2];
        CUDA_TRY(cudaMemcpyAsync(h, d_max.p, sizeof h, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        max_id = std::max(h[0], h[1]);
    }
    const int64_t n_ids = n_pairs ? (int64_t)max_id + 1 : 0;
    if (n_ids > (int64_t)1 << 31) throw BuildError{"device ingest needs ids below 2^31 (dense id space); use the host builder"};
    Buf<uint32_t> label((size_t)n_ids), orig((size_t)n_ids);
    int64_t n = 0;
    if (n_ids) {
        Buf<unsigned long long> first((size_t)n_ids), first_sorted((size_t)n_ids), d_n(1);
        Buf<uint32_t> ids((size_t)n_ids);
        CUDA_TRY(cudaMemsetAsync(first.p, 0xFF, sizeof(unsigned long long) * n_ids, st));
        CUDA_TRY(cudaMemsetAsync(d_n.p, 0, sizeof(unsigned long long), st));
        first_pos_kernel<<<grid_for(n_pairs), 256, 0, st>>>(u, v, n_pairs, first.p);
        LAUNCH_CHECK();
        iota_kernel<<<blocks_for(n_ids), 256, 0, st>>>(ids.p, n_ids);
        LAUNCH_CHECK();
        const int end_bit = std::min(64, bits_for(2 * (uint64_t)n_pairs + 1));     // NEVER has every bit set: sorts last within end_bit too
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, first.p, first_sorted.p, ids.p, orig.p, n_ids, 0, 64, st));
        ensure_tmp(tmp_bytes);
        (void)end_bit;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, first.p, first_sorted.p, ids.p, orig.p, n_ids, 0, 64, st));
        label_kernel<<<blocks_for(n_ids), 256, 0, st>>>(orig.p, first_sorted.p, n_ids, label.p, d_n.p);
        LAUNCH_CHECK();
        unsigned long long hn = 0;
        CUDA_TRY(cudaMemcpyAsync(&hn, d_n.p, sizeof hn, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n = (int64_t)hn;
    }

    // ---- degrees, partition
    Buf<uint32_t> ne((size_t)n), self((size_t)n);
    CUDA_TRY(cudaMemsetAsync(ne.p, 0, sizeof(uint32_t) * std::max<int64_t>(n, 1), st));
    CUDA_TRY(cudaMemsetAsync(self.p, 0, sizeof(uint32_t) * std::max<int64_t>(n, 1), st));
    if (n_pairs) { degree_kernel<<<grid_for(n_pairs), 256, 0, st>>>(u, v, n_pairs, label.p, ne.p, self.p); LAUNCH_CHECK(); }
    std::vector<int64_t> bounds((size_t)world + 1, 0);
    Buf<long long> d_bounds((size_t)world + 1);
    {
        Buf<unsigned long long> w((size_t)n + 1), prefix((size_t)n + 1);
        CUDA_TRY(cudaMemsetAsync(w.p, 0, sizeof(unsigned long long) * ((size_t)n + 1), st));
        if (n) { weight_kernel<<<blocks_for(n), 256, 0, st>>>(ne.p, n, w.p); LAUNCH_CHECK(); }
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, w.p, prefix.p, n + 1, st));
        ensure_tmp(tmp_bytes);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, w.p, prefix.p, n + 1, st));
        bounds_kernel<<<1, 128, 0, st>>>(prefix.p, n, world, d_bounds.p);
        LAUNCH_CHECK();
        std::vector<long long> hb((size_t)world + 1);
        CUDA_TRY(cudaMemcpyAsync(hb.data(), d_bounds.p, sizeof(long long) * hb.size(), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        for (int gq = 0; gq <= world; ++gq) bounds[(size_t)gq] = hb[(size_t)gq];
        for (int gq = 1; gq <= world; ++gq) bounds[(size_t)gq] = std::max(bounds[(size_t)gq], bounds[(size_t)gq - 1]);
        CUDA_TRY(cudaMemcpyAsync(d_bounds.p, bounds.data(), sizeof(long long) * hb.size(), cudaMemcpyHostToDevice, st));
    }
    int64_t block = 0;
    for (int gq = 0; gq < world; ++gq) block = std::max(block, bounds[(size_t)gq + 1] - bounds[(size_t)gq]);
    const int64_t r0 = bounds[(size_t)rank], r1 = bounds[(size_t)rank + 1], n_local = r1 - r0;
    const int64_t n_pad = world > 1 ? block * world : n;
    if (n_pad > 0xFFFFFFFFll) throw BuildError{"too many entities for 32-bit column indices"};

    // ---- entries of the local rows: count, emit, sort, merge
    int64_t n_entries = n_local;                                        // the diagonal
    if (n_local) {
        Buf<unsigned long long> d_sum(1);
        Buf<unsigned long long> w((size_t)n_local);
        weight_kernel<<<blocks_for(n_local), 256, 0, st>>>(ne.p + r0, n_local, w.p);
        LAUNCH_CHECK();
        CUDA_TRY(cub::DeviceReduce::Sum(nullptr, tmp_bytes, w.p, d_sum.p, n_local, st));
        ensure_tmp(tmp_bytes);
        CUDA_TRY(cub::DeviceReduce::Sum(tmp.p, tmp_bytes, w.p, d_sum.p, n_local, st));
        unsigned long long hs = 0;
        CUDA_TRY(cudaMemcpyAsync(&hs, d_sum.p, sizeof hs, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        n_entries = (int64_t)hs;                                        // sum of (ne + 1) over the local rows
    }
    const int col_bits = bits_for((uint64_t)std::max<int64_t>(n, 2) - 1), row_bits = bits_for((uint64_t)std::max<int64_t>(n_local, 2) - 1);
    int64_t nnz = 0;
    const size_t pad = 16;                                             // trailing padding, as the host upload path provides
    Buf<long long> rowptr((size_t)n_local + 1 + pad);
    Buf<uint32_t> col;
    Buf<float> left, sym;
    CUDA_TRY(cudaMemsetAsync(rowptr.p, 0, sizeof(long long) * ((size_t)n_local + 1 + pad), st));
    if (n_entries) {
        Buf<unsigned long long> keys((size_t)n_entries), keys_sorted((size_t)n_entries), counter(1);
        diag_keys_kernel<<<blocks_for(n_local), 256, 0, st>>>(r0, n_local, col_bits, keys.p);
        LAUNCH_CHECK();
        const unsigned long long start = (unsigned long long)n_local;
        CUDA_TRY(cudaMemcpyAsync(counter.p, &start, sizeof start, cudaMemcpyHostToDevice, st));
        if (n_pairs) {
            emit_kernel<<<grid_for(n_pairs), 256, 0, st>>>(u, v, n_pairs, label.p, r0, r1, col_bits, keys.p, counter.p);
            LAUNCH_CHECK();
        }
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys.p, keys_sorted.p, n_entries, 0, col_bits + row_bits, st));
        ensure_tmp(tmp_bytes);
        CUDA_TRY(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, keys.p, keys_sorted.p, n_entries, 0, col_bits + row_bits, st));
        keys.release();
        const int64_t n_blocks = (n_entries + DCH - 1) / DCH;
        Buf<unsigned long long> heads((size_t)n_blocks + 1), offs((size_t)n_blocks + 1);
        CUDA_TRY(cudaMemsetAsync(heads.p, 0, sizeof(unsigned long long) * ((size_t)n_blocks + 1), st));
        head_count_kernel<<<(unsigned)n_blocks, DCH_THREADS, 0, st>>>(keys_sorted.p, n_entries, heads.p);
        LAUNCH_CHECK();
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, heads.p, offs.p, n_blocks + 1, st));
        ensure_tmp(tmp_bytes);
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, heads.p, offs.p, n_blocks + 1, st));
        unsigned long long hn = 0;
        CUDA_TRY(cudaMemcpyAsync(&hn, offs.p + n_blocks, sizeof hn, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        nnz = (int64_t)hn;
        col.alloc((size_t)nnz + pad);
        left.alloc((size_t)nnz + pad);
        CUDA_TRY(cudaMemsetAsync(col.p + nnz, 0, sizeof(uint32_t) * pad, st));
        CUDA_TRY(cudaMemsetAsync(left.p + nnz, 0, sizeof(float) * pad, st));
        if (want_sym) { sym.alloc((size_t)nnz + pad); CUDA_TRY(cudaMemsetAsync(sym.p + nnz, 0, sizeof(float) * pad, st)); }
        compact_values_kernel<<<(unsigned)n_blocks, DCH_THREADS, 0, st>>>(keys_sorted.p, n_entries, offs.p, col_bits, r0, n_local, nnz,
                                                                        ne.p, self.p, d_bounds.p, world, block, rowptr.p, col.p,
                                                                        left.p, want_sym ? sym.p : nullptr);
        LAUNCH_CHECK();
        CUDA_TRY(cudaStreamSynchronize(st));                          // keys_sorted is freed at the end of this scope
    } else {
        col.alloc(pad); left.alloc(pad);
        CUDA_TRY(cudaMemsetAsync(col.p, 0, sizeof(uint32_t) * pad, st));
        CUDA_TRY(cudaMemsetAsync(left.p, 0, sizeof(float) * pad, st));
        if (want_sym) { sym.alloc(pad); CUDA_TRY(cudaMemsetAsync(sym.p, 0, sizeof(float) * pad, st)); }
    }

    Buf<float> row_sum((size_t)n);
    Buf<uint64_t> hash((size_t)n), hash_padded;
    if (n) {
        row_sum_kernel<<<blocks_for(n), 256, 0, st>>>(ne.p, self.p, n, row_sum.p);
        LAUNCH_CHECK();
        if (world > 1) { hash_padded.alloc((size_t)n_pad); CUDA_TRY(cudaMemsetAsync(hash_padded.p, 0, sizeof(uint64_t) * n_pad, st)); }
        hash_kernel<<<blocks_for(n), 256, 0, st>>>(orig.p, n, d_bounds.p, world, block, hash.p, world > 1 ? hash_padded.p : nullptr);
        LAUNCH_CHECK();
    }
    // host keeps the row offsets (small; the long-row schedule and the Python side need them)
    g->rowptr.assign((size_t)n_local + 1, 0);
    CUDA_TRY(cudaMemcpyAsync(g->rowptr.data(), rowptr.p, sizeof(int64_t) * ((size_t)n_local + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));

    g->n_rows = n_local;
    g->n_cols = world > 1 ? std::max<int64_t>(n_pad, 1) : n;
    g->row_offset = world > 1 ? rank * block : 0;
    g->device_only = true;
    g->nnz_device = nnz;
    g->n_global = n;
    g->shard_r0 = r0;
    dg->n_rows = g->n_rows; dg->n_cols = g->n_cols; dg->nnz = nnz; dg->row_offset = g->row_offset;
    dg->rowptr = (int64_t *)rowptr.take();
    dg->col = col.take();
    dg->left = left.take();
    dg->sym = want_sym ? sym.take() : nullptr;
    dg->hash = world > 1 ? hash_padded.take() : hash.take();             // what init reads: one hash per row of the gathered matrix
    dg->hash_rows = world > 1 ? n_pad : n;
    dg->row_sum_all = row_sum.take();
    dg->orig_ids = orig.take();
    dg->n_global = n;
    g->devs.push_back(dg.release());
    if (bounds_out) *bounds_out = bounds;
    return g;
}

}  // namespace cleora
