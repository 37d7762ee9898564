// Internal device-side declarations of libcleora_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <climits>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace cleora {

// Device image of a Graph: SoA CSR (8 B/nnz streamed per SpMM instead of the reference's 12 B AoS Edge,
// src/sparse_matrix.rs:73-78) with 64-bit row offsets (nnz of the 1.5 B-edge config exceeds 2^31).
struct DeviceGraph {
    int device = -1;
    int64_t n_rows = 0, n_cols = 0, nnz = 0, row_offset = 0;
    int64_t *rowptr = nullptr;
    uint32_t *col = nullptr;
    float *left = nullptr, *sym = nullptr;
    const float *host_sym = nullptr;       // symmetric values are uploaded on first use (the default loop never reads them)
    uint64_t *hash = nullptr;
    // long-row schedule: rows with more than long_threshold edges are split into chunks of long_chunk_edges, one warp
    // per chunk (kernels.cu); built at upload time
    // long_threshold applies to full-width rows (32 lanes per row); narrower kernels scale it down to
    // long_sched_threshold at the least, which is the degree above which a row enters the schedule at all
    int64_t n_long = 0, n_long_chunks = 0, long_threshold = INT64_MAX, long_chunk_edges = 0;
    int64_t long_sched_threshold = INT64_MAX, max_degree = 0;
    int64_t *long_rows = nullptr;          // [n_long] row indices
    int64_t *long_chunk_ptr = nullptr;     // [n_long + 1] first chunk of each long row
    int32_t *long_chunk_owner = nullptr;   // [n_long_chunks] index into long_rows
    uint32_t *row_order = nullptr;         // rows by descending degree (skewed graphs only): schedule of the narrow-row kernels
    float *rsum_left = nullptr, *rsum_sym = nullptr;   // A*1 per Markov type, built on first pipelined use
    // device-built graphs (graph_dev.cu): row sums of ALL entities, original integer ids by entity index, entity count
    // of the whole graph, and the number of entries in `hash` (n, or the padded row count of the gathered layout)
    float *row_sum_all = nullptr;
    uint32_t *orig_ids = nullptr;
    int64_t n_global = 0, hash_rows = 0;
    std::mutex lazy_mu;                    // guards the lazily built members (sym, rsum_*)
};

// Where a row-producing kernel (K1, K3, the row normaliser) stores its rows besides / instead of `out`.  The pointers
// may be other GPUs' buffers mapped through CUDA IPC (peer stores over NVLink): the collective is fused into the
// producer's epilogue.
//   mode 0  REPLICATE  the same full row also goes to extra[0 .. n_extra)            (all-gather of row blocks)
//   mode 1  SLICES     columns [h*slice_cols, (h+1)*slice_cols) of row r go to extra[h] at row (row_base + r) of an
//                      [rows x slice_cols] matrix, h = 0 .. n_extra-1                 (row-sharded -> column-sharded)
//   mode 2  OWNERS     the produced rows have `ld_cols`-wide destinations: row r goes to extra[r / block_rows] at row
//                      (r % block_rows), columns [col_off, col_off + d); `out` is unused (column-sharded -> row-sharded)
struct PeerOut {
    float *extra[8];
    int n_extra;
    int mode;
    int slice_cols;
    int col_off;
    int64_t row_base;
    int64_t block_rows;
    int64_t ld_cols;
};
enum { PEER_REPLICATE = 0, PEER_SLICES = 1, PEER_OWNERS = 2 };

void set_error(const std::string &msg);
extern std::atomic<int64_t> g_launches;

struct CudaFail {
    std::string msg;
};

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw ::cleora::CudaFail{std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                     ":" + std::to_string(__LINE__) + ")"};                              \
    } while (0)

#define LAUNCH_CHECK()                             \
    do {                                           \
        ::cleora::g_launches.fetch_add(1);         \
        CUDA_TRY(cudaGetLastError());              \
    } while (0)

// Growable device scratch, one per host thread and purpose.
struct Scratch {
    void *p = nullptr;
    size_t cap = 0;
    int device = -1;
    void *get(size_t bytes);
    void release();
};
struct Workspace {
    Scratch colsum_partials, gram_partials, sqdiff_partials, misc, spmm_partials, absmax_partials, chol;
    size_t bytes() const {
        return colsum_partials.cap + gram_partials.cap + sqdiff_partials.cap + misc.cap + spmm_partials.cap +
               absmax_partials.cap + chol.cap;
    }
    void release() {
        colsum_partials.release(); gram_partials.release(); sqdiff_partials.release(); misc.release();
        spmm_partials.release(); absmax_partials.release(); chol.release();
    }
};
Workspace &workspace();

// ---- launchers (kernels.cu); all enqueue on `st` and throw CudaFail on launch errors -------------------
void launch_init(const uint64_t *hash, int64_t n, int64_t d, int64_t seed, float *out, cudaStream_t st);
void launch_spmm(const DeviceGraph &g, const float *val, const float *x, int64_t d, float *out, const float *resid,
                 float alpha, float rw, int norm, cudaStream_t st, const PeerOut *peers = nullptr);
void launch_normalize(const float *x, int64_t n, int64_t d, int norm, float *out, cudaStream_t st);
// K1's fused row norm applied to existing rows, with K1's destinations (kernels.cu: normalize_rows_kernel)
bool normalize_rows_supported(int64_t d);
void launch_normalize_rows(const float *x, int64_t n, int64_t d, int norm, float *out, cudaStream_t st, const PeerOut *peers);
// Per-block maxima of |x| (device), a by-product of the column-sum pass that the integer Gram kernel needs for its
// fixed-point scale; count == 0 means "not produced" (the Gram launcher then makes its own pass).
struct AbsmaxPartials {
    const float *p = nullptr;
    int count = 0;
};
void launch_col_sums(const float *x, int64_t n, int64_t d, double *sums, bool accumulate, cudaStream_t st,
                     AbsmaxPartials *absmax = nullptr);
// ieee_f64: always the FP64 DMMA kernel.  The integer tensor-core path quantises with ONE global step (max|x| 2^-30):
// exact for the loop's row-normalised iterates, but a user matrix whose columns differ in scale by ~1e8 would lose its
// small columns -- the stand-alone whiten_embeddings entry point therefore asks for IEEE f64.
void launch_centered_gram(const float *x, int64_t n, int64_t d, const double *mean, double *cov, cudaStream_t st,
                          const AbsmaxPartials *absmax = nullptr, bool ieee_f64 = false);
void launch_whiten_apply(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T, int64_t dout,
                         float *out, cudaStream_t st);
bool whiten_apply_tc_supported(int64_t d, int64_t dout);
extern std::atomic<int> g_k3_asw;      // K3 A-tile layout: 0 = no swizzle (thread per row), 1 = SWIZZLE_128B (coalesced producers)
extern std::atomic<int> g_k3_bk;       // K3 stage shape: 32 = (BK 32, 2 stages), 16 = (BK 16, 4 stages); whiten_tc.cu
// upper_triangular: T has no entries below the diagonal (the Cholesky whitening transform) -- lets K3 skip the zero blocks
void launch_whiten_apply_tc(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T, int64_t dout,
                            float *out, int norm, const float *rowscale, cudaStream_t st, const PeerOut *peers = nullptr,
                            bool upper_triangular = false);
void launch_row_value_sums(const int64_t *rowptr, const float *val, int64_t n, float *out, cudaStream_t st);
void launch_sq_diff_sum(const float *a, const float *b, int64_t n, bool f64_diff, double *result, cudaStream_t st);
void launch_build_transform(const double *V, const double *w, int64_t d, int64_t dout, float *T, cudaStream_t st,
                            bool scaled = true);
void launch_scale_f64(double *v, int64_t n, double factor, cudaStream_t st);
void launch_f64_to_f32(const double *in, float *out, int64_t n, cudaStream_t st);
// Device-side integer ingest and synthetic pair generators (graph_dev.cu).
struct Graph;
std::unique_ptr<Graph> build_from_pairs_device(const uint32_t *u, const uint32_t *v, int64_t n_pairs,
                                               const std::string &column_name, int rank, int world, bool want_sym,
                                               cudaStream_t st, std::vector<int64_t> *bounds_out);
void synth_pairs_device(int kind, int64_t n_nodes, int64_t n_pairs, uint64_t seed, double alpha, uint32_t *u, uint32_t *v,
                        cudaStream_t st);
extern std::atomic<int> g_gram_needed_only;   // int8 Gram, d = 256: convert only the columns a tile reads (gram_tc.cu)
// Cholesky whitening (chol_whiten.cu): T = L^-T of cov = L L^T as f32; status[0] raised when cov is not safely SPD.
bool chol_whiten_supported(int64_t d);
void launch_chol_whiten(const double *cov, int64_t d, float *T, int *status, cudaStream_t st);

}  // namespace cleora
