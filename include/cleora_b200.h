/*
 * cleora_b200.h -- C ABI of libcleora_b200.so, the B200-native replacement for pycleora's hot path
 *   SparseMatrix::{left,symmetric}_markov_propagate -> row L2-normalise -> whiten_embeddings, wrapped by embed().
 *
 * This is the drop-in boundary: plain pointers and sizes, integer status returns, no exceptions and no
 * torch / numpy / pybind types.  It is what the reference's Rust host code (src/lib.rs #[pymethods]) would
 * bind through `extern "C"` instead of calling src/embedding.rs / src/sparse_matrix_builder.rs; the
 * reference-side binding is shown in INTEGRATION.md, and cleora_b200/pycleora.py is the same binding written
 * with ctypes (this image has no Rust toolchain).
 *
 * Conventions
 *   - every function returning `int` returns CLEORA_OK (0) or an error class; the message is available from
 *     cleora_last_error() (thread-local).  CLEORA_ERR_VALUE maps to Python ValueError, CLEORA_ERR_RUNTIME to
 *     RuntimeError -- the same classes the reference raises (src/lib.rs:39-42,186-189,220,335-338,465,472).
 *   - host-buffer entry points ("cleora_*") take HOST pointers owned by the caller; the library owns all
 *     device memory, copies in/out inside the call, and never aliases caller memory in results
 *     (the reference returns fresh numpy copies: src/lib.rs:46,251,363).
 *   - device-level entry points ("cleora_dev_*") take DEVICE pointers plus a cudaStream_t passed as void*;
 *     they only enqueue work.  They exist so a row-sharded multi-GPU loop (one process per GPU, collectives by
 *     torch.distributed/NCCL) is composed from the same kernels.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with CLEORA_ERR_CUDA.
 *   - matrices are C-contiguous row-major float32 [rows, d].
 */
#ifndef CLEORA_B200_H
#define CLEORA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLEORA_OK 0
#define CLEORA_ERR_VALUE 1   /* bad argument / shape / columns / unknown entity  -> ValueError   */
#define CLEORA_ERR_RUNTIME 2 /* (de)serialisation, internal invariant            -> RuntimeError */
#define CLEORA_ERR_CUDA 3    /* CUDA runtime error or no device                  -> RuntimeError */

/* MarkovType, src/embedding.rs:7-10 */
#define CLEORA_MARKOV_LEFT 0
#define CLEORA_MARKOV_SYMMETRIC 1

/* Row normalisation fused into the SpMM epilogue.
 *   L2_RUST  : x * (1 / max(sqrt(sum x^2), 1e-10))        src/embedding.rs:88-104 (embed_fast path)
 *   L2_NUMPY : x / max(sqrt(sum x^2), 1e-10)              pycleora/__init__.py:943-946 (default embed() path)
 *   L1_NUMPY : x / max(sum |x|, 1e-10)                    pycleora/__init__.py:947-950 */
#define CLEORA_NORM_NONE 0
#define CLEORA_NORM_L2_RUST 1
#define CLEORA_NORM_L2_NUMPY 2
#define CLEORA_NORM_L1_NUMPY 3

typedef struct cleora_graph cleora_graph_t; /* opaque: host CSR + entity tables + lazily built device copy */

const char *cleora_last_error(void);
const char *cleora_version(void);
/* Number of visible CUDA devices (0 when none / driver missing); never fails. */
int cleora_device_count(void);
/* Device used by subsequently created device state of this thread (cudaSetDevice). */
int cleora_set_device(int device);

/* ---------------------------------------------------------------------------------------------------------
 * Graph construction.  Replaces SparseMatrix::from_iterator (src/lib.rs:104-135) / from_files (:137-173) and
 * everything below them (src/pipeline.rs, src/entity.rs, src/sparse_matrix_builder.rs, src/configuration.rs).
 * CSR semantics are bit-exact with the reference for single-buffer accumulation order.
 * ------------------------------------------------------------------------------------------------------- */
/* `buf` holds n_lines UTF-8 strings back to back, line i = buf[offsets[i] .. offsets[i+1]). */
int cleora_graph_from_lines(const char *buf, const int64_t *offsets, int64_t n_lines, const char *columns,
                            int64_t hyperedge_trim_n, cleora_graph_t **out);
/* Only .tsv / .csv / .txt paths (src/lib.rs:148-158); empty lines skipped (src/pipeline.rs:212-214). */
int cleora_graph_from_files(const char *const *paths, int64_t n_paths, const char *columns,
                            int64_t hyperedge_trim_n, cleora_graph_t **out);
/* Direct integer ingest: the graph that `from_iterator(("{u} {v}" for u,v in pairs),
 * "complex::reflexive::<name>")` builds, without strings: entity index = first appearance, entity id =
 * decimal string of the integer, every pair adds 1/4+1/4 to M[u,v], M[v,u], M[u,u], M[v,v] and 1 to both row
 * sums.  u == v pairs are legal: all four ordered pairs of the line are (u,u), so they add 8 * 1/4 = 2 to M[u,u] and
 * 2 to row_sum[u] (src/sparse_matrix_builder.rs:170-233). */
int cleora_graph_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs, const char *column_name,
                            cleora_graph_t **out);
/* Integer hyperedges: line i has the members `members[offsets[i] .. offsets[i+1])` (ids of ONE column, given as
 * integers; their decimal strings are the entity ids).  Exactly the graph cleora_graph_from_lines builds from the
 * space-joined decimal strings under `columns` (e.g. "complex::reflexive::product"): same expansion, trimming and
 * accumulation order (src/sparse_matrix_builder.rs:170-233), without building strings in the caller. */
int cleora_graph_from_hyperedges(const uint32_t *members, const int64_t *offsets, int64_t n_lines, const char *columns,
                                 int64_t hyperedge_trim_n, cleora_graph_t **out);
/* The same ingest on the GPU (graph_dev.cu): `u`, `v` are DEVICE arrays on the current device; the CSR is built and
 * kept in HBM (no host copy: the 1.5 B-edge configuration has 24 GB of it; host accessors download on demand), bit for
 * bit the graph cleora_graph_from_pairs builds -- entity order, merged values, row sums, hashes.  shard_world > 1
 * builds ONE ROW SHARD directly: rows are split into shard_world contiguous blocks balanced by entry count
 * (`bounds_out`, int64[shard_world + 1], the same on every rank given the same pairs), this call keeps block
 * `shard_rank` with its column indices remapped to the padded gathered layout (owner * block + offset in owner,
 * block = largest block), n_cols = shard_world * block, row_offset = shard_rank * block.  want_sym == 0 skips the
 * symmetric values (4 B/nnz).  Ids must be dense (max id < 2^31).  Synchronises `stream` a few times (scalar counts). */
int cleora_dev_graph_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs, const char *column_name,
                                int shard_rank, int shard_world, int want_sym, void *stream, cleora_graph_t **out,
                                int64_t *bounds_out);
/* Synthetic pair streams for benchmarks, generated on the device with a counter-based RNG (pair i depends only on
 * (seed, i): every rank of a sharded run regenerates the same stream locally).  kind 0: endpoints i.i.d. uniform
 * (Erdos-Renyi multigraph); kind 1: Chung-Lu, endpoint weights (i + 10)^-alpha, ids decoupled from the weight rank by a
 * fixed pseudo-random permutation.  u != v always.  `u`, `v`: device uint32[n_pairs]. */
int cleora_dev_synth_pairs(int kind, int64_t n_nodes, int64_t n_pairs, uint64_t seed, double alpha, uint32_t *u,
                           uint32_t *v, void *stream);
/* Device pointer to the entity hashes cleora_dev_init reads: one per row of the (padded, for a device-built shard)
 * gathered matrix.  Uploads the graph if needed. */
int cleora_dev_graph_hashes(cleora_graph_t *g, const uint64_t **hash, int64_t *n_hash);
/* Entity count of the whole graph (== num_entities except for a device-built shard). */
int64_t cleora_graph_num_entities_global(const cleora_graph_t *g);
/* Adopt a prebuilt CSR (copied).  `n_rows` rows over `n_cols` columns; for a full graph n_rows == n_cols.  A
 * row shard of a larger graph has n_rows < n_cols and `row_offset` = global index of its first row (used by
 * the residual mix and by init).  `val_sym`, `row_sum`, `entity_hash` may be NULL. */
int cleora_graph_from_csr(const int64_t *rowptr, const uint32_t *col, const float *val_left, const float *val_sym,
                          const float *row_sum, const uint64_t *entity_hash, int64_t n_rows, int64_t n_cols,
                          int64_t row_offset, cleora_graph_t **out);
void cleora_graph_destroy(cleora_graph_t *g);
/* Drop the cached device copies (CSR stays on the host). */
int cleora_graph_release_device(cleora_graph_t *g);
/* Copy the host CSR into the current device's image again on `stream` (a cudaStream_t), reusing its buffers; uploads
 * it first if the device has no image yet.  For callers whose inputs arrive from the host every step. */
int cleora_graph_refresh_device(cleora_graph_t *g, void *stream);

/* --- introspection (getters of the pyclass: src/sparse_matrix.rs:56-66, src/lib.rs:175-240,254-318) ------- */
int64_t cleora_graph_num_entities(const cleora_graph_t *g); /* rows  (len(entity_ids)) */
int64_t cleora_graph_num_cols(const cleora_graph_t *g);
int64_t cleora_graph_num_edges(const cleora_graph_t *g);    /* nnz   (len(edges))      */
int cleora_graph_copy_csr(const cleora_graph_t *g, int64_t *rowptr, uint32_t *col, float *val_left, float *val_sym);
int cleora_graph_copy_row_sums(const cleora_graph_t *g, float *out);       /* entity_degrees */
int cleora_graph_copy_entity_hashes(const cleora_graph_t *g, uint64_t *out);
int cleora_graph_copy_column_ids(const cleora_graph_t *g, uint8_t *out);
int64_t cleora_graph_entity_ids_nbytes(const cleora_graph_t *g);
int cleora_graph_copy_entity_ids(const cleora_graph_t *g, char *buf, int64_t *offsets /* n+1 */);
/* entity_ids setter (#[pyo3(get, set)], src/sparse_matrix.rs:60): re-hashes, because
 * initialize_deterministically hashes the CURRENT ids (src/lib.rs:75). */
int cleora_graph_set_entity_ids(cleora_graph_t *g, const char *buf, const int64_t *offsets, int64_t n);
/* Restore the descriptor / column ids of an adopted CSR (used by __setstate__, src/lib.rs:470-475). */
int cleora_graph_set_descriptor(cleora_graph_t *g, int col_a_id, const char *col_a_name, int col_b_id,
                                const char *col_b_name);
int cleora_graph_set_column_ids(cleora_graph_t *g, const uint8_t *ids, int64_t n);
const char *cleora_graph_col_name(const cleora_graph_t *g, int which /* 0 = a, 1 = b */);
int cleora_graph_col_id(const cleora_graph_t *g, int which);
int64_t cleora_graph_find_entity(const cleora_graph_t *g, const char *id, int64_t id_len); /* -1 if absent */

/* XXH64(seed 0) of an entity id -- src/entity.rs:109-114. */
uint64_t cleora_hash_entity(const char *bytes, int64_t len);

/* ---------------------------------------------------------------------------------------------------------
 * Hot path, host buffers.  One call = one reference pymethod.
 * ------------------------------------------------------------------------------------------------------- */
/* initialize_deterministically (src/lib.rs:242-252): out[n, d]. */
int cleora_initialize_deterministically(cleora_graph_t *g, int64_t d, int64_t seed, float *out);
/* left_/symmetric_markov_propagate (src/lib.rs:86-102): x[x_rows, d] -> out[n, d];
 * CLEORA_ERR_VALUE "Embedding matrix has {} rows but graph has {} entities" when x_rows != n_cols. */
int cleora_markov_propagate(cleora_graph_t *g, const float *x, int64_t x_rows, int64_t d, int markov, float *out);
/* l2_normalize (src/lib.rs:414-424). */
int cleora_l2_normalize(const float *x, int64_t n, int64_t d, float *out);
/* embed_fast (src/lib.rs:320-364) and embed_fast_convergence (:366-412): init + device-resident loop. */
int cleora_embed_fast(cleora_graph_t *g, int64_t d, int64_t iters, int markov, int64_t seed,
                      float residual_weight, float *out);
int cleora_embed_fast_convergence(cleora_graph_t *g, int64_t d, int64_t max_iters, int markov, int64_t seed,
                                  float residual_weight, float convergence_threshold, float *out,
                                  int64_t *iters_done);
/* whiten_embeddings (pycleora/__init__.py:130-164): x[n, d] -> out[n, n_components], 1 <= n_components <= d
 * (CLEORA_ERR_VALUE otherwise; the binding resolves Python's None / slice semantics before the call). */
int cleora_whiten_embeddings(const float *x, int64_t n, int64_t d, int64_t n_components, float *out);
/* normalization="spectral" (pycleora/__init__.py:951-956): out = x @ V, V = eigenvectors of x^T x (f64 Gram) in
 * descending order -- U*S of the SVD of x up to column signs.  x is expected row-normalised; host [n, d]. */
int cleora_spectral_rotate(const float *x, int64_t n, int64_t d, float *out);
/* The loop body of embed() when it cannot take the Rust fast path (pycleora/__init__.py:97-125), kept on the
 * device for all iterations: propagate -> residual -> normalise -> whiten -> rmse early stop.
 * x0 == NULL: deterministic init from `seed`.  residual_weight is the Python float (double).  `x0` and `out` may
 * also be DEVICE pointers (the copies use cudaMemcpyDefault), which keeps the result in HBM.
 * `timings_ms` (optional, 8 doubles): h2d, init, spmm, stats, eigh, apply, rmse, d2h accumulated over the call. */
int cleora_embed(cleora_graph_t *g, const float *x0, int64_t d, int64_t iters, int markov, int64_t seed,
                 double residual_weight, double convergence_threshold, int normalization, int whiten,
                 float *out, int64_t *iters_done, double *timings_ms);

/* Symmetric eigensolver used by the whitening step: `a` is the d x d covariance (row-major, f64, symmetric);
 * on return `w[d]` holds eigenvalues in ASCENDING order and `a` the eigenvectors as COLUMNS (a[i*d + k] =
 * component i of eigenvector k) -- numpy.linalg.eigh's contract (pycleora/__init__.py:145).  Return 0 on
 * success.  Default (NULL): cuSOLVER Dsyevd on the device, enqueued on the loop's stream (no host sync).  Set
 * CLEORA_B200_EIGH=numpy to make the Python binding install numpy's LAPACK eigh, the call the reference makes. */
typedef int (*cleora_eigh_fn)(double *a, double *w, int64_t d, void *user);
void cleora_set_eigh(cleora_eigh_fn fn, void *user);            /* process-wide */
/* Override for the calling thread only (takes precedence over cleora_set_eigh): mode 1 = use `fn` (NULL = cuSOLVER),
 * mode 0 = follow the process-wide setting again.  Lets a binding scope an eigensolver to one call. */
void cleora_set_eigh_thread(int mode, cleora_eigh_fn fn, void *user);

/* Tuning switches.
 * "pipeline_whiten" (default 1): in cleora_embed's default configuration (whiten, l2, no residual, no early stop,
 *   tensor-core apply available for d) compute W = A Y while the transform of Y is being built, using
 *   A (Y - 1 mu^T) T = (A Y - (A 1) mu^T) T; 0 keeps the reference's stage order exactly.
 * "chol_whiten" (default 1): iterates that never leave the loop (iterations 0 .. T-2 of a call without rmse early
 *   stop, l2 / none normalisation, d <= 512) are whitened with the Cholesky factor T = L^-T of the covariance instead
 *   of the PCA transform of pycleora/__init__.py:145-156; the loop body is equivariant under orthogonal
 *   right-multiplication, so the final iterate (always PCA-whitened) is the reference's.  Falls back to the
 *   eigensolver for the whole call when a covariance is not safely positive definite (lambda_min near the reference's
 *   1e-10 clamp).  0 = eigensolver in every iteration.
 * "gram_needed_cols" (default 1): the integer tensor-core Gram kernel (d = 256, 384, 512) loads, converts and stages
 *   only the columns a tile reads -- its 128-column row block and its 64-column stripe -- in a compact shared-memory
 *   layout whose size does not depend on d; 0 = stage all d columns (d <= 256 only; d = 384 / 512 then use the FP64
 *   DMMA kernel).  Results are bit-identical (exact integer arithmetic either way).
 * "k3_asw" (default 1): A tiles of the tensor-core apply kernel in the SWIZZLE_128B K-major layout, written by coalesced
 *   producers (8 lanes per 128-byte row); 0 = one thread per row, no-swizzle layout.  Identical results.
 * "k3_bk" (32 or 16): stage shape of the tensor-core apply kernel -- 32 floats of K per stage, 2 stages, or 16 floats and
 *   4 stages (same shared memory, refills overlap three stage times instead of one); results are identical. */
int cleora_set_option(const char *key, int64_t value);
int64_t cleora_get_option(const char *key);      /* -1 for an unknown key */

/* Pinned host staging memory for callers that want async copies (bench e2e). */
int cleora_host_alloc(size_t nbytes, void **out);
void cleora_host_free(void *p);

/* ---------------------------------------------------------------------------------------------------------
 * Device-level building blocks (row-sharded multi-GPU composition; all pointers are DEVICE pointers on the
 * current device, `stream` is a cudaStream_t).  They enqueue and return.
 * ------------------------------------------------------------------------------------------------------- */
/* Upload (once) and return device views of the graph's CSR. */
int cleora_dev_graph_prepare(cleora_graph_t *g);
/* K0: out[n, d] = init_value(hash[i], j, seed). */
int cleora_dev_init(const uint64_t *hash, int64_t n, int64_t d, int64_t seed, float *out, void *stream);
/* K1: out[r, :] = norm( alpha * sum_e val[e] * x[col[e], :] + rw * resid[r, :] ) for the graph's rows.
 * `x` is the FULL [n_cols, d] matrix; `resid` is NULL or the [n_rows, d] block of the previous iterate.
 * The un-normalised product is bit-identical to src/embedding.rs:52-86 (same f32 operation order). */
int cleora_dev_spmm(cleora_graph_t *g, int markov, const float *x, int64_t d, float *out, const float *resid,
                    float alpha, float rw, int normalization, void *stream);
/* K1 / K3 with the all-gather fused into the epilogue: every produced row is stored to `out` AND to the same row of
 * up to 7 `extra_outs` -- the other ranks' copies of the gathered matrix, mapped into this process through CUDA IPC
 * (peer stores over NVLink).  `out` and each extra pointer address the first row of THIS rank's block. */
int cleora_dev_spmm_push(cleora_graph_t *g, int markov, const float *x, int64_t d, float *out,
                         float *const *extra_outs, int n_extra, const float *resid, float alpha, float rw,
                         int normalization, void *stream);
int cleora_dev_whiten_apply_push(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                 int64_t dout, float *out, float *const *extra_outs, int n_extra, int normalization,
                                 const float *rowscale, void *stream);
/* Column-sharded multi-GPU loop (cleora_b200/sharded.py: ColumnShardedEmbedder).  Every rank holds the WHOLE graph and a
 * column slice X[:, g*ds .. (g+1)*ds) of the iterate; the dense stages run row-sharded (rank h owns rows
 * [h*block_rows, (h+1)*block_rows)).  The two transposes between the layouts are all-to-alls fused into the producing
 * kernels' epilogues as coalesced peer stores (`dests` are the ranks' buffers mapped through CUDA IPC, own buffer
 * included, in rank order):
 *  - cleora_dev_spmm_scatter: K1 on a slice x[n_cols, d] (d = ds); row r of the product is stored to
 *    dests[r / block_rows] at row r % block_rows, columns [col_off, col_off + d) of an ld_cols-wide matrix.  No fused
 *    row norm (the row is not complete here); residual mix as in cleora_dev_spmm with resid laid out like x.
 *  - cleora_dev_whiten_apply_slices / cleora_dev_normalize_slices: K3 / K1's row normalisation on this rank's row
 *    block x[n, d]; the full rows go to `out` and columns [h*ds, (h+1)*ds) of row r to dests[h] at row row_base + r of
 *    an [n_total, ds] matrix, ds = dout / n_dst (a multiple of 4).  cleora_dev_normalize_slices with n_dst == 0 only
 *    writes `out`. */
int cleora_dev_spmm_scatter(cleora_graph_t *g, int markov, const float *x, int64_t d, float *const *dests, int n_dst,
                            int64_t block_rows, int64_t ld_cols, int64_t col_off, const float *resid, float alpha,
                            float rw, void *stream);
int cleora_dev_whiten_apply_slices(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                   int64_t dout, float *out, float *const *dests, int n_dst, int64_t row_base,
                                   int normalization, const float *rowscale, int t_upper, void *stream);
int cleora_dev_normalize_slices(const float *x, int64_t n, int64_t d, int normalization, float *out,
                                float *const *dests, int n_dst, int64_t row_base, void *stream);
/* embed() on several GPUs of one box from ONE process and ONE call -- what SURVEY.md section 8(b)/(e) sketched as "devices
 * inside the handle": the binding passes a device list instead of launching one process per GPU.  Replaces the same
 * reference entry points as cleora_embed / cleora_embed_fast[_convergence] (pycleora/__init__.py:51-127,
 * src/lib.rs:320-412), with the same arguments, results and error behaviour; `normalization` = CLEORA_NORM_L2_RUST
 * selects the Rust fast path's semantics (no whitening, f32 rmse early stop), any other code the Python loop's.
 * Implementation (cleora_b200/csrc/multi_gpu.inl): the column-sharded loop above with one host thread per entry of
 * `devices` (entries may repeat: several ranks on one GPU), peer access instead of CUDA IPC, cross-stream events instead
 * of NCCL barriers, and the two small all-reduces done by every rank itself with peer loads in rank order (bit-identical
 * statistics on every GPU).  `x0` NULL (deterministic init) or host/device [n, d]; `out` host or device [n, d].
 * Needs d % n_devices == 0 with both d and d / n_devices in {8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 1024}
 * (cleora_embed_multi_supported), every device visible and peer-accessible; otherwise CLEORA_ERR_VALUE / _CUDA. */
int cleora_embed_multi(cleora_graph_t *g, const int *devices, int n_devices, const float *x0, int64_t d, int64_t iters,
                       int markov, int64_t seed, double residual_weight, double convergence_threshold, int normalization,
                       int whiten, float *out, int64_t *iters_done);
int cleora_embed_multi_supported(int64_t d, int n_devices);
/* Device memory that can be exported to the other ranks of the node (cudaMalloc + cudaIpc*); handles are 64 bytes. */
int cleora_dev_malloc(size_t nbytes, void **out);
int cleora_dev_free(void *p);
int cleora_ipc_get_handle(void *p, unsigned char *handle64);
int cleora_ipc_open(const unsigned char *handle64, void **out);
int cleora_ipc_close(void *p);
/* Row normalisation alone (l2_normalize and the postprocess of user-supplied matrices). */
int cleora_dev_normalize(const float *x, int64_t n, int64_t d, int normalization, float *out, void *stream);
/* K2a: sums[d] (f64) += column sums of x[n, d]  (deterministic two-stage reduction; `sums` is overwritten
 * when accumulate == 0). */
int cleora_dev_col_sums(const float *x, int64_t n, int64_t d, double *sums, int accumulate, void *stream);
/* K2b: cov[d, d] (f64, unscaled) = sum_r (x_r - mean)(x_r - mean)^T with mean given in f64; overwritten. */
int cleora_dev_centered_gram(const float *x, int64_t n, int64_t d, const double *mean, double *cov, void *stream);
/* K3: out[n, dout] = (x[n, d] - mean_f32[d]) @ T[d, dout]  (f32). */
int cleora_dev_whiten_apply(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                            int64_t dout, float *out, void *stream);
/* K3 with the pipelined loop's extras: out = rownorm_or_not( (x - rowscale[r] * mean_f32) @ T ).  `rowscale` NULL = 1;
 * `normalization` CLEORA_NORM_NONE or CLEORA_NORM_L2_NUMPY (fused in the tensor-core epilogue; needs the tcgen05
 * shape rules d % 32 == 0 and either dout % 32 == 0, dout <= 256 or dout % 64 == 0, dout <= 512; otherwise
 * CLEORA_ERR_VALUE).  t_upper != 0 promises that T (d == dout) has no entries below the diagonal -- true for the
 * transform of cleora_dev_chol_whiten -- and lets the kernel skip the zero blocks (about 44 % of the tensor work). */
int cleora_dev_whiten_apply_ex(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                               int64_t dout, float *out, int normalization, const float *rowscale, int t_upper,
                               void *stream);
/* rowscale[r] = sum of the Markov values of row r (the vector A*1), f32 [n_rows], device. */
int cleora_dev_row_scale(cleora_graph_t *g, int markov, float *out, void *stream);
/* 1 if cleora_dev_whiten_apply_ex can fuse (tensor-core path available for this shape), else 0. */
int cleora_whiten_apply_fusable(int64_t d, int64_t dout);
/* sum((a - b)^2) over n elements, f64 accumulation; result[0] overwritten.  f64_diff == 0: f32 difference and
 * square (src/embedding.rs:173-174); != 0: f64 difference and square (pycleora/__init__.py:975-976). */
int cleora_dev_sq_diff_sum(const float *a, const float *b, int64_t n, int f64_diff, double *result, void *stream);
/* cov (f64 [d, d], device, already divided by n-1) -> T (f32 [d, dout], device) on `stream`: cuSOLVER Dsyevd +
 * a transform-building kernel, no host synchronisation -- unless a host eigensolver is installed
 * (cleora_set_eigh), in which case the call makes one synchronous round trip. */
int cleora_dev_whiten_transform(const double *cov, int64_t d, int64_t dout, float *T, void *stream);
/* Cholesky whitening on `stream` (one-SM f64 kernel, no library call, no host synchronisation): cov (f64 [d, d], device,
 * already divided by n-1) = L L^T  ->  T = L^-T as f32 [d, d] (upper triangular, T^T cov T = I).  status (int[1],
 * device) is set to 1 -- never cleared -- when a pivot is not positive or trace(cov^-1) > 1e8, i.e. when the PCA
 * transform with its 1e-10 clamp (pycleora/__init__.py:155) might not be a whitening of the same matrix.  d <= 512. */
int cleora_dev_chol_whiten(const double *cov, int64_t d, float *T, int *status, void *stream);
/* Host step of the whitening: cov (f64, already divided by n-1) -> T (f32 [d, dout]) via the installed eigh. */
int cleora_whiten_transform_from_cov(const double *cov, int64_t d, int64_t dout, float *T);
/* The calling thread keeps its iterate buffers, whitening scratch and cuSOLVER handle between calls (re-creating
 * them costs ~0.2 s per call); this frees them. */
int cleora_release_workspace(void);
/* Bytes of device scratch currently held by the calling thread's workspace (diagnostics). */
int64_t cleora_dev_workspace_bytes(void);
/* Number of kernel launches issued by this library since process start (bench's gpu_launches). */
int64_t cleora_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CLEORA_B200_H */
