// In-process multi-GPU embed(): the column-sharded loop of DESIGN.md section 8 behind ONE C-ABI call
// (cleora_embed_multi), so that a binding without torch.distributed -- the Rust pymethod of INTEGRATION.md, or
// cleora_b200.embed(devices=[...]) -- can use the GPUs of one box.  Included by abi.cu inside namespace cleora::{anon}.
//
// One host thread per rank ("rank" = one entry of the caller's device list; entries may repeat, which runs several
// ranks on one GPU -- how the single-GPU test box exercises the whole choreography).  Ranks share an address space, so
// the peer buffers need no CUDA IPC: cudaDeviceEnablePeerAccess makes every rank's cudaMalloc'ed buffers addressable
// from every other GPU, and the kernels' epilogues store straight into them over NVLink exactly as in the
// multi-process loop (cleora_b200/colsharded.py).  No NCCL either:
//   * "every rank has finished what it enqueued" = each rank records an event on its stream, the host threads meet at a
//     barrier, every rank's stream then waits for the other ranks' events (enqueue-only; no host thread ever waits for
//     a GPU inside the loop);
//   * the all-reduce of the d + d*d statistics = after such a barrier every rank sums all ranks' partial buffers itself,
//     in rank order, with peer loads (peer_sum_kernel): W * (d*d + d) * 8 bytes over NVLink per rank, and bit-identical
//     results on every rank by construction, so the redundantly computed Cholesky factors agree without a broadcast.
// Layouts, kernels and the order of every floating-point operation are those of colsharded.py; `whiten=False` is
// therefore bit-identical to one GPU here as well.

struct MultiAborted {};

class AbortableBarrier {
    std::mutex m_;
    std::condition_variable cv_;
    const int count_;
    int waiting_ = 0;
    uint64_t gen_ = 0;
    bool failed_ = false;

public:
    explicit AbortableBarrier(int n) : count_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        if (failed_) throw MultiAborted{};
        const uint64_t g = gen_;
        if (++waiting_ == count_) {
            waiting_ = 0;
            ++gen_;
            cv_.notify_all();
            return;
        }
        cv_.wait(lk, [&] { return gen_ != g || failed_; });
        if (gen_ == g) throw MultiAborted{};         // released by abort(), not by the last arrival
    }
    void abort() {
        std::lock_guard<std::mutex> lk(m_);
        failed_ = true;
        cv_.notify_all();
    }
};

struct PeerSrc {
    const double *p[8];
    int n;
};
// dst[i] = (p[0][i] + p[1][i] + ... ) * factor, summed in rank order (the same order on every rank)
__global__ void peer_sum_kernel(PeerSrc src, double *__restrict__ dst, int64_t count, double factor) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        double acc = src.p[0][i];
        for (int r = 1; r < src.n; ++r) acc += src.p[r][i];
        dst[i] = acc * factor;
    }
}

struct MultiArgs {
    Graph *g;
    const float *x0;
    int64_t d, iters, seed;
    int markov, norm, whiten, rust;
    double residual_weight, convergence_threshold;
    float *out;
    EighChoice eigh;
};

struct MultiShared {
    int W = 0;
    int64_t n = 0, d = 0, ds = 0, block = 0, n_pad = 0;
    std::vector<int> devs;
    std::vector<float *> xb, wa, T;                 // per rank: X[:, slice] (all rows) | W rows of the own block | transform
    std::vector<double *> part;                     // per rank: partial statistics [d*d + d + 1]
    std::vector<cudaEvent_t> ev[2];                 // [parity][rank]
    std::unique_ptr<AbortableBarrier> bar;
    std::atomic<int> flag{0};                       // scratch for host-side votes (Cholesky status)
    std::atomic<int64_t> done{0};
    std::mutex err_mu;
    std::string err;
    int err_code = CLEORA_OK;
};

struct MultiRank {
    MultiShared &S;
    const MultiArgs &A;
    const int rank, dev;
    cudaStream_t st = nullptr, side = nullptr;
    cudaEvent_t stats_done = nullptr, t_ready = nullptr;
    DeviceGraph *dg = nullptr;
    const float *val = nullptr;
    int64_t r0 = 0, r1 = 0, n_local = 0;
    uint32_t barriers = 0;
    DevBuf<float> ya, ya2, prev, mean32;
    DevBuf<double> sums, cov;
    DevBuf<int> status;
    DeviceEigh eig;
    PinnedBuf<double> h_scalar;
    PinnedBuf<int> h_status;
    PinnedBuf<double> h_cov;
    PinnedBuf<float> h_T;
    float *result_rows = nullptr;

    MultiRank(MultiShared &s, const MultiArgs &a, int r) : S(s), A(a), rank(r), dev(s.devs[(size_t)r]) {}

    double *part_cov() { return S.part[(size_t)rank]; }
    double *part_sums() { return S.part[(size_t)rank] + S.d * S.d; }
    double *part_scalar() { return S.part[(size_t)rank] + S.d * S.d + S.d; }

    // ---- set-up: everything that allocates or uploads; ends with the device idle and all ranks' pointers published
    void setup() {
        CUDA_TRY(cudaSetDevice(dev));
        for (int q = 0; q < S.W; ++q) {
            const int other = S.devs[(size_t)q];
            if (other == dev) continue;
            int can = 0;
            CUDA_TRY(cudaDeviceCanAccessPeer(&can, dev, other));
            if (!can) throw CudaFail{"GPU " + std::to_string(dev) + " cannot access GPU " + std::to_string(other) + " (no peer path)"};
            const cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else CUDA_TRY(e);
        }
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, greatest));
        CUDA_TRY(cudaEventCreateWithFlags(&stats_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t_ready, cudaEventDisableTiming));
        for (int p = 0; p < 2; ++p) CUDA_TRY(cudaEventCreateWithFlags(&S.ev[p][(size_t)rank], cudaEventDisableTiming));
        dg = &device_graph(*A.g);
        val = values_of(*dg, A.markov);
        if (!A.x0 && !dg->hash && S.n) value_error("graph has no entity hashes");
        r0 = std::min<int64_t>((int64_t)rank * S.block, S.n);
        r1 = std::min<int64_t>(r0 + S.block, S.n);
        n_local = r1 - r0;
        const size_t blk = (size_t)std::max<int64_t>(S.block, 1) * (size_t)S.d;
        auto dev_alloc = [&](size_t bytes) {
            void *p = nullptr;
            CUDA_TRY(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
            CUDA_TRY(cudaMemset(p, 0, std::max<size_t>(bytes, 16)));
            return p;
        };
        S.xb[(size_t)rank] = (float *)dev_alloc((size_t)std::max<int64_t>(S.n_pad, 1) * (size_t)S.ds * sizeof(float));
        S.wa[(size_t)rank] = (float *)dev_alloc(blk * sizeof(float));
        S.T[(size_t)rank] = (float *)dev_alloc((size_t)S.d * S.d * sizeof(float));
        S.part[(size_t)rank] = (double *)dev_alloc(((size_t)S.d * S.d + S.d + 1) * sizeof(double));
        ya.alloc(blk); ya2.alloc(blk);
        CUDA_TRY(cudaMemset(ya.p, 0, blk * sizeof(float)));
        CUDA_TRY(cudaMemset(ya2.p, 0, blk * sizeof(float)));
        mean32.alloc((size_t)std::max<int64_t>(S.d, 1));
        sums.alloc((size_t)std::max<int64_t>(S.d, 1));
        cov.alloc((size_t)std::max<int64_t>(S.d * S.d, 1));
        status.alloc(1);
        CUDA_TRY(cudaMemset(status.p, 0, sizeof(int)));
        h_scalar.resize(1); h_status.resize(1);
        if (rank == 0 && A.eigh.fn) {                      // the caller's host eigensolver (e.g. numpy's LAPACK), rank 0 only
            t_eigh_mode = 1;
            t_eigh = A.eigh;
            h_cov.resize((size_t)S.d * S.d);
            h_T.resize((size_t)S.d * S.d);
        }
        CUDA_TRY(cudaDeviceSynchronize());
        S.bar->wait();                                     // every rank's buffers exist and are zeroed
    }
    void teardown() {
        cudaSetDevice(dev);
        cudaDeviceSynchronize();
        if (S.bar) { try { S.bar->wait(); } catch (const MultiAborted &) {} }   // nobody still stores into a peer's buffer
        for (float *p : {S.xb[(size_t)rank], S.wa[(size_t)rank], S.T[(size_t)rank]}) if (p) cudaFree(p);
        if (S.part[(size_t)rank]) cudaFree(S.part[(size_t)rank]);
        for (int p = 0; p < 2; ++p) if (S.ev[p][(size_t)rank]) cudaEventDestroy(S.ev[p][(size_t)rank]);
        if (stats_done) cudaEventDestroy(stats_done);
        if (t_ready) cudaEventDestroy(t_ready);
        if (st) cudaStreamDestroy(st);
        if (side) cudaStreamDestroy(side);
        workspace().release();                             // this thread's scratch dies with the thread
    }

    // ---- synchronisation and the small collectives
    void dev_barrier() {
        const int p = (int)(barriers++ & 1);
        CUDA_TRY(cudaEventRecord(S.ev[p][(size_t)rank], st));
        S.bar->wait();                                     // all events of this round are recorded ...
        for (int q = 0; q < S.W; ++q)
            if (q != rank) CUDA_TRY(cudaStreamWaitEvent(st, S.ev[p][(size_t)q], 0));   // ... and gate this rank's next kernels
    }
    void all_sum(size_t part_off, double *dst, int64_t count, double factor) {
        PeerSrc src{};
        src.n = S.W;
        for (int q = 0; q < S.W; ++q) src.p[q] = S.part[(size_t)q] + part_off;
        const unsigned blocks = (unsigned)std::max<int64_t>(1, std::min<int64_t>((count + 255) / 256, 148));
        peer_sum_kernel<<<blocks, 256, 0, st>>>(src, dst, count, factor);
        LAUNCH_CHECK();
    }
    PeerOut dests(const std::vector<float *> &bufs, int mode) {
        PeerOut p{};
        p.n_extra = S.W;
        p.mode = mode;
        for (int q = 0; q < S.W; ++q) p.extra[q] = bufs[(size_t)q];
        return p;
    }

    // ---- stages (same kernels and arguments as colsharded.py)
    void init_slice() {
        float *xb = S.xb[(size_t)rank];
        if (A.x0) {
            if (S.n) CUDA_TRY(cudaMemcpy2DAsync(xb, (size_t)S.ds * sizeof(float), A.x0 + (size_t)rank * S.ds, (size_t)S.d * sizeof(float),
                                                 (size_t)S.ds * sizeof(float), (size_t)S.n, cudaMemcpyDefault, st));
        } else {
            const int64_t seed_g = (int64_t)((uint64_t)A.seed + (uint64_t)((int64_t)rank * S.ds));   // i64 wrap, src/lib.rs:478-488
            launch_init(dg->hash, S.n, S.ds, seed_g, xb, st);
        }
    }
    void spmm(bool resid, float alpha, float rw) {
        PeerOut p = dests(S.wa, PEER_OWNERS);
        p.block_rows = std::max<int64_t>(S.block, 1);
        p.ld_cols = S.d;
        p.col_off = (int)(rank * S.ds);
        const float *xb = S.xb[(size_t)rank];
        launch_spmm(*dg, val, xb, S.ds, nullptr, resid ? xb : nullptr, alpha, rw, CLEORA_NORM_NONE, st, &p);
        dev_barrier();                                     // every rank's slice of W has landed in its owner's buffer
    }
    void normalize_slices(const float *x, int norm, float *out, bool to_slices) {
        if (!to_slices) { launch_normalize_rows(x, n_local, S.d, norm, out, st, nullptr); return; }
        PeerOut p = dests(S.xb, PEER_SLICES);
        p.slice_cols = (int)S.ds;
        p.row_base = r0;
        launch_normalize_rows(x, n_local, S.d, norm, out, st, &p);
    }
    void apply_slices(const float *x, const float *rowscale, float *out, bool upper) {
        PeerOut p = dests(S.xb, PEER_SLICES);
        p.slice_cols = (int)S.ds;
        p.row_base = r0;
        launch_whiten_apply_tc(x, n_local, S.d, mean32.p, S.T[(size_t)rank], S.d, out, CLEORA_NORM_L2_NUMPY, rowscale, st, &p, upper);
    }
    void stats(const float *y) {
        const int64_t d = S.d;
        AbsmaxPartials mx;
        if (n_local > 0) launch_col_sums(y, n_local, d, part_sums(), false, st, &mx);
        else CUDA_TRY(cudaMemsetAsync(part_sums(), 0, sizeof(double) * (size_t)d, st));
        dev_barrier();
        all_sum((size_t)(d * d), sums.p, d, 1.0 / (double)S.n);                       // mean (f64), identical on every rank
        if (n_local > 0) launch_centered_gram(y, n_local, d, sums.p, part_cov(), st, &mx);
        else CUDA_TRY(cudaMemsetAsync(part_cov(), 0, sizeof(double) * (size_t)(d * d), st));
        dev_barrier();                                     // also: every rank's slice stores of this iteration have landed
        all_sum(0, cov.p, d * d, 1.0 / (double)(S.n - 1));
        launch_f64_to_f32(sums.p, mean32.p, d, st);
    }
    // PCA transform of `cov` on rank 0, copied by the others (one eigensolve: eigenvector signs are the solver's choice)
    void pca_transform_shared() {
        const int64_t d = S.d;
        if (rank == 0) {
            float *T = S.T[0];
            if (current_eigh().fn) {
                CUDA_TRY(cudaMemcpyAsync(h_cov.data(), cov.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                transform_from_cov(h_cov.data(), d, d, h_T.data());
                CUDA_TRY(cudaMemcpyAsync(T, h_T.data(), sizeof(float) * d * d, cudaMemcpyHostToDevice, st));
            } else {
                eig.transform(cov.p, d, d, T, st);
            }
        }
        dev_barrier();
        if (rank != 0) CUDA_TRY(cudaMemcpyAsync(S.T[(size_t)rank], S.T[0], sizeof(float) * d * d, cudaMemcpyDefault, st));
    }
    // true when no rank's Cholesky step raised its flag (synchronises the stream; a host-side vote)
    bool chol_ok() {
        CUDA_TRY(cudaMemcpyAsync(h_status.data(), status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaStreamSynchronize(side));
        if (h_status.data()[0] != 0) S.flag.store(1);
        S.bar->wait();
        const bool ok = S.flag.load() == 0;
        S.bar->wait();                                     // everyone has read the vote before anyone reuses the flag
        return ok;
    }
    void clear_status() {
        CUDA_TRY(cudaMemsetAsync(status.p, 0, sizeof(int), st));
        if (rank == 0) S.flag.store(0);
    }

    // ---- the default whitened loop (colsharded.py: run_pipelined; abi.cu: embed_pipelined_once)
    bool run_pipelined() {
        const int64_t d = S.d;
        const float *rowscale = row_scale_of(*dg, A.markov) + r0;
        CUDA_TRY(cudaStreamSynchronize(nullptr));           // row_scale_of may have launched on the default stream
        clear_status();
        init_slice();
        float *y = ya.p, *y2 = ya2.p;
        spmm(false, 1.f, 0.f);
        normalize_slices(S.wa[(size_t)rank], CLEORA_NORM_L2_NUMPY, y, true);
        stats(y);
        for (int64_t it = 1; it < A.iters; ++it) {
            CUDA_TRY(cudaEventRecord(stats_done, st));
            CUDA_TRY(cudaStreamWaitEvent(side, stats_done, 0));
            launch_chol_whiten(cov.p, d, S.T[(size_t)rank], status.p, side);       // every rank, identical input
            CUDA_TRY(cudaEventRecord(t_ready, side));
            spmm(false, 1.f, 0.f);                                                 // W = A Y needs no T
            CUDA_TRY(cudaStreamWaitEvent(st, t_ready, 0));
            apply_slices(S.wa[(size_t)rank], rowscale, y2, true);
            stats(y2);
            std::swap(y, y2);
        }
        if (!chol_ok()) return false;
        pca_transform_shared();                                                    // the iterate that leaves the loop: PCA
        launch_whiten_apply(y, n_local, d, mean32.p, S.T[(size_t)rank], d, y2, st);
        result_rows = y2;
        if (rank == 0) S.done.store(A.iters);
        return true;
    }

    // ---- reference stage order, any configuration (colsharded.py: run)
    bool run_general(bool allow_chol) {
        const int64_t d = S.d, n = S.n;
        bool use_res;
        float alpha, rw;
        if (A.rust) {                                        // src/embedding.rs:116
            use_res = A.residual_weight > 0.0 && A.residual_weight < 1.0;
            rw = (float)A.residual_weight;
            alpha = 1.0f - rw;
        } else {                                             // pycleora/__init__.py:114
            use_res = A.residual_weight > 0.0;
            alpha = (float)(1.0 - A.residual_weight);
            rw = (float)A.residual_weight;
        }
        const bool conv = A.convergence_threshold > 0.0;
        const bool do_whiten = A.whiten != 0 && n > 1;
        const bool inner_chol = allow_chol && do_whiten && !conv && A.iters >= 2 && chol_eligible(d, A.norm);
        if (inner_chol) clear_status();
        init_slice();
        if (conv && !prev.p) prev.alloc((size_t)std::max<int64_t>(S.block, 1) * (size_t)d);
        float *cur = ya.p, *other = ya2.p;
        bool have_prev = false;
        int64_t done = 0;
        for (int64_t it = 0; it < A.iters; ++it) {
            spmm(use_res, alpha, rw);
            float *fresh;
            if (do_whiten) {
                normalize_slices(S.wa[(size_t)rank], A.norm, other, false);
                stats(other);
                if (inner_chol && it + 1 < A.iters) {
                    launch_chol_whiten(cov.p, d, S.T[(size_t)rank], status.p, st);
                } else {
                    if (inner_chol && !chol_ok()) return false;             // before the PCA iterate is produced
                    pca_transform_shared();
                }
                launch_whiten_apply(other, n_local, d, mean32.p, S.T[(size_t)rank], d, cur, st);
                normalize_slices(cur, CLEORA_NORM_NONE, other, true);       // A -> B copy of the new iterate
                fresh = cur;
            } else {
                normalize_slices(S.wa[(size_t)rank], A.norm, other, true);
                fresh = other;
            }
            dev_barrier();
            done = it + 1;
            bool stop = false;
            if (conv && have_prev) {
                if (n_local > 0) launch_sq_diff_sum(fresh, prev.p, n_local * d, !A.rust, part_scalar(), st);
                else CUDA_TRY(cudaMemsetAsync(part_scalar(), 0, sizeof(double), st));
                dev_barrier();
                all_sum((size_t)(d * d + d), sums.p, 1, 1.0);              // (sums is free here: stats are done)
                CUDA_TRY(cudaMemcpyAsync(h_scalar.data(), sums.p, sizeof(double), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                const double tot = h_scalar.data()[0];
                const double cnt = (double)n * (double)d;
                const double rmse = A.rust ? (double)std::sqrt((float)tot / (float)(uint64_t)((uint64_t)n * (uint64_t)d))
                                           : std::sqrt(tot / cnt);
                stop = rmse < A.convergence_threshold;
                dev_barrier();                              // partial scalars are read before anyone overwrites them
            }
            if (conv) {
                CUDA_TRY(cudaMemcpyAsync(prev.p, fresh, sizeof(float) * (size_t)std::max<int64_t>(n_local, 0) * (size_t)d,
                                         cudaMemcpyDeviceToDevice, st));
                have_prev = true;
            }
            result_rows = fresh;
            if (fresh == other) std::swap(cur, other);
            if (stop) break;
        }
        if (A.iters == 0) {                                  // no iteration: the result is the initial matrix
            dev_barrier();
            gather_initial();
        }
        if (rank == 0) S.done.store(done);
        return true;
    }
    // iters == 0: X0 rows of the own block, re-assembled from the column slices (each rank reads all slices)
    void gather_initial() {
        for (int q = 0; q < S.W; ++q)
            if (n_local > 0)
                CUDA_TRY(cudaMemcpy2DAsync(ya.p + (size_t)q * S.ds, (size_t)S.d * sizeof(float),
                                           S.xb[(size_t)q] + (size_t)r0 * S.ds, (size_t)S.ds * sizeof(float),
                                           (size_t)S.ds * sizeof(float), (size_t)n_local, cudaMemcpyDefault, st));
        result_rows = ya.p;
    }

    void run() {
        const bool pipelined = pipeline_eligible(S.n, S.d, A.iters, A.norm, A.whiten, A.residual_weight, A.convergence_threshold) &&
                               chol_eligible(S.d, CLEORA_NORM_L2_NUMPY) && !A.rust;
        bool ok = pipelined ? run_pipelined() : run_general(true);
        if (!ok) {                                           // a covariance was not safely SPD: eigensolver throughout
            dev_barrier();
            ok = run_general(false);
        }
        if (n_local > 0 && result_rows)
            CUDA_TRY(cudaMemcpyAsync(A.out + (size_t)r0 * S.d, result_rows, sizeof(float) * (size_t)n_local * S.d, cudaMemcpyDefault, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaStreamSynchronize(side));
        if (rank == 0) eig.check_info();
    }
};

bool multi_eligible(int64_t d, int n_dev) {
    if (n_dev < 1 || n_dev > 8 || d <= 0 || d % n_dev != 0) return false;
    const int64_t ds = d / n_dev;
    return normalize_rows_supported(d) && normalize_rows_supported(ds);      // same width table as launch_spmm's row kernels
}

void embed_multi(const MultiArgs &A, const int *devices, int n_dev, int64_t *iters_done) {
    require_device();
    Graph &g = *A.g;
    if (g.n_rows != g.n_cols) value_error("embed needs a full (square) graph, not a row shard");
    if (A.d < 0 || A.iters < 0) value_error("feature_dim and num_iterations must be non-negative");
    check_norm(A.norm);
    if (!multi_eligible(A.d, n_dev))
        value_error("multi-GPU embed needs 1..8 devices and a feature dimension that splits into equal column slices of a "
                    "supported width (8, 16, 32, 64, 96, 128, 192, 256, 384, 512, 1024); got d=" + std::to_string(A.d) +
                    ", devices=" + std::to_string(n_dev));
    int visible = 0;
    CUDA_TRY(cudaGetDeviceCount(&visible));
    MultiShared S;
    S.W = n_dev;
    S.devs.assign(devices, devices + n_dev);
    for (int dv : S.devs)
        if (dv < 0 || dv >= visible) value_error("device " + std::to_string(dv) + " is not visible (" + std::to_string(visible) + " GPUs)");
    S.n = g.n_rows; S.d = A.d; S.ds = A.d / n_dev;
    S.block = S.n ? (S.n + n_dev - 1) / n_dev : 0;
    S.n_pad = S.block * n_dev;
    S.xb.assign((size_t)n_dev, nullptr); S.wa.assign((size_t)n_dev, nullptr); S.T.assign((size_t)n_dev, nullptr);
    S.part.assign((size_t)n_dev, nullptr);
    for (int p = 0; p < 2; ++p) S.ev[p].assign((size_t)n_dev, nullptr);
    S.bar = std::make_unique<AbortableBarrier>(n_dev);
    int caller_dev = 0;
    CUDA_TRY(cudaGetDevice(&caller_dev));
    std::vector<std::thread> threads;
    for (int r = 0; r < n_dev; ++r)
        threads.emplace_back([&S, &A, r] {
            MultiRank me(S, A, r);
            auto fail = [&](int code, const std::string &msg) {
                { std::lock_guard<std::mutex> lk(S.err_mu); if (S.err_code == CLEORA_OK) { S.err_code = code; S.err = msg; } }
                S.bar->abort();
            };
            try {
                me.setup();
                me.run();
            } catch (const MultiAborted &) {
            } catch (const BuildError &e) { fail(CLEORA_ERR_VALUE, e.msg);
            } catch (const CudaFail &e) { fail(CLEORA_ERR_CUDA, e.msg);
            } catch (const std::exception &e) { fail(CLEORA_ERR_RUNTIME, e.what()); }
            me.teardown();
        });
    for (auto &t : threads) t.join();
    cudaSetDevice(caller_dev);
    if (S.err_code == CLEORA_ERR_VALUE) throw BuildError{S.err};
    if (S.err_code == CLEORA_ERR_CUDA) throw CudaFail{S.err};
    if (S.err_code != CLEORA_OK) throw std::runtime_error(S.err);
    if (iters_done) *iters_done = S.done.load();
}
