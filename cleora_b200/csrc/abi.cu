// extern "C" surface of libcleora_b200 (see include/cleora_b200.h): handles, error mapping, host-buffer entry
// points (one per reference pymethod) and the device-resident embed loops.
#include "../../include/cleora_b200.h"
#include "device.cuh"
#include "graph.hpp"

#include <cusolverDn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

struct cleora_graph : cleora::Graph {};

namespace cleora {

// ------------------------------------------------------------------------------------------------ errors / scratch
static thread_local std::string t_err;
void set_error(const std::string &msg) { t_err = msg; }
std::atomic<int64_t> g_launches{0};

void *Scratch::get(size_t bytes) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (p && (dev != device || bytes > cap)) release();
    if (!p) {
        cap = std::max<size_t>(bytes, 256);
        device = dev;
        CUDA_TRY(cudaMalloc(&p, cap));
    }
    return p;
}
void Scratch::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}
static int current_device() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    return dev;
}
// Scratch and cached state are per (host thread, device): a thread that switches devices (cleora_set_device) gets a
// separate set instead of launching on buffers, streams or handles that belong to the previous device.
Workspace &workspace() {
    static thread_local std::map<int, Workspace> ws;
    return ws[current_device()];
}

namespace {

template <class F>
int guarded(F &&f) {
    try {
        f();
        return CLEORA_OK;
    } catch (const BuildError &e) {
        set_error(e.msg);
        return CLEORA_ERR_VALUE;
    } catch (const CudaFail &e) {
        set_error(e.msg);
        return CLEORA_ERR_CUDA;
    } catch (const std::bad_alloc &) {
        set_error("out of host memory");
        return CLEORA_ERR_RUNTIME;
    } catch (const std::exception &e) {
        set_error(e.what());
        return CLEORA_ERR_RUNTIME;
    }
}

struct ValueError : BuildError {};
[[noreturn]] void value_error(const std::string &m) { throw BuildError{m}; }

void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw CudaFail{"no CUDA device available: libcleora_b200 has no CPU fallback"};
    }
}

// RAII device buffer
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void alloc(size_t count) {
        free();
        n = count;
        if (count) CUDA_TRY(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { free(); }
};

int64_t env_int64(const char *name, int64_t dflt) {
    const char *e = getenv(name);
    return (e && *e) ? atoll(e) : dflt;
}

void free_device_graph(DeviceGraph *dg) {
    if (!dg) return;
    cudaFree(dg->rowptr); cudaFree(dg->col); cudaFree(dg->left); cudaFree(dg->sym); cudaFree(dg->hash);
    cudaFree(dg->long_rows); cudaFree(dg->long_chunk_ptr); cudaFree(dg->long_chunk_owner);
    cudaFree(dg->rsum_left); cudaFree(dg->rsum_sym); cudaFree(dg->row_sum_all); cudaFree(dg->orig_ids); cudaFree(dg->row_order);
    delete dg;
}

// Long-row schedule (degree skew): rows above the threshold are processed chunk-wise by separate warps.
// The default threshold is deliberately high: a chunked f32 sum differs from the reference's sequential sum by
// ~sqrt(deg) ulp (measured 1e-5 relative at 30k edges, and that feeds back through 40 iterations on star-like graphs),
// so splitting is reserved for hubs whose sequential walk would dominate the launch.
void attach_long_row_schedule(DeviceGraph &d, const std::vector<int64_t> &rowptr) {
    DeviceGraph *dg = &d;
    const int64_t n_rows = (int64_t)rowptr.size() - 1;
    const int64_t threshold = env_int64("CLEORA_B200_LONG_ROW", 65536), chunk = env_int64("CLEORA_B200_LONG_CHUNK", 4096);
    std::vector<int64_t> rows, cptr{0};
    std::vector<int32_t> owner;
    const int64_t sched = std::max<int64_t>(1024, threshold / 16);     // narrow-row kernels split sooner (kernels.cu: launch_rows)
    int64_t max_degree = 0;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t deg = rowptr[(size_t)r + 1] - rowptr[(size_t)r];
        max_degree = std::max(max_degree, deg);
        if (deg > sched) {
            const int64_t nc = (deg + chunk - 1) / chunk;
            for (int64_t c = 0; c < nc; ++c) owner.push_back((int32_t)rows.size());
            rows.push_back(r);
            cptr.push_back(cptr.back() + nc);
        }
    }
    {   // Row schedule for the kernels that put several (narrow) rows into one warp: inside windows of 4096 consecutive
        // rows the rows are visited by descending degree, so the lane groups of a warp get rows of similar length, while
        // the CSR is still streamed window by window (a global degree order was measured slower: it turns the col/val
        // stream into random ~100-byte segments).  Only where the degrees are skewed enough to matter.
        int64_t max_deg = 0;
        for (int64_t r = 0; r < n_rows; ++r) max_deg = std::max(max_deg, rowptr[(size_t)r + 1] - rowptr[(size_t)r]);
        const int64_t nnz = n_rows ? rowptr[(size_t)n_rows] : 0;
        if (n_rows >= 1024 && n_rows < ((int64_t)1 << 32) && max_deg * n_rows > 4 * nnz && env_int64("CLEORA_B200_ROW_ORDER", 1) != 0) {
            const int64_t window = env_int64("CLEORA_B200_ROW_WINDOW", 4096);
            std::vector<uint32_t> order((size_t)n_rows);
            #pragma omp parallel for schedule(static)
            for (int64_t w0 = 0; w0 < n_rows; w0 += window) {
                const int64_t w1 = std::min(n_rows, w0 + window);
                for (int64_t r = w0; r < w1; ++r) order[(size_t)r] = (uint32_t)r;
                std::stable_sort(order.begin() + w0, order.begin() + w1, [&](uint32_t a, uint32_t b) {
                    return rowptr[(size_t)a + 1] - rowptr[(size_t)a] > rowptr[(size_t)b + 1] - rowptr[(size_t)b];
                });
            }
            CUDA_TRY(cudaMalloc((void **)&dg->row_order, order.size() * sizeof(uint32_t)));
            CUDA_TRY(cudaMemcpy(dg->row_order, order.data(), order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    dg->long_threshold = threshold;
    dg->long_sched_threshold = sched;
    dg->max_degree = max_degree;
    dg->long_chunk_edges = chunk;
    dg->n_long = (int64_t)rows.size();
    dg->n_long_chunks = (int64_t)owner.size();
    if (!rows.empty()) {
        CUDA_TRY(cudaMalloc((void **)&dg->long_rows, rows.size() * sizeof(int64_t)));
        CUDA_TRY(cudaMemcpy(dg->long_rows, rows.data(), rows.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc((void **)&dg->long_chunk_ptr, cptr.size() * sizeof(int64_t)));
        CUDA_TRY(cudaMemcpy(dg->long_chunk_ptr, cptr.data(), cptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc((void **)&dg->long_chunk_owner, owner.size() * sizeof(int32_t)));
        CUDA_TRY(cudaMemcpy(dg->long_chunk_owner, owner.data(), owner.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
}

// Upload the CSR once per (graph, device).  Arrays get 16 trailing elements of padding so tile loads never
// touch unmapped memory.
DeviceGraph &device_graph(Graph &g) {
    require_device();
    const int dev = current_device();
    std::lock_guard<std::recursive_mutex> lock(g.mu);
    for (DeviceGraph *have : g.devs)
        if (have->device == dev) return *have;
    if (g.device_only && g.col.empty())
        value_error("this graph was built on device " + std::to_string(g.devs.empty() ? -1 : g.devs[0]->device) +
                    " and has no host copy; call it with that device current");
    if (!g.host_pinned && env_int64("CLEORA_B200_PIN_CSR", 1) != 0 && g.nnz() >= (1 << 20)) {
        // page-lock the host CSR once: uploads then run at PCIe speed instead of through the pageable staging path
        auto pin = [](const void *p, size_t bytes) { if (bytes) { if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) cudaGetLastError(); } };
        pin(g.rowptr.data(), g.rowptr.size() * sizeof(int64_t));
        pin(g.col.data(), g.col.size() * sizeof(uint32_t));
        pin(g.left.data(), g.left.size() * sizeof(float));
        pin(g.sym.data(), g.sym.size() * sizeof(float));
        g.host_pinned = true;
    }
    auto *dg = new DeviceGraph();
    dg->device = dev;
    dg->n_rows = g.n_rows; dg->n_cols = g.n_cols; dg->nnz = g.nnz(); dg->row_offset = g.row_offset;
    try {
        const size_t pad = 16, nnz = (size_t)g.nnz();
        CUDA_TRY(cudaMalloc((void **)&dg->rowptr, (g.rowptr.size() + pad) * sizeof(int64_t)));
        CUDA_TRY(cudaMemcpy(dg->rowptr, g.rowptr.data(), g.rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc((void **)&dg->col, (nnz + pad) * sizeof(uint32_t)));
        CUDA_TRY(cudaMemset(dg->col, 0, (nnz + pad) * sizeof(uint32_t)));
        CUDA_TRY(cudaMemcpy(dg->col, g.col.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMalloc((void **)&dg->left, (nnz + pad) * sizeof(float)));
        CUDA_TRY(cudaMemset(dg->left, 0, (nnz + pad) * sizeof(float)));
        CUDA_TRY(cudaMemcpy(dg->left, g.left.data(), nnz * sizeof(float), cudaMemcpyHostToDevice));
        if (!g.sym.empty()) dg->host_sym = g.sym.data();       // uploaded by values_of() when first asked for
        if (!g.hash.empty()) {
            CUDA_TRY(cudaMalloc((void **)&dg->hash, g.hash.size() * sizeof(uint64_t)));
            CUDA_TRY(cudaMemcpy(dg->hash, g.hash.data(), g.hash.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        }
        attach_long_row_schedule(*dg, g.rowptr);
    } catch (...) {
        free_device_graph(dg);
        throw;
    }
    g.devs.push_back(dg);
    return *dg;
}

// Host arrays of a device-built graph, downloaded on first use by an accessor that needs them.
void materialize_host(Graph &g) {
    std::lock_guard<std::recursive_mutex> lock(g.mu);
    if (!g.device_only || g.devs.empty()) return;
    DeviceGraph &dg = *g.devs[0];
    if (g.col.empty() && g.nnz_device > 0) {
        const size_t nnz = (size_t)g.nnz_device;
        std::vector<uint32_t> col(nnz);
        std::vector<float> left(nnz), sym;
        CUDA_TRY(cudaMemcpy(col.data(), dg.col, nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(left.data(), dg.left, nnz * sizeof(float), cudaMemcpyDeviceToHost));
        if (dg.sym) { sym.resize(nnz); CUDA_TRY(cudaMemcpy(sym.data(), dg.sym, nnz * sizeof(float), cudaMemcpyDeviceToHost)); }
        g.col.swap(col); g.left.swap(left); g.sym.swap(sym);
    }
    if (g.row_sum.empty() && g.n_rows) {
        g.row_sum.resize((size_t)g.n_rows);
        CUDA_TRY(cudaMemcpy(g.row_sum.data(), dg.row_sum_all + g.shard_r0, (size_t)g.n_rows * sizeof(float), cudaMemcpyDeviceToHost));
        g.hash.resize((size_t)g.n_rows);
        CUDA_TRY(cudaMemcpy(g.hash.data(), dg.hash + g.row_offset, (size_t)g.n_rows * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        g.column_id.assign((size_t)g.n_rows, 0);
    }
    if (g.ids.empty() && g.n_rows && dg.orig_ids) {             // entity ids = decimal strings of the integer ids
        std::vector<uint32_t> orig((size_t)g.n_rows);
        CUDA_TRY(cudaMemcpy(orig.data(), dg.orig_ids + g.shard_r0, orig.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        g.ids.resize(orig.size());
        for (size_t i = 0; i < orig.size(); ++i) g.ids[i] = std::to_string(orig[i]);
    }
}

const float *values_of(DeviceGraph &dg, int markov) {
    if (markov == CLEORA_MARKOV_LEFT) return dg.left;
    if (markov == CLEORA_MARKOV_SYMMETRIC) {
        std::lock_guard<std::mutex> lock(dg.lazy_mu);
        if (!dg.sym) {
            if (!dg.host_sym) value_error("graph was created without symmetric Markov values");
            const size_t pad = 16, nnz = (size_t)dg.nnz;
            float *p = nullptr;
            CUDA_TRY(cudaMalloc((void **)&p, (nnz + pad) * sizeof(float)));
            if (cudaMemset(p, 0, (nnz + pad) * sizeof(float)) != cudaSuccess ||
                cudaMemcpy(p, dg.host_sym, nnz * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(p);
                throw CudaFail{std::string("upload of symmetric values failed: ") + cudaGetErrorString(cudaGetLastError())};
            }
            dg.sym = p;
        }
        return dg.sym;
    }
    value_error("Unknown propagation. Use 'left' or 'symmetric'.");
}

void check_norm(int norm) {
    if (norm < CLEORA_NORM_NONE || norm > CLEORA_NORM_L1_NUMPY) value_error("unknown normalization code");
}

// ------------------------------------------------------------------------------------------------ eigh
// Host eigensolver: a process-wide setting (cleora_set_eigh) and a per-thread override (cleora_set_eigh_thread) that
// bindings use to scope a choice to one call without racing other threads.
struct EighChoice {
    cleora_eigh_fn fn = nullptr;
    void *user = nullptr;
};
std::atomic<cleora_eigh_fn> g_eigh_fn{nullptr};
std::atomic<void *> g_eigh_user{nullptr};
thread_local int t_eigh_mode = 0;          // 0: follow the process-wide setting, 1: use t_eigh
thread_local EighChoice t_eigh;
EighChoice current_eigh() {
    if (t_eigh_mode) return t_eigh;
    return EighChoice{g_eigh_fn.load(), g_eigh_user.load()};
}

// Default eigensolver: cuSOLVER Dsyevd on the current device (a library call for the small d x d step; the
// reference's own GPU path does the same through torch.linalg.eigh, pycleora/__init__.py:990).
int eigh_cusolver(double *a, double *w, int64_t d, void *) {
    static thread_local std::map<int, cusolverDnHandle_t> handles;       // one per (thread, device)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    cusolverDnHandle_t &h = handles[dev];
    if (!h && cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) return 1;
    const int n = (int)d;
    int lwork = 0, info = 0;
    try {
        DevBuf<double> dA((size_t)d * d), dW((size_t)d);
        DevBuf<int> dInfo(1);
        CUDA_TRY(cudaMemcpy(dA.p, a, sizeof(double) * d * d, cudaMemcpyHostToDevice));
        if (cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.p, n, dW.p, &lwork) !=
            CUSOLVER_STATUS_SUCCESS)
            return 2;
        DevBuf<double> work((size_t)std::max(lwork, 1));
        if (cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA.p, n, dW.p, work.p, lwork,
                             dInfo.p) != CUSOLVER_STATUS_SUCCESS)
            return 3;
        CUDA_TRY(cudaMemcpy(&info, dInfo.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (info != 0) return 4;
        std::vector<double> colmajor((size_t)d * d);
        CUDA_TRY(cudaMemcpy(colmajor.data(), dA.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(w, dW.p, sizeof(double) * d, cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < d; ++i)
            for (int64_t k = 0; k < d; ++k) a[i * d + k] = colmajor[(size_t)(k * d + i)];   // eigvec k = column k
    } catch (const CudaFail &) {
        return 5;
    }
    return 0;
}

// pycleora/__init__.py:145-156: eigh -> descending order -> scale = 1/sqrt(max(lambda,1e-10)) -> (V*scale) as f32.
void transform_from_cov(const double *cov, int64_t d, int64_t dout, float *T, bool scaled = true) {
    std::vector<double> a(cov, cov + d * d), w((size_t)d);
    const EighChoice eh = current_eigh();
    int rc = eh.fn ? eh.fn(a.data(), w.data(), d, eh.user) : eigh_cusolver(a.data(), w.data(), d, nullptr);
    if (rc != 0) throw std::runtime_error("eigh failed with code " + std::to_string(rc));
    std::vector<double> scale((size_t)dout);
    for (int64_t k = 0; k < dout; ++k)                          // argsort(eigenvalues)[::-1] on ascending input: column d-1-k
        scale[(size_t)k] = scaled ? 1.0 / std::sqrt(std::max(w[(size_t)(d - 1 - k)], 1e-10)) : 1.0;
    for (int64_t i = 0; i < d; ++i) {
        const double *row = a.data() + i * d + (d - 1);
        float *t = T + i * dout;
        for (int64_t k = 0; k < dout; ++k) t[k] = (float)(row[-k] * scale[(size_t)k]);
    }
}

// Whitening of a device-resident matrix: Y[n,d] -> Z[n,dout].  Scratch: sums/cov (f64), mean32, T on device.
// Device-resident eigensolver state: cuSOLVER Dsyevd on the caller's stream, workspace sized once per d.
struct DeviceEigh {
    cusolverDnHandle_t h = nullptr;
    DevBuf<double> evec, eval, work;
    DevBuf<int> info;
    int64_t d = 0;
    int lwork = 0;
    void ensure(int64_t d_) {
        if (!h && cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) throw std::runtime_error("cusolverDnCreate failed");
        if (d == d_) return;
        d = d_;
        evec.alloc((size_t)d * d); eval.alloc((size_t)d); info.alloc(1);
        if (cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, evec.p, (int)d, eval.p,
                                        &lwork) != CUSOLVER_STATUS_SUCCESS)
            throw std::runtime_error("cusolverDnDsyevd_bufferSize failed");
        work.alloc((size_t)std::max(lwork, 1));
    }
    // cov (d x d, symmetric, device) -> T (d x dout f32, device); everything enqueued on `st`.
    void transform(const double *cov, int64_t d_, int64_t dout, float *T, cudaStream_t st, bool scaled = true) {
        ensure(d_);
        if (cusolverDnSetStream(h, st) != CUSOLVER_STATUS_SUCCESS) throw std::runtime_error("cusolverDnSetStream failed");
        CUDA_TRY(cudaMemcpyAsync(evec.p, cov, sizeof(double) * d * d, cudaMemcpyDeviceToDevice, st));
        if (cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)d, evec.p, (int)d, eval.p, work.p,
                             lwork, info.p) != CUSOLVER_STATUS_SUCCESS)
            throw std::runtime_error("cusolverDnDsyevd failed");
        launch_build_transform(evec.p, eval.p, d, dout, T, st, scaled);
    }
    void check_info() {
        int hinfo = 0;
        if (info.p) CUDA_TRY(cudaMemcpy(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (hinfo != 0) throw std::runtime_error("cuSOLVER Dsyevd did not converge (info=" + std::to_string(hinfo) + ")");
    }
    ~DeviceEigh() { if (h) cusolverDnDestroy(h); }
};

// Page-locked host staging (the covariance down / the transform up, once per iteration with a host eigensolver).
template <typename T>
struct PinnedBuf {
    T *p = nullptr;
    size_t n = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    void resize(size_t count) {
        if (count == n) return;
        release();
        if (count) CUDA_TRY(cudaHostAlloc((void **)&p, count * sizeof(T), cudaHostAllocDefault));
        n = count;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
    }
    T *data() { return p; }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

struct WhitenState {
    DevBuf<double> sums, cov;
    DevBuf<float> mean32, T;
    PinnedBuf<double> h_cov;
    PinnedBuf<float> h_T;
    DevBuf<int> status;                  // raised by launch_chol_whiten when the covariance is not safely SPD
    PinnedBuf<int> h_status;
    DeviceEigh eig;
    int64_t d = 0, dout = 0;
    void ensure(int64_t d_, int64_t dout_) {
        if (!status.p) { status.alloc(1); h_status.resize(1); }
        if (d == d_ && dout == dout_) return;
        d = d_; dout = dout_;
        sums.alloc((size_t)d); cov.alloc((size_t)d * d); mean32.alloc((size_t)d); T.alloc((size_t)d * dout);
        h_cov.resize((size_t)d * d); h_T.resize((size_t)d * dout);
    }
    void clear_status(cudaStream_t st) { CUDA_TRY(cudaMemsetAsync(status.p, 0, sizeof(int), st)); }
    // true when no Cholesky step of the loop raised the flag; synchronises `st`
    bool status_ok(cudaStream_t st) {
        CUDA_TRY(cudaMemcpyAsync(h_status.data(), status.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return h_status.data()[0] == 0;
    }
};

struct Phase {
    // optional per-phase device timing (cudaEvent pairs), summed at the end
    bool on = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[8];
    cudaStream_t st = nullptr;
    cudaEvent_t a = nullptr;
    void begin() { begin_on(st); }
    void end(int which) { end_on(which, st); }
    void begin_on(cudaStream_t s) {
        if (!on) return;
        cudaEventCreate(&a);
        cudaEventRecord(a, s);
    }
    void end_on(int which, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t b;
        cudaEventCreate(&b);
        cudaEventRecord(b, s);
        ev[which].push_back({a, b});
    }
    void collect(double *out) {
        if (!on) return;
        cudaDeviceSynchronize();
        for (int i = 0; i < 8; ++i)
            for (auto &p : ev[i]) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, p.first, p.second);
                out[i] += ms;
                cudaEventDestroy(p.first);
                cudaEventDestroy(p.second);
            }
    }
};
enum { PH_H2D = 0, PH_INIT, PH_SPMM, PH_STATS, PH_EIGH, PH_APPLY, PH_RMSE, PH_D2H };

// `cholesky`: use T = L^-T (chol_whiten.cu) instead of the PCA transform -- only for iterates that are never handed
// to the caller (see embed_reference_order); dout must equal d then.
void whiten_device(const float *Y, int64_t n, int64_t d, int64_t dout, float *Z, WhitenState &ws, cudaStream_t st,
                   Phase &ph, bool cholesky = false, bool ieee_f64 = false) {
    ws.ensure(d, dout);
    ph.begin();
    AbsmaxPartials mx;
    launch_col_sums(Y, n, d, ws.sums.p, false, st, &mx);
    launch_scale_f64(ws.sums.p, d, 1.0 / (double)n, st);                 // mean (f64)
    launch_centered_gram(Y, n, d, ws.sums.p, ws.cov.p, st, &mx, ieee_f64);
    launch_scale_f64(ws.cov.p, d * d, 1.0 / (double)(n - 1), st);        // cov *= 1/(n-1)
    launch_f64_to_f32(ws.sums.p, ws.mean32.p, d, st);                    // mean.astype(float32)
    ph.end(PH_STATS);
    ph.begin();
    if (cholesky) {
        launch_chol_whiten(ws.cov.p, d, ws.T.p, ws.status.p, st);
    } else if (current_eigh().fn) {     // host eigensolver installed by the binding (e.g. numpy's LAPACK): one round trip
        CUDA_TRY(cudaMemcpyAsync(ws.h_cov.data(), ws.cov.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        transform_from_cov(ws.h_cov.data(), d, dout, ws.h_T.data());
        CUDA_TRY(cudaMemcpyAsync(ws.T.p, ws.h_T.data(), sizeof(float) * d * dout, cudaMemcpyHostToDevice, st));
    } else {                 // default: cuSOLVER on the stream, no host synchronisation inside the loop
        ws.eig.transform(ws.cov.p, d, dout, ws.T.p, st);
    }
    ph.end(PH_EIGH);
    ph.begin();
    launch_whiten_apply(Y, n, d, ws.mean32.p, ws.T.p, dout, Z, st);
    ph.end(PH_APPLY);
}

// ---- per-thread persistent device state (cuSOLVER handle, whitening scratch, iterate buffers): creating these per
// call costs ~0.2 s (cusolverDnCreate, four 1 GB cudaMallocs), measured; cleora_release_workspace() frees them.
struct Persistent {
    WhitenState ws;
    DevBuf<float> buf[5];
    float *get(int i, size_t count) {
        if (buf[i].n < count) buf[i].alloc(count);
        return buf[i].p;
    }
    void release() {
        for (auto &b : buf) b.free();
        ws.sums.free(); ws.cov.free(); ws.mean32.free(); ws.T.free(); ws.status.free();
        ws.eig.evec.free(); ws.eig.eval.free(); ws.eig.work.free(); ws.eig.info.free();
        ws.h_cov.release(); ws.h_T.release(); ws.h_status.release();
        ws.d = ws.dout = 0; ws.eig.d = 0;
    }
};
struct SecondStream;
struct PerDevice {                       // everything a host thread caches on ONE device
    Persistent ps;
    std::shared_ptr<SecondStream> side;  // side stream + events of the pipelined loop
    DeviceEigh eig_misc;                 // cleora_dev_whiten_transform
};
PerDevice &per_device() {
    static thread_local std::map<int, PerDevice> state;
    return state[current_device()];
}
Persistent &persistent() { return per_device().ps; }

// ---- options -------------------------------------------------------------------------------------------------
std::atomic<int> g_opt_pipeline{1};
std::atomic<int> g_opt_chol{1};          // Cholesky whitening for iterates that stay inside the loop

// mean / covariance of Y (device) into ws.sums (mean, f64), ws.mean32, ws.cov (scaled by 1/(n-1))
void stats_device(const float *Y, int64_t n, int64_t d, WhitenState &ws, cudaStream_t st) {
    AbsmaxPartials mx;                                   // max|Y| rides on the column-sum pass
    launch_col_sums(Y, n, d, ws.sums.p, false, st, &mx);
    launch_scale_f64(ws.sums.p, d, 1.0 / (double)n, st);
    launch_centered_gram(Y, n, d, ws.sums.p, ws.cov.p, st, &mx);
    launch_scale_f64(ws.cov.p, d * d, 1.0 / (double)(n - 1), st);
    launch_f64_to_f32(ws.sums.p, ws.mean32.p, d, st);
}

const float *row_scale_of(DeviceGraph &dg, int markov) {
    const float *val = values_of(dg, markov);
    std::lock_guard<std::mutex> lock(dg.lazy_mu);
    float *&slot = markov == CLEORA_MARKOV_LEFT ? dg.rsum_left : dg.rsum_sym;
    if (!slot) {
        CUDA_TRY(cudaMalloc((void **)&slot, sizeof(float) * (size_t)std::max<int64_t>(dg.n_rows, 1)));
        launch_row_value_sums(dg.rowptr, val, dg.n_rows, slot, nullptr);
    }
    return slot;
}

// Default whitened loop with the eigensolve taken off the critical path.
//
// Faithful order per iteration:  Y = rownorm(A X);  (mu, C) = stats(Y);  T = eig(C);  X' = (Y - 1 mu^T) T.
// The next product is linear in X':  A X' = (A Y - (A 1) mu^T) T, so W = A Y does not need T and runs on the main
// stream while cuSOLVER works on a second stream; the tensor-core GEMM then applies T to (W - s mu^T), s = A 1, and
// normalises rows in its epilogue.  Same mathematics, different f32 rounding points (a few ulp); the final iterate
// is produced by the plain apply.  Used only where every stage has a kernel for it (see eligible()).
struct SecondStream {
    cudaStream_t s = nullptr;
    cudaEvent_t stats_done = nullptr, t_ready = nullptr;
    SecondStream() {
        // highest priority: cuSOLVER's chain of small kernels must get SM slots ahead of the SpMM's 125k CTAs
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, greatest));
        CUDA_TRY(cudaEventCreateWithFlags(&stats_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&t_ready, cudaEventDisableTiming));
    }
    ~SecondStream() {
        if (stats_done) cudaEventDestroy(stats_done);
        if (t_ready) cudaEventDestroy(t_ready);
        if (s) cudaStreamDestroy(s);
    }
};

bool pipeline_eligible(int64_t n, int64_t d, int64_t iters, int normalization, int whiten, double residual_weight,
                       double convergence_threshold) {
    return g_opt_pipeline.load() && whiten && n > 1 && iters >= 2 &&
           normalization == CLEORA_NORM_L2_NUMPY && residual_weight == 0.0 && convergence_threshold <= 0.0 &&
           whiten_apply_tc_supported(d, d);
}

bool chol_eligible(int64_t d, int normalization) {
    // row L1 norms are not invariant under an orthogonal change of basis, so only l2 / none keep the loop equivariant
    return g_opt_chol.load() && chol_whiten_supported(d) &&
           (normalization == CLEORA_NORM_L2_NUMPY || normalization == CLEORA_NORM_NONE);
}

// T (device, f32 d x d) from ws.cov through the eigensolver (host callback or cuSOLVER), enqueued on / synchronising `st`.
void pca_transform(WhitenState &ws, int64_t d, cudaStream_t st) {
    if (current_eigh().fn) {
        CUDA_TRY(cudaMemcpyAsync(ws.h_cov.data(), ws.cov.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        transform_from_cov(ws.h_cov.data(), d, d, ws.h_T.data());
        CUDA_TRY(cudaMemcpyAsync(ws.T.p, ws.h_T.data(), sizeof(float) * d * d, cudaMemcpyHostToDevice, st));
    } else {
        ws.eig.transform(ws.cov.p, d, d, ws.T.p, st);
    }
}

// One attempt of the pipelined loop.  inner_chol: iterations 1 .. T-1 whiten with the Cholesky factor (one-SM kernel
// on the side stream, no host involvement); otherwise with the eigensolver as in round 1.  Returns false when a
// Cholesky step flagged its covariance (the caller repeats the loop with the eigensolver; `cur` is not written
// before the flag has been read).
bool embed_pipelined_once(DeviceGraph &dg, const float *val, const float *rowscale, float *cur, float *y, float *w,
                          float *y2, int64_t n, int64_t d, int64_t iters, WhitenState &ws, SecondStream &B, Phase &ph,
                          bool inner_chol) {
    cudaStream_t A = nullptr;
    if (inner_chol) ws.clear_status(A);
    ph.begin();
    launch_spmm(dg, val, cur, d, y, nullptr, 1.f, 0.f, CLEORA_NORM_L2_NUMPY, A);
    ph.end(PH_SPMM);
    ph.begin();
    stats_device(y, n, d, ws, A);
    ph.end(PH_STATS);
    const bool host_eigh = !inner_chol && current_eigh().fn != nullptr;   // e.g. numpy's LAPACK, on the CPU beside the SpMM
    for (int64_t it = 1; it < iters; ++it) {
        CUDA_TRY(cudaEventRecord(B.stats_done, A));
        CUDA_TRY(cudaStreamWaitEvent(B.s, B.stats_done, 0));
        if (host_eigh) {
            CUDA_TRY(cudaMemcpyAsync(ws.h_cov.data(), ws.cov.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost, B.s));
            ph.begin();
            launch_spmm(dg, val, y, d, w, nullptr, 1.f, 0.f, CLEORA_NORM_NONE, A);  // W = A Y, enqueued before the host blocks
            ph.end(PH_SPMM);
            ph.begin_on(B.s);
            CUDA_TRY(cudaStreamSynchronize(B.s));                                  // covariance on the host
            transform_from_cov(ws.h_cov.data(), d, d, ws.h_T.data());              // CPU eigensolve || GPU SpMM
            CUDA_TRY(cudaMemcpyAsync(ws.T.p, ws.h_T.data(), sizeof(float) * d * d, cudaMemcpyHostToDevice, B.s));
            ph.end_on(PH_EIGH, B.s);
            CUDA_TRY(cudaEventRecord(B.t_ready, B.s));
        } else {
            ph.begin_on(B.s);
            if (inner_chol) launch_chol_whiten(ws.cov.p, d, ws.T.p, ws.status.p, B.s);   // T = L^-T || SpMM
            else ws.eig.transform(ws.cov.p, d, d, ws.T.p, B.s);                          // eigensolve || SpMM
            ph.end_on(PH_EIGH, B.s);
            CUDA_TRY(cudaEventRecord(B.t_ready, B.s));
            ph.begin();
            launch_spmm(dg, val, y, d, w, nullptr, 1.f, 0.f, CLEORA_NORM_NONE, A);  // W = A Y
            ph.end(PH_SPMM);
        }
        CUDA_TRY(cudaStreamWaitEvent(A, B.t_ready, 0));
        ph.begin();
        launch_whiten_apply_tc(w, n, d, ws.mean32.p, ws.T.p, d, y2, CLEORA_NORM_L2_NUMPY, rowscale, A, nullptr, inner_chol);
        ph.end(PH_APPLY);
        ph.begin();
        stats_device(y2, n, d, ws, A);
        ph.end(PH_STATS);
        std::swap(y, y2);
    }
    if (inner_chol && !ws.status_ok(A)) return false;
    ph.begin();
    pca_transform(ws, d, A);                                               // the iterate that leaves the loop: PCA
    ph.end(PH_EIGH);
    ph.begin();
    launch_whiten_apply(y, n, d, ws.mean32.p, ws.T.p, d, cur, A);          // X_T = (Y - 1 mu^T) T
    ph.end(PH_APPLY);
    CUDA_TRY(cudaStreamSynchronize(A));
    CUDA_TRY(cudaStreamSynchronize(B.s));
    return true;
}

void embed_pipelined(DeviceGraph &dg, const float *val, int markov, float *cur, float *y, float *w, float *y2, int64_t n,
                     int64_t d, int64_t iters, WhitenState &ws, Phase &ph, float **result) {
    PerDevice &pd = per_device();
    if (!pd.side) pd.side = std::make_shared<SecondStream>();
    ws.ensure(d, d);
    const float *rowscale = row_scale_of(dg, markov);
    const bool chol = chol_eligible(d, CLEORA_NORM_L2_NUMPY);
    if (!embed_pipelined_once(dg, val, rowscale, cur, y, w, y2, n, d, iters, ws, *pd.side, ph, chol))
        embed_pipelined_once(dg, val, rowscale, cur, y, w, y2, n, d, iters, ws, *pd.side, ph, false);
    *result = cur;
}

#include "multi_gpu.inl"

}  // namespace
}  // namespace cleora

using namespace cleora;

// ================================================================================================ misc
extern "C" const char *cleora_last_error(void) { return t_err.c_str(); }
extern "C" const char *cleora_version(void) { return "cleora_b200 0.1 (sm_100a)"; }
extern "C" int cleora_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int cleora_set_device(int device) {
    return guarded([&] { require_device(); CUDA_TRY(cudaSetDevice(device)); });
}
extern "C" uint64_t cleora_hash_entity(const char *bytes, int64_t len) { return xxh64(bytes, (size_t)len, 0); }
extern "C" void cleora_set_eigh(cleora_eigh_fn fn, void *user) { g_eigh_user.store(user); g_eigh_fn.store(fn); }
extern "C" void cleora_set_eigh_thread(int mode, cleora_eigh_fn fn, void *user) {
    t_eigh_mode = mode ? 1 : 0;
    t_eigh = EighChoice{mode ? fn : nullptr, mode ? user : nullptr};
}
extern "C" int cleora_set_option(const char *key, int64_t value) {
    return guarded([&] {
        const std::string k = key ? key : "";
        if (k == "pipeline_whiten") g_opt_pipeline.store(value != 0);
        else if (k == "chol_whiten") g_opt_chol.store(value != 0);
        else if (k == "k3_asw") g_k3_asw.store(value != 0);
        else if (k == "gram_needed_cols") g_gram_needed_only.store(value != 0);
        else if (k == "k3_bk") {
            if (value != 16 && value != 32) value_error("k3_bk must be 16 or 32");
            g_k3_bk.store((int)value);
        }
        else value_error("unknown option '" + k + "'");
    });
}
extern "C" int64_t cleora_get_option(const char *key) {
    const std::string k = key ? key : "";
    if (k == "pipeline_whiten") return g_opt_pipeline.load();
    if (k == "chol_whiten") return g_opt_chol.load();
    if (k == "k3_bk") return g_k3_bk.load();
    if (k == "k3_asw") return g_k3_asw.load();
    if (k == "gram_needed_cols") return g_gram_needed_only.load();
    return -1;
}
extern "C" int cleora_host_alloc(size_t nbytes, void **out) {
    return guarded([&] { require_device(); CUDA_TRY(cudaMallocHost(out, nbytes ? nbytes : 1)); });
}
extern "C" void cleora_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int64_t cleora_dev_workspace_bytes(void) {
    int64_t t = (int64_t)workspace().bytes();
    for (auto &b : persistent().buf) t += (int64_t)(b.n * sizeof(float));
    return t;
}
extern "C" int cleora_release_workspace(void) {
    return guarded([&] {
        persistent().release();
        workspace().release();
    });
}
extern "C" int64_t cleora_kernel_launch_count(void) { return g_launches.load(); }

// ================================================================================================ graph
extern "C" int cleora_graph_from_lines(const char *buf, const int64_t *offsets, int64_t n_lines, const char *columns,
                                       int64_t trim_n, cleora_graph_t **out) {
    return guarded([&] {
        auto g = build_from_lines(buf, offsets, n_lines, columns, trim_n);
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" int cleora_graph_from_files(const char *const *paths, int64_t n_paths, const char *columns, int64_t trim_n,
                                       cleora_graph_t **out) {
    return guarded([&] {
        if (n_paths <= 0) value_error("At least one file path is required");
        std::vector<std::string> ps;
        for (int64_t i = 0; i < n_paths; ++i) {
            std::string p = paths[i];
            auto ends = [&](const char *suf) { size_t l = std::strlen(suf); return p.size() >= l && p.compare(p.size() - l, l, suf) == 0; };
            if (!ends(".tsv") && !ends(".csv") && !ends(".txt"))
                value_error("Unsupported file format: " + p + ". Supported: .tsv, .csv, .txt");
            ps.push_back(p);
        }
        auto g = build_from_files(ps, columns, trim_n);
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" int cleora_graph_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs, const char *column_name,
                                       cleora_graph_t **out) {
    return guarded([&] {
        auto g = build_from_pairs(u, v, n_pairs, column_name ? column_name : "node");
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" int cleora_graph_from_hyperedges(const uint32_t *members, const int64_t *offsets, int64_t n_lines,
                                            const char *columns, int64_t trim_n, cleora_graph_t **out) {
    return guarded([&] {
        if (n_lines < 0 || !offsets || (n_lines && offsets[0] != 0)) value_error("bad hyperedge offsets");
        for (int64_t i = 0; i < n_lines; ++i)
            if (offsets[i + 1] < offsets[i]) value_error("hyperedge offsets must be non-decreasing");
        auto g = build_from_hyperedges(members, offsets, n_lines, columns, trim_n);
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" int cleora_dev_graph_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs, const char *column_name,
                                           int shard_rank, int shard_world, int want_sym, void *stream,
                                           cleora_graph_t **out, int64_t *bounds_out) {
    return guarded([&] {
        require_device();
        std::vector<int64_t> bounds;
        auto g = build_from_pairs_device(u, v, n_pairs, column_name ? column_name : "node", shard_rank, shard_world,
                                         want_sym != 0, (cudaStream_t)stream, &bounds);
        attach_long_row_schedule(*g->devs[0], g->rowptr);
        if (bounds_out) std::copy(bounds.begin(), bounds.end(), bounds_out);
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" int cleora_dev_synth_pairs(int kind, int64_t n_nodes, int64_t n_pairs, uint64_t seed, double alpha,
                                      uint32_t *u, uint32_t *v, void *stream) {
    return guarded([&] { require_device(); synth_pairs_device(kind, n_nodes, n_pairs, seed, alpha, u, v, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_graph_hashes(cleora_graph_t *g, const uint64_t **hash, int64_t *n_hash) {
    return guarded([&] {
        DeviceGraph &dg = device_graph(*g);
        *hash = dg.hash;
        *n_hash = dg.hash_rows ? dg.hash_rows : (dg.hash ? g->n_rows : 0);
    });
}
extern "C" int64_t cleora_graph_num_entities_global(const cleora_graph_t *g) { return g->device_only ? g->n_global : g->n_rows; }
extern "C" int cleora_graph_from_csr(const int64_t *rowptr, const uint32_t *col, const float *val_left,
                                     const float *val_sym, const float *row_sum, const uint64_t *entity_hash,
                                     int64_t n_rows, int64_t n_cols, int64_t row_offset, cleora_graph_t **out) {
    return guarded([&] {
        if (n_rows < 0 || n_cols < 0 || !rowptr) value_error("bad CSR shape");
        if (rowptr[0] != 0) value_error("rowptr[0] must be 0");
        for (int64_t r = 0; r < n_rows; ++r)
            if (rowptr[r + 1] < rowptr[r]) value_error("rowptr must be non-decreasing");
        const int64_t nnz = rowptr[n_rows];
        for (int64_t k = 0; k < nnz; ++k)
            if ((int64_t)col[k] >= n_cols) value_error("column index out of range");
        auto g = std::make_unique<Graph>();
        g->n_rows = n_rows; g->n_cols = n_cols; g->row_offset = row_offset;
        g->rowptr.assign(rowptr, rowptr + n_rows + 1);
        g->col.assign(col, col + nnz);
        g->left.assign(val_left, val_left + nnz);
        if (val_sym) g->sym.assign(val_sym, val_sym + nnz);
        if (row_sum) g->row_sum.assign(row_sum, row_sum + n_rows);
        if (entity_hash) g->hash.assign(entity_hash, entity_hash + n_rows);
        g->column_id.assign((size_t)n_rows, 0);
        *out = static_cast<cleora_graph_t *>(g.release());
    });
}
extern "C" void cleora_graph_destroy(cleora_graph_t *g) {
    if (!g) return;
    for (DeviceGraph *dg : g->devs) free_device_graph(dg);
    g->devs.clear();
    if (g->host_pinned) {
        for (const void *p : {(const void *)g->rowptr.data(), (const void *)g->col.data(), (const void *)g->left.data(),
                              (const void *)g->sym.data()})
            if (p && cudaHostUnregister(const_cast<void *>(p)) != cudaSuccess) cudaGetLastError();
    }
    delete static_cast<Graph *>(g);
}
extern "C" int cleora_graph_release_device(cleora_graph_t *g) {
    return guarded([&] {
        std::lock_guard<std::recursive_mutex> lock(g->mu);
        for (DeviceGraph *dg : g->devs) free_device_graph(dg);
        g->devs.clear();
    });
}
// Copy the host CSR into the device image again (same buffers: no allocation, no long-row re-scan) -- the per-step
// "inputs arrive from the host" leg of an end-to-end measurement or of a serving loop that reuses its allocations.
// Creates the image when the current device has none yet.
extern "C" int cleora_graph_refresh_device(cleora_graph_t *g, void *stream) {
    return guarded([&] {
        DeviceGraph &dg = device_graph(*g);
        cudaStream_t st = (cudaStream_t)stream;
        const size_t nnz = (size_t)g->nnz();
        CUDA_TRY(cudaMemcpyAsync(dg.rowptr, g->rowptr.data(), g->rowptr.size() * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dg.col, g->col.data(), nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(dg.left, g->left.data(), nnz * sizeof(float), cudaMemcpyHostToDevice, st));
        if (dg.hash) CUDA_TRY(cudaMemcpyAsync(dg.hash, g->hash.data(), g->hash.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        std::lock_guard<std::mutex> lock(dg.lazy_mu);
        if (dg.sym) CUDA_TRY(cudaMemcpyAsync(dg.sym, g->sym.data(), nnz * sizeof(float), cudaMemcpyHostToDevice, st));
    });
}
extern "C" int64_t cleora_graph_num_entities(const cleora_graph_t *g) { return g->n_rows; }
extern "C" int64_t cleora_graph_num_cols(const cleora_graph_t *g) { return g->n_cols; }
extern "C" int64_t cleora_graph_num_edges(const cleora_graph_t *g) { return g->nnz(); }
extern "C" int cleora_graph_copy_csr(const cleora_graph_t *g, int64_t *rowptr, uint32_t *col, float *left, float *sym) {
    return guarded([&] { materialize_host(*const_cast<cleora_graph_t *>(g));
        if (rowptr) std::copy(g->rowptr.begin(), g->rowptr.end(), rowptr);
        if (col) std::copy(g->col.begin(), g->col.end(), col);
        if (left) std::copy(g->left.begin(), g->left.end(), left);
        if (sym) {
            if (g->sym.empty() && g->nnz()) value_error("graph has no symmetric values");
            std::copy(g->sym.begin(), g->sym.end(), sym);
        }
    });
}
extern "C" int cleora_graph_copy_row_sums(const cleora_graph_t *g, float *out) {
    return guarded([&] { materialize_host(*const_cast<cleora_graph_t *>(g)); std::copy(g->row_sum.begin(), g->row_sum.end(), out); });
}
extern "C" int cleora_graph_copy_entity_hashes(const cleora_graph_t *g, uint64_t *out) {
    return guarded([&] { materialize_host(*const_cast<cleora_graph_t *>(g)); std::copy(g->hash.begin(), g->hash.end(), out); });
}
extern "C" int cleora_graph_copy_column_ids(const cleora_graph_t *g, uint8_t *out) {
    return guarded([&] { materialize_host(*const_cast<cleora_graph_t *>(g)); std::copy(g->column_id.begin(), g->column_id.end(), out); });
}
extern "C" int64_t cleora_graph_entity_ids_nbytes(const cleora_graph_t *g) {
    try { materialize_host(*const_cast<cleora_graph_t *>(g)); } catch (...) { return 0; }
    int64_t t = 0;
    for (const auto &s : g->ids) t += (int64_t)s.size();
    return t;
}
extern "C" int cleora_graph_copy_entity_ids(const cleora_graph_t *g, char *buf, int64_t *offsets) {
    return guarded([&] { materialize_host(*const_cast<cleora_graph_t *>(g));
        int64_t pos = 0;
        offsets[0] = 0;
        for (size_t i = 0; i < g->ids.size(); ++i) {
            std::memcpy(buf + pos, g->ids[i].data(), g->ids[i].size());
            pos += (int64_t)g->ids[i].size();
            offsets[i + 1] = pos;
        }
    });
}
extern "C" int cleora_graph_set_entity_ids(cleora_graph_t *g, const char *buf, const int64_t *offsets, int64_t n) {
    return guarded([&] {
        if (n != g->n_rows) value_error("entity_ids must keep its length (" + std::to_string(g->n_rows) + ")");
        materialize_host(*g);
        std::lock_guard<std::recursive_mutex> lock0(g->mu);
        g->ids.resize((size_t)n);
        g->hash.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            g->ids[(size_t)i].assign(buf + offsets[i], (size_t)(offsets[i + 1] - offsets[i]));
            g->hash[(size_t)i] = xxh64(buf + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), 0);
        }
        g->id_index.clear();
        std::lock_guard<std::recursive_mutex> lock(g->mu);
        for (DeviceGraph *dg : g->devs)
            if (dg->hash)                   // (cudaMemcpy with a device pointer needs no cudaSetDevice under UVA)
                CUDA_TRY(cudaMemcpy(dg->hash, g->hash.data(), sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice));
    });
}
extern "C" int cleora_graph_set_descriptor(cleora_graph_t *g, int col_a_id, const char *col_a_name, int col_b_id,
                                           const char *col_b_name) {
    return guarded([&] {
        g->desc.col_a_id = col_a_id; g->desc.col_b_id = col_b_id;
        g->desc.col_a_name = col_a_name ? col_a_name : ""; g->desc.col_b_name = col_b_name ? col_b_name : "";
    });
}
extern "C" int cleora_graph_set_column_ids(cleora_graph_t *g, const uint8_t *ids, int64_t n) {
    return guarded([&] {
        if (n != g->n_rows) value_error("column_ids must have one entry per entity");
        g->column_id.assign(ids, ids + n);
    });
}
extern "C" const char *cleora_graph_col_name(const cleora_graph_t *g, int which) {
    return which ? g->desc.col_b_name.c_str() : g->desc.col_a_name.c_str();
}
extern "C" int cleora_graph_col_id(const cleora_graph_t *g, int which) { return which ? g->desc.col_b_id : g->desc.col_a_id; }
extern "C" int64_t cleora_graph_find_entity(const cleora_graph_t *cg, const char *id, int64_t id_len) {
    Graph *g = const_cast<cleora_graph_t *>(cg);
    try { materialize_host(*g); } catch (...) { return -1; }
    std::lock_guard<std::recursive_mutex> lock(g->mu);      // the index is built lazily; lookups may come from several threads
    if (g->id_index.empty() && !g->ids.empty())
        for (size_t i = 0; i < g->ids.size(); ++i) g->id_index.emplace(g->ids[i], (int64_t)i);   // first wins
    auto it = g->id_index.find(std::string(id, (size_t)id_len));
    return it == g->id_index.end() ? -1 : it->second;
}

// ================================================================================================ device-level API
extern "C" int cleora_dev_graph_prepare(cleora_graph_t *g) { return guarded([&] { device_graph(*g); }); }
extern "C" int cleora_dev_init(const uint64_t *hash, int64_t n, int64_t d, int64_t seed, float *out, void *stream) {
    return guarded([&] { launch_init(hash, n, d, seed, out, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_spmm(cleora_graph_t *g, int markov, const float *x, int64_t d, float *out,
                               const float *resid, float alpha, float rw, int normalization, void *stream) {
    return guarded([&] {
        check_norm(normalization);
        DeviceGraph &dg = device_graph(*g);
        launch_spmm(dg, values_of(dg, markov), x, d, out, resid, alpha, rw, normalization, (cudaStream_t)stream);
    });
}
static PeerOut make_peers(float *const *extra, int n_extra) {
    if (n_extra < 0 || n_extra > 7) value_error("at most 7 extra destinations");
    PeerOut p{};
    p.n_extra = n_extra;
    for (int i = 0; i < n_extra; ++i) p.extra[i] = extra[i];
    return p;
}
extern "C" int cleora_dev_spmm_push(cleora_graph_t *g, int markov, const float *x, int64_t d, float *out,
                                    float *const *extra_outs, int n_extra, const float *resid, float alpha, float rw,
                                    int normalization, void *stream) {
    return guarded([&] {
        check_norm(normalization);
        DeviceGraph &dg = device_graph(*g);
        const PeerOut peers = make_peers(extra_outs, n_extra);
        launch_spmm(dg, values_of(dg, markov), x, d, out, resid, alpha, rw, normalization, (cudaStream_t)stream, &peers);
    });
}
extern "C" int cleora_dev_whiten_apply_push(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                            int64_t dout, float *out, float *const *extra_outs, int n_extra,
                                            int normalization, const float *rowscale, void *stream) {
    return guarded([&] {
        if (!whiten_apply_tc_supported(d, dout)) value_error("fused apply needs d % 32 == 0 and dout % 32 == 0, dout <= 256 (or dout % 64 == 0, dout <= 512)");
        const PeerOut peers = make_peers(extra_outs, n_extra);
        launch_whiten_apply_tc(x, n, d, mean_f32, T, dout, out, normalization, rowscale, (cudaStream_t)stream, &peers);
    });
}
// ---- column-sharded multi-GPU loop: the two all-to-alls are fused into the producers' epilogues ------------------
static PeerOut make_dests(float *const *dests, int n_dst, int mode) {
    if (n_dst < 1 || n_dst > 8) value_error("1 to 8 destinations");
    PeerOut p{};
    p.n_extra = n_dst;
    p.mode = mode;
    for (int i = 0; i < n_dst; ++i) {
        if (!dests[i]) value_error("null destination");
        p.extra[i] = dests[i];
    }
    return p;
}
extern "C" int cleora_dev_spmm_scatter(cleora_graph_t *g, int markov, const float *x, int64_t d, float *const *dests,
                                       int n_dst, int64_t block_rows, int64_t ld_cols, int64_t col_off, const float *resid,
                                       float alpha, float rw, void *stream) {
    return guarded([&] {
        DeviceGraph &dg = device_graph(*g);
        if (block_rows <= 0 || (dg.n_rows + block_rows - 1) / block_rows > n_dst) value_error("row blocks do not cover the graph");
        if (col_off < 0 || col_off % 4 != 0 || ld_cols % 4 != 0 || col_off + d > ld_cols) value_error("bad destination columns");
        PeerOut peers = make_dests(dests, n_dst, PEER_OWNERS);
        peers.block_rows = block_rows; peers.ld_cols = ld_cols; peers.col_off = (int)col_off;
        launch_spmm(dg, values_of(dg, markov), x, d, nullptr, resid, alpha, rw, CLEORA_NORM_NONE, (cudaStream_t)stream, &peers);
    });
}
extern "C" int cleora_dev_whiten_apply_slices(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                              int64_t dout, float *out, float *const *dests, int n_dst, int64_t row_base,
                                              int normalization, const float *rowscale, int t_upper, void *stream) {
    return guarded([&] {
        if (!whiten_apply_tc_supported(d, dout)) value_error("fused apply needs the tensor-core shape rules (see cleora_whiten_apply_fusable)");
        if (dout % n_dst != 0 || (dout / n_dst) % 4 != 0) value_error("column slices must be multiples of 4 columns");
        PeerOut peers = make_dests(dests, n_dst, PEER_SLICES);
        peers.slice_cols = (int)(dout / n_dst); peers.row_base = row_base;
        launch_whiten_apply_tc(x, n, d, mean_f32, T, dout, out, normalization, rowscale, (cudaStream_t)stream, &peers, t_upper != 0);
    });
}
extern "C" int cleora_dev_normalize_slices(const float *x, int64_t n, int64_t d, int normalization, float *out,
                                           float *const *dests, int n_dst, int64_t row_base, void *stream) {
    return guarded([&] {
        check_norm(normalization);
        if (!normalize_rows_supported(d)) value_error("feature dimension has no vectorised row mapping");
        if (!out) value_error("the local copy is required");
        if (n_dst == 0) {                                           // local rows only
            launch_normalize_rows(x, n, d, normalization, out, (cudaStream_t)stream, nullptr);
            return;
        }
        if (d % n_dst != 0 || (d / n_dst) % 4 != 0) value_error("column slices must be multiples of 4 columns");
        PeerOut peers = make_dests(dests, n_dst, PEER_SLICES);
        peers.slice_cols = (int)(d / n_dst); peers.row_base = row_base;
        launch_normalize_rows(x, n, d, normalization, out, (cudaStream_t)stream, &peers);
    });
}
// ---- peer memory plumbing (CUDA IPC): buffers that other ranks' kernels write into ------------------------------
extern "C" int cleora_dev_malloc(size_t nbytes, void **out) {
    return guarded([&] { require_device(); CUDA_TRY(cudaMalloc(out, nbytes ? nbytes : 1)); });
}
extern "C" int cleora_dev_free(void *p) { return guarded([&] { if (p) CUDA_TRY(cudaFree(p)); }); }
extern "C" int cleora_ipc_get_handle(void *p, unsigned char *handle64) {
    return guarded([&] {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
        cudaIpcMemHandle_t h;
        CUDA_TRY(cudaIpcGetMemHandle(&h, p));
        std::memcpy(handle64, &h, 64);
    });
}
extern "C" int cleora_ipc_open(const unsigned char *handle64, void **out) {
    return guarded([&] {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handle64, 64);
        CUDA_TRY(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    });
}
extern "C" int cleora_ipc_close(void *p) { return guarded([&] { if (p) CUDA_TRY(cudaIpcCloseMemHandle(p)); }); }
extern "C" int cleora_dev_normalize(const float *x, int64_t n, int64_t d, int normalization, float *out, void *stream) {
    return guarded([&] { check_norm(normalization); launch_normalize(x, n, d, normalization, out, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_col_sums(const float *x, int64_t n, int64_t d, double *sums, int accumulate, void *stream) {
    return guarded([&] { launch_col_sums(x, n, d, sums, accumulate != 0, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_centered_gram(const float *x, int64_t n, int64_t d, const double *mean, double *cov,
                                        void *stream) {
    return guarded([&] { launch_centered_gram(x, n, d, mean, cov, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_whiten_apply(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                       int64_t dout, float *out, void *stream) {
    return guarded([&] { launch_whiten_apply(x, n, d, mean_f32, T, dout, out, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_whiten_apply_ex(const float *x, int64_t n, int64_t d, const float *mean_f32, const float *T,
                                          int64_t dout, float *out, int normalization, const float *rowscale,
                                          int t_upper, void *stream) {
    return guarded([&] {
        if (normalization == CLEORA_NORM_NONE && rowscale == nullptr) {
            launch_whiten_apply(x, n, d, mean_f32, T, dout, out, (cudaStream_t)stream);
            return;
        }
        if (!whiten_apply_tc_supported(d, dout)) value_error("fused apply needs d % 32 == 0 and dout % 32 == 0, dout <= 256 (or dout % 64 == 0, dout <= 512)");
        launch_whiten_apply_tc(x, n, d, mean_f32, T, dout, out, normalization, rowscale, (cudaStream_t)stream, nullptr, t_upper != 0);
    });
}
extern "C" int cleora_dev_row_scale(cleora_graph_t *g, int markov, float *out, void *stream) {
    return guarded([&] {
        DeviceGraph &dg = device_graph(*g);
        launch_row_value_sums(dg.rowptr, values_of(dg, markov), dg.n_rows, out, (cudaStream_t)stream);
    });
}
extern "C" int cleora_whiten_apply_fusable(int64_t d, int64_t dout) { return whiten_apply_tc_supported(d, dout) ? 1 : 0; }
extern "C" int cleora_dev_sq_diff_sum(const float *a, const float *b, int64_t n, int f64_diff, double *result,
                                      void *stream) {
    return guarded([&] { launch_sq_diff_sum(a, b, n, f64_diff != 0, result, (cudaStream_t)stream); });
}
extern "C" int cleora_dev_whiten_transform(const double *cov, int64_t d, int64_t dout, float *T, void *stream) {
    return guarded([&] {
        if (dout <= 0 || dout > d) value_error("n_components must be in [1, d]");
        cudaStream_t st = (cudaStream_t)stream;
        if (current_eigh().fn) {
            std::vector<double> h((size_t)d * d);
            std::vector<float> hT((size_t)d * dout);
            CUDA_TRY(cudaMemcpyAsync(h.data(), cov, sizeof(double) * d * d, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            transform_from_cov(h.data(), d, dout, hT.data());
            CUDA_TRY(cudaMemcpyAsync(T, hT.data(), sizeof(float) * d * dout, cudaMemcpyHostToDevice, st));
            CUDA_TRY(cudaStreamSynchronize(st));
        } else {
            per_device().eig_misc.transform(cov, d, dout, T, st);
        }
    });
}
extern "C" int cleora_dev_chol_whiten(const double *cov, int64_t d, float *T, int *status, void *stream) {
    return guarded([&] {
        if (!chol_whiten_supported(d)) value_error("Cholesky whitening needs 1 <= d <= 512");
        launch_chol_whiten(cov, d, T, status, (cudaStream_t)stream);
    });
}
extern "C" int cleora_whiten_transform_from_cov(const double *cov, int64_t d, int64_t dout, float *T) {
    return guarded([&] {
        if (dout <= 0 || dout > d) value_error("n_components must be in [1, d]");
        transform_from_cov(cov, d, dout, T);
    });
}

// ================================================================================================ host-buffer API
extern "C" int cleora_initialize_deterministically(cleora_graph_t *g, int64_t d, int64_t seed, float *out) {
    return guarded([&] {
        if (d < 0) value_error("feature_dim must be non-negative");
        DeviceGraph &dg = device_graph(*g);
        if (!dg.hash && g->n_rows) value_error("graph has no entity hashes");
        const size_t cnt = (size_t)g->n_rows * (size_t)d;
        DevBuf<float> x(cnt);
        launch_init(dg.hash, g->n_rows, d, seed, x.p, nullptr);
        CUDA_TRY(cudaMemcpy(out, x.p, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

extern "C" int cleora_markov_propagate(cleora_graph_t *g, const float *x, int64_t x_rows, int64_t d, int markov,
                                       float *out) {
    return guarded([&] {
        if (x_rows != g->n_cols)                                                       // src/lib.rs:38-43
            value_error("Embedding matrix has " + std::to_string(x_rows) + " rows but graph has " +
                        std::to_string(g->n_cols) + " entities");
        DeviceGraph &dg = device_graph(*g);
        const float *val = values_of(dg, markov);
        DevBuf<float> dx((size_t)x_rows * d), dout((size_t)g->n_rows * d);
        CUDA_TRY(cudaMemcpy(dx.p, x, dx.n * sizeof(float), cudaMemcpyHostToDevice));
        launch_spmm(dg, val, dx.p, d, dout.p, nullptr, 1.f, 0.f, CLEORA_NORM_NONE, nullptr);
        CUDA_TRY(cudaMemcpy(out, dout.p, dout.n * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

extern "C" int cleora_l2_normalize(const float *x, int64_t n, int64_t d, float *out) {
    return guarded([&] {
        require_device();
        DevBuf<float> dx((size_t)n * d), dy((size_t)n * d);
        CUDA_TRY(cudaMemcpy(dx.p, x, dx.n * sizeof(float), cudaMemcpyHostToDevice));
        launch_normalize(dx.p, n, d, CLEORA_NORM_L2_RUST, dy.p, nullptr);
        CUDA_TRY(cudaMemcpy(out, dy.p, dy.n * sizeof(float), cudaMemcpyDeviceToHost));
    });
}

// embed_full / embed_full_with_convergence (src/embedding.rs:106-188) on the device: X stays in HBM.
static void embed_fast_impl(cleora_graph_t *g, int64_t d, int64_t iters, int markov, int64_t seed, float rw,
                            float threshold, bool with_conv, float *out, int64_t *iters_done) {
    if (g->n_rows != g->n_cols) value_error("embed needs a full (square) graph, not a row shard");
    if (d < 0 || iters < 0) value_error("feature_dim and num_iterations must be non-negative");
    DeviceGraph &dg = device_graph(*g);
    const float *val = values_of(dg, markov);
    if (!dg.hash && g->n_rows) value_error("graph has no entity hashes");
    const int64_t n = g->n_rows;
    const size_t cnt = (size_t)n * (size_t)d;
    Persistent &ps = persistent();
    DevBuf<double> dsum(1);
    float *src = ps.get(0, cnt), *dst = ps.get(1, cnt);
    launch_init(dg.hash, n, d, seed, src, nullptr);
    const bool use_res = rw > 0.0f && rw < 1.0f;                 // embedding.rs:116
    const float alpha = 1.0f - rw;
    const bool check = with_conv && threshold > 0.0f;
    int64_t actual = iters;
    for (int64_t it = 0; it < iters; ++it) {
        launch_spmm(dg, val, src, d, dst, use_res ? src : nullptr, alpha, rw, CLEORA_NORM_L2_RUST, nullptr);
        std::swap(src, dst);                                     // src = newest iterate, dst = previous
        if (check && it > 0) {
            launch_sq_diff_sum(src, dst, (int64_t)cnt, false, dsum.p, nullptr);
            double h = 0.0;
            CUDA_TRY(cudaMemcpy(&h, dsum.p, sizeof(double), cudaMemcpyDeviceToHost));
            const float rmse = std::sqrt((float)h / (float)(uint64_t)cnt);
            if (rmse < threshold) { actual = it + 1; break; }
        }
    }
    CUDA_TRY(cudaMemcpy(out, src, cnt * sizeof(float), cudaMemcpyDeviceToHost));
    if (iters_done) *iters_done = actual;
}

extern "C" int cleora_embed_fast(cleora_graph_t *g, int64_t d, int64_t iters, int markov, int64_t seed,
                                 float residual_weight, float *out) {
    return guarded([&] { embed_fast_impl(g, d, iters, markov, seed, residual_weight, 0.f, false, out, nullptr); });
}
extern "C" int cleora_embed_fast_convergence(cleora_graph_t *g, int64_t d, int64_t max_iters, int markov, int64_t seed,
                                             float residual_weight, float convergence_threshold, float *out,
                                             int64_t *iters_done) {
    return guarded([&] {
        embed_fast_impl(g, d, max_iters, markov, seed, residual_weight, convergence_threshold, true, out, iters_done);
    });
}

extern "C" int cleora_whiten_embeddings(const float *x, int64_t n, int64_t d, int64_t n_components, float *out) {
    return guarded([&] {
        require_device();
        if (n_components <= 0 || n_components > d) value_error("n_components must be in [1, feature dimension]");
        const int64_t dout = n_components;
        if (n <= 1) {                                            // pycleora/__init__.py:132-133
            if (n == 1) std::memcpy(out, x, sizeof(float) * (size_t)d);
            return;
        }
        DevBuf<float> dx((size_t)n * d), dz((size_t)n * dout);
        CUDA_TRY(cudaMemcpy(dx.p, x, dx.n * sizeof(float), cudaMemcpyHostToDevice));
        WhitenState &ws = persistent().ws;
        Phase ph;
        whiten_device(dx.p, n, d, dout, dz.p, ws, nullptr, ph, false, true);      // arbitrary user data: IEEE f64 covariance
        CUDA_TRY(cudaMemcpy(out, dz.p, dz.n * sizeof(float), cudaMemcpyDeviceToHost));
        ws.eig.check_info();
    });
}

// normalization="spectral" (pycleora/__init__.py:951-956): rows are l2-normalised (the caller does that in the loop) and
// replaced by U*S of their SVD, i.e. rotated by the right singular vectors = the eigenvectors of X^T X in descending
// order.  The loop body is equivariant under that rotation, so the binding applies it once, to the iterate that leaves
// the loop (cleora_b200/__init__.py).  x, out: host [n, d].
extern "C" int cleora_spectral_rotate(const float *x, int64_t n, int64_t d, float *out) {
    return guarded([&] {
        require_device();
        if (n == 0 || d == 0) return;
        DevBuf<float> dx((size_t)n * d), dz((size_t)n * d);
        CUDA_TRY(cudaMemcpy(dx.p, x, dx.n * sizeof(float), cudaMemcpyHostToDevice));
        WhitenState &ws = persistent().ws;
        ws.ensure(d, d);
        CUDA_TRY(cudaMemsetAsync(ws.sums.p, 0, sizeof(double) * d, nullptr));            // centre = 0: the Gram matrix X^T X
        CUDA_TRY(cudaMemsetAsync(ws.mean32.p, 0, sizeof(float) * d, nullptr));
        launch_centered_gram(dx.p, n, d, ws.sums.p, ws.cov.p, nullptr, nullptr);
        if (current_eigh().fn) {
            CUDA_TRY(cudaMemcpy(ws.h_cov.data(), ws.cov.p, sizeof(double) * d * d, cudaMemcpyDeviceToHost));
            transform_from_cov(ws.h_cov.data(), d, d, ws.h_T.data(), false);
            CUDA_TRY(cudaMemcpy(ws.T.p, ws.h_T.data(), sizeof(float) * d * d, cudaMemcpyHostToDevice));
        } else {
            ws.eig.transform(ws.cov.p, d, d, ws.T.p, nullptr, false);
        }
        launch_whiten_apply(dx.p, n, d, ws.mean32.p, ws.T.p, d, dz.p, nullptr);
        CUDA_TRY(cudaMemcpy(out, dz.p, dz.n * sizeof(float), cudaMemcpyDeviceToHost));
        ws.eig.check_info();
    });
}

// The Python loop of embed() (pycleora/__init__.py:97-125), device-resident.
extern "C" int cleora_embed(cleora_graph_t *g, const float *x0, int64_t d, int64_t iters, int markov, int64_t seed,
                            double residual_weight, double convergence_threshold, int normalization, int whiten,
                            float *out, int64_t *iters_done, double *timings_ms) {
    return guarded([&] {
        check_norm(normalization);
        if (g->n_rows != g->n_cols) value_error("embed needs a full (square) graph, not a row shard");
        if (d < 0 || iters < 0) value_error("feature_dim and num_iterations must be non-negative");
        DeviceGraph &dg = device_graph(*g);
        const float *val = values_of(dg, markov);
        const int64_t n = g->n_rows;
        const size_t cnt = (size_t)n * (size_t)d;
        Phase ph;
        ph.on = timings_ms != nullptr;
        if (timings_ms) std::fill(timings_ms, timings_ms + 8, 0.0);
        const bool conv = convergence_threshold > 0.0;
        const bool do_whiten = whiten != 0 && n > 1;
        Persistent &ps = persistent();
        DevBuf<double> dsum(1);
        float *cur = ps.get(0, cnt), *y = ps.get(1, cnt), *w = (conv && do_whiten) ? ps.get(2, cnt) : nullptr;
        auto load_x0 = [&] {
            if (x0) {
                ph.begin();
                CUDA_TRY(cudaMemcpyAsync(cur, x0, cnt * sizeof(float), cudaMemcpyDefault, nullptr));   // host or device
                ph.end(PH_H2D);
            } else {
                if (!dg.hash && n) value_error("graph has no entity hashes");
                ph.begin();
                launch_init(dg.hash, n, d, seed, cur, nullptr);
                ph.end(PH_INIT);
            }
        };
        load_x0();
        const bool use_res = residual_weight > 0.0;              // pycleora/__init__.py:114 (no < 1 guard)
        const float alpha = (float)(1.0 - residual_weight), rwf = (float)residual_weight;
        WhitenState &ws = ps.ws;
        int64_t done = 0;
        float *result = cur;
        const bool pipelined = pipeline_eligible(n, d, iters, normalization, whiten, residual_weight, convergence_threshold);
        if (pipelined) {
            embed_pipelined(dg, val, markov, cur, y, ps.get(3, cnt), ps.get(4, cnt), n, d, iters, ws, ph, &result);
            done = iters;
        }
        // Reference stage order.  Without the rmse early stop no intermediate iterate is visible to the caller, so
        // iterations 0 .. T-2 may whiten with the Cholesky factor (chol_whiten.cu: same final iterate, no eigensolve);
        // with it, every iterate is compared element-wise with its predecessor and keeps the PCA basis.
        auto reference_order = [&](bool inner_chol) -> bool {
            if (inner_chol) { ws.ensure(d, d); ws.clear_status(nullptr); }
            for (int64_t it = 0; it < iters; ++it) {
                ph.begin();
                launch_spmm(dg, val, cur, d, y, use_res ? cur : nullptr, alpha, rwf, normalization, nullptr);
                ph.end(PH_SPMM);
                float *fresh;
                if (do_whiten) {
                    fresh = conv ? w : cur;                      // without rmse the old iterate can be overwritten
                    const bool last = it + 1 == iters;
                    if (inner_chol && last && !ws.status_ok(nullptr)) return false;     // before the last overwrite
                    whiten_device(y, n, d, d, fresh, ws, nullptr, ph, inner_chol && !last);
                } else {
                    fresh = y;
                }
                done = it + 1;
                bool stop = false;
                if (conv && it > 0) {                            // _compute_rmse, pycleora/__init__.py:974-976
                    ph.begin();
                    launch_sq_diff_sum(fresh, cur, (int64_t)cnt, true, dsum.p, nullptr);
                    double h = 0.0;
                    CUDA_TRY(cudaMemcpy(&h, dsum.p, sizeof(double), cudaMemcpyDeviceToHost));
                    ph.end(PH_RMSE);
                    stop = std::sqrt(h / (double)cnt) < convergence_threshold;
                }
                if (fresh == w) std::swap(cur, w);
                else if (fresh == y) std::swap(cur, y);
                result = cur;
                if (stop) break;
            }
            return true;
        };
        if (!pipelined) {
            const bool inner_chol = do_whiten && !conv && iters >= 2 && chol_eligible(d, normalization);
            if (!reference_order(inner_chol)) {                  // a covariance was not safely SPD: eigensolver throughout
                cur = ps.get(0, cnt); y = ps.get(1, cnt);
                load_x0();
                reference_order(false);
            }
        }
        ph.begin();
        CUDA_TRY(cudaMemcpyAsync(out, result, cnt * sizeof(float), cudaMemcpyDefault, nullptr));   // host or device
        ph.end(PH_D2H);
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        ws.eig.check_info();
        ph.collect(timings_ms);
        if (iters_done) *iters_done = done;
    });
}

// embed() on several GPUs of this box from ONE process and ONE call (multi_gpu.inl): the column-sharded loop with peer
// stores over NVLink, host threads instead of ranks, events instead of NCCL.  normalization == CLEORA_NORM_L2_RUST selects
// the Rust fast path's semantics (embed_full[_with_convergence], src/embedding.rs:106-188) as cleora_embed_fast does.
extern "C" int cleora_embed_multi(cleora_graph_t *g, const int *devices, int n_devices, const float *x0, int64_t d,
                                  int64_t iters, int markov, int64_t seed, double residual_weight,
                                  double convergence_threshold, int normalization, int whiten, float *out,
                                  int64_t *iters_done) {
    return guarded([&] {
        if (!devices || n_devices < 1) value_error("at least one device is required");
        MultiArgs a{};
        a.g = g; a.x0 = x0; a.d = d; a.iters = iters; a.seed = seed;
        a.markov = markov; a.norm = normalization; a.whiten = whiten;
        a.rust = normalization == CLEORA_NORM_L2_RUST ? 1 : 0;
        a.residual_weight = residual_weight; a.convergence_threshold = convergence_threshold;
        a.out = out;
        a.eigh = current_eigh();
        embed_multi(a, devices, n_devices, iters_done);
    });
}
extern "C" int cleora_embed_multi_supported(int64_t d, int n_devices) { return multi_eligible(d, n_devices) ? 1 : 0; }
