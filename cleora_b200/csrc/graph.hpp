// Internal host-side data model of libcleora_b200 (not part of the public ABI).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <unordered_map>
#include <memory>
#include <mutex>

namespace cleora {

// Column spec, reference: src/configuration.rs:12-17.
struct Column {
    std::string name;
    bool complex_ = false;
    bool reflexive = false;
};

// Reference: src/sparse_matrix.rs:48-54.
struct Descriptor {
    int col_a_id = 0, col_b_id = 0;
    std::string col_a_name, col_b_name;
};

struct DeviceGraph;   // defined in device.cuh

// Host image of the reference's SparseMatrix (src/sparse_matrix.rs:56-66) in SoA/CSR form:
// slices -> rowptr (explicit n+1, 64-bit), AoS Edge -> col / left / sym arrays.
struct Graph {
    Descriptor desc;
    int64_t n_rows = 0, n_cols = 0, row_offset = 0;
    std::vector<int64_t> rowptr;          // n_rows + 1
    std::vector<uint32_t> col;            // nnz
    std::vector<float> left, sym;         // nnz (sym may be empty for adopted CSR)
    std::vector<float> row_sum;           // n_rows  (Entity.row_sum)
    std::vector<uint64_t> hash;           // n_rows  (xxh64 of the CURRENT entity id)
    std::vector<uint8_t> column_id;       // n_rows
    std::vector<std::string> ids;         // n_rows (may be empty for adopted CSR)
    std::unordered_map<std::string, int64_t> id_index;   // built lazily by find()
    std::vector<DeviceGraph *> devs;      // lazily uploaded copies, one per device that used the graph; owned
    std::recursive_mutex mu;              // guards the lazy caches (devs and their lazily built members, id_index):
                                          // ctypes drops the GIL, so two Python threads may share one graph
    bool host_pinned = false;             // CSR arrays page-locked (cudaHostRegister) for fast (re-)uploads
    // Graphs built on the device (graph_dev.cu) keep col / left / sym / row_sum / hash in HBM only (host `rowptr` is
    // filled); the host accessors download them on first use (abi.cu: materialize_host).
    bool device_only = false;
    int64_t nnz_device = 0;               // nnz while the host arrays are not materialised
    int64_t n_global = 0, shard_r0 = 0;   // device-built shards: entity count of the whole graph, first global row
    int64_t nnz() const { return device_only && col.empty() ? nnz_device : (int64_t)col.size(); }
};

struct BuildError {
    std::string msg;
};

uint64_t xxh64(const void *data, size_t len, uint64_t seed);

// from_iterator / from_files semantics; throw BuildError for the reference's ValueError cases.
std::unique_ptr<Graph> build_from_lines(const char *buf, const int64_t *offsets, int64_t n_lines,
                                        const std::string &columns, int64_t trim_n);
std::unique_ptr<Graph> build_from_hyperedges(const uint32_t *members, const int64_t *offsets, int64_t n_lines,
                                             const std::string &columns, int64_t trim_n);
std::unique_ptr<Graph> build_from_files(const std::vector<std::string> &paths, const std::string &columns,
                                        int64_t trim_n);
std::unique_ptr<Graph> build_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs,
                                        const std::string &column_name);

}  // namespace cleora
