// Host-side graph construction for libcleora_b200: string hyperedges (or integer pairs) -> CSR with the
// reference's exact semantics.  Independent C++ implementation (the CPU oracle under oracle/ is a separate C
// restatement; the two are cross-checked bit-for-bit by tests/test_graph_build.py).
//
// Reference behaviour implemented here (paths relative to /root/reference):
//   column spec          src/configuration.rs:19-70      parse_fields / validate_column_modifiers
//   one relation only    src/sparse_matrix.rs:5-46       create_sparse_matrix_descriptor
//   line parsing         src/pipeline.rs:223-240         parse_line (TAB > comma > single column; split(' '))
//   entity hashing       src/entity.rs:109-114           XXH64 seed 0 of the UTF-8 bytes
//   entity indexing      src/sparse_matrix_builder.rs:59-70   first appearance, keyed by hash
//   hyperedge expansion  src/sparse_matrix_builder.rs:170-233 row_sum / value / trim / symmetric pairs
//   reduce               src/sparse_matrix_builder.rs:275-343 sort (row,col), left & symmetric Markov values
// Accumulation order = input order through ONE buffer (the reference's multi-buffer order is
// non-deterministic; results coincide whenever the f32 sums are exact).
#include "graph.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>

namespace cleora {

// ------------------------------------------------------------------------------------------------ XXH64
namespace {
constexpr uint64_t PR1 = 0x9E3779B185EBCA87ULL, PR2 = 0xC2B2AE3D27D4EB4FULL, PR3 = 0x165667B19E3779F9ULL,
                   PR4 = 0x85EBCA77C2B2AE63ULL, PR5 = 0x27D4EB2F165667C5ULL;
inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t load64(const unsigned char *p) { uint64_t v; std::memcpy(&v, p, 8); return v; }
inline uint32_t load32(const unsigned char *p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
inline uint64_t lane_round(uint64_t acc, uint64_t input) { return rotl(acc + input * PR2, 31) * PR1; }
inline uint64_t lane_merge(uint64_t h, uint64_t v) { return (h ^ lane_round(0, v)) * PR1 + PR4; }
}  // namespace

uint64_t xxh64(const void *data, size_t len, uint64_t seed) {
    const unsigned char *p = static_cast<const unsigned char *>(data);
    const unsigned char *const end = p + len;
    uint64_t h;
    if (len >= 32) {
        uint64_t a = seed + PR1 + PR2, b = seed + PR2, c = seed, d = seed - PR1;
        for (; p + 32 <= end; p += 32) {
            a = lane_round(a, load64(p));
            b = lane_round(b, load64(p + 8));
            c = lane_round(c, load64(p + 16));
            d = lane_round(d, load64(p + 24));
        }
        h = rotl(a, 1) + rotl(b, 7) + rotl(c, 12) + rotl(d, 18);
        h = lane_merge(lane_merge(lane_merge(lane_merge(h, a), b), c), d);
    } else {
        h = seed + PR5;
    }
    h += static_cast<uint64_t>(len);
    for (; p + 8 <= end; p += 8) h = rotl(h ^ lane_round(0, load64(p)), 27) * PR1 + PR4;
    if (p + 4 <= end) { h = rotl(h ^ (load32(p) * PR1), 23) * PR2 + PR3; p += 4; }
    for (; p < end; ++p) h = rotl(h ^ (*p * PR5), 11) * PR1;
    h ^= h >> 33; h *= PR2; h ^= h >> 29; h *= PR3; h ^= h >> 32;
    return h;
}

// ------------------------------------------------------------------------------------------------ text helpers
namespace {

// Length in bytes of a Unicode White_Space scalar starting at s[i], 0 if none (Rust char::is_whitespace).
inline int white_at(const char *s, int64_t i, int64_t end) {
    const unsigned char c = (unsigned char)s[i];
    if (c == ' ' || (c >= 9 && c <= 13)) return 1;
    if (c == 0xC2 && i + 1 < end) {
        const unsigned char d = (unsigned char)s[i + 1];
        return (d == 0x85 || d == 0xA0) ? 2 : 0;
    }
    if (i + 2 < end && (c == 0xE1 || c == 0xE2 || c == 0xE3)) {
        const unsigned char d = (unsigned char)s[i + 1], e = (unsigned char)s[i + 2];
        if (c == 0xE1) return (d == 0x9A && e == 0x80) ? 3 : 0;
        if (c == 0xE3) return (d == 0x80 && e == 0x80) ? 3 : 0;
        if (d == 0x80) return ((e >= 0x80 && e <= 0x8A) || e == 0xA8 || e == 0xA9 || e == 0xAF) ? 3 : 0;
        if (d == 0x81) return e == 0x9F ? 3 : 0;
    }
    return 0;
}

struct View {
    int64_t b, e;
};

inline View trimmed(const char *s, View v) {   // str::trim
    while (v.b < v.e) { int w = white_at(s, v.b, v.e); if (!w) break; v.b += w; }
    while (v.e > v.b) {
        int w = 0;
        for (int k = 1; k <= 3 && v.e - k >= v.b; ++k)
            if (white_at(s, v.e - k, v.e) == k) { w = k; break; }
        if (!w) break;
        v.e -= w;
    }
    return v;
}

inline bool ieq(const std::string &a, const char *lit) {
    size_t n = std::strlen(lit);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; ++i) {
        char c = a[i];
        if (c >= 'A' && c <= 'Z') c = char(c - 'A' + 'a');
        if (c != lit[i]) return false;
    }
    return true;
}

std::vector<std::string> split_str(const std::string &s, const std::string &sep) {
    std::vector<std::string> out;
    size_t pos = 0;
    for (;;) {
        size_t q = s.find(sep, pos);
        if (q == std::string::npos) { out.push_back(s.substr(pos)); break; }
        out.push_back(s.substr(pos, q - pos));
        pos = q + sep.size();
    }
    return out;
}

std::vector<Column> parse_columns(const std::string &spec) {
    std::vector<Column> cols;
    for (const std::string &field : split_str(spec, " ")) {
        std::vector<std::string> parts = split_str(field, "::");
        Column c;
        if (parts.size() > 1) {
            c.name = parts.back();
            for (size_t k = 0; k + 1 < parts.size(); ++k) {
                if (ieq(parts[k], "complex")) c.complex_ = true;
                else if (ieq(parts[k], "reflexive")) c.reflexive = true;
                else throw BuildError{"Unrecognized column field modifier: " + parts[k]};
            }
        } else {
            c.name = field;
        }
        cols.push_back(c);
    }
    for (const Column &c : cols)
        if (c.reflexive && !c.complex_)
            throw BuildError{"A field cannot be REFLEXIVE but NOT COMPLEX. It does not make sense: " + c.name};
    return cols;
}

Descriptor single_relation(const std::vector<Column> &cols) {
    std::vector<Descriptor> rel;
    const int nf = (int)cols.size();
    int reflexive_seen = 0;
    for (int i = 0; i < nf; ++i)
        for (int j = i; j < nf; ++j) {
            if (i < j) rel.push_back(Descriptor{i, j, cols[i].name, cols[j].name});
            else if (cols[i].reflexive) rel.push_back(Descriptor{i, nf + reflexive_seen++, cols[i].name, cols[j].name});
        }
    if (rel.size() != 1)
        throw BuildError{"More than one relation! Adjust your columns so there is only one relation."};
    return rel[0];
}

// Flat hash map u64 -> u32 slot index (linear probing); keys are already well mixed hashes or (row,col) pairs.
class SlotMap {
   public:
    explicit SlotMap(size_t cap_pow2 = 1024) : keys_(cap_pow2), vals_(cap_pow2, EMPTY) {}
    static constexpr uint32_t EMPTY = 0xFFFFFFFFu;
    // returns reference to the slot value; EMPTY means newly inserted (caller must assign).
    uint32_t &at(uint64_t key) {
        if ((size_ + 1) * 4 > keys_.size() * 3) grow();
        size_t mask = keys_.size() - 1, i = mix(key) & mask;
        while (vals_[i] != EMPTY && keys_[i] != key) i = (i + 1) & mask;
        if (vals_[i] == EMPTY) { keys_[i] = key; ++size_; }
        return vals_[i];
    }

   private:
    static size_t mix(uint64_t x) {
        x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ULL; x ^= x >> 32;
        return (size_t)x;
    }
    void grow() {
        std::vector<uint64_t> ok; std::vector<uint32_t> ov;
        ok.swap(keys_); ov.swap(vals_);
        keys_.assign(ok.size() * 2, 0); vals_.assign(ok.size() * 2, EMPTY);
        size_ = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ov[i] != EMPTY) at(ok[i]) = ov[i];
    }
    std::vector<uint64_t> keys_;
    std::vector<uint32_t> vals_;
    size_t size_ = 0;
};

// One SparseMatrixBuffer + SyncNodeIndexerBuilder, fed line by line.
class Builder {
   public:
    Builder(const std::string &columns, int64_t trim_n) : cols_(parse_columns(columns)), trim_n_(trim_n) {
        desc_ = single_relation(cols_);
    }

    void add_line(const char *s, int64_t b, int64_t e) {
        View line = trimmed(s, View{b, e});
        // ---- parse_line
        char sep = 0;
        if (std::memchr(s + line.b, '\t', size_t(line.e - line.b))) sep = '\t';
        else if (std::memchr(s + line.b, ',', size_t(line.e - line.b))) sep = ',';
        cells_.clear(); tokens_.clear();
        int64_t pos = line.b;
        for (;;) {
            int64_t cell_end = line.e;
            if (sep)
                if (const void *q = std::memchr(s + pos, sep, size_t(line.e - pos))) cell_end = (const char *)q - s;
            View cell{pos, cell_end};
            if (sep == ',') cell = trimmed(s, cell);
            cells_.push_back((int64_t)tokens_.size());
            int64_t t = cell.b;
            for (;;) {
                const void *q = std::memchr(s + t, ' ', size_t(cell.e - t));
                int64_t te = q ? (const char *)q - s : cell.e;
                tokens_.push_back(View{t, te});
                if (!q) break;
                t = te + 1;
            }
            if (!sep || cell_end == line.e) break;
            pos = cell_end + 1;
        }
        cells_.push_back((int64_t)tokens_.size());
        if (cells_.size() - 1 != cols_.size()) return;   // "Wrong number of columns" -> line skipped

        // ---- process_row_and_get_edges
        nodes_.clear();
        View span[2] = {{0, 0}, {0, 0}};
        int reflexive_seen = 0;
        for (size_t ci = 0; ci < cols_.size(); ++ci) {
            int64_t t0 = cells_[ci], t1 = cols_[ci].complex_ ? cells_[ci + 1] : cells_[ci] + 1;
            const int64_t first = (int64_t)nodes_.size();
            for (int64_t t = t0; t < t1; ++t) nodes_.push_back(index_of(s + tokens_[t].b, size_t(tokens_[t].e - tokens_[t].b), (uint8_t)ci));
            const View sp{first, (int64_t)nodes_.size()};
            if (ci < 2) span[ci] = sp;
            if (cols_[ci].complex_ && cols_[ci].reflexive) {
                size_t rid = cols_.size() + size_t(reflexive_seen++);
                if (rid < 2) span[rid] = sp;
            }
        }
        hyperedge(span[desc_.col_a_id], span[desc_.col_b_id]);
    }

    std::unique_ptr<Graph> finish() {
        auto g = std::make_unique<Graph>();
        const int64_t n = (int64_t)ids_.size();
        g->desc = desc_;
        g->n_rows = g->n_cols = n;
        // sort (row, col): counting sort on rows, then comparison sort inside each row
        std::vector<int64_t> rowptr(size_t(n) + 1, 0);
        for (uint64_t k : edge_key_) rowptr[(k >> 32) + 1]++;
        std::partial_sum(rowptr.begin(), rowptr.end(), rowptr.begin());
        const size_t nnz = edge_key_.size();
        std::vector<std::pair<uint32_t, float>> sorted(nnz);
        {
            std::vector<int64_t> cursor(rowptr.begin(), rowptr.end() - 1);
            for (size_t i = 0; i < nnz; ++i)
                sorted[size_t(cursor[edge_key_[i] >> 32]++)] = {uint32_t(edge_key_[i]), edge_val_[i]};
        }
        g->col.resize(nnz); g->left.resize(nnz); g->sym.resize(nnz);
        #pragma omp parallel for schedule(dynamic, 1024)
        for (int64_t r = 0; r < n; ++r) {
            auto b = sorted.begin() + rowptr[r], e = sorted.begin() + rowptr[r + 1];
            std::sort(b, e, [](const auto &x, const auto &y) { return x.first < y.first; });
            const float rs = row_sum_[r];
            for (auto it = b; it != e; ++it) {
                const size_t k = size_t(it - sorted.begin());
                g->col[k] = it->first;
                g->left[k] = it->second / rs;
                g->sym[k] = it->second / std::sqrt(rs * row_sum_[it->first]);
            }
        }
        g->rowptr = std::move(rowptr);
        g->row_sum = std::move(row_sum_);
        g->hash = std::move(hashes_);
        g->column_id = std::move(column_id_);
        g->ids = std::move(ids_);
        return g;
    }

   private:
    uint32_t index_of(const char *tok, size_t len, uint8_t column) {
        const uint64_t h = xxh64(tok, len, 0);
        uint32_t &slot = by_hash_.at(h);
        if (slot == SlotMap::EMPTY) {
            slot = (uint32_t)ids_.size();
            ids_.emplace_back(tok, len);
            hashes_.push_back(h);
            column_id_.push_back(column);
            row_sum_.push_back(0.0f);
            occurrence_.push_back(0u);
        }
        return slot;
    }

    void hyperedge(View a, View b) {
        const uint32_t la = uint32_t(a.e - a.b), lb = uint32_t(b.e - b.b);
        for (int64_t i = a.b; i < a.e; ++i) { occurrence_[nodes_[i]] += lb; row_sum_[nodes_[i]] += 1.0f / float(lb); }
        for (int64_t i = b.b; i < b.e; ++i) { occurrence_[nodes_[i]] += la; row_sum_[nodes_[i]] += 1.0f / float(la); }
        const float value = 1.0f / float(uint32_t(la * lb));
        size_t ha, hb;
        side(a, A_, ha);
        side(b, B_, hb);
        combine(0, ha, 0, hb, value);              // high x high
        combine(0, ha, hb, B_.size(), value);      // high x low
        combine(ha, A_.size(), 0, hb, value);      // low  x high     (low x low dropped)
    }

    // get_high_low_nodes: the trim_n most frequent nodes first.  Rust's select_nth_unstable leaves the order of
    // ties unspecified; this implementation keeps input order among ties (stable) -- DESIGN.md "known deviations".
    void side(View v, std::vector<uint32_t> &out, size_t &n_high) {
        out.assign(nodes_.begin() + v.b, nodes_.begin() + v.e);
        n_high = out.size();
        if ((int64_t)out.size() > trim_n_) {
            std::stable_sort(out.begin(), out.end(), [&](uint32_t x, uint32_t y) { return occurrence_[x] > occurrence_[y]; });
            n_high = size_t(trim_n_);
        }
    }

    void combine(size_t a0, size_t a1, size_t b0, size_t b1, float value) {
        for (size_t i = a0; i < a1; ++i)
            for (size_t j = b0; j < b1; ++j) {
                add((uint64_t(A_[i]) << 32) | B_[j], value);
                add((uint64_t(B_[j]) << 32) | A_[i], value);
            }
    }

    void add(uint64_t key, float value) {
        uint32_t &slot = by_pair_.at(key);
        if (slot == SlotMap::EMPTY) {
            if (edge_key_.size() >= 0xFFFFFFFEull) throw BuildError{"graph too large for the string builder (>= 2^32 edges)"};
            slot = (uint32_t)edge_key_.size();
            edge_key_.push_back(key);
            edge_val_.push_back(0.0f);
        }
        edge_val_[slot] += value;
    }

    std::vector<Column> cols_;
    Descriptor desc_;
    int64_t trim_n_;
    SlotMap by_hash_{1 << 12}, by_pair_{1 << 14};
    std::vector<std::string> ids_;
    std::vector<uint64_t> hashes_;
    std::vector<uint8_t> column_id_;
    std::vector<float> row_sum_;
    std::vector<uint32_t> occurrence_;
    std::vector<uint64_t> edge_key_;
    std::vector<float> edge_val_;
    // per-line scratch
    std::vector<int64_t> cells_;
    std::vector<View> tokens_;
    std::vector<uint32_t> nodes_, A_, B_;
};

}  // namespace

std::unique_ptr<Graph> build_from_lines(const char *buf, const int64_t *offsets, int64_t n_lines,
                                        const std::string &columns, int64_t trim_n) {
    Builder b(columns, trim_n);
    for (int64_t i = 0; i < n_lines; ++i) b.add_line(buf, offsets[i], offsets[i + 1]);
    return b.finish();
}

// Integer hyperedges: line i = members[offsets[i] .. offsets[i+1]) as decimal ids of ONE column, i.e. exactly the
// graph build_from_lines makes of the space-joined decimal strings -- the text of each line is formatted here and fed
// through the same Builder (same hashing, indexing, trimming and accumulation order), only the Python-side string
// handling is skipped.
std::unique_ptr<Graph> build_from_hyperedges(const uint32_t *members, const int64_t *offsets, int64_t n_lines,
                                             const std::string &columns, int64_t trim_n) {
    Builder b(columns, trim_n);
    std::string text;
    char dec[12];
    for (int64_t i = 0; i < n_lines; ++i) {
        text.clear();
        for (int64_t k = offsets[i]; k < offsets[i + 1]; ++k) {
            if (k > offsets[i]) text.push_back(' ');
            uint32_t x = members[k];
            int len = 0;
            do { dec[len++] = char('0' + x % 10); x /= 10; } while (x);
            while (len) text.push_back(dec[--len]);
        }
        b.add_line(text.data(), 0, (int64_t)text.size());
    }
    return b.finish();
}

std::unique_ptr<Graph> build_from_files(const std::vector<std::string> &paths, const std::string &columns,
                                        int64_t trim_n) {
    Builder b(columns, trim_n);
    for (const std::string &p : paths) {
        std::ifstream in(p, std::ios::binary);
        if (!in) continue;                       // "Cannot open file" is logged and skipped (pipeline.rs:193-199)
        std::string line;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();   // BufRead::lines strips "\r\n"
            if (line.empty()) continue;          // pipeline.rs:212-214
            b.add_line(line.data(), 0, (int64_t)line.size());
        }
    }
    return b.finish();
}

// Integer ingest: identical result to feeding "{u} {v}" lines to a `complex::reflexive` column, without strings.
// All f32 quantities are sums of 0.5 / 1.0 (exact below 2^23 repeats), so accumulation order is irrelevant.
std::unique_ptr<Graph> build_from_pairs(const uint32_t *u, const uint32_t *v, int64_t n_pairs,
                                        const std::string &column_name) {
    auto g = std::make_unique<Graph>();
    g->desc = Descriptor{0, 1, column_name, column_name};
    uint32_t max_id = 0;
    #pragma omp parallel for reduction(max : max_id) schedule(static)
    for (int64_t i = 0; i < n_pairs; ++i) max_id = std::max(max_id, std::max(u[i], v[i]));
    // first-appearance relabel, in parallel: first[x] = smallest position (2i for u[i], 2i+1 for v[i]) at which x
    // occurs; entities ordered by that position get labels 0, 1, 2, ... (identical to a sequential scan)
    const size_t n_ids = n_pairs ? size_t(max_id) + 1 : 0;
    constexpr uint64_t NEVER = ~0ull;
    std::vector<uint64_t> first(n_ids, NEVER);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_pairs; ++i) {
        for (int side = 0; side < 2; ++side) {
            const uint32_t x = side ? v[i] : u[i];
            const uint64_t pos = uint64_t(2 * i + side);
            uint64_t cur = __atomic_load_n(&first[x], __ATOMIC_RELAXED);
            while (pos < cur && !__atomic_compare_exchange_n(&first[x], &cur, pos, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
        }
    }
    std::vector<uint32_t> original;
    original.reserve(n_ids);
    for (size_t x = 0; x < n_ids; ++x)
        if (first[x] != NEVER) original.push_back(uint32_t(x));
    std::sort(original.begin(), original.end(), [&](uint32_t a, uint32_t b) { return first[a] < first[b]; });
    std::vector<uint32_t> label(n_ids, 0xFFFFFFFFu);
    const int64_t n = (int64_t)original.size();
    #pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) label[original[size_t(k)]] = uint32_t(k);
    { std::vector<uint64_t>().swap(first); }
    g->n_rows = g->n_cols = n;
    // entries per row (upper bound before merging duplicates), then bucket; weights are exact multiples of 0.5, so
    // the order in which a row's entries arrive does not change any sum
    std::vector<int64_t> start(size_t(n) + 1, 0);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_pairs; ++i) {
        const uint32_t a = label[u[i]], b = label[v[i]];
        if (a == b) { __atomic_fetch_add(&start[a + 1], 1, __ATOMIC_RELAXED); }
        else { __atomic_fetch_add(&start[a + 1], 2, __ATOMIC_RELAXED); __atomic_fetch_add(&start[b + 1], 2, __ATOMIC_RELAXED); }
    }
    std::partial_sum(start.begin(), start.end(), start.begin());
    std::vector<std::pair<uint32_t, float>> ent((size_t)start[(size_t)n]);
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        #pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n_pairs; ++i) {
            const uint32_t a = label[u[i]], b = label[v[i]];
            if (a == b) ent[size_t(__atomic_fetch_add(&cur[a], 1, __ATOMIC_RELAXED))] = {a, 2.0f};
            else {
                const int64_t pa = __atomic_fetch_add(&cur[a], 2, __ATOMIC_RELAXED);
                ent[size_t(pa)] = {a, 0.5f}; ent[size_t(pa + 1)] = {b, 0.5f};
                const int64_t pb = __atomic_fetch_add(&cur[b], 2, __ATOMIC_RELAXED);
                ent[size_t(pb)] = {b, 0.5f}; ent[size_t(pb + 1)] = {a, 0.5f};
            }
        }
    }
    g->row_sum.assign(size_t(n), 0.0f);
    std::vector<int64_t> cnt(size_t(n), 0);
    #pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t r = 0; r < n; ++r) {
        auto b = ent.begin() + start[r], e = ent.begin() + start[r + 1];
        std::sort(b, e, [](const auto &x, const auto &y) { return x.first < y.first; });
        auto w = b;
        float rs = 0.0f;
        for (auto it = b; it != e;) {
            uint32_t c = it->first; float s = 0.0f;
            for (; it != e && it->first == c; ++it) s += it->second;
            *w++ = {c, s};
            // row_sum: +1 per endpoint occurrence of a distinct-node pair (= 2 * weight of the off-diagonal
            // entries' share): every pair contributes 0.5 to the diagonal and 0.5 to one off-diagonal of this
            // row, and 1.0 to row_sum; self pairs contribute 2.0 to both.  Hence row_sum == sum of the row.
            rs += s;
        }
        g->row_sum[size_t(r)] = rs;
        cnt[size_t(r)] = w - b;
    }
    g->rowptr.assign(size_t(n) + 1, 0);
    for (int64_t r = 0; r < n; ++r) g->rowptr[size_t(r) + 1] = g->rowptr[size_t(r)] + cnt[size_t(r)];
    const size_t nnz = size_t(g->rowptr[size_t(n)]);
    g->col.resize(nnz); g->left.resize(nnz); g->sym.resize(nnz);
    #pragma omp parallel for schedule(dynamic, 4096)
    for (int64_t r = 0; r < n; ++r) {
        const float rs = g->row_sum[size_t(r)];
        int64_t src = start[r];
        for (int64_t k = g->rowptr[size_t(r)]; k < g->rowptr[size_t(r) + 1]; ++k, ++src) {
            g->col[size_t(k)] = ent[size_t(src)].first;
            g->left[size_t(k)] = ent[size_t(src)].second / rs;
            g->sym[size_t(k)] = ent[size_t(src)].second / std::sqrt(rs * g->row_sum[ent[size_t(src)].first]);
        }
    }
    g->column_id.assign(size_t(n), 0);
    g->hash.resize(size_t(n));
    g->ids.resize(size_t(n));
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        char tmp[16];
        int len = std::snprintf(tmp, sizeof tmp, "%u", original[size_t(i)]);
        g->ids[size_t(i)].assign(tmp, size_t(len));
        g->hash[size_t(i)] = xxh64(tmp, size_t(len), 0);
    }
    return g;
}

}  // namespace cleora
