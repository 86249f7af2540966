// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// HDF5-free C binder over the *unmodified* reference PUFFINN headers, compiled from where they lie
// under /root/reference/libpuffinn/include by oracle/Makefile into oracle/_ref/libpuffinn_ref.so.
// It replaces the reference's own libpuffinn-ffi/c_binder.cpp (which needs hdf5.h) for three jobs:
//   1. pin oracle/clann_oracle.c (the plain-C restatement) against the real reference,
//   2. generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
//   3. serve as the "reference" CPU baseline arm of bench.py (cpu_baseline.kind == "reference").
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Built with -fno-access-control so private members of puffinn::Index (lsh_maps, filterer, hash_source)
// can be read for golden dumps without patching any reference header.
//
// The CLANN layer (Rust: src/core/{gmm,index,heap}.rs, src/metricdata/angulardata.rs) cannot be built
// here (no cargo); ref_clann_* below restates its control flow over *reference* PUFFINN indices so the
// reference arm exercises the real L0/L1 code. Citations are file:line into /root/reference.

#include "puffinn.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <queue>
#include <sstream>
#include <vector>

using RefIndex = puffinn::Index<puffinn::CosineSimilarity>;

namespace {

struct Handle {
    RefIndex* index;
    int dim;
};

// ndarray 0.16.1 `unrolled_dot` (numeric_util.rs), the kernel behind ArrayBase::dot for f32 without BLAS:
// eight independent partial sums over chunks of 8, combined (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail.
// Rust never contracts a*b+c into an FMA, hence the volatile-free explicit mul/add under -ffp-contract=off.
__attribute__((optimize("fp-contract=off"))) float ndarray_dot(const float* x, const float* y, size_t len) {
    float p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    size_t i = 0;
    for (; i + 8 <= len; i += 8) {
        p0 = p0 + x[i + 0] * y[i + 0];
        p1 = p1 + x[i + 1] * y[i + 1];
        p2 = p2 + x[i + 2] * y[i + 2];
        p3 = p3 + x[i + 3] * y[i + 3];
        p4 = p4 + x[i + 4] * y[i + 4];
        p5 = p5 + x[i + 5] * y[i + 5];
        p6 = p6 + x[i + 6] * y[i + 6];
        p7 = p7 + x[i + 7] * y[i + 7];
    }
    float sum = 0;
    sum = sum + (p0 + p4);
    sum = sum + (p1 + p5);
    sum = sum + (p2 + p6);
    sum = sum + (p3 + p7);
    for (; i < len; i++) sum = sum + x[i] * y[i];
    return sum;
}

// angulardata.rs:29-35 — distance_point re-derives the query norm with a sequential iterator sum.
__attribute__((optimize("fp-contract=off"))) float query_norm(const float* q, size_t d) {
    float s = 0;
    for (size_t i = 0; i < d; i++) s = s + q[i] * q[i];
    return std::sqrt(s);
}

struct ClannRef {
    const float* data;  // borrowed, row-major n x d
    uint64_t n;
    uint32_t d;
    uint32_t L;
    uint32_t k;
    float delta;
    uint64_t seed_base;
    std::vector<float> norms;                    // angulardata.rs:12-19
    std::vector<uint64_t> centers;               // ClusterCenter.center_idx
    std::vector<float> radii;                    // ClusterCenter.radius
    std::vector<std::vector<uint64_t>> members;  // ClusterCenter.assignment
    std::vector<uint8_t> brute;                  // index.rs:204-205
    std::vector<RefIndex*> indices;              // puffinn_indices (lazily built, see ref_clann_search)
    // counters of the last query
    uint64_t last_visited = 0, last_distcomp = 0, last_candidates = 0;
    double build_seconds = 0;
};

__attribute__((optimize("fp-contract=off"))) float dist_point(const ClannRef& c, uint64_t i, const float* q) {
    // angulardata.rs:29-35
    float dot = ndarray_dot(c.data + i * c.d, q, c.d);
    float nq = query_norm(q, c.d);
    float cs = dot / (c.norms[i] * nq);
    return 1.0f - cs;
}

// heap.rs:5-49 — max-heap on (OrderedFloat distance, point_index); derived Ord compares distance then index.
struct Elem {
    float dist;
    uint64_t idx;
    bool operator<(const Elem& o) const { return dist < o.dist || (dist == o.dist && idx < o.idx); }
};
struct TopK {
    std::priority_queue<Elem> heap;
    size_t cap;
    explicit TopK(size_t k) : cap(k) {}
    bool add(Elem e) {  // heap.rs:23-36
        if (heap.size() < cap) {
            heap.push(e);
        } else if (!heap.empty()) {
            if (e.dist < heap.top().dist) {
                heap.pop();
                heap.push(e);
            } else {
                return false;
            }
        }
        return true;
    }
};

void ensure_index(ClannRef& c, size_t ci) {
    if (c.indices[ci] || c.brute[ci]) return;
    double t0 = omp_get_wtime();
    // Reproducible functions per cluster (the reference seeds from the wall clock, typedefs.hpp:17).
    puffinn::get_default_random_generator().seed(c.seed_base + ci);
    auto* idx = new RefIndex(c.d);  // c_binder.cpp:39-50
    for (uint64_t p : c.members[ci]) {
        // puffinn.rs:41-46 -> c_binder.cpp:63-66
        idx->insert(std::vector<float>(c.data + p * c.d, c.data + (p + 1) * c.d));
    }
    idx->rebuild(c.L);  // c_binder.cpp:53-60
    c.indices[ci] = idx;
    c.build_seconds += omp_get_wtime() - t0;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- single PUFFINN index (L1)

void ref_seed(uint64_t s) { puffinn::get_default_random_generator().seed(s); }

void* ref_index_create(int d) {
    try {
        return new Handle{new RefIndex((unsigned)d), d};
    } catch (...) {
        return nullptr;
    }
}

void ref_index_free(void* h) {
    auto* hh = static_cast<Handle*>(h);
    if (!hh) return;
    delete hh->index;
    delete hh;
}

int ref_index_insert(void* h, const float* p, int d) {
    auto* hh = static_cast<Handle*>(h);
    try {
        hh->index->insert(std::vector<float>(p, p + d));
        return 0;
    } catch (...) {
        return -1;
    }
}

uint64_t ref_index_rebuild(void* h, unsigned L) {
    auto* hh = static_cast<Handle*>(h);
    try {
        return hh->index->rebuild(L);
    } catch (...) {
        return 0;
    }
}

// Returns number of ids written (<= cap). metrics[4] = {distance_computations, candidates, hash_length, considered_maps}
// of this query (performance.hpp:30-36); hash_length/considered_maps stay 0 when no depth stopped.
int ref_index_search(void* h, const float* q, unsigned k, float recall, float max_sim, uint32_t* out, int cap,
                     uint32_t* metrics) {
    auto* hh = static_cast<Handle*>(h);
    puffinn::g_performance_metrics.clear();
    std::vector<uint32_t> res;
    try {
        res = hh->index->search(std::vector<float>(q, q + hh->dim), k, recall, max_sim);
    } catch (...) {
        return -1;
    }
    auto qm = puffinn::g_performance_metrics.get_query_metrics();
    if (metrics) {
        metrics[0] = metrics[1] = metrics[2] = metrics[3] = 0;
        if (!qm.empty()) {  // the brute-force path (collection.hpp:550-555) returns before new_query()
            auto& m = qm.back();
            metrics[0] = m.distance_computations;
            metrics[1] = m.candidates;
            metrics[2] = m.hash_length;
            metrics[3] = m.considered_maps;
        }
    }
    int n = std::min<int>(cap, (int)res.size());
    for (int i = 0; i < n; i++) out[i] = res[i];
    return n;
}

// The same through Index::search's FilterType argument (collection.hpp:22-34,324-334): 0 = Default, 1 = None, 2 = Simple.
int ref_index_search_filter(void* h, const float* q, unsigned k, float recall, float max_sim, int filter_type, uint32_t* out, int cap,
                            uint32_t* metrics) {
    auto* hh = static_cast<Handle*>(h);
    puffinn::g_performance_metrics.clear();
    const puffinn::FilterType ft = filter_type == 1   ? puffinn::FilterType::None
                                   : filter_type == 2 ? puffinn::FilterType::Simple
                                                      : puffinn::FilterType::Default;
    std::vector<uint32_t> res;
    try {
        res = hh->index->search(std::vector<float>(q, q + hh->dim), k, recall, max_sim, ft);
    } catch (...) {
        return -1;
    }
    auto qm = puffinn::g_performance_metrics.get_query_metrics();
    if (metrics) {
        metrics[0] = metrics[1] = metrics[2] = metrics[3] = 0;
        if (!qm.empty()) {  // the brute-force path (collection.hpp:550-555) returns before new_query()
            auto& m = qm.back();
            metrics[0] = m.distance_computations;
            metrics[1] = m.candidates;
            metrics[2] = m.hash_length;
            metrics[3] = m.considered_maps;
        }
    }
    int n = std::min<int>(cap, (int)res.size());
    for (int i = 0; i < n; i++) out[i] = res[i];
    return n;
}

// Index::serialize (collection.hpp:185-203). Returns the stream size; copies min(size, cap) bytes.
uint64_t ref_index_serialize(void* h, uint8_t* dst, uint64_t cap) {
    auto* hh = static_cast<Handle*>(h);
    std::ostringstream ss;
    hh->index->serialize(ss, false);
    std::string s = ss.str();
    if (dst) std::memcpy(dst, s.data(), std::min<uint64_t>(cap, s.size()));
    return s.size();
}

void* ref_index_deserialize(const uint8_t* src, uint64_t len) {
    try {
        std::istringstream ss(std::string(reinterpret_cast<const char*>(src), len));
        auto* idx = new RefIndex(ss);
        return new Handle{idx, (int)idx->dataset.get_description().args};
    } catch (...) {
        return nullptr;
    }
}

uint32_t ref_index_size(void* h) { return static_cast<Handle*>(h)->index->get_size(); }
uint32_t ref_index_storage_len(void* h) { return static_cast<Handle*>(h)->index->dataset.get_description().storage_len; }

// --- golden dumps through private members (needs -fno-access-control)

// Stored Q15 row of point i (dataset.hpp:88-90); writes storage_len values.
void ref_dump_point(void* h, uint32_t i, int16_t* out) {
    auto* ix = static_cast<Handle*>(h)->index;
    auto sl = ix->dataset.get_description().storage_len;
    std::memcpy(out, ix->dataset[i], sl * sizeof(int16_t));
}

// Q15 form of an arbitrary float vector (format/unit_vector.hpp:61-89).
void ref_store_q15(void* h, const float* v, int16_t* out) {
    auto* hh = static_cast<Handle*>(h);
    auto desc = hh->index->dataset.get_description();
    auto st = puffinn::to_stored_type<puffinn::UnitVectorFormat>(std::vector<float>(v, v + hh->dim), desc);
    std::memcpy(out, st.get(), desc.storage_len * sizeof(int16_t));
}

// L table codes of a float query with this index's functions (independent.hpp:70-86).
int ref_query_codes(void* h, const float* q, uint64_t* out) {
    auto* hh = static_cast<Handle*>(h);
    if (!hh->index->hash_source) return -1;
    auto desc = hh->index->dataset.get_description();
    auto st = puffinn::to_stored_type<puffinn::UnitVectorFormat>(std::vector<float>(q, q + hh->dim), desc);
    std::vector<uint64_t> codes;
    hh->index->hash_source->hash_repetitions(st.get(), codes);
    std::memcpy(out, codes.data(), codes.size() * sizeof(uint64_t));
    return (int)codes.size();
}

// 32 sketches of a float query (filterer.hpp:99-102).
void ref_query_sketches(void* h, const float* q, uint64_t* out) {
    auto* hh = static_cast<Handle*>(h);
    auto desc = hh->index->dataset.get_description();
    auto st = puffinn::to_stored_type<puffinn::UnitVectorFormat>(std::vector<float>(q, q + hh->dim), desc);
    puffinn::QuerySketches qs;
    hh->index->filterer.sketch(st.get(), qs);
    std::memcpy(out, qs.query_sketches.data(), 32 * sizeof(uint64_t));
}

// Stored sketches of point i (filterer.hpp:113-115).
void ref_point_sketches(void* h, uint32_t i, uint64_t* out) {
    auto* ix = static_cast<Handle*>(h)->index;
    for (int s = 0; s < 32; s++) out[s] = ix->filterer.get_sketch(i, s);
}

// Table t as stored (padded): returns len = n + 24 and copies hashes/indices (prefixmap.hpp:76-79).
uint64_t ref_table(void* h, uint32_t t, uint32_t* hashes, uint32_t* indices) {
    auto* ix = static_cast<Handle*>(h)->index;
    auto& m = ix->lsh_maps[t];
    if (hashes) std::memcpy(hashes, m.hashes.data(), m.hashes.size() * 4);
    if (indices) std::memcpy(indices, m.indices.data(), m.indices.size() * 4);
    return m.hashes.size();
}

// Anchor + the 24 per-depth ranges of every table for one float query, exactly as SearchBuffers would see them
// (prefixmap.hpp:250-304, collection.hpp:628-667). anchors[L]; ranges[24][L][2] = padded (start,end) positions.
int ref_query_ranges(void* h, const float* q, uint32_t* anchors, uint32_t* ranges) {
    auto* hh = static_cast<Handle*>(h);
    auto* ix = hh->index;
    if (!ix->hash_source) return -1;
    auto desc = ix->dataset.get_description();
    auto st = puffinn::to_stored_type<puffinn::UnitVectorFormat>(std::vector<float>(q, q + hh->dim), desc);
    std::vector<uint64_t> codes;
    ix->hash_source->hash_repetitions(st.get(), codes);
    size_t L = ix->lsh_maps.size();
    std::vector<puffinn::PrefixMapQuery> qo;
    for (size_t t = 0; t < L; t++) {
        qo.push_back(ix->lsh_maps[t].create_query(codes[t]));
        anchors[t] = qo[t].prefix_start;
    }
    for (int it = 0; it < 24; it++) {
        for (size_t t = 0; t < L; t++) {
            auto r = ix->lsh_maps[t].get_next_range(qo[t]);
            const uint32_t* base = ix->lsh_maps[t].indices.data();
            ranges[(it * L + t) * 2 + 0] = (uint32_t)(r.first - base);
            ranges[(it * L + t) * 2 + 1] = (uint32_t)(r.second - base);
        }
    }
    return (int)L;
}

// Q15 similarity of a float query to stored point i (cosine.hpp:19-23).
float ref_similarity(void* h, const float* q, uint32_t i) {
    auto* hh = static_cast<Handle*>(h);
    auto desc = hh->index->dataset.get_description();
    auto st = puffinn::to_stored_type<puffinn::UnitVectorFormat>(std::vector<float>(q, q + hh->dim), desc);
    return puffinn::CosineSimilarity::compute_similarity(st.get(), hh->index->dataset[i], desc);
}

// failure_probability (independent.hpp:108-119) and get_max_sketch_diff (filterer.hpp:108-111) as the index computes them.
float ref_failure_probability(void* h, unsigned depth, unsigned tables, unsigned max_tables, float sim) {
    return static_cast<Handle*>(h)->index->hash_source->failure_probability(depth, tables, max_tables, sim);
}
unsigned ref_max_sketch_diff(void* h, float sim) { return static_cast<Handle*>(h)->index->filterer.get_max_sketch_diff(sim); }

// --- reference known-answer helpers that need no index (format_test.hpp, math_test.hpp, maxbuffer_test.hpp)
int16_t ref_to_q15(float v) { return puffinn::UnitVectorFormat::to_16bit_fixed_point(v); }
float ref_from_q15(int16_t v) { return puffinn::UnitVectorFormat::from_16bit_fixed_point(v); }
int16_t ref_dot_i16(const int16_t* a, const int16_t* b, unsigned n) {
    // the AVX2 path needs 32-byte alignment (math.hpp:16-17)
    std::vector<int16_t> buf(2 * n + 64);
    auto al = [](int16_t* p) { return reinterpret_cast<int16_t*>((reinterpret_cast<uintptr_t>(p) + 31) & ~uintptr_t(31)); };
    int16_t* x = al(buf.data());
    int16_t* y = al(x + n);
    std::memcpy(x, a, n * 2);
    std::memcpy(y, b, n * 2);
    return puffinn::dot_product_i16(x, y, n);
}
int16_t ref_dot_i16_simple(const int16_t* a, const int16_t* b, unsigned n) { return puffinn::dot_product_i16_simple(a, b, n); }
void ref_fht(float* buf, int log_n) {
    alignas(64) float tmp[4096];
    std::memcpy(tmp, buf, sizeof(float) << log_n);
    fht(tmp, log_n);
    std::memcpy(buf, tmp, sizeof(float) << log_n);
}
// Drives a reference MaxBuffer(k) with n (idx, value) inserts; writes best entries; returns their count and minval.
int ref_maxbuffer(unsigned k, const uint32_t* ids, const float* vals, int n, uint32_t* out_ids, float* out_vals, float* minval) {
    puffinn::MaxBuffer mb(k);
    for (int i = 0; i < n; i++) mb.insert(ids[i], vals[i]);
    float mv_before = mb.smallest_value();
    auto e = mb.best_entries();
    for (size_t i = 0; i < e.size(); i++) {
        out_ids[i] = e[i].first;
        out_vals[i] = e[i].second;
    }
    if (minval) *minval = mv_before;
    return (int)e.size();
}

// ---------------------------------------------------------------- CLANN layer over reference PUFFINN (L3 restated)

// index.rs:71-91 — K = max(1, floor(f64(factor_f32) * sqrt(n)))
uint64_t ref_clann_num_clusters(float factor, uint64_t n) {
    double v = std::floor((double)factor * std::sqrt((double)n));
    uint64_t k = (uint64_t)v;
    return k < 1 ? 1 : k;
}

// gmm.rs:21-62 over AngularData (angulardata.rs:12-43). centers[K], assignment[n], radii[K] (K' = min(n,K) used).
uint64_t ref_clann_gmm(const float* data, uint64_t n, uint32_t d, uint64_t K, uint64_t* centers, uint64_t* assignment, float* radii) {
    std::vector<float> norms(n);
    for (uint64_t i = 0; i < n; i++) norms[i] = std::sqrt(ndarray_dot(data + i * d, data + i * d, d));
    if (n <= K) {
        for (uint64_t i = 0; i < n; i++) {
            centers[i] = i;
            assignment[i] = i;
            radii[i] = 0;
        }
        return n;
    }
    auto all_dist = [&](uint64_t j, std::vector<float>& out) {
        for (uint64_t i = 0; i < n; i++)
            out[i] = 1.0f - (ndarray_dot(data + i * d, data + j * d, d) / (norms[i] * norms[j]));
    };
    std::vector<float> dist(n), nd(n);
    for (uint64_t i = 0; i < n; i++) assignment[i] = 0;
    centers[0] = 0;
    all_dist(0, dist);
    for (uint64_t c = 1; c < K; c++) {
        uint64_t far = 0;
        float m = dist[0];
        for (uint64_t i = 1; i < n; i++)
            if (dist[i] > m) {
                far = i;
                m = dist[i];
            }
        centers[c] = far;
        all_dist(far, nd);
        for (uint64_t i = 0; i < n; i++)
            if (nd[i] < dist[i]) {
                assignment[i] = c;
                dist[i] = nd[i];
            }
    }
    for (uint64_t c = 0; c < K; c++) radii[c] = 0;
    for (uint64_t i = 0; i < n; i++) radii[assignment[i]] = std::max(radii[assignment[i]], dist[i]);
    return K;
}

// Creates the CLANN state from a given clustering (data is borrowed and must outlive the handle).
// PUFFINN indices are built lazily on first visit so a bounded query sample only pays for the clusters it touches;
// build time is accumulated separately and never counted as search time.
void* ref_clann_create(const float* data, uint64_t n, uint32_t d, uint32_t L, uint32_t k, float delta, uint64_t K,
                       const uint64_t* centers, const uint64_t* assignment, const float* radii, uint64_t seed_base) {
    auto* c = new ClannRef();
    c->data = data;
    c->n = n;
    c->d = d;
    c->L = L;
    c->k = k;
    c->delta = delta;
    c->seed_base = seed_base;
    c->norms.resize(n);
    for (uint64_t i = 0; i < n; i++) c->norms[i] = std::sqrt(ndarray_dot(data + i * d, data + i * d, d));
    c->centers.assign(centers, centers + K);
    c->radii.assign(radii, radii + K);
    c->members.resize(K);
    for (uint64_t i = 0; i < n; i++) c->members[assignment[i]].push_back(i);  // index.rs:188-192
    c->brute.resize(K);
    for (uint64_t ci = 0; ci < K; ci++) c->brute[ci] = c->members[ci].size() < 100 || c->members[ci].size() < k;  // index.rs:204-205
    c->indices.assign(K, nullptr);
    return c;
}

void ref_clann_build_all(void* h) {
    auto* c = static_cast<ClannRef*>(h);
    for (size_t ci = 0; ci < c->indices.size(); ci++)
        if (!c->members[ci].empty()) ensure_index(*c, ci);
}

void ref_clann_build_cluster(void* h, uint64_t ci) { ensure_index(*static_cast<ClannRef*>(h), ci); }

double ref_clann_build_seconds(void* h) { return static_cast<ClannRef*>(h)->build_seconds; }

uint64_t ref_clann_cluster_serialize(void* h, uint64_t ci, uint8_t* dst, uint64_t cap) {
    auto* c = static_cast<ClannRef*>(h);
    ensure_index(*c, ci);
    if (!c->indices[ci]) return 0;
    std::ostringstream ss;
    c->indices[ci]->serialize(ss, false);
    std::string s = ss.str();
    if (dst) std::memcpy(dst, s.data(), std::min<uint64_t>(cap, s.size()));
    return s.size();
}

// index.rs:311-439. Returns the number of (dist,id) pairs written, ascending by distance (heap.rs:42-48).
// order_out (optional, K entries) receives the cluster visiting order (index.rs:592-616).
// search_seconds (optional) accumulates the time spent excluding lazy index builds.
int ref_clann_search(void* h, const float* q, uint64_t* out_ids, float* out_dists, uint64_t* order_out, double* search_seconds) {
    auto* c = static_cast<ClannRef*>(h);
    size_t K = c->centers.size();
    double t_start = omp_get_wtime();
    double built_before = c->build_seconds;
    c->last_visited = c->last_distcomp = c->last_candidates = 0;

    std::vector<std::pair<size_t, float>> cd(K);
    for (size_t ci = 0; ci < K; ci++) cd[ci] = {ci, dist_point(*c, c->centers[ci], q)};
    std::stable_sort(cd.begin(), cd.end(), [](const auto& a, const auto& b) { return a.second < b.second; });  // slice::sort_by is stable
    if (order_out)
        for (size_t i = 0; i < K; i++) order_out[i] = cd[i].first;

    TopK pq(c->k);
    float max_dist = std::numeric_limits<float>::infinity();
    for (size_t oi = 0; oi < K; oi++) {
        size_t ci = cd[oi].first;
        if (!pq.heap.empty()) {
            Elem top = pq.heap.top();
            max_dist = top.dist;
            float cmin = dist_point(*c, c->centers[ci], q) - c->radii[ci];  // index.rs:351-352
            if (cmin > top.dist) break;                                       // index.rs:353-360
        }
        c->last_visited++;
        if (c->brute[ci]) {
            // index.rs:666-685 then :369-376
            TopK local(c->k);
            for (uint64_t p : c->members[ci]) local.add({dist_point(*c, p, q), p});
            std::vector<Elem> l;
            while (!local.heap.empty()) {
                l.push_back(local.heap.top());
                local.heap.pop();
            }
            std::stable_sort(l.begin(), l.end(), [](const Elem& a, const Elem& b) { return a.dist < b.dist; });
            for (auto& e : l) pq.add(e);
            // counters follow PUFFINN's own (performance.hpp:72-86): brute-force clusters add nothing
        } else {
            ensure_index(*c, ci);
            float max_sim = 1.0f - max_dist / 2.0f;  // puffinn_types.rs:77-79
            puffinn::g_performance_metrics.clear();
            auto res = c->indices[ci]->search(std::vector<float>(q, q + c->d), c->k, c->delta, max_sim);  // c_binder.cpp:75-76
            auto qm = puffinn::g_performance_metrics.get_query_metrics();
            c->last_distcomp += qm.back().distance_computations;
            c->last_candidates += qm.back().candidates;
            for (uint32_t local : res) {
                if (local >= c->members[ci].size()) continue;  // index.rs:639-646 would raise IndexOutOfBounds
                uint64_t p = c->members[ci][local];
                pq.add({dist_point(*c, p, q), p});  // index.rs:402-416
            }
        }
    }
    std::vector<Elem> l;
    while (!pq.heap.empty()) {
        l.push_back(pq.heap.top());
        pq.heap.pop();
    }
    // heap.rs:42-48: BinaryHeap::iter() order then stable sort by distance. iter() order is an implementation detail of
    // std's binary heap; ties in distance are the only place it shows, and those are "ties within 1e-5" for parity.
    std::reverse(l.begin(), l.end());
    std::stable_sort(l.begin(), l.end(), [](const Elem& a, const Elem& b) { return a.dist < b.dist; });
    for (size_t i = 0; i < l.size(); i++) {
        out_ids[i] = l[i].idx;
        out_dists[i] = l[i].dist;
    }
    if (search_seconds) *search_seconds += (omp_get_wtime() - t_start) - (c->build_seconds - built_before);
    return (int)l.size();
}

void ref_clann_last_counters(void* h, uint64_t* visited, uint64_t* distcomp, uint64_t* candidates) {
    auto* c = static_cast<ClannRef*>(h);
    *visited = c->last_visited;
    *distcomp = c->last_distcomp;
    *candidates = c->last_candidates;
}

void ref_clann_free(void* h) {
    auto* c = static_cast<ClannRef*>(h);
    if (!c) return;
    for (auto* p : c->indices) delete p;
    delete c;
}

int ref_omp_threads() { return omp_get_max_threads(); }

}  // extern "C"
