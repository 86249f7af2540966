/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See clann_oracle.c for the header comment and the pinning status. */
#ifndef CLANN_ORACLE_H
#define CLANN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_SKETCHES 32
#define ORC_SKETCH_BITS 64
#define ORC_MAX_HASHBITS 24
#define ORC_SEGMENT 12
#define ORC_PREFIX_BITS 13
#define ORC_EST_BINS 201

/* One PUFFINN "function set": what Index::serialize stores besides the data. */
typedef struct {
    uint32_t d, sl, m, bpf, fph, cut, L, rotations;
    int16_t* planes; /* 2048 x sl, Q15 SimHash hyperplanes, function 64*s+b -> bit 63-b of sketch s */
    int8_t* signs;   /* (L*fph) x (rotations * 2^m) */
    float* est;      /* (m+2) x 201 collision estimates */
    float eps;
} orc_functions;

typedef struct {
    orc_functions fn;
    uint32_t n;
    int16_t* q15;       /* n x sl */
    uint64_t* sketches; /* n x 32 */
    uint32_t* hashes;   /* L x (n+24), padded like prefixmap.hpp:215-226 */
    uint32_t* indices;  /* L x (n+24) */
    uint32_t* prefix_index; /* L x 8193 */
} orc_index;

typedef struct {
    uint32_t distance_computations;
    uint32_t candidates;
    uint32_t stop_depth; /* 0 = ran through depth 1 without stopping */
    uint32_t stop_table;
    uint32_t n_batches;
    float kth;               /* MaxBuffer minval at exit */
    uint32_t max_sketch_diff; /* at exit */
    uint32_t* passing;       /* optional: concatenated passing-filter ids in evaluation order */
    uint64_t passing_cap, passing_len;
    uint32_t* batch_sizes;   /* optional: entries per "empty the buffer" step */
    uint64_t batch_cap;
} orc_trace;

/* --- L0 */
int16_t orc_to_q15(float v);
float orc_from_q15(int16_t v);
uint32_t orc_storage_len(uint32_t d);
uint32_t orc_ceil_log(uint32_t v);
void orc_store_q15(const float* v, uint32_t d, uint32_t sl, int16_t* out);
int16_t orc_dot_i16(const int16_t* a, const int16_t* b, uint32_t n);
float orc_similarity(const int16_t* a, const int16_t* b, uint32_t sl);
void orc_fht(float* buf, uint32_t m);
uint32_t orc_fht_hash(const orc_functions* fn, uint32_t fidx, const int16_t* q15);
void orc_codes(const orc_functions* fn, const int16_t* q15, uint32_t* out /*L*/);
void orc_sketch(const orc_functions* fn, const int16_t* q15, uint64_t* out /*32*/);
void orc_sort_pairs_24(const uint32_t* hashes_in, const uint32_t* idx_in, uint32_t n, uint32_t* hashes_out, uint32_t* idx_out);
float orc_failure_probability(const orc_functions* fn, uint32_t hash_length, uint64_t tables, uint64_t max_tables, float kth);
uint32_t orc_max_sketch_diff(float kth);

/* MaxBuffer driver for the reference's known-answer tests. Returns number of entries; minval before the final filter. */
int orc_maxbuffer_run(uint32_t k, const uint32_t* ids, const float* vals, int n, uint32_t* out_ids, float* out_vals, float* minval);

/* --- L1 */
orc_index* orc_index_build(const orc_functions* fn /*deep-copied*/, const float* data, uint32_t n);
orc_index* orc_index_import(const uint8_t* stream, uint64_t len); /* Index::serialize bytes */
void orc_index_free(orc_index* ix);
uint32_t orc_anchor(const orc_index* ix, uint32_t t, uint32_t hash);
/* ranges[24][L][2] padded positions, anchors[L] */
void orc_query_ranges(const orc_index* ix, const uint32_t* codes, uint32_t* anchors, uint32_t* ranges);
/* Index::search: returns count of ids (best first) */
int orc_index_search(const orc_index* ix, const float* q, uint32_t k, float recall, float max_sim, uint32_t* out, orc_trace* tr);
/* Index::search with FilterType::None (1) / FilterType::Simple (2), collection.hpp:671-765 (both ignore max_sim); 0 = the
 * default path */
int orc_index_search_filter(const orc_index* ix, const float* q, uint32_t k, float recall, float max_sim, int filter_type,
                            uint32_t* out, orc_trace* tr);

/* --- L3 (CLANN) */
uint64_t orc_num_clusters(float factor, uint64_t n);
float orc_ndarray_dot(const float* x, const float* y, size_t len);
float orc_distance_point(const float* row, float row_norm, const float* q, uint32_t d);
uint64_t orc_gmm(const float* data, uint64_t n, uint32_t d, uint64_t K, uint64_t* centers, uint64_t* assignment, float* radii);

int orc_topk_run(uint32_t k, const float* dists, const uint64_t* ids, int n, float* out_d, uint64_t* out_ids, uint8_t* added,
                 uint64_t* top_id, float* top_dist);
void orc_sort_clusters(const float* data, uint32_t d, const uint64_t* centers, uint64_t K, const float* q, uint64_t* order);

typedef struct orc_clann orc_clann;
/* fns: either 1 shared function set or K per-cluster sets (n_fns in {1, K}); streams: optional per-cluster Index::serialize
   blobs (then fns is ignored for those clusters). data is borrowed. */
orc_clann* orc_clann_create(const float* data, uint64_t n, uint32_t d, uint32_t k, float delta, uint64_t K,
                            const uint64_t* centers, const uint64_t* assignment, const float* radii);
int orc_clann_set_cluster_stream(orc_clann* c, uint64_t ci, const uint8_t* stream, uint64_t len);
int orc_clann_build_cluster(orc_clann* c, uint64_t ci, const orc_functions* fn);
int orc_clann_search(orc_clann* c, const float* q, uint64_t* out_ids, float* out_dists, uint64_t* order_out,
                     uint64_t* counters /*visited, distcomp, candidates*/);
/* the same search with the reference's per-visit metric rows (metrics/mod.rs:84-112): visit_log[v*3 + ..] = {cluster,
 * points_added, cluster_distance_computations} for the first visit_cap visits */
int orc_clann_search_visits(orc_clann* c, const float* q, uint64_t* out_ids, float* out_dists, uint64_t* order_out,
                            uint64_t* counters, uint64_t* visit_log, uint64_t visit_cap);
void orc_clann_free(orc_clann* c);

#ifdef __cplusplus
}
#endif
#endif
