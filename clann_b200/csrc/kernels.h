// Host-callable launchers of the CUDA kernels (all asynchronous on `stream`). Device pointers unless noted.
#pragma once

#include "common.cuh"

namespace clann {

// ---- build: CLANN layer (gmm.rs, angulardata.rs)
void launch_row_norms(const float* data, uint64_t n, uint32_t d, float* norms, cudaStream_t s);
// One greedy k-center pass for centre number c (gmm.rs:40-53). keys[K]: packed arg-max of pass c is written to keys[c];
// the centre of pass c (c > 0) is decoded from keys[c-1].
// cc: scratch of K floats (distances from the pass's centre to the earlier ones, for the triangle-inequality filter); may be null.
// Processes rows [row0, row1) of the full arrays (a rank's share of a sharded clustering; 0, n otherwise).
void launch_gmm_pass(const float* data, const float* norms, uint64_t row0, uint64_t row1, uint32_t d, uint32_t c, uint64_t* keys,
                     float* dist, uint32_t* assign, float* cc, cudaStream_t s);
void launch_gmm_finish(const uint64_t* keys, uint32_t K, uint64_t n, const float* dist, const uint32_t* assign,
                       uint32_t* centers, float* radii, uint32_t* sizes, cudaStream_t s);
void launch_widen_f16(const uint16_t* in, uint64_t count, float* out, cudaStream_t s);
void launch_gather_rows(const float* data, const uint32_t* rows, uint32_t count, uint32_t d, float* out, cudaStream_t s);

// ---- build: PUFFINN layer
// format/unit_vector.hpp:61-89 for rows perm[row] (perm may be null = identity) -> q15[row*sl..]
void launch_store_q15(const float* data, const uint32_t* perm, uint64_t rows, uint32_t d, uint32_t sl, int16_t* q15, cudaStream_t s);
// filterer.hpp:76-102: sketches[out_row*32 + s] for every row of every tile. planes: [fset][2048][sl] Q15.
void launch_sketch(const int16_t* q15, const RowTile* tiles, uint32_t n_tiles, const int16_t* planes, uint32_t sl,
                   uint64_t* sketches, cudaStream_t s);
// independent.hpp:70-86 over crosspolytope.hpp:187-209: codes[fset*fset_stride + t*code_stride + out_row].
// signbits: [fset][L*fph][3][max(1, npts/32) words], bit i set = sign -1.
void launch_codes(const int16_t* q15, const RowTile* tiles, uint32_t n_tiles, const uint32_t* signbits, HashGeom g,
                  uint32_t* codes, uint64_t code_stride, uint64_t fset_stride, cudaStream_t s);

// The same projection on the tensor pipe (kernels_tc.cu): exact integer dot products from int8 slices by tcgen05.mma.kind::i8
// into TMEM, the sign decided wherever the rounding slack cannot change it, the remaining ~1.4 % of the pairs evaluated the
// reference's way. Bit-identical to launch_sketch. Tiles hold up to 128 consecutive rows of one function set.
struct SketchTcTile {
    uint32_t in_row0, out_row0, count, fset;
};
bool sketch_tc_supported(uint32_t sl);
uint32_t sketch_tc_kp(uint32_t sl);  // slice row length: SL rounded up to 128; a slice row is 2 * kp bytes (high bytes, low bytes)
void launch_q15_slices(const int16_t* q15, uint64_t rows, uint32_t sl, uint8_t* out, cudaStream_t s);
void launch_sketch_tc(const void* tiles, uint32_t n_tiles, const uint8_t* row_slices, uint32_t slice_row_base, uint64_t slice_rows,
                      const uint8_t* plane_slices, uint32_t n_fsets, const int16_t* q15, const int16_t* planes, uint32_t sl,
                      uint64_t* sketches, cudaStream_t s);

struct SortSegment {
    uint64_t base;          // offset of the segment in keys / idx arrays
    uint64_t scratch_base;  // offset in the scratch arrays (only used when len > smem capacity)
    uint32_t len;
    uint32_t pad;
};
// sorthash.hpp:133-194: stable LSD radix sort (3 byte passes) of every segment by its 24-bit key; payload = position in
// the segment. keys are sorted in place, idx receives the payload. max_len = largest segment.
// Stable partition of the rows by cluster id (index.rs:188-192): perm[offsets[c] + j] = j-th row (ascending) assigned to cluster c.
// counts = scratch of n_chunks * K u32 (one private counter row per warp-sized chunk of the rows).
void launch_assign_partition(const uint32_t* assign, uint64_t n, uint32_t K, const uint64_t* offsets, uint32_t* counts, uint32_t n_chunks,
                             uint32_t* perm, cudaStream_t s);
void launch_segment_sort(const SortSegment* segs, uint32_t n_segs, uint32_t max_len, uint32_t* keys, uint32_t* idx,
                         uint32_t* scratch_keys, uint32_t* scratch_idx, cudaStream_t s);
uint32_t segment_sort_smem_capacity();  // segments up to this length are sorted entirely in shared memory

// Bucket directory over the top 12 bits of the sorted codes (the role of PrefixMap::prefix_index, prefixmap.hpp:86,231-240,
// at 12 instead of 13 bits): dir[(c*L + t)*kDirEntries + b] = first position, inside cluster c of table t, whose code has a
// top 12 bits >= b (b = 0..4096; entry 4096 = cluster size). Built after the sort; brute-force / foreign clusters are skipped.
constexpr uint32_t kDirBits = 12;
constexpr uint32_t kDirEntries = (1u << kDirBits) + 1;
void launch_build_dir(const uint32_t* tbl_hash, uint64_t n, const uint64_t* offsets, const uint8_t* skip, uint32_t K, uint32_t L,
                      uint32_t* dir, cudaStream_t s);

// crosspolytope.hpp:16-88 Monte-Carlo collision estimates, est[(m+2)*201]
void launch_cp_estimates(uint32_t m, uint32_t reps, uint64_t seed, float* est, uint32_t* scratch_counts, cudaStream_t s);

// ---- search
struct SearchParams {
    // geometry
    HashGeom g;
    uint32_t k, K, n_fsets;
    uint64_t n;
    uint32_t stop_words;  // u32 words per (depth, bin) row of the stop table = ceil((L+1)/32)
    // index (device)
    const float* data;         // [n][d] original rows
    const float* norms;        // [n]
    const uint32_t* perm;      // [n] cluster-sorted row -> point id
    const uint64_t* offsets;   // [K+1]
    const int16_t* q15;        // [n][sl] cluster-sorted
    const uint64_t* sketches;  // [n][32]
    const uint32_t* tbl_hash;  // cluster-major: table t of cluster c at table_base(offsets[c], nc, L, t) (common.cuh)
    const uint32_t* tbl_idx;   // same layout
    const uint32_t* tbl_dir;   // [K][L][kDirEntries] bucket directory over the top 12 code bits
    const float* center_rows;  // [K][d]
    const float* center_norms; // [K]
    const float* radii;        // [K]
    const uint8_t* brute;      // [K]
    const uint32_t* fset_of;   // [K]
    const uint8_t* owner;      // [K] owning shard (multi-GPU), all 0 on one GPU
    const uint32_t* stop;      // [n_fsets][24][201][stop_words] stop decision bit t (independent.hpp:108-119, host glibc pow)
    const uint8_t* msd;        // [65536] max_sketch_diff by (dot + 32768) (filterer.hpp:108-111, host glibc acosf)
    const uint32_t* msd_thr;   // [64] msd_thr[m] = smallest v with msd[v] <= m (msd is non-increasing): msd[v] = #{m : msd_thr[m] > v}
    uint32_t shard_rank;
    uint32_t max_cluster;      // largest cluster size (sizes the per-CTA similarity memo)
    uint32_t prefetch_rows;    // probe kernel: bulk-prefetch the Q15 rows of a batch into the L2 before gathering them
    uint32_t reserve_sms;      // probe kernel: CTAs that land on the last reserve_sms SMs exit at once (room for other streams' kernels)
};

constexpr uint32_t kFsMeta = 52;  // u32 words of stream metadata per query (QueryBatch::fs_meta)

struct QueryBatch {
    uint64_t nq;
    const float* queries;   // [nq][d]
    float* qnorm;           // [nq]
    int16_t* q15;           // [nq][sl]
    uint32_t* codes;        // [n_fsets][L][nq]   (table-major per function set)
    uint64_t* sketches;     // [n_fsets][nq][32]
    float* cdist;           // [nq][K] distance to every centre (unsorted; the probe kernel walks it in key order)
    // Tensor-pipe screen (launch_center_order with the tf32 GEMM): cdist[q][c] is the reference's exact fp32 value wherever it is
    // below exact_limit[q]; every other entry is a lower bound of the exact value that is itself >= exact_limit[q]. The probe
    // re-evaluates the whole row exactly when its walk reaches the limit (and sets it to +inf). null = every entry is exact.
    float* exact_limit;     // [nq]
    uint32_t* first;        // [nq] nearest cluster (scratch: sorted in place to derive qperm)
    uint32_t* qperm;        // [nq] work order: queries sorted by nearest cluster
    // per-query running state (also the multi-GPU exchange unit): see kernels_search.cu
    uint8_t* state;
    uint32_t* work_counter; // [1]
    // similarity memo scratch owned by the index workspace: memo_slots regions of memo_stride u16 (one per resident warp /
    // CTA of the probe kernel); null = no memo (the largest cluster would make it too big)
    uint16_t* memo;
    uint64_t memo_stride;
    uint32_t memo_slots;
    // dense first-visit similarities (launch_dense_sims): dense[q * dense_stride + local id] = Q15 similarity (dot + 32768) of
    // query q to every row of its nearest cluster, i.e. the memo of its first visit filled in advance; null = not computed
    uint16_t* dense;
    uint64_t dense_stride;
    // first-visit anchors and ranges (launch_first_ranges): pre_anchor[q * L + t] = anchor of query q in table t of its nearest
    // cluster; pre_range[(q * 24 + depth - 1) * L + t] = segments of the range at `depth` | upward << 31; null = not computed
    // batch statistics for the host's choice of schedule (no synchronisation: the host reads the page-locked words whenever it
    // next looks): stats_dev[0] = running sum of clusters visited, [1] = finished-block ticket; the last block of k_finish
    // publishes {sum of clusters visited, queries} of the batch to stats_host (mapped host memory); null = not collected
    unsigned long long* stats_dev;
    volatile unsigned long long* stats_host;
    uint32_t* pre_anchor;
    uint32_t* pre_range;
    // pre_lcp[q * L + t] = the stride-12 common-prefix samples of table_anchor (up.x, up.y, dn.x, dn.y); with pre_range null
    // the probe takes anchors and samples from here and evaluates the ranges itself, depth by depth, as far as it gets
    uint4* pre_lcp;
    // First-visit candidate stream (launch_first_stream): the candidates of every query's first visit laid out in the order
    // search_maps consumes them — depth by depth, ring sweep by ring sweep — with the sketch test already evaluated as a
    // Hamming distance. Per query q (stride fs_cap segments; a segment = 4 candidates = one ring slot, collection.hpp:802-808):
    //   fs_idx[(q*fs_cap + s)*4 + j]  local id of candidate j of stream segment s (u16: clusters up to 65 536 rows)
    //   fs_hd[q*fs_cap + s]           byte j = popcount(sketch[id_j][slot] ^ query_sketch[slot]), slot = s mod 32 (filterer.hpp:28-31)
    //   fs_tab[(q*fs_cap + s) / 32]   table of the first segment of the ring sweep that starts at s (stop rule, collection.hpp:927-943)
    //   fs_meta[q*kFsMeta + ..]       [depth-1] = segments of the depth's stream, [24 + depth-1] = its first segment in the
    //                                 stream, [48] = lowest depth present (25 = none): deeper levels are evaluated by the probe
    // The block of a depth starts at a multiple of 32 segments. How deep the stream goes is a prediction (the first depth at
    // whose end the stop rule fires for the cluster's true k-th similarity); it never changes a result: the probe takes what
    // is there and evaluates the rest itself. null = not computed.
    uint16_t* fs_idx;
    uint32_t* fs_hd;
    uint8_t* fs_tab;
    uint32_t* fs_meta;
    uint32_t fs_cap;
    // cluster-sharded search (index.cu search_sharded): first_is_own = every query's nearest cluster is owned by this rank (round one);
    // shard_packed[q] = (order_bits(bound) << 32) | (0xffffffff - clusters consumed in round one), the input of round two
    bool first_is_own;
    const unsigned long long* shard_packed;
    // per-visit log for the reference's cluster-granularity metrics (metrics/mod.rs:84-112, result_schema.sql:73-90), opt-in
    // (clann_set_option "visit_log" = rows kept per query): visit_log[(q * visit_cap + v) * 4 + ..] = {cluster + 1, points_added
    // (heap adds that returned true, index.rs:367-372,405-410), cluster_distance_computations (prune-test evaluation + PUFFINN's
    // counter or the brute-force list length, index.rs:348,378,421), nanoseconds spent in the visit}; null = not recorded
    uint32_t* visit_log;
    uint32_t visit_cap;
    // outputs
    uint32_t* out_ids;      // [nq][k]
    float* out_dists;       // [nq][k]
    uint32_t* out_counts;   // [nq]
    // counters [nq]
    unsigned long long* cnt_candidates;
    unsigned long long* cnt_distcomp;
    uint32_t* cnt_visited;
};

// Process-global tuning knobs for A/B measurements (the defaults are the shipped configuration). The first read of a key
// takes its value from the environment variable CLANN_TUNE_<KEY> when set; clann_tune() overrides it at any time.
int64_t tune_get(const char* key, int64_t dflt);
void tune_set(const char* key, int64_t value);

uint64_t query_state_bytes(uint32_t k);
void launch_query_tiles(uint64_t nq, uint32_t F, RowTile* sk_tiles, RowTile* code_tiles, SketchTcTile* tc_tiles, SortSegment* seg,
                        cudaStream_t s);
void launch_prep_queries(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
// Tensor-pipe centre scoring (kernels_tc.cu): approx[q*K + c] = 1 - dot_tf32(q, c) / (|q| |c|), within kCentreEps of the exact value.
constexpr float kCentreEps = 0.004f;   // tf32 operands: |error| <= 2^-9 |q||c| on the dot, i.e. 0.00195 on the cosine; doubled
constexpr uint32_t kCentreExact = 32;  // candidates per query evaluated exactly by launch_center_refine
bool center_gemm_tc_supported(uint32_t d);
void launch_center_gemm_tc(const float* queries, const float* qnorm, uint64_t nq, const float* center_rows, const float* center_norms,
                           uint32_t K, uint32_t d, float* approx, cudaStream_t s);
// cdist + first (nearest cluster); the caller then sorts `first` with launch_segment_sort to obtain qperm.
void launch_center_order(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
void launch_init_state(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
// Q15 similarity of every query to EVERY row of its nearest cluster (b.first sorted, b.qperm = the matching query ids), streamed
// cluster by cluster with every lane busy, into b.dense: the rerank of a query's first visit — 83 % of all visits on the
// glove-100 shape — then only looks similarities up instead of gathering rows. Returns false when the geometry is unsupported.
bool launch_dense_sims(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
bool dense_sims_supported(const SearchParams& p);
// Anchors (prefixmap.hpp:36-57) and the ranges of all 24 depths (prefixmap.hpp:267-304) of every query in its nearest cluster,
// one thread per (query, table) instead of a dependent chain inside the probe; needs b.codes, sorted b.first and b.qperm.
void launch_first_ranges(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
// The first-visit candidate stream of every query (QueryBatch::fs_*): anchors, per-depth ranges, table indices and the sketch
// test of every candidate of the first visit, as one warp per query with every load independent of the search state — the
// probe kernel then replays the sequential decisions from a linear stream. Needs b.dense (launch_dense_sims) for the depth
// prediction, sorted b.first and b.qperm. Returns false when the geometry is unsupported (then nothing was written).
bool launch_first_stream(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
// Trace export for the parity tests: anchors[nq*L] and ranges[nq*24*L*2] of the batch's queries in `cluster`, in the
// reference's padded table coordinates (needs b.codes of the last search_begin).
void launch_export_ranges(const SearchParams& p, const QueryBatch& b, uint32_t cluster, uint32_t* anchors, uint32_t* ranges,
                          cudaStream_t s);
// stop_at_foreign: 0 = walk every cluster (one GPU holds the whole index); 1 = advance every unfinished query through the consecutive
// clusters this shard owns and stop at the first foreign one (stepping protocol; round one of the sharded search); 2 = round two of
// the sharded search: walk the whole order from the start, skip what round one consumed and what other shards own, prune with the
// agreed bound (QueryBatch::shard_packed).
void launch_probe(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s);
void launch_probe_warp(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s);  // one warp per query
// Cluster-sharded search (index.cu search_sharded): routing, bound exchange and result merge helpers.
void launch_shard_select_owned(const uint32_t* first_all, const uint8_t* owner, uint32_t rank, uint64_t nq, uint32_t* list, uint32_t* count,
                               cudaStream_t s);
void launch_shard_select_open(const unsigned long long* packed, uint64_t nq, uint32_t* list, unsigned long long* packed_local,
                              uint32_t* count, cudaStream_t s);
void launch_shard_select_mine(const float* cdist, const float* exact_limit, const float* radii, const uint8_t* owner, uint32_t rank,
                              uint32_t K, const uint32_t* list_in, const unsigned long long* packed_in, uint32_t count_in,
                              uint32_t* list_out, unsigned long long* packed_out, uint32_t* count_out, cudaStream_t s);
void launch_shard_gather_rows(const float* all, const uint32_t* list, uint32_t count, uint32_t d, float* out, cudaStream_t s);
void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t s);
void launch_shard_pack_bounds(const uint8_t* state, uint32_t k, const uint32_t* list, uint32_t count, unsigned long long* packed,
                              cudaStream_t s);
void launch_shard_collect(const uint8_t* state, uint32_t k, const uint32_t* list, uint32_t count, unsigned long long* top, bool merge,
                          unsigned long long* counters, cudaStream_t s);
void launch_shard_final_merge(const unsigned long long* all, uint32_t world, uint64_t nq, uint32_t k, uint32_t* out_ids, float* out_dists,
                              uint32_t* out_counts, cudaStream_t s);
void launch_merge_states(const SearchParams& p, const QueryBatch& b, const uint8_t* all_states, int world, uint32_t* active, cudaStream_t s);
void launch_finish(const SearchParams& p, const QueryBatch& b, cudaStream_t s);
uint32_t probe_memo_slots();  // upper bound on the memo regions any probe launch uses on the current device

// Single PUFFINN index query (legacy CPUFFINN_search_cosine / clann_puffinn_search): one cluster, explicit recall / max_sim, Q15
// brute force when n < 100 (collection.hpp:550-555). filter_type = the reference's FilterType (collection.hpp:22-34): 0 Default
// (search_maps), 1 None, 2 Simple (:671-765; both ignore max_sim). out_ids[k] local ids best first, out_count, out_stop =
// (depth << 16 | table index) where the stop rule fired, 0 = never.
void launch_puffinn_search(const SearchParams& p, const QueryBatch& b, const uint32_t* stop_table, float max_sim, int filter_type,
                           uint32_t* out_ids, uint32_t* out_count, uint32_t* out_distcomp, uint32_t* out_stop, cudaStream_t s);

}  // namespace clann
