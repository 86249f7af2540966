/* clann_b200 — C ABI of the B200-native CLANN build + search hot path.
 *
 * This header is the drop-in boundary. It has two halves:
 *
 *  (1) The eight legacy `CPUFFINN_*` symbols, byte-compatible with the reference's
 *      /root/reference/libpuffinn-ffi/c_binder.h:10-27 (Rust declarations: src/puffinn_binds/puffinn_sys.rs:8-52),
 *      so the unmodified Rust crate can link `libclann_b200.so` instead of compiling c_binder.cpp.
 *  (2) A batched `clann_*` ABI that moves all of ClusteredIndex::{new,build,search}
 *      (/root/reference/src/core/index.rs:71-91,177-289,311-439) behind the boundary, because one FFI call per
 *      (query, cluster) with a host-side early exit between calls cannot feed a GPU. lib.rs's init_with_config / build /
 *      search (src/lib.rs:118,142,183) become one-line wrappers over it.
 *
 * Plain pointers and sizes only; no C++ or torch types. Every entry point returns 0 on success or a negative
 * clann_status; clann_last_error() gives the message (thread-local). No exception crosses this boundary.
 * There is no CPU fallback: every call needs a CUDA device and fails with CLANN_ERR_CUDA without one.
 */
#ifndef CLANN_B200_H
#define CLANN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------------
 * (2) batched ABI
 * ------------------------------------------------------------------------------------------------------------------ */

typedef struct clann_index clann_index; /* opaque; owns device memory */

/* Mirrors the hot-path fields of Config (src/core/config.rs:17-35). dataset_name / metrics_output are host-side
 * concerns of the Rust crate and stay there. */
typedef struct {
    uint64_t num_tables;       /* L, PUFFINN tables per cluster                                  (config.rs:19) */
    float num_clusters_factor; /* K = max(1, floor(f64(factor) * sqrt(n)))        (config.rs:22, index.rs:78-80) */
    uint64_t k;                /* neighbours returned                                            (config.rs:25) */
    float delta;               /* target recall                                                  (config.rs:28) */
} clann_config;

typedef enum {
    CLANN_OK = 0,
    CLANN_ERR_DATA = -1,       /* ClusteredIndexError::DataError            (errors.rs:10-11) e.g. empty dataset */
    CLANN_ERR_CONFIG = -2,     /* ClusteredIndexError::ConfigError          (errors.rs:7-8)                      */
    CLANN_ERR_CREATION = -3,   /* ClusteredIndexError::PuffinnCreationError (errors.rs:19-20)                    */
    CLANN_ERR_SEARCH = -4,     /* ClusteredIndexError::PuffinnSearchError   (errors.rs:22-23)                    */
    CLANN_ERR_NOT_BUILT = -5,  /* ClusteredIndexError::IndexNotFound        (errors.rs:25-26)                    */
    CLANN_ERR_BOUNDS = -6,     /* ClusteredIndexError::IndexOutOfBounds     (errors.rs:28-29)                    */
    CLANN_ERR_SERIALIZE = -7,  /* ClusteredIndexError::SerializeError       (errors.rs:34-35)                    */
    CLANN_ERR_CUDA = -8,       /* no device / CUDA runtime failure (no reference counterpart: it has no GPU path) */
    CLANN_ERR_ARG = -9         /* null pointer or malformed argument                                             */
} clann_status;

/* ClusteredIndex::new (index.rs:71-91) behind init_with_config (lib.rs:118-124). `data` is n x d row-major f32 in HOST
 * memory; it is copied to the device. n == 0 -> CLANN_ERR_DATA ("empty dataset"). Uses the current CUDA device. */
int clann_init_with_config(const float* data, uint64_t n, uint32_t d, const clann_config* config, clann_index** out);

/* The same with the rows as IEEE half precision (dtype 1; BASELINE.json's "100M x 96 fp16" configuration) and / or already
 * resident on the device (on_device != 0: `data` is a device pointer, copied device to device). fp16 rows are widened to f32
 * exactly on the device: the index is the one the reference builds from the same rows widened to f32. dtype 0 = f32. */
int clann_init_with_config_ex(const void* data, uint64_t n, uint32_t d, const clann_config* config, int dtype, int on_device,
                              clann_index** out);

/* Options, to be set between init and build. Unknown key -> CLANN_ERR_ARG.
 *   "seed"            seed of the function generator (the reference seeds from the wall clock, typedefs.hpp:17)
 *   "function_sets"   0 (default) = one hash/sketch function set shared by all clusters (query hashed once);
 *                     1 = one set per cluster, as the reference does (collection.hpp:128-135,265-268)
 *   "strict"          1 (default) = reproduce search_maps' candidate order and quirks exactly;
 *   "shard_rank", "shard_count"   cluster ownership for multi-GPU: this index only probes clusters c with
 *                     owner(c) == shard_rank (longest-processing-time assignment on cluster sizes), default 0 / 1  */
int clann_set_option(clann_index* index, const char* key, int64_t value);

/* Config.delta of a built index without rebuilding it (only the stop rule reads it, collection.hpp:927-943): the recall sweep of
 * BASELINE.json's last configuration searches one index at delta = 0.8 / 0.9 / 0.95. */
int clann_set_delta(clann_index* index, float delta);

/* Parity mode: impose a clustering instead of running greedy k-center (centers[K] = point ids, assignment[n] = cluster
 * of every point, radii[K]); and import the function set (SimHash planes, FHT signs, collision estimates) of one
 * cluster from the bytes of puffinn::Index::serialize (collection.hpp:185-203). Importing implies function_sets = 1. */
int clann_set_clustering(clann_index* index, uint64_t K, const uint64_t* centers, const uint64_t* assignment, const float* radii);
int clann_import_reference(clann_index* index, uint64_t cluster, const void* blob, uint64_t len);
/* Same, from raw arrays: planes[2048*SL] Q15, signs[L*fph*3*2^m] (+1/-1), est[(m+2)*201]. cluster == UINT64_MAX sets
 * the shared set. */
int clann_set_functions(clann_index* index, uint64_t cluster, const int16_t* planes, const int8_t* signs, const float* est);

/* ClusteredIndex::build (index.rs:177-289): greedy k-center, per-cluster Q15 rows, sketches, table codes, sorted tables. */
int clann_build(clann_index* index);

/* ClusteredIndex::search (index.rs:311-439) for a batch of nq queries (HOST pointers, nq x d row-major f32).
 * ids[nq*k] (0xFFFFFFFF pad), dists[nq*k] (+inf pad), counts[nq] = pairs returned per query (<= k), ascending distance. */
int clann_search(clann_index* index, const float* queries, uint64_t nq, uint32_t* ids, float* dists, uint32_t* counts);
/* Same with DEVICE pointers (queries and outputs already resident in HBM); asynchronous on `stream` (a cudaStream_t). */
int clann_search_device(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists,
                        uint32_t* d_counts, void* stream);
/* Batch pipelining for a stream of batches (no counterpart in the reference, which answers one query at a time): the same
 * search, issued on one of three internal streams (knob pipeline_depth) with its own workspace, NOT ordered after the caller's streams, so that
 * consecutive batches overlap on the device (the next batch's hashing beside the current probe, its probe in the SMs the
 * current probe's last wave leaves idle). The query buffer must be complete when the call is made and every batch in
 * flight needs its own output buffers. clann_search_flush makes `stream` wait for every batch issued so far; results are
 * complete once that stream has been synchronised. Counters and the search profile are those of the stream-ordered calls. */
int clann_search_device_async(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists,
                              uint32_t* d_counts);
int clann_search_flush(clann_index* index, void* stream);
/* The same pipelining with HOST buffers (page-locked, or the copies degrade to synchronous ones): host-to-device copy,
 * search and the three device-to-host copies of one batch are queued on the batch's internal stream and the call returns;
 * the output buffers are complete after clann_search_wait, which blocks the host until every batch issued so far is done. */
int clann_search_async(clann_index* index, const float* queries, uint64_t nq, uint32_t* ids, float* dists, uint32_t* counts);
int clann_search_wait(clann_index* index);

/* Multi-GPU stepping (one process per GPU, clusters sharded by owner). clann_search_begin prepares the batch (query
 * hashing, centre ordering); each clann_search_step advances every unfinished query through the consecutive clusters
 * this rank owns and stops at the first cluster owned by another rank; the caller then exchanges the per-query state
 * (clann_state_bytes() bytes per query, device pointer from clann_state_ptr) with an all-gather and calls
 * clann_search_merge, which adopts for every query the state of the rank that advanced it. *active_out = queries not
 * yet finished anywhere. clann_search_end writes the results. */
int clann_search_begin(clann_index* index, const float* d_queries, uint64_t nq, void* stream);
int clann_search_step(clann_index* index, void* stream);
uint64_t clann_state_bytes(const clann_index* index);
void* clann_state_ptr(clann_index* index);
int clann_search_merge(clann_index* index, const void* d_all_states, int world, uint64_t* active_out, void* stream);
int clann_search_end(clann_index* index, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, void* stream);

/* Cluster-sharded search, the multi-GPU mode BASELINE.json's north_star names: one process per GPU, each index built
 * with options shard_count = world and shard_rank = rank over the SAME data (every rank derives the same clustering and the same
 * longest-processing-time ownership of clusters, and builds tables for its own clusters only). Every rank then calls
 * clann_search_sharded with the same global batch (device pointers) and receives the complete results. Inside: each rank scores its
 * slice of the queries against the replicated centres, one all-gather routes every query to the owner of its nearest cluster, that
 * rank runs the reference's loop for as long as the walk stays in its own clusters, one all-reduce(min) of 8 bytes per query
 * publishes the bound reached, every rank visits its own clusters among those the bound does not prune for the queries still open,
 * and one all-gather of nq x k x (distance, id) feeds a k-way merge. Visits are a superset of the single-GPU search's (recall >=).
 * The transport is NCCL over NVLink (clann_comm_unique_id on one rank, broadcast the 128 bytes by any means, clann_comm_init on
 * all — libnccl.so.2 is resolved at run time), or any two collectives the caller supplies (clann_set_collectives; e.g. MPI, or the
 * in-process transport of the tests). The exact stepping protocol above remains available. */
typedef int (*clann_allgather_fn)(void* ctx, const void* d_send, void* d_recv, uint64_t bytes_per_rank, void* stream);
typedef int (*clann_allreduce_min_u64_fn)(void* ctx, void* d_buf, uint64_t count, void* stream);
int clann_comm_unique_id(uint8_t* out, uint64_t cap /* >= 128 */);
int clann_comm_init(clann_index* index, int rank, int world, const uint8_t* unique_id);
int clann_set_collectives(clann_index* index, clann_allgather_fn allgather, clann_allreduce_min_u64_fn allreduce_min, void* ctx);
int clann_search_sharded(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts,
                         void* stream);
/* Two whole batches of nq queries each in flight: the phases of the two searches are interleaved on two internal streams, so
 * that one batch's waits (count read-backs, collectives, the latency-bound second round) overlap the other's first round — the
 * batch pipelining of clann_search_device_async for the sharded search. Results are those of two clann_search_sharded calls. */
int clann_search_sharded_pair(clann_index* index, const float* d_queries_a, const float* d_queries_b, uint64_t nq, uint32_t* d_ids_a,
                              float* d_dists_a, uint32_t* d_counts_a, uint32_t* d_ids_b, float* d_dists_b, uint32_t* d_counts_b, void* stream);
/* The same for 1 to 4 batches (host arrays of device pointers, one entry per batch). */
int clann_search_sharded_multi(clann_index* index, int n_batches, const float* const* d_queries, uint64_t nq, uint32_t* const* d_ids,
                               float* const* d_dists, uint32_t* const* d_counts, void* stream);
/* Streaming form: a software pipeline across calls. Every call takes one new batch and advances each of the up to four batches in
 * flight by one phase (route | round one | scoring of the open queries | round two + merge), newest first, each on its own
 * internal stream: the latency-bound second round of one batch runs in the shadow of a later batch's first round, a count the host
 * reads back was produced a whole call earlier, and a collective that waits for a slower rank stalls only its own batch. The
 * query and output buffers of a batch must stay untouched until three more batches have been submitted or clann_search_sharded_flush
 * has returned; its results are then complete in stream order on the stream of that call. Every rank must make the same sequence
 * of submit / flush calls (the collectives are issued in call order). Results per batch are those of clann_search_sharded. */
int clann_search_sharded_submit(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists,
                                uint32_t* d_counts, void* stream);
int clann_search_sharded_flush(clann_index* index, void* stream);
/* queries routed to this rank in round one / still open in round two of the last clann_search_sharded; phase_ms[6] (may be NULL) =
 * device time of its phases: route scoring, route all-gather + selection, round one, bound all-reduce + selection, round two, merge */
int clann_shard_stats(clann_index* index, uint64_t* routed_round_one, uint64_t* open_round_two, float* phase_ms);

/* Per-query counters of the last search, the same quantities the reference keeps (performance.hpp:72-86 plus the
 * cluster count): any pointer may be NULL. Arrays of nq. */
int clann_get_counters(clann_index* index, uint64_t nq, uint64_t* candidates, uint64_t* distance_computations,
                       uint32_t* clusters_visited);

/* Golden-dump access for parity tests (device -> host copies). *size receives the byte size; dst may be NULL to query it. */
typedef enum {
    CLANN_X_NUM_CLUSTERS = 0, /* u64 K */
    CLANN_X_CENTERS = 1,      /* u64[K]  centre point ids */
    CLANN_X_ASSIGNMENT = 2,   /* u64[n]  cluster of every point */
    CLANN_X_RADII = 3,        /* f32[K] */
    CLANN_X_OFFSETS = 4,      /* u64[K+1] cluster c owns rows [off[c], off[c+1]) of the cluster-sorted arrays */
    CLANN_X_PERM = 5,         /* u32[n]  cluster-sorted row -> point id (the concatenated ClusterCenter.assignment) */
    CLANN_X_Q15 = 6,          /* i16[n_c*SL] of cluster `arg` */
    CLANN_X_SKETCHES = 7,     /* u64[n_c*32] of cluster `arg` */
    CLANN_X_TABLE_HASHES = 8, /* u32[L*n_c] sorted hashes of cluster `arg`, table-major, unpadded */
    CLANN_X_TABLE_INDICES = 9,/* u32[L*n_c] matching local point indices */
    CLANN_X_BRUTE = 10,       /* u8[K] brute-force flag (index.rs:204-205) */
    CLANN_X_NORMS = 11,       /* f32[n] row norms (angulardata.rs:12-19) */
    CLANN_X_EST = 12,         /* f32[(m+2)*201] collision estimates of function set of cluster `arg` */
    CLANN_X_QUERY_CODES = 13, /* u32[nq*L] table codes of the last search batch for the function set of cluster `arg` */
    CLANN_X_QUERY_SKETCHES = 14, /* u64[nq*32] likewise */
    CLANN_X_CLUSTER_ORDER = 15,  /* u32[nq*K] visiting order of the last search batch (index.rs:592-616) */
    CLANN_X_BUILD_MS = 16,    /* f64[5] last build: gmm phase, hashing (store+sketch+codes), table sort, total, the K k-center
                                 passes alone (device ms) */
    CLANN_X_TABLE_DIR = 17,   /* u32[L*4097] bucket directory of cluster `arg`: entry b of table t = first position whose top 12
                                 code bits are >= b, entry 4096 = cluster size (the role of PrefixMap::prefix_index,
                                 prefixmap.hpp:86,231-240, at 12 instead of 13 bits), table-major */
    CLANN_X_REFERENCE_STREAM = 18, /* bytes of puffinn::Index::serialize (collection.hpp:185-203) for cluster `arg`: what the CPU
                                 reference would write for the same rows and functions, loadable by Index(std::istream&) —
                                 persistence compatible with the reference without HDF5 (CPUFFINN_save_index's payload) */
    /* order-free probe traces of the last search batch against cluster `arg` (SURVEY.md 8b clann_export_trace), in the
       reference's coordinates: positions in the table padded by 12 sentinels each side (prefixmap.hpp:215-226) */
    CLANN_X_QUERY_ANCHORS = 19, /* u32[nq*L]       PrefixMapQuery anchors (prefixmap.hpp:36-57,250-260) */
    CLANN_X_QUERY_RANGES = 20,  /* u32[nq*24*L*2]  {start, end} returned by call it = 0..23 of get_next_range
                                   (prefixmap.hpp:267-304; depth 24 - it), layout [query][it][table][2] */
    CLANN_X_STOP_POINTS = 21,   /* u32[nq*2]       {depth, table index} at which the stop rule (collection.hpp:927-943) ended the
                                   last PUFFINN visit of each query of the last batch; {0, 0} = the visit ran out of depths */
    CLANN_X_VISIT_LOG = 22      /* u32[nq*cap*4]   the cluster granularity of the reference's RunMetrics (metrics/mod.rs:84-112,
                                   result_schema.sql:73-90) for the last clann_search / clann_search_device batch, recorded when
                                   clann_set_option(index, "visit_log", cap) was set before the search (cap rows per query, 0 =
                                   off): row v of query q = {cluster + 1 (0 = no such visit), n_candidates (heap adds that
                                   returned true, index.rs:367-372,405-410), cluster_distance_computations (index.rs:348,378,
                                   421), nanoseconds the visit took on the device} */
} clann_export_what;
int clann_export(clann_index* index, int what, uint64_t arg, void* dst, uint64_t cap, uint64_t* size);

/* Kernel-level timing of the last clann_search / clann_search_device call: ms[0] = query prep + hashing + ordering,
 * ms[1] = probe kernel (the roofline kernel), launches = kernels launched. Measured with CUDA events on the call's stream. */
int clann_last_search_profile(clann_index* index, float* ms, uint32_t* launches);

/* Schedule selection. The search picks between two probe schedules that return the same bits: dense first-visit
 * similarities + first-visit anchors ahead of a 16-warp probe, or the 24-warp gather probe alone; likewise between the tensor-pipe
 * centre screen and the all-exact centre kernel. The choice is made per call from {clusters visited, queries} of the last batch
 * that FINISHED on this index (published by the device into mapped host memory without synchronisation): more than 3 clusters
 * per query selects the gather probe, more than 8 the all-exact centre kernel. Results, counters and exports never depend on it;
 * the time of a call can depend on what ran before it. clann_tune("dense_adaptive", 0) pins the first of each pair. */

/* Process-global launch-parameter knob for A/B measurements of the kernels ("probe_occ", "probe_warps", "probe_ctas",
 * "probe_nomemo", "dense_sims", "first_ranges", "first_stream", "tc_sketch", "tc_center", "pipeline_depth", ...; the full list
 * is in INTEGRATION.md). Never changes a result. The
 * environment variable CLANN_TUNE_<KEY> seeds a key that was not set. Not part of the reference's interface. */
int clann_tune(const char* key, int64_t value);

const char* clann_last_error(void);
void clann_destroy(clann_index* index);

/* ------------------------------------------------------------------------------------------------------------------
 * (1) legacy per-cluster ABI (c_binder.h:14-26). Same signatures, same ownership, two hardening deviations:
 *     results are always max(k,1) words, 0xFFFFFFFF-padded (the reference lets Rust over-read, puffinn.rs:108-114),
 *     and no C++ exception crosses the boundary (the reference throws through extern "C", c_binder.cpp:8,15).
 *     Persistence: save_index appends the bytes of puffinn::Index::serialize (collection.hpp:185-203, byte-compatible
 *     with the reference) as record "index_{id}" of a flat file, load_from_file reads such a record back (NULL on failure);
 *     the reference keeps the same bytes in an HDF5 dataset of the same name (no HDF5 in this build).
 * ------------------------------------------------------------------------------------------------------------------ */
struct CPUFFINN;
typedef struct CPUFFINN CPUFFINN;

CPUFFINN* CPUFFINN_load_from_file(const char* file_name, const char* dataset_name);          /* c_binder.cpp:4-36   */
CPUFFINN* CPUFFINN_index_create(const char* dataset_type, int dataset_args);                 /* c_binder.cpp:39-50  */
uint64_t CPUFFINN_index_rebuild(CPUFFINN* index, unsigned int num_maps);                     /* c_binder.cpp:53-60  */
void CPUFFINN_index_insert_cosine(CPUFFINN* index, float* point, int dimension);             /* c_binder.cpp:63-66  */
uint32_t* CPUFFINN_search_cosine(CPUFFINN* index, float* query, unsigned int k, float recall, float max_sim,
                                 int dimension);                                             /* c_binder.cpp:69-96  */
unsigned int CPUFFINN_get_distance_computations(void);                                       /* c_binder.cpp:98-100 */
void CPUFFINN_clear_distance_computations(void);                                             /* c_binder.cpp:102-104*/
void CPUFFINN_save_index(CPUFFINN* index, const char* file_name, int index_number);          /* c_binder.cpp:106-146*/

/* puffinn::Index::search with its FilterType argument (collection.hpp:22-34,324-334,568-594), which c_binder.cpp never passes and
 * CLANN therefore never reaches: filter_type 0 = Default (search_maps; what CPUFFINN_search_cosine runs), 1 = None
 * (search_maps_no_filter, :671-714), 2 = Simple (search_maps_simple_filter, :717-765); 1 and 2 ignore max_sim and count no
 * distance computations, like the reference. `query` = the index's dimension floats (host). out_ids[k] (caller-owned, host,
 * 0xFFFFFFFF-padded) best first, *out_count their number, *out_stop_depth (optional) = the prefix length at which the stop
 * rule fired (performance.hpp hash_length), 0 = never. Returns a clann_status. */
int clann_puffinn_search(CPUFFINN* index, const float* query, uint32_t k, float recall, float max_sim, int filter_type,
                         uint32_t* out_ids, uint32_t* out_count, uint32_t* out_stop_depth);

#ifdef __cplusplus
}
#endif
#endif /* CLANN_B200_H */
