// Search-side kernels: query preparation, centre ordering, and the probe kernel (one warp per query) that walks the
// clusters in centre-distance order and, inside each cluster, reproduces puffinn::Index::search_maps exactly:
// anchors, per-depth ranges, the 32-slot sketch-filter ring (ring slot == warp lane), the 128-entry passing buffer,
// the Q15 rerank gather, the 2k-slot MaxBuffer, the filter threshold update and the delta stop rule.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

static std::map<std::string, int64_t>& tune_table() {
    static std::map<std::string, int64_t> t;
    return t;
}
static std::mutex g_tune_mutex;

int64_t tune_get(const char* key, int64_t dflt) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    auto& t = tune_table();
    auto it = t.find(key);
    if (it != t.end()) return it->second;
    std::string env = std::string("CLANN_TUNE_") + key;
    for (auto& ch : env) ch = (char)toupper((unsigned char)ch);
    const char* e = getenv(env.c_str());
    const int64_t v = e ? atoll(e) : dflt;
    t[key] = v;
    return v;
}

void tune_set(const char* key, int64_t value) {
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    tune_table()[key] = value;
}

// Tile lists of the row-parallel hashing kernels for a query batch of nq rows and F function sets, and the one-segment descriptor of
// the work-order sort — written on the device so that a new batch size needs no host round trip.
__global__ void __launch_bounds__(256) k_query_tiles(uint64_t nq, uint32_t F, RowTile* __restrict__ sk_tiles, RowTile* __restrict__ code_tiles,
                                                     SketchTcTile* __restrict__ tc_tiles, SortSegment* __restrict__ seg) {
    const uint32_t per_f = (uint32_t)((nq + 31) / 32), per_f_tc = (uint32_t)((nq + 127) / 128);
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && seg) *seg = SortSegment{0, 0, (uint32_t)nq, 0};
    if (i < per_f * F) {
        const uint32_t f = i / per_f;
        const uint64_t q0 = (uint64_t)(i % per_f) * 32;
        const uint32_t cnt = (uint32_t)(nq - q0 < 32 ? nq - q0 : 32);
        sk_tiles[i] = RowTile{(uint32_t)q0, (uint32_t)(f * nq + q0), cnt, f, 0, 0, 0};     // sketches: out row = fset * nq + q
        code_tiles[i] = RowTile{(uint32_t)q0, (uint32_t)q0, cnt, f, 0, 0, 0};              // codes: [fset][table][q]
    }
    if (tc_tiles && i < per_f_tc * F) {
        const uint32_t f = i / per_f_tc;
        const uint64_t q0 = (uint64_t)(i % per_f_tc) * 128;
        tc_tiles[i] = SketchTcTile{(uint32_t)q0, (uint32_t)(f * nq + q0), (uint32_t)(nq - q0 < 128 ? nq - q0 : 128), f};
    }
}

void launch_query_tiles(uint64_t nq, uint32_t F, RowTile* sk_tiles, RowTile* code_tiles, SketchTcTile* tc_tiles, SortSegment* seg,
                        cudaStream_t s) {
    const uint32_t n = (uint32_t)((nq + 31) / 32) * F;
    k_query_tiles<<<(n + 255) / 256 + 1, 256, 0, s>>>(nq, F, sk_tiles, code_tiles, tc_tiles, seg);
}

uint64_t query_state_bytes(uint32_t k) { return sizeof(QueryStateHeader) + (uint64_t)k * 8; }

// ------------------------------------------------------------------------------------------------ query prep

// Q15 form of the query (collection.hpp:331-333 -> unit_vector.hpp:61-89) and its fp32 norm as distance_point derives it
// (src/metricdata/angulardata.rs:31: sequential sum of squares, no FMA).
__global__ void __launch_bounds__(128) k_prep_queries(const float* __restrict__ queries, uint64_t nq, uint32_t d, uint32_t sl,
                                                      int16_t* __restrict__ q15, float* __restrict__ qnorm) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const float* v = queries + q * d;
    float acc = 0.0f, accn = 0.0f;
    const uint32_t body = d & ~3u;
    for (uint32_t i = 0; i < d; i++) {
        float x = v[i];
        accn = __fadd_rn(accn, __fmul_rn(x, x));
        acc = (i < body) ? __fadd_rn(acc, __fmul_rn(x, x)) : __fmaf_rn(x, x, acc);
    }
    qnorm[q] = __fsqrt_rn(accn);
    const float len = __fsqrt_rn(acc);
    int16_t* out = q15 + q * sl;
    for (uint32_t i = 0; i < d; i++) {
        float x = v[i];
        if (len != 0.0f) x = __fdiv_rn(x, len);
        out[i] = to_q15(x);
    }
    for (uint32_t i = d; i < sl; i++) out[i] = 0;
}

// src/core/index.rs:592-600 — distance from every query to every centre (the "GEMM" of the CLANN layer, computed in the
// exact fp32 order of angulardata.rs:29-35). One CTA = 32 queries x all centres; centre rows stream through shared
// memory in chunks of 64; lane = centre (padded rows: conflict-free), query row = warp-uniform broadcast.
// The stable ascending order of index.rs:609-613 is NOT materialised: the probe kernel picks "the smallest
// (distance, cluster) key above the last one" on demand, which is the same order.
__global__ void __launch_bounds__(256) k_center_dist(const float* __restrict__ queries, const float* __restrict__ qnorm, uint64_t nq,
                                                     const float* __restrict__ center_rows, const float* __restrict__ center_norms,
                                                     uint32_t K, uint32_t d, float* __restrict__ cdist) {
    extern __shared__ float s_f[];
    const uint32_t stride = d | 1;        // odd row stride
    float* s_q = s_f;                      // [32][stride]
    float* s_c = s_f + 32 * stride;        // [64][stride]
    const uint64_t q0 = (uint64_t)blockIdx.x * 32;
    const uint32_t nqt = (uint32_t)((nq - q0) < 32 ? (nq - q0) : 32);
    for (uint32_t e = threadIdx.x; e < 32 * d; e += blockDim.x) {
        uint32_t j = e / d, i = e % d;
        s_q[j * stride + i] = j < nqt ? queries[(q0 + j) * d + i] : 0.0f;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (uint32_t cb = 0; cb < K; cb += 64) {
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < 64 * d; e += blockDim.x) {
            uint32_t c = e / d, i = e % d;
            s_c[c * stride + i] = (cb + c) < K ? center_rows[(uint64_t)(cb + c) * d + i] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (uint32_t h = 0; h < 2; h++) {
            const uint32_t c = cb + h * 32 + lane;
            if (c >= K) continue;
            const float cn = center_norms[c];
            for (uint32_t j = warp; j < nqt; j += 8)
                cdist[(q0 + j) * K + c] = distance_point(s_c + (h * 32 + lane) * stride, cn, s_q + j * stride, qnorm[q0 + j], d);
        }
    }
}

// Register-tiled version of the same computation (d <= 256): a warp owns four queries, a lane two centres of the 64-centre
// chunk, so eight dot products advance together and every shared-memory word read feeds several of them (the plain tile
// above spends two LDS per multiply-add and is bound by shared-memory bandwidth). Centre rows are stored transposed with a
// 65-word pitch (conflict-free for the transposing store and for the lane-contiguous read); query rows are read as
// float4 broadcasts. The arithmetic and its order are those of ndarray's unrolled_dot (common.cuh), per (query, centre).
__global__ void __launch_bounds__(256) k_center_dist_tiled(const float* __restrict__ queries, const float* __restrict__ qnorm, uint64_t nq,
                                                           const float* __restrict__ center_rows, const float* __restrict__ center_norms,
                                                           uint32_t K, uint32_t d, float* __restrict__ cdist) {
    extern __shared__ __align__(16) float s_f[];
    const uint32_t dq = (d + 3) & ~3u;     // query pitch (16-byte aligned rows)
    float* s_q = s_f;                       // [32][dq]
    float* s_ct = s_f + 32 * dq;            // [d][65]
    const uint64_t q0 = (uint64_t)blockIdx.x * 32;
    const uint32_t nqt = (uint32_t)((nq - q0) < 32 ? (nq - q0) : 32);
    for (uint32_t e = threadIdx.x; e < 32 * dq; e += blockDim.x) {
        uint32_t j = e / dq, i = e % dq;
        s_q[e] = (j < nqt && i < d) ? queries[(q0 + j) * d + i] : 0.0f;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t full = d & ~7u;
    // blockIdx.y strides over the 64-centre chunks: small CTAs keep the last wave of the grid short
    for (uint32_t cb = blockIdx.y * 64; cb < K; cb += gridDim.y * 64) {
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < 64 * d; e += blockDim.x) {
            uint32_t c = e / d, i = e % d;
            s_ct[i * 65 + c] = (cb + c) < K ? center_rows[(uint64_t)(cb + c) * d + i] : 0.0f;
        }
        __syncthreads();
        float p[4][2][8];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int u = 0; u < 8; u++) p[j][h][u] = 0.0f;
        for (uint32_t i0 = 0; i0 < full; i0 += 8) {
            float cv[2][8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                cv[0][u] = s_ct[(i0 + u) * 65 + lane];
                cv[1][u] = s_ct[(i0 + u) * 65 + lane + 32];
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float4 qa = *reinterpret_cast<const float4*>(s_q + (warp * 4 + j) * dq + i0);
                const float4 qb = *reinterpret_cast<const float4*>(s_q + (warp * 4 + j) * dq + i0 + 4);
                const float qv[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int h = 0; h < 2; h++)
#pragma unroll
                    for (int u = 0; u < 8; u++) p[j][h][u] = __fadd_rn(p[j][h][u], __fmul_rn(cv[h][u], qv[u]));
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t jq = warp * 4 + j;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t c = cb + h * 32 + lane;
                float sum = 0.0f;
                sum = __fadd_rn(sum, __fadd_rn(p[j][h][0], p[j][h][4]));
                sum = __fadd_rn(sum, __fadd_rn(p[j][h][1], p[j][h][5]));
                sum = __fadd_rn(sum, __fadd_rn(p[j][h][2], p[j][h][6]));
                sum = __fadd_rn(sum, __fadd_rn(p[j][h][3], p[j][h][7]));
                for (uint32_t i = full; i < d; i++) sum = __fadd_rn(sum, __fmul_rn(s_ct[i * 65 + h * 32 + lane], s_q[jq * dq + i]));
                if (jq < nqt && c < K) {
                    const float cs = __fdiv_rn(sum, __fmul_rn(center_norms[c], qnorm[q0 + jq]));  // angulardata.rs:29-35
                    cdist[(q0 + jq) * K + c] = __fsub_rn(1.0f, cs);
                }
            }
        }
    }
}

// Fallback for very wide rows (the tiles above would not fit in shared memory): one thread per (query, centre).
__global__ void __launch_bounds__(256) k_center_dist_simple(const float* __restrict__ queries, const float* __restrict__ qnorm, uint64_t nq,
                                                            const float* __restrict__ center_rows, const float* __restrict__ center_norms,
                                                            uint32_t K, uint32_t d, float* __restrict__ cdist) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * K) return;
    uint64_t q = i / K;
    uint32_t c = (uint32_t)(i % K);
    cdist[i] = distance_point(center_rows + (uint64_t)c * d, center_norms[c], queries + q * d, qnorm[q], d);
}

// Work-order key of every query: its nearest centre (first cluster of its visiting order) in the low 20 bits, so that
// queries that start in the same cluster run next to each other and share the cluster's rows, sketches and tables through
// the L2, and above it 15 - min(15, e) where e predicts how many clusters the query will visit: the clusters whose ball
// reaches closer than the nearest centre (distance - radius <= nearest distance; the prune test of index.rs:342-361 with the
// nearest-centre distance standing in for the k-th neighbour distance, which is not known yet). Sorting by the key starts
// the predicted-longest queries first. OFF by default (knob order_longest_first): measured on B200 (glove-100 shape,
// profiles/exp/visit_stats.py) the predictor is useless — queries visit at most 6 clusters, the cost of a query is its
// candidate count (p99 = 1.9x mean) and correlates 0.02 with e — while the coarser cluster grouping costs L2 hits (DRAM reads
// 10.3 -> 14.0 GB for 3.08 -> 3.03 ms). The 29 % of warp slots the probe kernel leaves idle are its last wave: the queue is
// empty after ~1.7 ms and the queries still in flight take up to one query latency more. The order never changes a result.
__global__ void __launch_bounds__(256) k_first_cluster(const float* __restrict__ cdist, const float* __restrict__ radii, uint64_t nq, uint32_t K,
                                                       int longest_first, uint32_t* __restrict__ first) {
    const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    unsigned long long best = ~0ull;
    for (uint32_t c = lane_id(); c < K; c += 32) {
        unsigned long long key = ((unsigned long long)float_order_bits(cdist[q * K + c]) << 32) | c;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t < best ? t : best;
    }
    uint32_t key = (uint32_t)best;
    if (longest_first && K <= (1u << 20)) {
        const float nearest = float_from_order_bits((uint32_t)(best >> 32));
        uint32_t reach = 0;
        for (uint32_t c = lane_id(); c < K; c += 32) reach += (cdist[q * K + c] - radii[c] <= nearest) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) reach += __shfl_xor_sync(0xffffffffu, reach, o);
        key |= (15u - (reach < 15u ? reach : 15u)) << 20;
    }
    if (lane_id() == 0) first[q] = key;
}

// k-th smallest (k >= 1) of n u32 keys read through `key(i)`, by a four-pass byte radix select over a warp-private histogram.
template <typename KeyFn>
__device__ __forceinline__ uint32_t warp_kth_smallest_u32(KeyFn key, uint32_t n, uint32_t k, uint32_t* hist) {
    const uint32_t lane = lane_id();
    uint32_t prefix = 0, want = k;
#pragma unroll 1
    for (int pass = 0; pass < 4; pass++) {
        const uint32_t shift = 24 - 8 * pass;
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < n; i += 32) {
            const uint32_t v = key(i);
            if (pass == 0 || (v >> (shift + 8)) == prefix) atomicAdd(&hist[(v >> shift) & 0xffu], 1u);
        }
        __syncwarp();
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) mine += hist[8 * lane + j];
        uint32_t pre = mine;  // entries in the bins of this lane and of every lower lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
            if ((int)lane >= o) pre += t;
        }
        const uint32_t below = pre - mine;
        const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, below < want && want <= pre)) - 1;
        uint32_t bin = 0, rem = 0;
        if (lane == owner) {
            uint32_t acc = below;
            for (int j = 0; j < 8; j++) {
                const uint32_t h = hist[8 * lane + j];
                if (acc + h >= want) {
                    bin = 8 * lane + j;
                    rem = want - acc;
                    break;
                }
                acc += h;
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        rem = __shfl_sync(0xffffffffu, rem, owner);
        __syncwarp();
        prefix = (prefix << 8) | bin;
        want = rem;
    }
    return prefix;
}

// After the tensor-pipe screen: one warp per query. The kCentreExact centres with the smallest approximate distance are
// evaluated exactly (distance_point: the reference's fp32 arithmetic, index.rs:592-600 over angulardata.rs:29-35) and stored;
// every other entry becomes approx - kCentreEps, a lower bound of its exact value; exact_limit[q] = (the smallest
// approximate distance that was not selected) - kCentreEps, so that every stored value below the limit is exact and every
// value at or above it belongs to a centre whose exact distance is at or above it too. first[q] = the nearest centre (exact;
// if the best exact value does not beat the limit the whole row is evaluated exactly here).
__global__ void __launch_bounds__(256) k_center_refine(const float* __restrict__ queries, const float* __restrict__ qnorm, uint64_t nq,
                                                       const float* __restrict__ center_rows, const float* __restrict__ center_norms,
                                                       uint32_t K, uint32_t d, float* __restrict__ cdist, float* __restrict__ exact_limit,
                                                       uint32_t* __restrict__ first) {
    __shared__ uint32_t s_hist[8][256];
    __shared__ uint32_t s_sel[8][kCentreExact];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t q = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nq) return;
    float* cd = cdist + q * K;
    const float* qv = queries + q * d;
    const float qn = qnorm[q];
    float limit = INFINITY;
    if (K > kCentreExact) {
        // tau = the (kCentreExact + 1)-th smallest approximate distance; selected = entries strictly below it (at most kCentreExact)
        const uint32_t tau_bits = warp_kth_smallest_u32([&](uint32_t i) { return float_order_bits(cd[i]); }, K, kCentreExact + 1, s_hist[warp]);
        const float tau = float_from_order_bits(tau_bits);
        uint32_t nsel = 0;
        for (uint32_t c0 = 0; c0 < K; c0 += 32) {
            const uint32_t c = c0 + lane;
            const float a = c < K ? cd[c] : INFINITY;
            const bool sel = c < K && float_order_bits(a) < tau_bits;
            const uint32_t bal = __ballot_sync(0xffffffffu, sel);
            if (sel) s_sel[warp][nsel + __popc(bal & ((1u << lane) - 1u))] = c;
            if (c < K && !sel) cd[c] = a - kCentreEps;
            nsel += __popc(bal);
        }
        __syncwarp();
        unsigned long long best = ~0ull;
        if (lane < nsel) {
            const uint32_t c = s_sel[warp][lane];
            const float e = distance_point(center_rows + (uint64_t)c * d, center_norms[c], qv, qn, d);
            cd[c] = e;
            best = ((unsigned long long)float_order_bits(e) << 32) | c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
            best = t < best ? t : best;
        }
        limit = tau - kCentreEps;
        if (nsel > 0 && float_from_order_bits((uint32_t)(best >> 32)) < limit) {
            if (lane == 0) {
                first[q] = (uint32_t)best;
                exact_limit[q] = limit;
            }
            return;
        }
    }
    // small K, NaN rows (zero query) or a nearest centre that does not clear the limit: the whole row, exactly
    unsigned long long best = ~0ull;
    for (uint32_t c = lane; c < K; c += 32) {
        const float e = distance_point(center_rows + (uint64_t)c * d, center_norms[c], qv, qn, d);
        cd[c] = e;
        const unsigned long long key = ((unsigned long long)float_order_bits(e) << 32) | c;
        best = key < best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t < best ? t : best;
    }
    if (lane == 0) {
        first[q] = (uint32_t)best;
        exact_limit[q] = INFINITY;
    }
}

__global__ void k_init_state(uint8_t* __restrict__ state, uint64_t nq, uint64_t state_bytes, uint32_t* work_counter) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *work_counter = 0;
    if (q >= nq) return;
    QueryStateHeader* h = reinterpret_cast<QueryStateHeader*>(state + q * state_bytes);
    h->next_pos = 0;
    h->done = 0;
    h->heap_len = 0;
    h->visited = 0;
    h->candidates = 0;
    h->distcomp = 0;
    h->last_key = 0;
    h->stop_point = 0;
}

// Per-warp scratch carved out of dynamic shared memory.
struct WarpSmem {
    uint32_t* pass_idx;        // [kPassingCap] passing-filter ids (collection.hpp:783)
    uint16_t* pass_sim;        // [kPassingCap] their Q15 similarity as dot + 32768
    uint16_t* unk;             // [kPassingCap] positions of the passing list whose similarity is not memoised yet
    uint32_t* anchor;          // [L] lower bound of the query code in each table (prefixmap.hpp:36-57), unpadded position
    uint2* lcp_up;             // [L] 8 bytes: common-prefix length with the code at anchor + 12 j, j = 0..7
    uint2* lcp_dn;             // [L] 8 bytes: same at anchor - 1 - 12 j
    uint32_t* code;            // [L] query code per table
    uint32_t* start;           // [L] range start of the current depth
    uint32_t* segbase;         // [L+1] exclusive prefix of 4-entry segment counts of the current depth
    uint16_t* wstart;          // [kWinCap] table that holds segment 32 w of the current depth's stream (one entry per ring sweep)
    unsigned long long* mb;    // [P2K] MaxBuffer slots (maxbuffer.hpp:20): (sim16 << 32) | local id
    unsigned long long* heap;  // [k] TopKClosestHeap (src/core/heap.rs): (order_bits(dist) << 32) | point id
    unsigned long long* loc;   // [k] local heap of a brute-force cluster (index.rs:671)
};

constexpr uint32_t kWinCap = 64;  // ring sweeps per depth with a table hint (8 192 candidates); later sweeps search all tables

__host__ __device__ inline uint32_t warp_smem_bytes(uint32_t L, uint32_t k) {
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint32_t b = kPassingCap * 4 + kPassingCap * 2 * 2;  // pass_idx, pass_sim, unk
    b += L * 4;                                       // anchor
    b += L * 8 * 2;                                   // lcp_up, lcp_dn
    b += L * 4 * 2;                                   // code, start
    b += (L + 1) * 4;                                 // segbase
    b += kWinCap * 2;                                 // wstart
    b = (b + 15) & ~15u;
    b += p2k * 8 + k * 8 * 2;
    return (b + 15) & ~15u;
}

__device__ __forceinline__ WarpSmem carve(uint8_t* base, uint32_t L, uint32_t k) {
    WarpSmem w;
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint8_t* p = base;
    w.lcp_up = reinterpret_cast<uint2*>(p); p += L * 8;
    w.lcp_dn = reinterpret_cast<uint2*>(p); p += L * 8;
    w.pass_idx = reinterpret_cast<uint32_t*>(p); p += kPassingCap * 4;
    w.anchor = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.code = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.start = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.segbase = reinterpret_cast<uint32_t*>(p); p += (L + 1) * 4;
    w.pass_sim = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    w.unk = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    w.wstart = reinterpret_cast<uint16_t*>(p); p += kWinCap * 2;
    p = base + (((uint32_t)(p - base) + 15) & ~15u);
    w.mb = reinterpret_cast<unsigned long long*>(p); p += p2k * 8;
    w.heap = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    w.loc = reinterpret_cast<unsigned long long*>(p);
    return w;
}

// ------------------------------------------------------------------------------------------------ probe of one cluster

// Q15 similarities (as dot + 32768) of the first `count` ids in sm.pass_idx -> sm.pass_sim.
// memo (one u16 per local id, 0 = unknown; may be null) returns similarities already computed during this visit: the
// reference rescans nested ranges at every depth, so more than half of its distance computations are repeats. The rows
// that are missing are gathered with 128-bit loads, G lanes per row (the HBM-bound rerank gather, cosine.hpp:19-23 /
// math.hpp:11-44).
template <int G>
__device__ __forceinline__ void rerank(const WarpSmem& sm, uint32_t count, const int16_t* __restrict__ rows, uint32_t sl,
                                       const int16_t* __restrict__ qrow_smem, const int (&qreg)[8], bool qreg_valid, uint16_t* memo,
                                       bool prefetch_rows = false) {
    constexpr int CPI = 32 / G;  // candidates per warp iteration
    const uint32_t lane = lane_id();
    const uint32_t sub = lane % G;
    const uint32_t grp = lane / G;
    const uint32_t cpr = sl / 8;  // 16-byte chunks per row
    uint32_t nunk = 0;
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < count;
        const uint32_t m = (valid && memo) ? (uint32_t)memo[sm.pass_idx[i]] : 0u;
        const bool need = valid && m == 0;
        if (valid && m) sm.pass_sim[i] = (uint16_t)m;
        const uint32_t bal = __ballot_sync(0xffffffffu, need);
        if (need) sm.unk[nunk + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)i;
        nunk += __popc(bal);
    }
    __syncwarp();
    if (prefetch_rows) {
        // every row of the batch is requested into the L2 at once (TMA bulk prefetch, no destination register), so the
        // gather below pays one DRAM latency for the batch instead of one per round of eight rows
        for (uint32_t i = lane; i < nunk; i += 32) {
            const int16_t* src = rows + (uint64_t)sm.pass_idx[sm.unk[i]] * sl;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(sl * 2) : "memory");
        }
    }
    for (uint32_t base = 0; base < nunk; base += CPI * 4) {
        int part[4];
        uint4 w[4];
        uint32_t pos[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t cand = base + u * CPI + grp;
            pos[u] = cand < nunk ? sm.unk[cand] : 0xffffffffu;
            w[u] = make_uint4(0, 0, 0, 0);
            if (pos[u] != 0xffffffffu && sub < cpr) {
                const uint4* src = reinterpret_cast<const uint4*>(rows + (uint64_t)sm.pass_idx[pos[u]] * sl) + sub;
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(w[u].x), "=r"(w[u].y), "=r"(w[u].z), "=r"(w[u].w)
                             : "l"(src));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int s = 0;
            if (qreg_valid) {
                s += q15_mul(unpack_lo(w[u].x), qreg[0]); s += q15_mul(unpack_hi(w[u].x), qreg[1]);
                s += q15_mul(unpack_lo(w[u].y), qreg[2]); s += q15_mul(unpack_hi(w[u].y), qreg[3]);
                s += q15_mul(unpack_lo(w[u].z), qreg[4]); s += q15_mul(unpack_hi(w[u].z), qreg[5]);
                s += q15_mul(unpack_lo(w[u].w), qreg[6]); s += q15_mul(unpack_hi(w[u].w), qreg[7]);
            }
            part[u] = s;
        }
        if (!qreg_valid) {
            // rows wider than 32 chunks (d > 256): loop over the remaining chunks with the query read from shared memory
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int s = 0;
                if (pos[u] != 0xffffffffu) {
                    const uint4* src = reinterpret_cast<const uint4*>(rows + (uint64_t)sm.pass_idx[pos[u]] * sl);
                    for (uint32_t ch = sub; ch < cpr; ch += G) {
                        uint4 a = __ldg(src + ch);
                        uint4 b = *reinterpret_cast<const uint4*>(qrow_smem + ch * 8);
                        s += q15_mul(unpack_lo(a.x), unpack_lo(b.x)); s += q15_mul(unpack_hi(a.x), unpack_hi(b.x));
                        s += q15_mul(unpack_lo(a.y), unpack_lo(b.y)); s += q15_mul(unpack_hi(a.y), unpack_hi(b.y));
                        s += q15_mul(unpack_lo(a.z), unpack_lo(b.z)); s += q15_mul(unpack_hi(a.z), unpack_hi(b.z));
                        s += q15_mul(unpack_lo(a.w), unpack_lo(b.w)); s += q15_mul(unpack_hi(a.w), unpack_hi(b.w));
                    }
                }
                part[u] = s;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int s = part[u];
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (sub == 0 && pos[u] != 0xffffffffu) {
                const uint16_t sim16 = (uint16_t)(s + 32768);
                sm.pass_sim[pos[u]] = sim16;
                if (memo) memo[sm.pass_idx[pos[u]]] = sim16;
            }
        }
    }
    __syncwarp();
}

// One PUFFINN query against cluster c (collection.hpp:543-601 -> search_maps :768-948). On return sm.mb[0..cnt) holds the
// best entries, best first (maxbuffer.hpp:79-96). `codes` points at this query's code of table 0 (stride code_stride).
// AHEAD (the 128-register instantiations): anchors for three tables per lane in lockstep, and the table indices of the next
// ring sweep are requested before the sketch words of the current one, so a sweep costs one memory round trip instead of two.
// This query's slice of the first-visit candidate stream (QueryBatch::fs_*); meta == nullptr = no stream.
struct FsView {
    const uint2* idx;      // 4 x u16 local ids per segment
    const uint32_t* hd;    // 4 x u8 Hamming distances per segment
    const uint8_t* tab;    // table of the first segment of every ring sweep
    const uint32_t* meta;  // kFsMeta words
};

template <int G, bool AHEAD = false, bool STREAM = false>
__device__ uint32_t probe_cluster(const SearchParams& p, const WarpSmem& sm, uint32_t c, const uint32_t* __restrict__ codes,
                                  uint64_t code_stride, uint64_t my_sketch, const uint32_t* __restrict__ stop, float max_sim,
                                  const int16_t* qrow_smem, const int (&qreg)[8], bool qreg_valid, uint16_t* memo, ProbeCounters& ctr,
                                  bool memo_prefilled = false, unsigned long long* memo_bar = nullptr, uint32_t memo_phase = 0,
                                  const uint32_t* __restrict__ pre_anchor = nullptr, const uint32_t* __restrict__ pre_range = nullptr,
                                  const uint4* __restrict__ pre_lcp = nullptr, FsView fs = FsView{nullptr, nullptr, nullptr, nullptr}) {
    const uint32_t L = p.g.L, k = p.k;
    const uint32_t lane = lane_id();
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    const int16_t* rows = p.q15 + off * p.g.sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;

    uint32_t inserted = 0, minval16 = 0, max_diff = kSketchBits;  // maxbuffer.hpp:53-55, filterer.hpp:101
    if (memo && !memo_prefilled) {
        uint4* mz = reinterpret_cast<uint4*>(memo);
        for (uint32_t i = lane; i < (nc + 7) / 8; i += 32) mz[i] = make_uint4(0, 0, 0, 0);
    }

    // --- first-visit candidate stream: the depths down to fs_lo are replayed from it; anchors and ranges are only computed if
    // the visit goes deeper than the stream (or there is none)
    const uint32_t fs_lo = (STREAM && fs.meta) ? __ldg(fs.meta + 48) : (uint32_t)kMaxHashBits + 1;
    bool anchors_ready = fs_lo > (uint32_t)kMaxHashBits;
    // --- SearchBuffers ctor (collection.hpp:642-645): anchor per table + 8 stride-12 samples each way
    if (!anchors_ready) {
    } else if (AHEAD && pre_range) {  // anchors and every depth's range computed in advance by k_first_ranges
        for (uint32_t t = lane; t < L; t += 32) sm.anchor[t] = __ldg(pre_anchor + t);
    } else if (AHEAD && pre_lcp) {  // anchors and samples computed in advance; ranges evaluated here as far as the visit gets
        for (uint32_t t = lane; t < L; t += 32) {
            const uint4 l = __ldg(pre_lcp + t);
            sm.code[t] = codes[(uint64_t)t * code_stride];
            sm.anchor[t] = __ldg(pre_anchor + t);
            sm.lcp_up[t] = make_uint2(l.x, l.y);
            sm.lcp_dn[t] = make_uint2(l.z, l.w);
        }
    } else if constexpr (AHEAD) anchors_lockstep(p, sm, c, off, nc, codes, code_stride);
    else for (uint32_t t = lane; t < L; t += 32) {
        const uint32_t h = codes[(uint64_t)t * code_stride];
        uint32_t A;
        uint2 up, dn;
        table_anchor(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc, h, A, up, dn);
        sm.code[t] = h;
        sm.anchor[t] = A;
        sm.lcp_up[t] = up;
        sm.lcp_dn[t] = dn;
    }
    __syncwarp();

    if (memo_bar) pc_mbar_wait(memo_bar, memo_phase);  // the pre-filled memo was bulk-copied to shared memory behind the anchors

    bool stopped = false;
    ctr.stop_point = 0;
    for (uint32_t depth = kMaxHashBits; depth > 0 && !stopped; depth--) {
        const bool streamed = STREAM && depth >= fs_lo;
        uint32_t soff = 0;  // first segment of this depth's block in the stream
        // --- fill_ranges (collection.hpp:650-667) with get_next_range (prefixmap.hpp:267-304) in closed form
        uint32_t running = 0;
        if (streamed) {
            running = __ldg(fs.meta + depth - 1);
            soff = __ldg(fs.meta + kMaxHashBits + depth - 1);
        } else if (!anchors_ready) {  // the visit outlived the stream: from here on it is probed the ordinary way
            if constexpr (AHEAD) anchors_lockstep(p, sm, c, off, nc, codes, code_stride);
            anchors_ready = true;
        }
        for (uint32_t t0 = 0; t0 < L && !streamed; t0 += 32) {
            const uint32_t t = t0 + lane;
            uint32_t nseg = 0;
            if (AHEAD && pre_range) {
                if (t < L) {
                    const uint32_t w = __ldg(pre_range + (size_t)(depth - 1) * L + t);
                    nseg = w & 0x7fffffffu;
                    sm.start[t] = (w >> 31) ? sm.anchor[t] : sm.anchor[t] - 4u * nseg;  // downward ranges end at the anchor
                }
            } else if (t < L)
                sm.start[t] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc,
                                          sm.code[t], sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg);
            uint32_t total;
            uint32_t ex = warp_excl_scan(nseg, total);
            if (t < L) {
                const uint32_t sb = running + ex;
                sm.segbase[t] = sb;
                // every ring sweep starts at a multiple of 32 segments: note which table holds that segment
                uint32_t w = (sb + 31u) >> 5, wend = (sb + nseg + 31u) >> 5;
                wend = wend < kWinCap ? wend : kWinCap;
                for (; w < wend; w++) sm.wstart[w] = (uint16_t)t;
            }
            running += total;
        }
        if (lane == 0 && !streamed) sm.segbase[L] = running;
        __syncwarp();
        const uint32_t S = running;
        if (S <= kRing) continue;  // the initial ring fill swallows the whole stream (collection.hpp:802-810)

        // segment number -> (table, position): the table is the last one whose prefix is <= s (upper_bound(segbase, s) - 1); the
        // sweep's window [32 w, 32 w + 32) lies between the tables noted for w and w + 1, so the search is over that span only
        auto locate = [&](uint32_t s, uint32_t& t_out) -> uint64_t {
            const uint32_t w = s >> 5;
            uint32_t lo = 0, hi = L - 1;
            if (w < kWinCap) {
                lo = sm.wstart[w];
                if (w + 1 < kWinCap && ((w + 1) << 5) < S) hi = sm.wstart[w + 1];
            }
            uint32_t a = lo + 1, len = hi - lo;  // upper_bound over segbase[lo + 1 .. hi]
            while (len > 0) {
                uint32_t half = len >> 1, mid = a + half;
                if (sm.segbase[mid] <= s) { a = mid + 1; len -= half + 1; } else { len = half; }
            }
            uint32_t t = a - 1;
            t_out = t;
            return table_base(off, nc, L, t) + sm.start[t] + 4 * (s - sm.segbase[t]);
        };

        uint32_t base = 0;  // first stream segment held by the ring
        bool have = false;  // AHEAD: n0..n3 (and nh) hold this lane's segment of the full ring [base, base + 32)
        uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0, nh = 0;
        // one lane's segment of the stream: ids as 4 x u16, Hamming distances as 4 x u8
        auto stream_load = [&](uint32_t s, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3, uint32_t& h4) {
            const uint2 iv = __ldg(fs.idx + soff + s);
            h4 = __ldg(fs.hd + soff + s);
            a0 = iv.x & 0xffffu; a1 = iv.x >> 16; a2 = iv.y & 0xffffu; a3 = iv.y >> 16;
        };
        do {
            uint32_t np = 0;
            uint32_t missing = (base + kRing > S) ? (base + kRing - S > kRing ? kRing : base + kRing - S) : 0;
            while (np < kFilterBuffer && missing == 0) {  // collection.hpp:813-866: a full ring sweep, slot == lane
                uint32_t t;
                uint32_t v0, v1, v2, v3, h4 = 0;
                if (AHEAD && have) {
                    v0 = n0; v1 = n1; v2 = n2; v3 = n3; h4 = nh;
                } else if (streamed) {
                    stream_load(base + lane, v0, v1, v2, v3, h4);
                } else {
                    const uint32_t* seg = p.tbl_idx + locate(base + lane, t);
                    v0 = __ldg(seg); v1 = __ldg(seg + 1); v2 = __ldg(seg + 2); v3 = __ldg(seg + 3);
                }
                if constexpr (AHEAD) {
                    have = base + 2 * kRing <= S;
                    if (have) {
                        if (streamed) {
                            stream_load(base + kRing + lane, n0, n1, n2, n3, nh);
                        } else {
                            const uint32_t* nseg = p.tbl_idx + locate(base + kRing + lane, t);
                            n0 = __ldg(nseg); n1 = __ldg(nseg + 1); n2 = __ldg(nseg + 2); n3 = __ldg(nseg + 3);
                        }
                    }
                }
                uint32_t p0, p1, p2, p3;
                if (streamed) {  // the sketch test was evaluated ahead of time; only the threshold is applied here
                    p0 = (h4 & 0xffu) <= max_diff; p1 = ((h4 >> 8) & 0xffu) <= max_diff;
                    p2 = ((h4 >> 16) & 0xffu) <= max_diff; p3 = (h4 >> 24) <= max_diff;
                } else {
                    uint64_t s0 = __ldg(sk + ((uint64_t)v0 << 5 | lane)), s1 = __ldg(sk + ((uint64_t)v1 << 5 | lane));
                    uint64_t s2 = __ldg(sk + ((uint64_t)v2 << 5 | lane)), s3 = __ldg(sk + ((uint64_t)v3 << 5 | lane));
                    p0 = (uint32_t)__popcll(s0 ^ my_sketch) <= max_diff; p1 = (uint32_t)__popcll(s1 ^ my_sketch) <= max_diff;
                    p2 = (uint32_t)__popcll(s2 ^ my_sketch) <= max_diff; p3 = (uint32_t)__popcll(s3 ^ my_sketch) <= max_diff;
                }
                uint32_t cnt = p0 + p1 + p2 + p3, total;
                uint32_t pos = np + warp_excl_scan(cnt, total);
                if (p0) sm.pass_idx[pos++] = v0;
                if (p1) sm.pass_idx[pos++] = v1;
                if (p2) sm.pass_idx[pos++] = v2;
                if (p3) sm.pass_idx[pos++] = v3;
                np += total;
                ctr.candidates += kRing * 4;
                base += kRing;
                missing = (base + kRing > S) ? (base + kRing - S > kRing ? kRing : base + kRing - S) : 0;
            }
            // tail (collection.hpp:869-903): the not-yet-tested ring slots, in descending slot order, tested with the point
            // index itself in place of its sketch (:890-893)
            {
                const uint32_t live = kRing - missing;  // slots 0..live-1 hold real segments base+slot
                uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0;
                if (lane < live) {
                    if (AHEAD && have) {  // a full ring whose indices are already here
                        v0 = n0; v1 = n1; v2 = n2; v3 = n3;
                    } else if (streamed) {
                        uint32_t h4;
                        stream_load(base + lane, v0, v1, v2, v3, h4);
                    } else {
                        uint32_t t;
                        const uint32_t* seg = p.tbl_idx + locate(base + lane, t);
                        v0 = __ldg(seg); v1 = __ldg(seg + 1); v2 = __ldg(seg + 2); v3 = __ldg(seg + 3);
                    }
                    p0 = (uint32_t)__popcll((uint64_t)v0 ^ my_sketch) <= max_diff;
                    p1 = (uint32_t)__popcll((uint64_t)v1 ^ my_sketch) <= max_diff;
                    p2 = (uint32_t)__popcll((uint64_t)v2 ^ my_sketch) <= max_diff;
                    p3 = (uint32_t)__popcll((uint64_t)v3 ^ my_sketch) <= max_diff;
                }
                uint32_t cnt = p0 + p1 + p2 + p3, total;
                uint32_t ex = warp_excl_scan(cnt, total);
                uint32_t pos = np + (total - ex - cnt);  // entries of higher slots come first
                if (p0) sm.pass_idx[pos++] = v0;
                if (p1) sm.pass_idx[pos++] = v1;
                if (p2) sm.pass_idx[pos++] = v2;
                if (p3) sm.pass_idx[pos++] = v3;
                np += total;
                ctr.candidates += 4 * live;
            }
            __syncwarp();
            // empty the buffer (collection.hpp:909-925)
            rerank<G>(sm, np, rows, p.g.sl, qrow_smem, qreg, qreg_valid, memo, p.prefetch_rows != 0);
            maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, np);
            ctr.distcomp += np;
            max_diff = p.msd[minval16 < 65536u ? minval16 : 65535u];  // filterer.hpp:108-111
            // stop rule (collection.hpp:927-943)
            uint32_t pulled = base + kRing;
            uint32_t table_idx = L;
            if (streamed) {
                if (pulled < S) table_idx = __ldg(fs.tab + ((soff + pulled) >> 5));
            } else if (pulled < S) {
                if ((pulled >> 5) < kWinCap) {
                    table_idx = sm.wstart[pulled >> 5];  // pulled is a multiple of the ring size: the table noted for its sweep
                } else {
                    uint32_t lo = 0, len = L;
                    while (len > 0) {
                        uint32_t half = len >> 1, mid = lo + half;
                        if (sm.segbase[mid] <= pulled) { lo = mid + 1; len -= half + 1; } else { len = half; }
                    }
                    table_idx = lo - 1;
                }
            }
            float kth = __fdiv_rn((float)minval16, 65536.0f);
            float sim = kth < max_sim ? max_sim : kth;  // std::max(kth, max_sim)
            uint32_t bin = (uint32_t)__fdiv_rn(sim, 0.005f);  // crosspolytope.hpp:116-118
            bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
            uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (table_idx >> 5));
            if ((word >> (table_idx & 31)) & 1u) {
                stopped = true;
                ctr.stop_point = depth << 16 | table_idx;
                break;
            }
        } while (base + kRing < S);
    }
    // best_indices (collection.hpp:598, maxbuffer.hpp:79-96)
    maxbuffer_filter(sm.mb, P, k, inserted, minval16);
    return inserted;
}

// Q15 brute force of a whole cluster (collection.hpp:524-541), used by the legacy ABI when n < 100 (:550-555).
template <int G>
__device__ uint32_t probe_bruteforce_q15(const SearchParams& p, const WarpSmem& sm, uint32_t c, const int16_t* qrow_smem,
                                         const int (&qreg)[8], bool qreg_valid) {
    const uint32_t k = p.k;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    const int16_t* rows = p.q15 + off * p.g.sl;
    uint32_t inserted = 0, minval16 = 0;
    for (uint32_t base = 0; base < nc; base += kFilterBuffer) {
        uint32_t cnt = nc - base < (uint32_t)kFilterBuffer ? nc - base : (uint32_t)kFilterBuffer;
        for (uint32_t i = lane_id(); i < cnt; i += 32) sm.pass_idx[i] = base + i;
        __syncwarp();
        rerank<G>(sm, cnt, rows, p.g.sl, qrow_smem, qreg, qreg_valid, nullptr);
        maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, cnt);
    }
    maxbuffer_filter(sm.mb, P, k, inserted, minval16);
    return inserted;
}

// FilterType::None (SIMPLE = false, collection.hpp:671-714) and FilterType::Simple (SIMPLE = true, :717-765) of one PUFFINN
// query: no ring, no passing buffer, no tail, no max_sim. Every entry of every non-empty range of a depth, tables in order,
// goes to MaxBuffer::insert — in the Simple variant after the sketch test of slot `range_idx % 32`, range_idx counting the
// non-empty ranges (fill_ranges drops the empty ones, :660), with the threshold refreshed after each range (:739-740). One
// stop test per depth with table_idx = last_tables = L. The warp takes a range 128 entries at a time (lane = four consecutive
// entries, compacted in order), so the inserts happen in the reference's order; neither variant counts candidates or distance
// computations (only :865,904,921 do). On return sm.mb[0..cnt) holds the best entries, best first.
template <int G, bool SIMPLE>
__device__ uint32_t probe_cluster_plain(const SearchParams& p, const WarpSmem& sm, uint32_t c, const uint32_t* __restrict__ codes,
                                        uint64_t code_stride, uint64_t my_sketch, const uint32_t* __restrict__ stop,
                                        const int16_t* qrow_smem, const int (&qreg)[8], bool qreg_valid, ProbeCounters& ctr) {
    const uint32_t L = p.g.L, k = p.k;
    const uint32_t lane = lane_id();
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    const int16_t* rows = p.q15 + off * p.g.sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    uint32_t inserted = 0, minval16 = 0, max_diff = kSketchBits;

    for (uint32_t t = lane; t < L; t += 32) {  // SearchBuffers ctor (collection.hpp:642-645)
        const uint32_t h = codes[(uint64_t)t * code_stride];
        uint32_t A;
        uint2 up, dn;
        table_anchor(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc, h, A, up, dn);
        sm.code[t] = h;
        sm.anchor[t] = A;
        sm.lcp_up[t] = up;
        sm.lcp_dn[t] = dn;
    }
    __syncwarp();

    ctr.stop_point = 0;
    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        for (uint32_t t = lane; t < L; t += 32) {  // fill_ranges; segbase[t] = this range's length in 4-entry segments
            uint32_t nseg = 0;
            sm.start[t] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc,
                                      sm.code[t], sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg);
            sm.segbase[t] = nseg;
        }
        __syncwarp();
        uint32_t range_idx = 0;
        for (uint32_t t = 0; t < L; t++) {
            const uint32_t len = 4u * sm.segbase[t];
            if (len == 0) continue;
            const uint32_t slot = range_idx & (kNumSketches - 1);
            const uint64_t qsk = __shfl_sync(0xffffffffu, my_sketch, slot);
            const uint32_t* idx = p.tbl_idx + table_base(off, nc, L, t) + sm.start[t];
            for (uint32_t e0 = 0; e0 < len; e0 += kFilterBuffer) {
                const uint32_t e = e0 + 4 * lane;
                uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0;
                if (e < len) {  // len is a multiple of 4: a lane holds four entries or none
                    v0 = __ldg(idx + e); v1 = __ldg(idx + e + 1); v2 = __ldg(idx + e + 2); v3 = __ldg(idx + e + 3);
                    if (SIMPLE) {
                        const uint64_t s0 = __ldg(sk + ((uint64_t)v0 << 5 | slot)), s1 = __ldg(sk + ((uint64_t)v1 << 5 | slot));
                        const uint64_t s2 = __ldg(sk + ((uint64_t)v2 << 5 | slot)), s3 = __ldg(sk + ((uint64_t)v3 << 5 | slot));
                        p0 = (uint32_t)__popcll(s0 ^ qsk) <= max_diff; p1 = (uint32_t)__popcll(s1 ^ qsk) <= max_diff;
                        p2 = (uint32_t)__popcll(s2 ^ qsk) <= max_diff; p3 = (uint32_t)__popcll(s3 ^ qsk) <= max_diff;
                    } else {
                        p0 = p1 = p2 = p3 = 1;
                    }
                }
                uint32_t total;
                uint32_t pos = warp_excl_scan(p0 + p1 + p2 + p3, total);
                if (p0) sm.pass_idx[pos++] = v0;
                if (p1) sm.pass_idx[pos++] = v1;
                if (p2) sm.pass_idx[pos++] = v2;
                if (p3) sm.pass_idx[pos++] = v3;
                __syncwarp();
                if (total) {
                    rerank<G>(sm, total, rows, p.g.sl, qrow_smem, qreg, qreg_valid, nullptr);
                    maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, total);
                }
                __syncwarp();
            }
            if (SIMPLE) max_diff = p.msd[minval16 < 65536u ? minval16 : 65535u];  // :739-740
            range_idx++;
        }
        // stop rule (:697-713, :748-764): the k-th similarity alone, all L tables of this depth done
        const float kth = __fdiv_rn((float)minval16, 65536.0f);
        uint32_t bin = (uint32_t)__fdiv_rn(kth, 0.005f);
        bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
        const uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (L >> 5));
        if ((word >> (L & 31)) & 1u) {
            ctr.stop_point = depth << 16 | L;
            break;
        }
        __syncwarp();
    }
    maxbuffer_filter(sm.mb, P, k, inserted, minval16);
    return inserted;
}

// ------------------------------------------------------------------------------------------------ CLANN search loop

// src/core/index.rs:311-439 — one warp per query (queries are pulled from a global counter, so cheap queries make room
// for expensive ones). Warps are persistent; grid = multiple of the SM count.
// Cold path of the tensor-pipe centre screen: every distance of one query's row with the reference's arithmetic (one warp).
__device__ __noinline__ void exact_center_row(const float* __restrict__ center_rows, const float* __restrict__ center_norms, uint32_t K,
                                              uint32_t d, const float* __restrict__ qv, float qn, float* __restrict__ cd) {
    for (uint32_t cc = lane_id(); cc < K; cc += 32) cd[cc] = distance_point(center_rows + (uint64_t)cc * d, center_norms[cc], qv, qn, d);
}

template <int G, int OCC, bool DENSE, bool STREAM>
__global__ void __launch_bounds__(256, OCC) k_probe(SearchParams p, QueryBatch b, uint32_t warp_bytes, int stop_at_foreign,
                                                    uint16_t* memo_base, uint64_t memo_stride, uint32_t smem_memo_cap) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    // The persistent grid fills every SM for the whole launch, so kernels of other streams (the other batches in flight of the
    // sharded search: their collectives, selections, hashing) could only start in its last wave. With reserve_sms the CTAs that land
    // on the last SMs leave at once — the work queue does not care who drains it — and those SMs stay free.
    if (p.reserve_sms) {
        uint32_t smid, nsmid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
        if (smid + p.reserve_sms >= nsmid) return;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    // per warp: [query row][WarpSmem][DENSE: memo of smem_memo_cap u16 + mbarrier]
    const uint32_t memo_extra = (DENSE && smem_memo_cap) ? smem_memo_cap * 2 + 16 : 0;
    uint8_t* wbase = s_dyn + (size_t)warp * (warp_bytes + p.g.sl * 2 + memo_extra);
    const WarpSmem sm = carve(wbase + p.g.sl * 2, p.g.L, p.k);
    uint16_t* memo_s = reinterpret_cast<uint16_t*>(wbase + p.g.sl * 2 + warp_bytes);
    unsigned long long* memo_bar = reinterpret_cast<unsigned long long*>(wbase + p.g.sl * 2 + warp_bytes + smem_memo_cap * 2);
    uint32_t memo_phase = 0;
    if (DENSE && smem_memo_cap) {
        if (lane == 0) pc_mbar_init(memo_bar, 1);
        __syncwarp();
    }
    int16_t* qrow = reinterpret_cast<int16_t*>(wbase);  // this warp's query in Q15
    const uint64_t state_bytes = sizeof(QueryStateHeader) + (uint64_t)p.k * 8;
    const uint32_t cpr = p.g.sl / 8;
    const bool qreg_valid = cpr <= 32;
    uint16_t* memo = memo_base ? memo_base + ((uint64_t)blockIdx.x * (blockDim.x >> 5) + warp) * memo_stride : nullptr;

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(b.work_counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= b.nq) break;
        q = b.qperm[q];  // queries sorted by their nearest cluster
        QueryStateHeader* st = reinterpret_cast<QueryStateHeader*>(b.state + (uint64_t)q * state_bytes);
        if (st->done) continue;
        unsigned long long* st_heap = reinterpret_cast<unsigned long long*>(st + 1);
        uint32_t heap_len = st->heap_len;
        uint32_t pos = st->next_pos;
        unsigned long long last_key = st->last_key;
        uint32_t visited = st->visited;
        ProbeCounters ctr{st->candidates, st->distcomp, (uint32_t)((st->stop_point >> 32) << 16 | (st->stop_point & 0xffffu))};
        for (uint32_t i = lane; i < heap_len; i += 32) sm.heap[i] = st_heap[i];
        // query row -> shared memory and this lane's 16-byte chunk -> registers
        for (uint32_t i = lane; i < p.g.sl / 2; i += 32)
            reinterpret_cast<uint32_t*>(qrow)[i] = reinterpret_cast<const uint32_t*>(b.q15 + (uint64_t)q * p.g.sl)[i];
        __syncwarp();
        int qreg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (qreg_valid && (lane % G) < cpr) {
            uint4 w = *reinterpret_cast<const uint4*>(qrow + (lane % G) * 8);
            qreg[0] = unpack_lo(w.x); qreg[1] = unpack_hi(w.x); qreg[2] = unpack_lo(w.y); qreg[3] = unpack_hi(w.y);
            qreg[4] = unpack_lo(w.z); qreg[5] = unpack_hi(w.z); qreg[6] = unpack_lo(w.w); qreg[7] = unpack_hi(w.w);
        }
        const float* qv = b.queries + (uint64_t)q * p.g.d;
        const float qn = b.qnorm[q];
        bool done = false, foreign = false;

        float* cd = b.cdist + (uint64_t)q * p.K;
        float exact_limit = b.exact_limit ? b.exact_limit[q] : INFINITY;
        // cluster-sharded second round (stop_at_foreign == 2): the bound every rank agreed on after the first round and the number
        // of clusters the first round consumed (all on the rank that owns the query's nearest cluster)
        float ext_bound = INFINITY;
        uint32_t pos0 = 0;
        if (stop_at_foreign == 2) {
            const unsigned long long pk = b.shard_packed[q];
            ext_bound = float_from_order_bits((uint32_t)(pk >> 32));
            pos0 = 0xffffffffu - (uint32_t)pk;
        }
        for (; pos < p.K; pos++) {
            // next cluster of the stable ascending centre-distance order (index.rs:592-616): smallest key above last_key
            unsigned long long nk;
            for (;;) {
                nk = ~0ull;
                for (uint32_t cc = lane; cc < p.K; cc += 32) {
                    unsigned long long key = ((unsigned long long)float_order_bits(cd[cc]) << 32) | cc;
                    if ((pos == 0 || key > last_key) && key < nk) nk = key;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    unsigned long long t = __shfl_xor_sync(0xffffffffu, nk, o);
                    nk = t < nk ? t : nk;
                }
                // tensor-pipe screen: values below exact_limit are exact and precede everything else; once the walk reaches the
                // limit the row is re-evaluated with the reference's arithmetic (the keys consumed so far were exact and stay)
                if (!(float_from_order_bits((uint32_t)(nk >> 32)) >= exact_limit)) break;
                exact_center_row(p.center_rows, p.center_norms, p.K, p.g.d, qv, qn, cd);
                exact_limit = INFINITY;
                if (lane == 0) b.exact_limit[q] = INFINITY;
                __syncwarp();
            }
            const uint32_t c = (uint32_t)nk;
            float max_dist = INFINITY;
            if (stop_at_foreign == 2) {
                // The prune test of index.rs:342-361 against a bound that is never below the reference's running k-th distance:
                // the agreed bound of round one, tightened by this rank's own k-th distance once its local heap is full. The walk
                // therefore ends no earlier than the reference's, and visits this rank's clusters among the ones the reference
                // would visit (a superset: recall >= the reference's).
                if (heap_len >= p.k) {
                    unsigned long long top = topk_peek(sm.heap, heap_len);
                    max_dist = float_from_order_bits((uint32_t)(top >> 32));
                }
                max_dist = fminf(max_dist, ext_bound);
                float cmin = __fsub_rn(float_from_order_bits((uint32_t)(nk >> 32)), p.radii[c]);
                if (cmin > max_dist) {
                    done = true;
                    break;
                }
                if (pos < pos0 || p.owner[c] != p.shard_rank) {  // consumed in round one, or another rank's cluster
                    last_key = nk;
                    continue;
                }
            } else if (heap_len > 0) {  // index.rs:342-361
                unsigned long long top = topk_peek(sm.heap, heap_len);
                max_dist = float_from_order_bits((uint32_t)(top >> 32));
                float cmin = __fsub_rn(float_from_order_bits((uint32_t)(nk >> 32)), p.radii[c]);
                if (cmin > max_dist) {
                    done = true;
                    break;
                }
            }
            if (stop_at_foreign == 1 && p.owner[c] != p.shard_rank) {
                foreign = true;
                break;
            }
            last_key = nk;
            visited++;
            // cluster-granularity metrics (opt-in): the prune test above was an evaluation whenever the heap held something
            // (the start values are parked in the log row itself, so nothing stays live across probe_cluster)
            if (b.visit_log && lane == 0 && visited <= b.visit_cap) {
                uint32_t* row = b.visit_log + ((uint64_t)q * b.visit_cap + (visited - 1)) * 4;
                row[0] = c + 1;
                row[2] = (uint32_t)ctr.distcomp - ((stop_at_foreign != 2 && heap_len > 0) ? 1u : 0u);  // index.rs:348
                row[3] = (uint32_t)global_timer_ns();
            }
            uint32_t v_added = 0, v_brute = 0;
            const uint64_t off = p.offsets[c];
            const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
            if (p.brute[c]) {
                // index.rs:364-378 with brute_force_search :666-685: members in assignment order into a local top-k, then merge
                uint32_t loc_len = 0;
                for (uint32_t base = 0; base < nc; base += 32) {
                    uint32_t j = base + lane;
                    float dist = 0.0f;
                    uint32_t pid = 0;
                    if (j < nc) {
                        pid = p.perm[off + j];
                        dist = distance_point(p.data + (uint64_t)pid * p.g.d, p.norms[pid], qv, qn, p.g.d);
                    }
                    uint32_t lim = nc - base < 32 ? nc - base : 32;
                    for (uint32_t l = 0; l < lim; l++) {
                        float dl = __shfl_sync(0xffffffffu, dist, l);
                        uint32_t il = __shfl_sync(0xffffffffu, pid, l);
                        topk_add(sm.loc, loc_len, p.k, dl, il);
                    }
                }
                // to_list (heap.rs:42-48): ascending by distance; ties are in std's heap order there, by id here
                const uint32_t P = next_pow2(2 * p.k) < 32 ? 32 : next_pow2(2 * p.k);
                for (uint32_t i = lane; i < P; i += 32) sm.mb[i] = i < loc_len ? ~sm.loc[i] : 0ull;  // ~ turns asc into desc
                __syncwarp();
                warp_sort_desc(sm.mb, P);
                for (uint32_t i = 0; i < loc_len; i++) {
                    unsigned long long key = ~sm.mb[i];
                    v_added += topk_add(sm.heap, heap_len, p.k, float_from_order_bits((uint32_t)(key >> 32)), (uint32_t)key) ? 1u : 0u;
                }
                v_brute = loc_len;  // index.rs:378
            } else {
                const uint32_t fs = p.fset_of[c];
                const float max_sim = __fsub_rn(1.0f, __fdiv_rn(max_dist, 2.0f));  // puffinn_types.rs:77-79
                const uint32_t* codes = b.codes + (uint64_t)fs * p.g.L * b.nq + q;
                const uint64_t my_sketch = b.sketches[((uint64_t)fs * b.nq + q) * kNumSketches + lane];
                const uint32_t* stop = p.stop + (uint64_t)fs * kMaxHashBits * kEstBins * p.stop_words;
                uint16_t* use_memo = (memo && nc <= memo_stride) ? memo : nullptr;
                // first visit of a fresh query: its similarities to the whole cluster were computed in advance (launch_dense_sims)
                // (DENSE is a template parameter: the instantiation without it keeps the register allocation it had)
                const bool prefilled = DENSE && pos == 0;
                if (prefilled) use_memo = b.dense + (uint64_t)q * b.dense_stride;
                unsigned long long* wait_bar = nullptr;
                if (DENSE && smem_memo_cap && nc <= smem_memo_cap) {
                    // the memo of this visit lives in shared memory: the pre-filled one arrives by one TMA bulk copy while the
                    // anchors are computed; later visits zero it (probe_cluster)
                    if (prefilled) {
                        const uint32_t bytes = (nc * 2 + 15) & ~15u;
                        __syncwarp();
                        if (lane == 0) {
                            pc_mbar_expect_tx(memo_bar, bytes);
                            pc_bulk_g2s(memo_s, use_memo, bytes, memo_bar);
                        }
                        wait_bar = memo_bar;
                    }
                    use_memo = memo_s;
                }
                const uint32_t* pre_a = (prefilled && b.pre_anchor) ? b.pre_anchor + (uint64_t)q * p.g.L : nullptr;
                const uint32_t* pre_r = (prefilled && b.pre_range) ? b.pre_range + (uint64_t)q * kMaxHashBits * p.g.L : nullptr;
                const uint4* pre_l = (prefilled && b.pre_lcp && !b.pre_range) ? b.pre_lcp + (uint64_t)q * p.g.L : nullptr;
                FsView fsv{nullptr, nullptr, nullptr, nullptr};
                if (STREAM && prefilled && b.fs_meta) {
                    const uint64_t sbase = (uint64_t)q * b.fs_cap;
                    fsv.idx = reinterpret_cast<const uint2*>(b.fs_idx) + sbase;
                    fsv.hd = b.fs_hd + sbase;
                    fsv.tab = b.fs_tab + (sbase >> 5);
                    fsv.meta = b.fs_meta + (uint64_t)q * kFsMeta;
                }
                uint32_t cnt = probe_cluster<G, DENSE, STREAM>(p, sm, c, codes, b.nq, my_sketch, stop, max_sim, qrow, qreg, qreg_valid, use_memo, ctr,
                                                       prefilled, wait_bar, memo_phase, pre_a, pre_r, pre_l, fsv);
                if (wait_bar) memo_phase ^= 1u;
                // map_candidates + fp32 distance + heap (index.rs:392-416); results are visited best-first
                for (uint32_t base = 0; base < cnt; base += 32) {
                    uint32_t j = base + lane;
                    float dist = 0.0f;
                    uint32_t pid = 0;
                    if (j < cnt) {
                        pid = p.perm[off + (uint32_t)sm.mb[j]];
                        dist = distance_point(p.data + (uint64_t)pid * p.g.d, p.norms[pid], qv, qn, p.g.d);
                    }
                    uint32_t lim = cnt - base < 32 ? cnt - base : 32;
                    for (uint32_t l = 0; l < lim; l++) {
                        float dl = __shfl_sync(0xffffffffu, dist, l);
                        uint32_t il = __shfl_sync(0xffffffffu, pid, l);
                        v_added += topk_add(sm.heap, heap_len, p.k, dl, il) ? 1u : 0u;
                    }
                }
            }
            if (b.visit_log && lane == 0 && visited <= b.visit_cap) {
                uint32_t* row = b.visit_log + ((uint64_t)q * b.visit_cap + (visited - 1)) * 4;
                row[1] = v_added;
                row[2] = (uint32_t)ctr.distcomp + v_brute - row[2];  // index.rs:378,421
                row[3] = (uint32_t)global_timer_ns() - row[3];
            }
        }
        if (pos >= p.K) done = true;
        (void)foreign;
        __syncwarp();
        for (uint32_t i = lane; i < heap_len; i += 32) st_heap[i] = sm.heap[i];
        if (lane == 0) {
            st->heap_len = heap_len;
            st->next_pos = pos;
            st->last_key = last_key;
            st->visited = visited;
            st->done = done ? 1u : 0u;
            st->candidates = ctr.candidates;
            st->distcomp = ctr.distcomp;
            st->stop_point = ((unsigned long long)(ctr.stop_point >> 16) << 32) | (ctr.stop_point & 0xffffu);
        }
        __syncwarp();
    }
}

// Multi-GPU: adopt, for every query, the state of the rank that advanced it furthest (exactly one rank advances a
// query per step: the owner of its next cluster).
__global__ void k_merge_states(uint8_t* __restrict__ mine, const uint8_t* __restrict__ all, int world, uint64_t nq,
                               uint64_t state_bytes, uint32_t* __restrict__ active, uint32_t* work_counter) {
    uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q == 0) *work_counter = 0;
    if (q >= nq) return;
    int best = 0;
    uint32_t best_pos = 0, best_done = 0;
    for (int r = 0; r < world; r++) {
        const QueryStateHeader* h = reinterpret_cast<const QueryStateHeader*>(all + ((uint64_t)r * nq + q) * state_bytes);
        if (r == 0 || h->done > best_done || (h->done == best_done && h->next_pos > best_pos)) {
            best = r;
            best_pos = h->next_pos;
            best_done = h->done;
        }
    }
    const uint64_t* src = reinterpret_cast<const uint64_t*>(all + ((uint64_t)best * nq + q) * state_bytes);
    uint64_t* dst = reinterpret_cast<uint64_t*>(mine + q * state_bytes);
    for (uint64_t i = 0; i < state_bytes / 8; i++) dst[i] = src[i];
    if (!best_done) atomicAdd(active, 1u);
}

// ------------------------------------------------------------------------------------------------ cluster-sharded search

// Round-one routing: the global ids of the queries whose nearest cluster this rank owns.
__global__ void __launch_bounds__(256) k_shard_select_owned(const uint32_t* __restrict__ first_all, const uint8_t* __restrict__ owner,
                                                            uint32_t rank, uint64_t nq, uint32_t* __restrict__ list, uint32_t* count) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = gid < nq && owner[first_all[gid]] == rank;
    const uint32_t bal = __ballot_sync(0xffffffffu, mine);
    uint32_t base = 0;
    if (lane_id() == 0 && bal) base = atomicAdd(count, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) list[base + __popc(bal & ((1u << lane_id()) - 1u))] = (uint32_t)gid;
}

// Round-two routing: the queries no rank has finished (agreed bound >= 0), with their packed (bound, consumed) word.
__global__ void __launch_bounds__(256) k_shard_select_open(const unsigned long long* __restrict__ packed, uint64_t nq,
                                                           uint32_t* __restrict__ list, unsigned long long* __restrict__ packed_local,
                                                           uint32_t* count) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long pk = gid < nq ? packed[gid] : 0ull;
    const bool open = gid < nq && (uint32_t)(pk >> 32) >= 0x80000000u;  // order bits of a non-negative float
    const uint32_t bal = __ballot_sync(0xffffffffu, open);
    uint32_t base = 0;
    if (lane_id() == 0 && bal) base = atomicAdd(count, (uint32_t)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (open) {
        const uint32_t slot = base + __popc(bal & ((1u << lane_id()) - 1u));
        list[slot] = (uint32_t)gid;
        packed_local[slot] = pk;
    }
}

// Round two, second cut: of the open queries, those for which this rank owns at least one cluster the agreed bound does not prune
// (distance to the centre - radius <= bound, index.rs:342-361 with the bound of round one). Only these are hashed and probed
// here; the test ignores the order of the walk, so it keeps a superset of the queries the probe would actually serve.
__global__ void __launch_bounds__(256) k_shard_select_mine(const float* __restrict__ cdist, const float* __restrict__ exact_limit,
                                                           const float* __restrict__ radii, const uint8_t* __restrict__ owner,
                                                           uint32_t rank, uint32_t K, const uint32_t* __restrict__ list_in,
                                                           const unsigned long long* __restrict__ packed_in, uint32_t count_in,
                                                           uint32_t* __restrict__ list_out, unsigned long long* __restrict__ packed_out,
                                                           uint32_t* count_out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = lane_id();
    if (warp >= count_in) return;
    const unsigned long long pk = packed_in[warp];
    const float bound = float_from_order_bits((uint32_t)(pk >> 32));
    const float* cd = cdist + (uint64_t)warp * K;
    // entries at or above exact_limit are lower bounds of the exact distance (tensor-pipe screen): testing them keeps a superset
    (void)exact_limit;
    bool any = false;
    for (uint32_t c = lane; c < K; c += 32) any |= owner[c] == rank && !(cd[c] - radii[c] > bound);
    if (__any_sync(0xffffffffu, any) && lane == 0) {
        const uint32_t slot = atomicAdd(count_out, 1u);
        list_out[slot] = list_in[warp];
        packed_out[slot] = pk;
    }
}

__global__ void __launch_bounds__(256) k_shard_gather_rows(const float* __restrict__ all, const uint32_t* __restrict__ list, uint32_t count,
                                                           uint32_t d, float* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (uint64_t)count * d) out[i] = all[(uint64_t)list[i / d] * d + i % d];
}

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long* __restrict__ p, uint64_t n, unsigned long long v) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// What a rank tells the others about the queries it advanced in round one: a bound on the k-th distance that holds whatever the
// other ranks find — the heap top once the heap is full (index.rs:342-361 peeks earlier; an unfilled heap bounds nothing), -1 for
// a finished query — and how many clusters of the visiting order it consumed. min over ranks of
// (order_bits(bound) << 32) | (0xffffffff - consumed) picks the advancing rank's word (the others hold +inf / 0).
__global__ void __launch_bounds__(256) k_shard_pack_bounds(const uint8_t* __restrict__ state, uint64_t state_bytes, uint32_t k,
                                                           const uint32_t* __restrict__ list, uint32_t count,
                                                           unsigned long long* __restrict__ packed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const QueryStateHeader* h = reinterpret_cast<const QueryStateHeader*>(state + (uint64_t)i * state_bytes);
    const unsigned long long* heap = reinterpret_cast<const unsigned long long*>(h + 1);
    uint32_t bits = float_order_bits(INFINITY);
    if (h->done) bits = float_order_bits(-1.0f);
    else if (h->heap_len >= k) {
        uint32_t top = 0;
        for (uint32_t j = 0; j < h->heap_len; j++) top = max(top, (uint32_t)(heap[j] >> 32));
        bits = top;
    }
    packed[list[i]] = ((unsigned long long)bits << 32) | (0xffffffffu - h->next_pos);
}

// This rank's candidates of every query it touched, ascending, into top[gid * k .. + k) (~0 = empty); one warp per local query.
// merge != 0: the row already holds the sorted candidates of an earlier round. Counters are accumulated per global query.
__global__ void __launch_bounds__(256) k_shard_collect(const uint8_t* __restrict__ state, uint64_t state_bytes, uint32_t k,
                                                       const uint32_t* __restrict__ list, uint32_t count, unsigned long long* __restrict__ top,
                                                       int merge, unsigned long long* __restrict__ counters) {
    extern __shared__ unsigned long long s_cand[];  // [warps][2k]
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp;
    if (i >= count) return;
    const uint32_t gid = list[i];
    const QueryStateHeader* h = reinterpret_cast<const QueryStateHeader*>(state + (uint64_t)i * state_bytes);
    const unsigned long long* heap = reinterpret_cast<const unsigned long long*>(h + 1);
    unsigned long long* cand = s_cand + (size_t)warp * 2 * k;
    unsigned long long* row = top + (uint64_t)gid * k;
    for (uint32_t j = lane; j < k; j += 32) {
        cand[j] = j < h->heap_len ? heap[j] : ~0ull;
        cand[k + j] = merge ? row[j] : ~0ull;
    }
    __syncwarp();
    for (uint32_t j = lane; j < 2 * k; j += 32) {
        const unsigned long long key = cand[j];
        if (key == ~0ull) continue;
        uint32_t rank = 0;
        for (uint32_t o = 0; o < 2 * k; o++) rank += (cand[o] < key) || (cand[o] == key && o < j);
        if (rank < k) row[rank] = key;
    }
    if (!merge) {
        uint32_t valid = h->heap_len < k ? h->heap_len : k;
        for (uint32_t j = valid + lane; j < k; j += 32) row[j] = ~0ull;
    }
    if (lane == 0 && counters) {
        counters[(uint64_t)gid * 3 + 0] += h->candidates;
        counters[(uint64_t)gid * 3 + 1] += h->distcomp;
        counters[(uint64_t)gid * 3 + 2] += h->visited;
    }
}

// k-way merge of the ranks' candidate lists (all[r * nq * k + q * k ..]) into the final results (heap.rs:42-48 order); one warp per query.
__global__ void __launch_bounds__(256) k_shard_final_merge(const unsigned long long* __restrict__ all, uint32_t world, uint64_t nq, uint32_t k,
                                                           uint32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                                                           uint32_t* __restrict__ out_counts) {
    const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    const uint32_t lane = lane_id();
    const uint32_t total = world * k;
    uint32_t valid = 0;
    for (uint32_t j = lane; j < total; j += 32) {
        const unsigned long long key = all[(uint64_t)(j / k) * nq * k + q * k + j % k];
        if (key == ~0ull) continue;
        valid++;
        uint32_t rank = 0;
        for (uint32_t o = 0; o < total; o++) {
            const unsigned long long other = all[(uint64_t)(o / k) * nq * k + q * k + o % k];
            rank += (other < key) || (other == key && o < j);
        }
        if (rank < k) {
            out_ids[q * k + rank] = (uint32_t)key;
            out_dists[q * k + rank] = float_from_order_bits((uint32_t)(key >> 32));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    const uint32_t cnt = valid < k ? valid : k;
    for (uint32_t j = cnt + lane; j < k; j += 32) {
        out_ids[q * k + j] = 0xffffffffu;
        out_dists[q * k + j] = INFINITY;
    }
    if (lane == 0) out_counts[q] = cnt;
}

void launch_shard_select_owned(const uint32_t* first_all, const uint8_t* owner, uint32_t rank, uint64_t nq, uint32_t* list, uint32_t* count,
                               cudaStream_t s) {
    if (nq) k_shard_select_owned<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(first_all, owner, rank, nq, list, count);
}
void launch_shard_select_open(const unsigned long long* packed, uint64_t nq, uint32_t* list, unsigned long long* packed_local,
                              uint32_t* count, cudaStream_t s) {
    if (nq) k_shard_select_open<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(packed, nq, list, packed_local, count);
}
void launch_shard_select_mine(const float* cdist, const float* exact_limit, const float* radii, const uint8_t* owner, uint32_t rank,
                              uint32_t K, const uint32_t* list_in, const unsigned long long* packed_in, uint32_t count_in,
                              uint32_t* list_out, unsigned long long* packed_out, uint32_t* count_out, cudaStream_t s) {
    if (count_in)
        k_shard_select_mine<<<(unsigned)(((uint64_t)count_in * 32 + 255) / 256), 256, 0, s>>>(cdist, exact_limit, radii, owner, rank, K, list_in,
                                                                                          packed_in, count_in, list_out, packed_out, count_out);
}
void launch_shard_gather_rows(const float* all, const uint32_t* list, uint32_t count, uint32_t d, float* out, cudaStream_t s) {
    const uint64_t n = (uint64_t)count * d;
    if (n) k_shard_gather_rows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(all, list, count, d, out);
}
void launch_fill_u64(unsigned long long* p, uint64_t n, unsigned long long v, cudaStream_t s) {
    if (n) k_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, n, v);
}
void launch_shard_pack_bounds(const uint8_t* state, uint32_t k, const uint32_t* list, uint32_t count, unsigned long long* packed,
                              cudaStream_t s) {
    if (count) k_shard_pack_bounds<<<(count + 255) / 256, 256, 0, s>>>(state, query_state_bytes(k), k, list, count, packed);
}
void launch_shard_collect(const uint8_t* state, uint32_t k, const uint32_t* list, uint32_t count, unsigned long long* top, bool merge,
                          unsigned long long* counters, cudaStream_t s) {
    if (!count) return;
    const size_t smem = (size_t)8 * 2 * k * sizeof(unsigned long long);
    k_shard_collect<<<(count + 7) / 8, 256, smem, s>>>(state, query_state_bytes(k), k, list, count, top, merge ? 1 : 0, counters);
}
void launch_shard_final_merge(const unsigned long long* all, uint32_t world, uint64_t nq, uint32_t k, uint32_t* out_ids, float* out_dists,
                              uint32_t* out_counts, cudaStream_t s) {
    if (nq) k_shard_final_merge<<<(unsigned)((nq * 32 + 255) / 256), 256, 0, s>>>(all, world, nq, k, out_ids, out_dists, out_counts);
}

// heap.rs:42-48 — results ascending by distance; pads with 0xFFFFFFFF / +inf.
__global__ void __launch_bounds__(128) k_finish(const uint8_t* __restrict__ state, uint64_t nq, uint32_t k, uint64_t state_bytes,
                                                uint32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                                                uint32_t* __restrict__ out_counts, unsigned long long* __restrict__ cnt_cand,
                                                unsigned long long* __restrict__ cnt_dc, uint32_t* __restrict__ cnt_vis,
                                                unsigned long long* stats_dev, volatile unsigned long long* stats_host) {
    // one warp per query; selection sort by repeated minimum is fine for small k, bitonic otherwise is unnecessary here
    const uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t my_visited = 0;
    if (q < nq) {
    const QueryStateHeader* h = reinterpret_cast<const QueryStateHeader*>(state + q * state_bytes);
    const unsigned long long* heap = reinterpret_cast<const unsigned long long*>(h + 1);
    const uint32_t len = h->heap_len;
    const uint32_t lane = lane_id();
    // rank of each entry = number of strictly smaller keys (keys are unique unless an id repeats; ties broken by slot)
    for (uint32_t i = lane; i < len; i += 32) {
        unsigned long long key = heap[i];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < len; j++) {
            unsigned long long o = heap[j];
            rank += (o < key) || (o == key && j < i);
        }
        out_ids[q * k + rank] = (uint32_t)key;
        out_dists[q * k + rank] = float_from_order_bits((uint32_t)(key >> 32));
    }
    for (uint32_t i = len + lane; i < k; i += 32) {
        out_ids[q * k + i] = 0xffffffffu;
        out_dists[q * k + i] = INFINITY;
    }
    if (lane == 0) {
        out_counts[q] = len;
        if (cnt_cand) cnt_cand[q] = h->candidates;
        if (cnt_dc) cnt_dc[q] = h->distcomp;
        if (cnt_vis) cnt_vis[q] = h->visited;
        my_visited = h->visited;
    }
    }
    // batch statistics: clusters visited summed per block, the last block to finish publishes the batch total to the host
    if (stats_dev) {
        __shared__ uint32_t s_vis;
        if (threadIdx.x == 0) s_vis = 0;
        __syncthreads();
        if (my_visited) atomicAdd(&s_vis, my_visited);
        __syncthreads();
        if (threadIdx.x == 0) {
            atomicAdd(&stats_dev[0], (unsigned long long)s_vis);
            __threadfence();
            const unsigned long long ticket = atomicAdd(&stats_dev[1], 1ull);
            if (ticket == gridDim.x - 1) {
                stats_host[0] = atomicExch(&stats_dev[0], 0ull);
                stats_host[1] = nq;
                stats_dev[1] = 0;
                __threadfence_system();
            }
        }
    }
}

// Legacy single-index query (c_binder.cpp:69-96 -> collection.hpp:324-334): one warp, cluster 0, explicit max_sim.
template <int G>
__global__ void __launch_bounds__(32) k_puffinn_search(SearchParams p, QueryBatch b, const uint32_t* stop, float max_sim,
                                                       int filter_type, uint32_t warp_bytes, uint32_t* out_ids, uint32_t* out_count,
                                                       uint32_t* out_distcomp, uint32_t* out_stop) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const uint32_t lane = lane_id();
    const WarpSmem sm = carve(s_dyn + p.g.sl * 2, p.g.L, p.k);
    (void)warp_bytes;
    int16_t* qrow = reinterpret_cast<int16_t*>(s_dyn);
    for (uint32_t i = lane; i < p.g.sl / 2; i += 32) reinterpret_cast<uint32_t*>(qrow)[i] = reinterpret_cast<const uint32_t*>(b.q15)[i];
    __syncwarp();
    const uint32_t cpr = p.g.sl / 8;
    const bool qreg_valid = cpr <= 32;
    int qreg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (qreg_valid && (lane % G) < cpr) {
        uint4 w = *reinterpret_cast<const uint4*>(qrow + (lane % G) * 8);
        qreg[0] = unpack_lo(w.x); qreg[1] = unpack_hi(w.x); qreg[2] = unpack_lo(w.y); qreg[3] = unpack_hi(w.y);
        qreg[4] = unpack_lo(w.z); qreg[5] = unpack_hi(w.z); qreg[6] = unpack_lo(w.w); qreg[7] = unpack_hi(w.w);
    }
    ProbeCounters ctr{0, 0, 0};
    uint32_t cnt;
    if (p.n < 100) {  // collection.hpp:550-555
        cnt = probe_bruteforce_q15<G>(p, sm, 0, qrow, qreg, qreg_valid);
    } else {
        const uint64_t my_sketch = b.sketches[lane];
        if (filter_type == 1)  // FilterType::None
            cnt = probe_cluster_plain<G, false>(p, sm, 0, b.codes, 1, my_sketch, stop, qrow, qreg, qreg_valid, ctr);
        else if (filter_type == 2)  // FilterType::Simple
            cnt = probe_cluster_plain<G, true>(p, sm, 0, b.codes, 1, my_sketch, stop, qrow, qreg, qreg_valid, ctr);
        else
            cnt = probe_cluster<G>(p, sm, 0, b.codes, 1, my_sketch, stop, max_sim, qrow, qreg, qreg_valid, nullptr, ctr);
    }
    for (uint32_t i = lane; i < cnt; i += 32) out_ids[i] = (uint32_t)sm.mb[i];
    if (lane == 0) {
        *out_count = cnt;
        *out_distcomp = (uint32_t)ctr.distcomp;
        *out_stop = ctr.stop_point;
    }
}

// ------------------------------------------------------------------------------------------------ first-visit anchors and ranges

// One thread per (work item w, table t): the anchor of query qperm[w] in table t of its nearest cluster first[w] and the range of
// every depth, exactly as the probe kernel derives them (table_anchor / table_range, probe_common.cuh) — but as 840 000
// independent threads instead of three dependent rounds per warp at the head of every visit and a round of directory reads or
// binary searches at every depth.
__global__ void __launch_bounds__(256) k_first_ranges(SearchParams p, QueryBatch b) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = p.g.L;
    if (i >= b.nq * L) return;
    const uint32_t w = (uint32_t)(i / L), t = (uint32_t)(i % L);
    const uint32_t q = b.qperm[w], c = b.first[w];
    if (p.brute[c]) return;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t h = b.codes[((uint64_t)p.fset_of[c] * L + t) * b.nq + q];
    const uint32_t* H = p.tbl_hash + table_base(off, nc, L, t);
    const uint32_t* dir = p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries;
    uint32_t A;
    uint2 up, dn;
    table_anchor(H, dir, nc, h, A, up, dn);
    b.pre_anchor[(uint64_t)q * L + t] = A;
    if (b.pre_lcp) b.pre_lcp[(uint64_t)q * L + t] = make_uint4(up.x, up.y, dn.x, dn.y);
    if (!b.pre_range) return;
    uint32_t* out = b.pre_range + (uint64_t)q * kMaxHashBits * L + t;
    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        uint32_t nseg;
        const uint32_t start = table_range(H, dir, nc, h, A, up, dn, depth, nseg);
        out[(size_t)(depth - 1) * L] = nseg | (start == A ? 0x80000000u : 0u);  // start == A: upward (or empty: nseg == 0)
    }
}

// Order-free trace of the probe (SURVEY.md 8b clann_export_trace): the anchor of every query of the batch in every table of
// cluster c (prefixmap.hpp:36-57,250-260) and the range get_next_range returns at each of its 24 calls (prefixmap.hpp:267-304),
// in the reference's own coordinates (positions in the table padded by 12 sentinels each side): anchors[q*L + t],
// ranges[((q*24 + it)*L + t)*2 + {0,1}] = {start, end} of call it = 0..23 (depth 24 - it). The same table_anchor /
// table_range the probe kernel uses, so the parity tests pin them directly against the reference's golden arrays.
__global__ void __launch_bounds__(256) k_export_ranges(SearchParams p, QueryBatch b, uint32_t c, uint32_t* __restrict__ anchors,
                                                       uint32_t* __restrict__ ranges) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t L = p.g.L;
    if (i >= b.nq * L) return;
    const uint32_t q = (uint32_t)(i / L), t = (uint32_t)(i % L);
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t h = b.codes[((uint64_t)p.fset_of[c] * L + t) * b.nq + q];
    const uint32_t* H = p.tbl_hash + table_base(off, nc, L, t);
    const uint32_t* dir = p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries;
    uint32_t A;
    uint2 up, dn;
    table_anchor(H, dir, nc, h, A, up, dn);
    anchors[(uint64_t)q * L + t] = A + kSegment;
    for (uint32_t it = 0; it < (uint32_t)kMaxHashBits; it++) {
        uint32_t nseg;
        const uint32_t start = table_range(H, dir, nc, h, A, up, dn, kMaxHashBits - it, nseg);
        uint32_t* out = ranges + (((uint64_t)q * kMaxHashBits + it) * L + t) * 2;
        out[0] = start + kSegment;
        out[1] = start + kSegment + 4 * nseg;
    }
}

void launch_export_ranges(const SearchParams& p, const QueryBatch& b, uint32_t cluster, uint32_t* anchors, uint32_t* ranges,
                          cudaStream_t s) {
    if (b.nq == 0) return;
    const uint64_t threads = b.nq * p.g.L;
    k_export_ranges<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p, b, cluster, anchors, ranges);
}

void launch_first_ranges(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0 || !b.pre_anchor || (!b.pre_range && !b.pre_lcp)) return;
    const uint64_t threads = b.nq * p.g.L;
    k_first_ranges<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(p, b);
}

// ------------------------------------------------------------------------------------------------ first-visit candidate stream

// Per-warp scratch of k_first_stream (the fields anchors_lockstep fills, plus range starts, segment prefix and a histogram).
struct StreamSmem {
    uint2* lcp_up;
    uint2* lcp_dn;
    uint32_t* anchor;
    uint32_t* code;
    uint32_t* start;
    uint32_t* segbase;
    uint32_t* hist;  // [256]
};

__host__ __device__ inline uint32_t stream_smem_bytes(uint32_t L) {
    uint32_t b = L * (8 + 8 + 4 + 4 + 4) + (L + 1) * 4;
    b = (b + 15) & ~15u;
    return b + 256 * 4;
}

// k-th largest of n u16 values (n < 65 536 * 2) by a two-pass byte radix select over a warp-private histogram; 0 when n < k.
// `v` is 4-byte aligned. This is the largest value MaxBuffer's k-th entry can ever take in this cluster (maxbuffer.hpp:25-46).
__device__ __forceinline__ uint32_t warp_kth_largest_u16(const uint16_t* __restrict__ v, uint32_t n, uint32_t k, uint32_t* hist) {
    if (k == 0 || n < k) return 0;
    const uint32_t lane = lane_id();
    const uint32_t* v32 = reinterpret_cast<const uint32_t*>(v);
    uint32_t prefix = 0, want = k;  // rank (from the top) still to be located inside the current prefix class
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        for (uint32_t i = lane; i < 256; i += 32) hist[i] = 0;
        __syncwarp();
        for (uint32_t i = lane; i < (n + 1) / 2; i += 32) {
            const uint32_t w = __ldg(v32 + i);
            const uint32_t a = w & 0xffffu, bb = w >> 16;
            if (pass == 0) {
                atomicAdd(&hist[a >> 8], 1u);
                if (2 * i + 1 < n) atomicAdd(&hist[bb >> 8], 1u);
            } else {
                if ((a >> 8) == prefix) atomicAdd(&hist[a & 0xffu], 1u);
                if (2 * i + 1 < n && (bb >> 8) == prefix) atomicAdd(&hist[bb & 0xffu], 1u);
            }
        }
        __syncwarp();
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) mine += hist[8 * lane + j];
        uint32_t suf = mine;  // entries in the bins of this lane and of every higher lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
            if (lane + o < 32) suf += t;
        }
        const uint32_t above = suf - mine;
        const uint32_t owner = __ffs(__ballot_sync(0xffffffffu, above < want && want <= suf)) - 1;
        uint32_t bin = 0, rem = 0;
        if (lane == owner) {
            uint32_t acc = above;
            for (int j = 7; j >= 0; j--) {
                const uint32_t h = hist[8 * lane + j];
                if (acc + h >= want) {
                    bin = 8 * lane + j;
                    rem = want - acc;
                    break;
                }
                acc += h;
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        rem = __shfl_sync(0xffffffffu, rem, owner);
        __syncwarp();
        if (pass == 0) {
            prefix = bin;
            want = rem;
        } else {
            return (prefix << 8) | bin;
        }
    }
    return 0;
}

// One warp per work item (query qperm[w], its nearest cluster first[w]): everything of the query's first visit that does not
// depend on the search state, written out in the order search_maps consumes it (QueryBatch::fs_*). Per depth 24, 23, ...:
// the ranges of all tables (fill_ranges, collection.hpp:650-667 — the same table_anchor / table_range as the probe), then
// for every 4-entry segment of the concatenated ranges its table indices and, for ring slot = segment mod 32 (= lane), the
// Hamming distance of each candidate's sketch word to the query's (filterer.hpp:28-31). All loads are independent of one
// another across segments, so they overlap freely — inside the probe the same loads form a chain of two dependent round trips per
// ring sweep. The stream stops after the first depth at whose end the stop rule (collection.hpp:927-943) fires for the
// cluster's true k-th similarity: the lagging k-th value of MaxBuffer can only be lower, so the visit cannot end earlier.
__global__ void __launch_bounds__(256, 3) k_first_stream(SearchParams p, QueryBatch b, uint32_t warp_bytes) {
    extern __shared__ __align__(16) uint8_t s_stream[];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= b.nq) return;
    const uint32_t q = b.qperm[w], c = b.first[w];
    uint32_t* meta = b.fs_meta + (uint64_t)q * kFsMeta;
    const uint32_t L = p.g.L;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    if (p.brute[c] || nc > 65536u || nc > b.dense_stride) {
        if (lane == 0) meta[48] = kMaxHashBits + 1;  // nothing streamed: the probe does it all
        return;
    }
    StreamSmem sm;
    {
        uint8_t* ptr = s_stream + (size_t)warp * warp_bytes;
        sm.lcp_up = reinterpret_cast<uint2*>(ptr); ptr += L * 8;
        sm.lcp_dn = reinterpret_cast<uint2*>(ptr); ptr += L * 8;
        sm.anchor = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.code = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.start = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.segbase = reinterpret_cast<uint32_t*>(ptr);
        sm.hist = reinterpret_cast<uint32_t*>(s_stream + (size_t)warp * warp_bytes + (((L * 28 + (L + 1) * 4) + 15) & ~15u));
    }
    const uint32_t fsid = p.fset_of[c];
    const uint32_t* codes = b.codes + (uint64_t)fsid * L * b.nq + q;
    const uint64_t my_sketch = b.sketches[((uint64_t)fsid * b.nq + q) * kNumSketches + lane];
    const uint32_t* stop = p.stop + (uint64_t)fsid * kMaxHashBits * kEstBins * p.stop_words;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    anchors_lockstep(p, sm, c, off, nc, codes, b.nq);
    // the similarity bin the stop rule will look at once MaxBuffer holds the cluster's true top k
    const uint32_t kth16 = warp_kth_largest_u16(b.dense + (uint64_t)q * b.dense_stride, nc, p.k, sm.hist);
    uint32_t bin = (uint32_t)__fdiv_rn(__fdiv_rn((float)kth16, 65536.0f), 0.005f);
    bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;

    const uint64_t sbase = (uint64_t)q * b.fs_cap;
    uint2* out_idx = reinterpret_cast<uint2*>(b.fs_idx) + sbase;
    uint32_t* out_hd = b.fs_hd + sbase;
    uint8_t* out_tab = b.fs_tab + (sbase >> 5);
    uint32_t cum = 0, lowest = kMaxHashBits + 1;
    uint32_t my_S = 0, my_off = 0;  // lane d-1 keeps the record of depth d
    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        uint32_t running = 0;
        for (uint32_t t0 = 0; t0 < L; t0 += 32) {
            const uint32_t t = t0 + lane;
            uint32_t nseg = 0;
            if (t < L)
                sm.start[t] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc,
                                          sm.code[t], sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg);
            uint32_t total;
            const uint32_t ex = warp_excl_scan(nseg, total);
            if (t < L) sm.segbase[t] = running + ex;
            running += total;
        }
        if (lane == 0) sm.segbase[L] = running;
        __syncwarp();
        const uint32_t S = running;
        const uint32_t S32 = (S + 31u) & ~31u;
        if (S > (uint32_t)kRing && cum + S32 > b.fs_cap) break;  // no room: the probe evaluates this depth and the deeper ones
        if (lane == depth - 1) {
            my_S = S;
            my_off = cum;
        }
        lowest = depth;
        if (S <= (uint32_t)kRing) continue;  // skipped by search_maps (collection.hpp:802-810): nothing to stream, no stop check
        for (uint32_t s0 = 0; s0 < S; s0 += 2 * kRing) {
            uint32_t v[2][4], tb[2];
            bool valid[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t s = s0 + u * kRing + lane;
                valid[u] = s < S;
                tb[u] = 0;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0;
                if (valid[u]) {
                    uint32_t lo = 0, len = L;  // upper_bound(segbase, s) - 1 over segbase[0..L)
                    while (len > 0) {
                        const uint32_t half = len >> 1, mid = lo + half;
                        if (sm.segbase[mid] <= s) { lo = mid + 1; len -= half + 1; } else { len = half; }
                    }
                    const uint32_t t = lo - 1;
                    tb[u] = t;
                    const uint32_t* seg = p.tbl_idx + table_base(off, nc, L, t) + sm.start[t] + 4 * (s - sm.segbase[t]);
                    v[u][0] = __ldg(seg); v[u][1] = __ldg(seg + 1); v[u][2] = __ldg(seg + 2); v[u][3] = __ldg(seg + 3);
                }
            }
            uint64_t sw[2][4];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) sw[u][j] = valid[u] ? __ldg(sk + ((uint64_t)v[u][j] << 5 | lane)) : 0ull;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!valid[u]) continue;
                const uint32_t s = s0 + u * kRing + lane;
                const uint32_t h4 = (uint32_t)__popcll(sw[u][0] ^ my_sketch) | (uint32_t)__popcll(sw[u][1] ^ my_sketch) << 8 |
                                    (uint32_t)__popcll(sw[u][2] ^ my_sketch) << 16 | (uint32_t)__popcll(sw[u][3] ^ my_sketch) << 24;
                out_idx[cum + s] = make_uint2(v[u][0] | v[u][1] << 16, v[u][2] | v[u][3] << 16);
                out_hd[cum + s] = h4;
                if (lane == 0) out_tab[(cum + s) >> 5] = (uint8_t)tb[u];
            }
        }
        cum += S32;
        // would the stop rule fire at the end of this depth (table index L, every table consumed) for the cluster's true k-th value?
        const uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (L >> 5));
        if ((word >> (L & 31)) & 1u) break;
        __syncwarp();  // sm.start of this depth must not be overwritten while another lane still reads it
    }
    if (lane < (uint32_t)kMaxHashBits) {
        meta[lane] = my_S;
        meta[kMaxHashBits + lane] = my_off;
    }
    if (lane == 0) meta[48] = lowest;
}

// The same stream produced by one CTA of four warps per query (knob first_stream = 2): every table has its own thread for the
// anchor and for the range of each depth (no three-rounds-per-lane serialisation), the segments of a depth are spread over 128
// threads, and — because a query is done four times sooner — the queries in flight at any time span a quarter of the clusters, so
// that their sketches and tables stay in the L2.
constexpr uint32_t kFsThreads = 128;

__global__ void __launch_bounds__(kFsThreads, 8) k_first_stream_cta(SearchParams p, QueryBatch b) {
    extern __shared__ __align__(16) uint8_t s_stream[];
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_sel[4];  // {bin, remaining rank, S, stop flag}
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = lane_id();
    const uint64_t w = blockIdx.x;
    if (w >= b.nq) return;
    const uint32_t q = b.qperm[w], c = b.first[w];
    uint32_t* meta = b.fs_meta + (uint64_t)q * kFsMeta;
    const uint32_t L = p.g.L;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    if (p.brute[c] || nc > 65536u || nc > b.dense_stride) {
        if (tid == 0) meta[48] = kMaxHashBits + 1;
        return;
    }
    StreamSmem sm;
    {
        uint8_t* ptr = s_stream;
        sm.lcp_up = reinterpret_cast<uint2*>(ptr); ptr += L * 8;
        sm.lcp_dn = reinterpret_cast<uint2*>(ptr); ptr += L * 8;
        sm.anchor = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.code = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.start = reinterpret_cast<uint32_t*>(ptr); ptr += L * 4;
        sm.segbase = reinterpret_cast<uint32_t*>(ptr); ptr += (L + 1) * 4;
        sm.hist = reinterpret_cast<uint32_t*>(ptr);  // [L] segments per table of the current depth
    }
    const uint32_t fsid = p.fset_of[c];
    const uint32_t* codes = b.codes + (uint64_t)fsid * L * b.nq + q;
    const uint64_t my_sketch = b.sketches[((uint64_t)fsid * b.nq + q) * kNumSketches + lane];
    const uint32_t* stop = p.stop + (uint64_t)fsid * kMaxHashBits * kEstBins * p.stop_words;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    // ---- anchors: one thread per table
    for (uint32_t t = tid; t < L; t += kFsThreads) {
        const uint32_t h = __ldg(codes + (uint64_t)t * b.nq);
        uint32_t A;
        uint2 up, dn;
        table_anchor(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc, h, A, up, dn);
        sm.code[t] = h;
        sm.anchor[t] = A;
        sm.lcp_up[t] = up;
        sm.lcp_dn[t] = dn;
    }
    // ---- k-th largest dense similarity of the cluster (two byte passes over a block-wide histogram)
    uint32_t kth16 = 0;
    if (p.k > 0 && nc >= p.k) {
        const uint32_t* v32 = reinterpret_cast<const uint32_t*>(b.dense + (uint64_t)q * b.dense_stride);
        uint32_t prefix = 0, want = p.k;
        for (int pass = 0; pass < 2; pass++) {
            for (uint32_t i = tid; i < 256; i += kFsThreads) s_hist[i] = 0;
            __syncthreads();
            for (uint32_t i = tid; i < (nc + 1) / 2; i += kFsThreads) {
                const uint32_t wv = __ldg(v32 + i);
                const uint32_t a = wv & 0xffffu, bb = wv >> 16;
                if (pass == 0) {
                    atomicAdd(&s_hist[a >> 8], 1u);
                    if (2 * i + 1 < nc) atomicAdd(&s_hist[bb >> 8], 1u);
                } else {
                    if ((a >> 8) == prefix) atomicAdd(&s_hist[a & 0xffu], 1u);
                    if (2 * i + 1 < nc && (bb >> 8) == prefix) atomicAdd(&s_hist[bb & 0xffu], 1u);
                }
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t mine = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) mine += s_hist[8 * lane + j];
                uint32_t suf = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_down_sync(0xffffffffu, suf, o);
                    if (lane + o < 32) suf += t;
                }
                const uint32_t above = suf - mine;
                if (above < want && want <= suf) {
                    uint32_t acc = above;
                    for (int j = 7; j >= 0; j--) {
                        const uint32_t h = s_hist[8 * lane + j];
                        if (acc + h >= want) {
                            s_sel[0] = 8 * lane + j;
                            s_sel[1] = want - acc;
                            break;
                        }
                        acc += h;
                    }
                }
            }
            __syncthreads();
            if (pass == 0) {
                prefix = s_sel[0];
                want = s_sel[1];
            } else {
                kth16 = (prefix << 8) | s_sel[0];
            }
            __syncthreads();
        }
    } else {
        __syncthreads();
    }
    uint32_t bin = (uint32_t)__fdiv_rn(__fdiv_rn((float)kth16, 65536.0f), 0.005f);
    bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;

    const uint64_t sbase = (uint64_t)q * b.fs_cap;
    uint2* out_idx = reinterpret_cast<uint2*>(b.fs_idx) + sbase;
    uint32_t* out_hd = b.fs_hd + sbase;
    uint8_t* out_tab = b.fs_tab + (sbase >> 5);
    uint32_t cum = 0, lowest = kMaxHashBits + 1;
    uint32_t my_S = 0, my_off = 0;  // thread d-1 keeps the record of depth d
    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        for (uint32_t t = tid; t < L; t += kFsThreads) {
            uint32_t nseg = 0;
            sm.start[t] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc,
                                      sm.code[t], sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg);
            sm.hist[t] = nseg;
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t running = 0;
            for (uint32_t t0 = 0; t0 < L; t0 += 32) {
                const uint32_t t = t0 + lane;
                const uint32_t nseg = t < L ? sm.hist[t] : 0u;
                uint32_t total;
                const uint32_t ex = warp_excl_scan(nseg, total);
                if (t < L) sm.segbase[t] = running + ex;
                running += total;
            }
            if (lane == 0) {
                sm.segbase[L] = running;
                s_sel[2] = running;
            }
        }
        __syncthreads();
        const uint32_t S = s_sel[2];
        const uint32_t S32 = (S + 31u) & ~31u;
        if (S > (uint32_t)kRing && cum + S32 > b.fs_cap) break;
        if (tid == depth - 1) {
            my_S = S;
            my_off = cum;
        }
        lowest = depth;
        if (S <= (uint32_t)kRing) continue;
        for (uint32_t s0 = 0; s0 < S; s0 += 2 * kFsThreads) {
            uint32_t v[2][4], tb[2];
            bool valid[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint32_t s = s0 + u * kFsThreads + tid;
                valid[u] = s < S;
                tb[u] = 0;
                v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0;
                if (valid[u]) {
                    uint32_t lo = 0, len = L;  // upper_bound(segbase, s) - 1 over segbase[0..L)
                    while (len > 0) {
                        const uint32_t half = len >> 1, mid = lo + half;
                        if (sm.segbase[mid] <= s) { lo = mid + 1; len -= half + 1; } else { len = half; }
                    }
                    const uint32_t t = lo - 1;
                    tb[u] = t;
                    const uint32_t* seg = p.tbl_idx + table_base(off, nc, L, t) + sm.start[t] + 4 * (s - sm.segbase[t]);
                    v[u][0] = __ldg(seg); v[u][1] = __ldg(seg + 1); v[u][2] = __ldg(seg + 2); v[u][3] = __ldg(seg + 3);
                }
            }
            uint64_t sw[2][4];
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int j = 0; j < 4; j++) sw[u][j] = valid[u] ? __ldg(sk + ((uint64_t)v[u][j] << 5 | lane)) : 0ull;
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!valid[u]) continue;
                const uint32_t s = s0 + u * kFsThreads + tid;
                const uint32_t h4 = (uint32_t)__popcll(sw[u][0] ^ my_sketch) | (uint32_t)__popcll(sw[u][1] ^ my_sketch) << 8 |
                                    (uint32_t)__popcll(sw[u][2] ^ my_sketch) << 16 | (uint32_t)__popcll(sw[u][3] ^ my_sketch) << 24;
                out_idx[cum + s] = make_uint2(v[u][0] | v[u][1] << 16, v[u][2] | v[u][3] << 16);
                out_hd[cum + s] = h4;
                if (lane == 0) out_tab[(cum + s) >> 5] = (uint8_t)tb[u];
            }
        }
        cum += S32;
        const uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (L >> 5));
        if ((word >> (L & 31)) & 1u) break;
        __syncthreads();  // sm.start of this depth is still being read by slower threads; the next depth overwrites it
    }
    if (tid < (uint32_t)kMaxHashBits) {
        meta[tid] = my_S;
        meta[kMaxHashBits + tid] = my_off;
    }
    if (tid == 0) meta[48] = lowest;
}

bool launch_first_stream(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0 || !b.fs_meta || !b.fs_idx || !b.fs_hd || !b.fs_tab || !b.dense || b.fs_cap < 64) return false;
    if (p.g.L > 255) return false;  // fs_tab holds the table index in a byte
    if (tune_get("first_stream", 0) == 2) {  // one CTA of four warps per query
        const size_t smem_cta = (size_t)p.g.L * (8 + 8 + 4 + 4 + 4 + 4 + 4) + 16;
        k_first_stream_cta<<<(unsigned)b.nq, kFsThreads, smem_cta, s>>>(p, b);
        return true;
    }
    const uint32_t wb = stream_smem_bytes(p.g.L);
    const uint32_t warps = 8;
    const size_t smem = (size_t)warps * wb;
    if (smem > 200 * 1024) return false;
    if (smem > 48 * 1024) ensure_dynamic_smem((const void*)(k_first_stream), smem);
    k_first_stream<<<(unsigned)((b.nq + warps - 1) / warps), warps * 32, smem, s>>>(p, b, wb);
    return true;
}

// ------------------------------------------------------------------------------------------------ dense first-visit similarities

constexpr uint32_t kDenseRows = 128;   // rows of a cluster staged per CTA
constexpr uint32_t kDenseQueries = 32;  // queries of the cluster staged per pass
constexpr int kDenseNQ = 4;             // queries multiplied against one row block by a warp at a time

// One CTA = (cluster c, tile of kDenseRows rows). The tile is staged in shared memory once (row pitch padded by 16 bytes:
// conflict-free when every lane reads its own row) and the queries whose nearest cluster is c — a contiguous run of the sorted
// work order — are staged as int32, kDenseQueries at a time. A work item = (block of 32 rows, group of kDenseNQ queries): lane =
// row, the row's 16-byte chunk is loaded and unpacked once and multiplied against the kDenseNQ queries (broadcast reads), so
// the multiply-round-accumulate of math.hpp:37-44 (3 instructions) is nearly all that is issued: 12.4 warp instructions per
// (query, row) against 33 in the gather path of the probe kernel. Results go out as 64-byte stores of 32 similarities.
__global__ void __launch_bounds__(256, 4) k_dense_sims(SearchParams p, QueryBatch b) {
    extern __shared__ __align__(16) uint8_t s_dense[];
    const uint32_t c = blockIdx.x, sl = p.g.sl, cpr = sl / 8, pitch = sl * 2 + 16;
    if (p.brute[c]) return;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t r0 = blockIdx.y * kDenseRows;
    if (r0 >= nc || nc > b.dense_stride) return;
    // queries whose work-order key is c: [lo, hi) in the sorted key array
    uint32_t lo = 0, len = (uint32_t)b.nq;
    while (len > 0) {
        uint32_t half = len >> 1;
        if (b.first[lo + half] < c) { lo += half + 1; len -= half + 1; } else { len = half; }
    }
    uint32_t hi = lo;
    len = (uint32_t)b.nq - lo;
    while (len > 0) {
        uint32_t half = len >> 1;
        if (b.first[hi + half] <= c) { hi += half + 1; len -= half + 1; } else { len = half; }
    }
    if (hi == lo) return;
    uint8_t* s_tile = s_dense;                                                   // [kDenseRows][pitch]
    int* s_q = reinterpret_cast<int*>(s_dense + (size_t)kDenseRows * pitch);     // [kDenseQueries][sl]
    const uint32_t rows_here = nc - r0 < kDenseRows ? nc - r0 : kDenseRows;
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.q15 + (off + r0) * sl);
        for (uint32_t i = threadIdx.x; i < kDenseRows * cpr; i += blockDim.x) {
            const uint32_t r = i / cpr, ch = i % cpr;
            const uint4 v = r < rows_here ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(s_tile + (size_t)r * pitch + ch * 16) = v;
        }
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t nblk = (rows_here + 31) / 32;
    for (uint32_t qbase = lo; qbase < hi; qbase += kDenseQueries) {
        const uint32_t nqh = hi - qbase < kDenseQueries ? hi - qbase : kDenseQueries;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nqh * (sl / 2); i += blockDim.x) {
            const uint32_t j = i / (sl / 2), e = i % (sl / 2);
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(b.q15 + (uint64_t)b.qperm[qbase + j] * sl) + e);
            s_q[j * sl + 2 * e] = unpack_lo(w);
            s_q[j * sl + 2 * e + 1] = unpack_hi(w);
        }
        __syncthreads();
        const uint32_t ngrp = (nqh + kDenseNQ - 1) / kDenseNQ;
        for (uint32_t item = warp; item < nblk * ngrp; item += blockDim.x >> 5) {
            const uint32_t blk = item % nblk, qg = item / nblk;
            const uint4* row = reinterpret_cast<const uint4*>(s_tile + (size_t)(blk * 32 + lane) * pitch);
            const int4* qp[kDenseNQ];
            int acc[kDenseNQ];
#pragma unroll
            for (int j = 0; j < kDenseNQ; j++) {
                const uint32_t qi = qg * kDenseNQ + j < nqh ? qg * kDenseNQ + j : nqh - 1;
                qp[j] = reinterpret_cast<const int4*>(s_q + qi * sl);
                acc[j] = 0;
            }
            for (uint32_t ch = 0; ch < cpr; ch++) {
                const uint4 v = row[ch];
                const int a0 = unpack_lo(v.x), a1 = unpack_hi(v.x), a2 = unpack_lo(v.y), a3 = unpack_hi(v.y);
                const int a4 = unpack_lo(v.z), a5 = unpack_hi(v.z), a6 = unpack_lo(v.w), a7 = unpack_hi(v.w);
#pragma unroll
                for (int j = 0; j < kDenseNQ; j++) {
                    const int4 x = qp[j][2 * ch], y = qp[j][2 * ch + 1];
                    int s = acc[j];
                    s += q15_mul(a0, x.x); s += q15_mul(a1, x.y); s += q15_mul(a2, x.z); s += q15_mul(a3, x.w);
                    s += q15_mul(a4, y.x); s += q15_mul(a5, y.y); s += q15_mul(a6, y.z); s += q15_mul(a7, y.w);
                    acc[j] = s;
                }
            }
            const uint32_t r = blk * 32 + lane;
            if (r < rows_here) {
#pragma unroll
                for (int j = 0; j < kDenseNQ; j++) {
                    const uint32_t qi = qg * kDenseNQ + j;
                    if (qi < nqh) b.dense[(uint64_t)b.qperm[qbase + qi] * b.dense_stride + r0 + r] = (uint16_t)(acc[j] + 32768);
                }
            }
        }
    }
}

bool dense_sims_supported(const SearchParams& p) { return p.g.sl / 8 <= 32 && tune_get("dense_sims", 1) != 0; }

bool launch_dense_sims(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0 || !b.dense || !dense_sims_supported(p)) return false;
    if (tune_get("order_longest_first", 0) != 0) return false;  // the sorted keys are not plain cluster ids then
    const size_t smem = (size_t)kDenseRows * (p.g.sl * 2 + 16) + (size_t)kDenseQueries * p.g.sl * sizeof(int);
    dim3 grid(p.K, (p.max_cluster + kDenseRows - 1) / kDenseRows);
    if (grid.y == 0) return false;
    if (smem > 48 * 1024) ensure_dynamic_smem((const void*)(k_dense_sims), smem);
    k_dense_sims<<<grid, 256, smem, s>>>(p, b);
    return true;
}

// ------------------------------------------------------------------------------------------------ launchers

void launch_prep_queries(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0) return;
    k_prep_queries<<<(unsigned)((b.nq + 127) / 128), 128, 0, s>>>(b.queries, b.nq, p.g.d, p.g.sl, b.q15, b.qnorm);
}

void launch_center_order(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0) return;
    if (b.exact_limit && center_gemm_tc_supported(p.g.d) && tune_get("order_longest_first", 0) == 0) {
        // nq x K x d on the tensor pipe as a screen, then the exact evaluation of each query's nearest candidates
        launch_center_gemm_tc(b.queries, b.qnorm, b.nq, p.center_rows, p.center_norms, p.K, p.g.d, b.cdist, s);
        k_center_refine<<<(unsigned)((b.nq + 7) / 8), 256, 0, s>>>(b.queries, b.qnorm, b.nq, p.center_rows, p.center_norms, p.K, p.g.d,
                                                                    b.cdist, b.exact_limit, b.first);
        return;
    }
    const uint32_t stride = p.g.d | 1;
    size_t smem = (size_t)96 * stride * sizeof(float);
    if (p.g.d <= 256) {
        const size_t tsmem = ((size_t)32 * ((p.g.d + 3) & ~3u) + (size_t)p.g.d * 65) * sizeof(float);
        if (tsmem > 48 * 1024) ensure_dynamic_smem((const void*)(k_center_dist_tiled), tsmem);
        const uint32_t nchunks = (p.K + 63) / 64;
        dim3 tgrid((unsigned)((b.nq + 31) / 32), nchunks < 8 ? nchunks : 8);
        k_center_dist_tiled<<<tgrid, 256, tsmem, s>>>(b.queries, b.qnorm, b.nq, p.center_rows, p.center_norms,
                                                                             p.K, p.g.d, b.cdist);
    } else if (smem <= 160 * 1024) {
        if (smem > 48 * 1024) ensure_dynamic_smem((const void*)(k_center_dist), smem);
        k_center_dist<<<(unsigned)((b.nq + 31) / 32), 256, smem, s>>>(b.queries, b.qnorm, b.nq, p.center_rows, p.center_norms, p.K,
                                                                     p.g.d, b.cdist);
    } else {
        k_center_dist_simple<<<(unsigned)((b.nq * p.K + 255) / 256), 256, 0, s>>>(b.queries, b.qnorm, b.nq, p.center_rows,
                                                                                 p.center_norms, p.K, p.g.d, b.cdist);
    }
    k_first_cluster<<<(unsigned)((b.nq * 32 + 255) / 256), 256, 0, s>>>(b.cdist, p.radii, b.nq, p.K,
                                                                        tune_get("order_longest_first", 0) != 0 ? 1 : 0, b.first);
}

void launch_init_state(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    uint64_t threads = b.nq > 0 ? b.nq : 1;
    k_init_state<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(b.state, b.nq, query_state_bytes(p.k), b.work_counter);
}

static int rerank_group(uint32_t sl) {
    uint32_t cpr = sl / 8;
    int g = 2;
    while ((uint32_t)g < cpr && g < 32) g <<= 1;
    return g;
}

template <int G, int OCC, bool DENSE, bool STREAM = false>
static void launch_probe_go(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s) {
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev;
        CLANN_CUDA(cudaGetDevice(&dev));
        CLANN_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const uint32_t wb = warp_smem_bytes(p.g.L, p.k);
    uint32_t per_warp = wb + p.g.sl * 2;
    // warps per CTA: as many as keep OCC CTAs of shared memory on one SM, at most 8 (knob: probe_warps)
    uint32_t warps = (uint32_t)tune_get("probe_warps", 8);
    if (warps < 1 || warps > 8) warps = 8;
    while (warps > 1 && (size_t)warps * per_warp * OCC > 200 * 1024) warps >>= 1;
    // DENSE: the memo of a visit in shared memory when the largest cluster fits beside the rest (knob probe_smem_memo)
    uint32_t smem_memo_cap = 0;
    if (DENSE && tune_get("probe_smem_memo", 1) != 0) {
        const uint32_t cap = (p.max_cluster + 7u) & ~7u;
        if (cap > 0 && (size_t)warps * (per_warp + cap * 2 + 16) * OCC <= 224 * 1024) {  // 228 KB per SM - 1 KB per CTA, some slack
            smem_memo_cap = cap;
            per_warp += cap * 2 + 16;
        }
    }
    size_t smem = (size_t)warps * per_warp;
    if (smem > 227 * 1024) throw std::invalid_argument("num_tables / k too large for the probe kernel's shared memory");
    ensure_dynamic_smem((const void*)(k_probe<G, OCC, DENSE, STREAM>), smem);
    int ctas_per_sm = 0;
    CLANN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_probe<G, OCC, DENSE, STREAM>, warps * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int cap = (int)tune_get("probe_ctas", 0);  // knob: fewer resident queries = smaller L2 working set
    if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = cap;
    uint64_t want = (b.nq + warps - 1) / warps;
    uint64_t grid = (uint64_t)sm_count * ctas_per_sm;  // persistent: a whole number of CTAs per SM
    if (want < grid) grid = want ? want : 1;
    // per-warp similarity memo (u16 per local id of the cluster being probed), from the index workspace
    const bool no_memo = tune_get("probe_nomemo", 0) != 0;  // A/B knob
    uint16_t* use = (!no_memo && b.memo && (uint64_t)grid * warps <= b.memo_slots) ? b.memo : nullptr;
    const uint64_t stride = b.memo_stride;
    SearchParams pp = p;
    pp.prefetch_rows = (uint32_t)tune_get("probe_prefetch_rows", 0);  // A/B knob (measured: 3.22 vs 3.12 ms, off by default)
    // knob shard_reserve_sms: SMs the probe of a sharded search leaves to the other batches in flight (their NCCL kernels need whole SMs)
    if (stop_at_foreign) {
        const int64_t r = tune_get("shard_reserve_sms", 0);
        pp.reserve_sms = (r > 0 && r < sm_count / 2) ? (uint32_t)r : 0u;
    }
    k_probe<G, OCC, DENSE, STREAM><<<(unsigned)grid, warps * 32, smem, s>>>(pp, b, wb, stop_at_foreign, use, stride, smem_memo_cap);
}

template <int G>
static void launch_probe_g(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s) {
    // With the first visit's similarities computed in advance (b.dense) the kernel is lighter and two CTAs per SM with 128
    // registers beat three with 80 (measured 2.82 vs 2.94 ms including the dense kernel); without, three (3.08 vs 3.36 ms).
    // dense first-visit similarities need the query's nearest cluster on this rank: always true unsharded, and in the first round of
    // the sharded search (queries are routed to the owner of their nearest cluster); the stepping protocol gives no such guarantee
    const bool dense = b.dense != nullptr && (!stop_at_foreign || (stop_at_foreign == 1 && b.first_is_own));
    int occ = (int)tune_get("probe_occ", 0);  // knob: resident CTAs per SM the kernel is compiled for (0 = the default above)
    if (occ < 2 || occ > 3) occ = dense ? 2 : 3;  // (four CTAs of 64 registers spill: 5.5 ms, instantiation dropped)
    if (dense && b.fs_meta) {  // the opt-in first-visit candidate stream has its own instantiation (it costs registers)
        launch_probe_go<G, 2, true, true>(p, b, stop_at_foreign, s);
    } else if (dense) {
        if (occ == 2) launch_probe_go<G, 2, true>(p, b, stop_at_foreign, s);
        else launch_probe_go<G, 3, true>(p, b, stop_at_foreign, s);
    } else {
        if (occ == 2) launch_probe_go<G, 2, false>(p, b, stop_at_foreign, s);
        else launch_probe_go<G, 3, false>(p, b, stop_at_foreign, s);
    }
}

void launch_probe(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s) {
    static int64_t fetch_set = 0;
    const int64_t fetch = tune_get("l2_fetch", 0);  // knob: cudaLimitMaxL2FetchGranularity in bytes (32/64/128), 0 = leave alone
    if (fetch != fetch_set && fetch > 0) {
        CLANN_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch));
        fetch_set = fetch;
    }
    launch_probe_warp(p, b, stop_at_foreign, s);
}

void launch_probe_warp(const SearchParams& p, const QueryBatch& b, int stop_at_foreign, cudaStream_t s) {
    if (b.nq == 0) return;
    switch (rerank_group(p.g.sl)) {
        case 2: launch_probe_g<2>(p, b, stop_at_foreign, s); break;
        case 4: launch_probe_g<4>(p, b, stop_at_foreign, s); break;
        case 8: launch_probe_g<8>(p, b, stop_at_foreign, s); break;
        case 16: launch_probe_g<16>(p, b, stop_at_foreign, s); break;
        default: launch_probe_g<32>(p, b, stop_at_foreign, s); break;
    }
}

uint32_t probe_memo_slots() {
    int dev = 0, sms = 0;
    CLANN_CUDA(cudaGetDevice(&dev));
    CLANN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    return (uint32_t)sms * 32u;  // at most 32 resident probe warps (warp kernel) or 8 CTAs (CTA kernel) per SM
}

void launch_merge_states(const SearchParams& p, const QueryBatch& b, const uint8_t* all_states, int world, uint32_t* active,
                         cudaStream_t s) {
    CLANN_CUDA(cudaMemsetAsync(active, 0, sizeof(uint32_t), s));
    uint64_t threads = b.nq > 0 ? b.nq : 1;
    k_merge_states<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(b.state, all_states, world, b.nq, query_state_bytes(p.k), active,
                                                                      b.work_counter);
}

void launch_finish(const SearchParams& p, const QueryBatch& b, cudaStream_t s) {
    if (b.nq == 0) return;
    uint64_t threads = b.nq * 32;
    k_finish<<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(b.state, b.nq, p.k, query_state_bytes(p.k), b.out_ids, b.out_dists,
                                                               b.out_counts, b.cnt_candidates, b.cnt_distcomp, b.cnt_visited, b.stats_dev,
                                                               b.stats_host);
}

template <int G>
static void launch_puffinn_g(const SearchParams& p, const QueryBatch& b, const uint32_t* stop, float max_sim, int filter_type,
                             uint32_t* out_ids, uint32_t* out_count, uint32_t* out_distcomp, uint32_t* out_stop, cudaStream_t s) {
    const uint32_t wb = warp_smem_bytes(p.g.L, p.k);
    size_t smem = (size_t)wb + p.g.sl * 2;
    if (smem > 227 * 1024) throw std::invalid_argument("num_tables / k too large for the probe kernel's shared memory");
    ensure_dynamic_smem((const void*)(k_puffinn_search<G>), smem);
    k_puffinn_search<G><<<1, 32, smem, s>>>(p, b, stop, max_sim, filter_type, wb, out_ids, out_count, out_distcomp, out_stop);
}

void launch_puffinn_search(const SearchParams& p, const QueryBatch& b, const uint32_t* stop_table, float max_sim, int filter_type,
                           uint32_t* out_ids, uint32_t* out_count, uint32_t* out_distcomp, uint32_t* out_stop, cudaStream_t s) {
    switch (rerank_group(p.g.sl)) {
        case 2: launch_puffinn_g<2>(p, b, stop_table, max_sim, filter_type, out_ids, out_count, out_distcomp, out_stop, s); break;
        case 4: launch_puffinn_g<4>(p, b, stop_table, max_sim, filter_type, out_ids, out_count, out_distcomp, out_stop, s); break;
        case 8: launch_puffinn_g<8>(p, b, stop_table, max_sim, filter_type, out_ids, out_count, out_distcomp, out_stop, s); break;
        case 16: launch_puffinn_g<16>(p, b, stop_table, max_sim, filter_type, out_ids, out_count, out_distcomp, out_stop, s); break;
        default: launch_puffinn_g<32>(p, b, stop_table, max_sim, filter_type, out_ids, out_count, out_distcomp, out_stop, s); break;
    }
}

}  // namespace clann
