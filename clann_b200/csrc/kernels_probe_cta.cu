// k_probe_cta — the CLANN search loop with ONE CTA (4 warps) PER QUERY.
//
// Same semantics, bit for bit, as the one-warp-per-query kernel in kernels_search.cu (which stays as the simple
// restatement and serves the legacy single-query ABI); what changes is how the work of one query is spread:
//   * 4x fewer queries are in flight for the same SM occupancy, and queries are scheduled in nearest-cluster order, so
//     the clusters being probed at any moment (rows + sketches + tables, ~3 MB each) stay resident in the 126 MB L2;
//   * the four warps evaluate four consecutive ring sweeps of search_maps concurrently (speculatively: sweeps beyond
//     the point where the 128-entry passing buffer fills are discarded and redone with the updated filter threshold,
//     exactly as the sequential reference would see them);
//   * the Q15 rerank of a batch is spread over all 128 threads: similarities already computed during this visit are
//     reused (the reference rescans nested ranges at every depth, so >50% of its distance computations are repeats —
//     the counter still counts them), missing rows are gathered by the TMA unit with one bulk asynchronous copy per
//     row (cp.async.bulk global -> shared, completion on an mbarrier) into padded shared memory and each thread then
//     reduces one whole row against the query without shuffles.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <stdlib.h>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

constexpr int kCtaWarps = 4;
constexpr int kCtaThreads = kCtaWarps * 32;
constexpr uint32_t kStageRows = 64;  // Q15 rows gathered per bulk-copy round
constexpr int kTabCap = 1024;  // coarse segment -> table map: one entry per 32 segments, first 32768 segments of a depth

struct CtaCtrl {
    unsigned long long nk, top;
    unsigned long long mbar;  // mbarrier of the bulk row copies
    uint32_t work, heap_len, inserted, minval16, max_diff, stopped, result_cnt, tail_cnt, unk_cnt;
    uint32_t spec_cnt[kCtaWarps];
    uint32_t warp_tot[kCtaWarps];
};

struct CtaSmem {
    CtaCtrl* ctrl;
    uint2* lcp_up;              // [L]
    uint2* lcp_dn;              // [L]
    unsigned long long* mb;     // [P2K]
    unsigned long long* heap;   // [k]
    unsigned long long* loc;    // [k]
    uint32_t* anchor;           // [L]
    uint32_t* code;             // [L]
    uint32_t* start;            // [L]
    uint32_t* segbase;          // [L+1]
    uint32_t* spec;             // [kCtaWarps][128] speculative passing lists (also scratch for brute-force clusters)
    uint32_t* pass_idx;         // [kPassingCap]
    int* qrow;                  // [sl] query, Q15 widened to int32
    uint16_t* pass_sim;         // [kPassingCap]
    uint16_t* unk;              // [kPassingCap] positions in pass_idx whose similarity is not memoised yet
    uint16_t* tab32;            // [kTabCap] table holding segment 32*i of the current depth
    uint8_t* stage;             // [stage_rows][stage_stride] gathered Q15 rows
};

__host__ __device__ inline uint32_t stage_stride_bytes(uint32_t sl) { return ((sl / 8) | 1u) * 16u; }  // odd number of 16-byte units

__host__ __device__ inline uint32_t cta_smem_bytes(uint32_t L, uint32_t k, uint32_t sl, uint32_t stage_rows) {
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint32_t b = 128;                                   // ctrl
    b += sl * 4;                                        // qrow (16-byte aligned: sl is a multiple of 16)
    b += stage_rows * stage_stride_bytes(sl);           // stage (16-byte aligned)
    b += L * 8 * 2 + p2k * 8 + k * 8 * 2;               // lcp_up, lcp_dn, mb, heap, loc
    b += L * 4 * 3 + (L + 1) * 4 + kCtaWarps * 128 * 4 + kPassingCap * 4;  // anchor, code, start, segbase, spec, pass_idx
    b += kPassingCap * 2 * 2 + kTabCap * 2;             // pass_sim, unk, tab32
    return (b + 15) & ~15u;
}

__device__ __forceinline__ CtaSmem carve_cta(uint8_t* base, uint32_t L, uint32_t k, uint32_t sl, uint32_t stage_rows) {
    CtaSmem s;
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint8_t* p = base;
    s.ctrl = reinterpret_cast<CtaCtrl*>(p); p += 128;
    s.qrow = reinterpret_cast<int*>(p); p += sl * 4;
    s.stage = p; p += stage_rows * stage_stride_bytes(sl);
    s.lcp_up = reinterpret_cast<uint2*>(p); p += L * 8;
    s.lcp_dn = reinterpret_cast<uint2*>(p); p += L * 8;
    s.mb = reinterpret_cast<unsigned long long*>(p); p += p2k * 8;
    s.heap = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    s.loc = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    s.anchor = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.code = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.start = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.segbase = reinterpret_cast<uint32_t*>(p); p += (L + 1) * 4;
    s.spec = reinterpret_cast<uint32_t*>(p); p += kCtaWarps * 128 * 4;
    s.pass_idx = reinterpret_cast<uint32_t*>(p); p += kPassingCap * 4;
    s.pass_sim = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    s.unk = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    s.tab32 = reinterpret_cast<uint16_t*>(p);
    return s;
}

// --- mbarrier + bulk asynchronous copy (TMA unit, SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Q15 similarities (dot + 32768) of pass_idx[0..np) -> pass_sim, by the whole CTA (collection.hpp:909-920,
// cosine.hpp:19-23, math.hpp:11-44). memo[local id] caches the similarities of this visit (0 = not yet known).
// `phase` is the running parity of the row-copy mbarrier (uniform across the CTA).
__device__ __forceinline__ void rerank_cta(const CtaSmem& sm, uint32_t np, const int16_t* __restrict__ rows, uint32_t sl,
                                           uint16_t* __restrict__ memo, uint32_t stage_rows, uint32_t& phase) {
    const uint32_t tid = threadIdx.x;
    const uint32_t cpr = sl / 8;
    const uint32_t sstride = stage_stride_bytes(sl);
    // (ctrl->unk_cnt was reset before the barrier that precedes this call)
    for (uint32_t i = tid; i < np; i += kCtaThreads) {
        uint32_t m = memo ? memo[sm.pass_idx[i]] : 0u;
        if (m) {
            sm.pass_sim[i] = (uint16_t)m;
        } else {
            uint32_t pos = atomicAdd(&sm.ctrl->unk_cnt, 1u);
            sm.unk[pos] = (uint16_t)i;
        }
    }
    __syncthreads();
    const uint32_t nunk = sm.ctrl->unk_cnt;
    for (uint32_t cb = 0; cb < nunk; cb += stage_rows) {
        const uint32_t nrows = nunk - cb < stage_rows ? nunk - cb : stage_rows;
        // gather: one bulk copy per row, issued by the thread that will reduce it
        if (tid == 0) mbar_arrive_expect_tx(&sm.ctrl->mbar, nrows * sl * 2);
        for (uint32_t r = tid; r < nrows; r += kCtaThreads) {
            const uint32_t id = sm.pass_idx[sm.unk[cb + r]];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the slot vs the async write
            bulk_copy_g2s(sm.stage + r * sstride, rows + (uint64_t)id * sl, sl * 2, &sm.ctrl->mbar);
        }
        mbar_wait(&sm.ctrl->mbar, phase);
        phase ^= 1u;
        for (uint32_t r = tid; r < nrows; r += kCtaThreads) {
            const uint4* row = reinterpret_cast<const uint4*>(sm.stage + r * sstride);
            const int4* qv = reinterpret_cast<const int4*>(sm.qrow);
            int s = 0;
            for (uint32_t chn = 0; chn < cpr; chn++) {
                uint4 w = row[chn];
                int4 a = qv[2 * chn], b = qv[2 * chn + 1];
                s += q15_mul(unpack_lo(w.x), a.x); s += q15_mul(unpack_hi(w.x), a.y);
                s += q15_mul(unpack_lo(w.y), a.z); s += q15_mul(unpack_hi(w.y), a.w);
                s += q15_mul(unpack_lo(w.z), b.x); s += q15_mul(unpack_hi(w.z), b.y);
                s += q15_mul(unpack_lo(w.w), b.z); s += q15_mul(unpack_hi(w.w), b.w);
            }
            const uint16_t sim16 = (uint16_t)(s + 32768);
            const uint32_t i = sm.unk[cb + r];
            sm.pass_sim[i] = sim16;
            if (memo) memo[sm.pass_idx[i]] = sim16;
        }
        __syncthreads();
    }
}

// One PUFFINN query against cluster c by the whole CTA (collection.hpp:543-601 -> search_maps :768-948).
// Returns the number of results; sm.mb[0..cnt) holds them best first.
__device__ uint32_t probe_cluster_cta(const SearchParams& p, const CtaSmem& sm, uint32_t c, const uint32_t* __restrict__ codes,
                                      uint64_t code_stride, const uint64_t* __restrict__ qsketch, const uint32_t* __restrict__ stop,
                                      float max_sim, uint16_t* __restrict__ memo, uint32_t stage_rows, uint32_t& phase,
                                      ProbeCounters& ctr) {
    const uint32_t L = p.g.L, k = p.k;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    const int16_t* rows = p.q15 + off * p.g.sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    CtaCtrl* ctrl = sm.ctrl;

    if (memo) {
        uint4* mz = reinterpret_cast<uint4*>(memo);
        for (uint32_t i = tid; i < (nc + 7) / 8; i += kCtaThreads) mz[i] = make_uint4(0, 0, 0, 0);
    }
    // --- SearchBuffers ctor (collection.hpp:642-645): one thread per table
    for (uint32_t t = tid; t < L; t += kCtaThreads) {
        const uint32_t h = codes[(uint64_t)t * code_stride];
        const uint32_t* H = p.tbl_hash + (uint64_t)t * p.n + off;
        uint32_t lo = 0, len = nc;
        while (len > 0) {
            uint32_t half = len >> 1, mid = lo + half;
            if (__ldg(H + mid) < h) { lo = mid + 1; len -= half + 1; } else { len = half; }
        }
        sm.code[t] = h;
        sm.anchor[t] = lo;
        uint32_t up[8], dn[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint32_t pu = lo + kSegment * j;
            up[j] = pu < nc ? lcp24(__ldg(H + pu), h) : 0u;
            int64_t pd = (int64_t)lo - 1 - kSegment * j;
            dn[j] = pd >= 0 ? lcp24(__ldg(H + pd), h) : 0u;
        }
        sm.lcp_up[t] = make_uint2(up[0] | up[1] << 8 | up[2] << 16 | up[3] << 24, up[4] | up[5] << 8 | up[6] << 16 | up[7] << 24);
        sm.lcp_dn[t] = make_uint2(dn[0] | dn[1] << 8 | dn[2] << 16 | dn[3] << 24, dn[4] | dn[5] << 8 | dn[6] << 16 | dn[7] << 24);
    }
    const uint64_t my_sketch = __ldg(qsketch + lane);  // ring slot == lane, in every warp
    if (tid == 0) {
        ctrl->inserted = 0;
        ctrl->minval16 = 0;
        ctrl->max_diff = kSketchBits;  // filterer.hpp:101
        ctrl->stopped = 0;
    }
    __syncthreads();

    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        if (ctrl->stopped) break;
        // --- fill_ranges (collection.hpp:650-667) with get_next_range (prefixmap.hpp:267-304) in closed form
        const uint32_t it = kMaxHashBits + 1 - depth;
        const uint32_t dir_bit = 1u << (it >= 2 ? it - 2 : 0);
        uint32_t running = 0;
        for (uint32_t t0 = 0; t0 < L; t0 += kCtaThreads) {
            const uint32_t t = t0 + tid;
            uint32_t nseg = 0;
            if (t < L) {
                const uint32_t h = sm.code[t];
                const uint32_t A = sm.anchor[t];
                const uint32_t* H = p.tbl_hash + (uint64_t)t * p.n + off;
                int64_t start, end;
                if ((h & dir_bit) == 0) {
                    uint32_t j = lead_count(sm.lcp_up[t], depth);
                    if (j == 8) {
                        uint32_t lo = A + 8 * kSegment, len = nc > lo ? nc - lo : 0;
                        while (len > 0) {
                            uint32_t half = len >> 1, mid = lo + half;
                            if (lcp24(__ldg(H + mid), h) >= depth) { lo = mid + 1; len -= half + 1; } else { len = half; }
                        }
                        j = (lo - A + kSegment - 1) / kSegment;
                    }
                    start = A;
                    end = (int64_t)A + (int64_t)kSegment * j;
                    if (end >= (int64_t)nc) end = (end - kSegment) > start ? (end - kSegment) : start;
                } else {
                    uint32_t j = lead_count(sm.lcp_dn[t], depth);
                    if (j == 8) {
                        uint32_t hi = A >= 8 * kSegment ? A - 8 * kSegment : 0;
                        uint32_t lo = 0, len = hi;
                        while (len > 0) {
                            uint32_t half = len >> 1, mid = lo + half;
                            if (lcp24(__ldg(H + mid), h) < depth) { lo = mid + 1; len -= half + 1; } else { len = half; }
                        }
                        j = (A - lo + kSegment - 1) / kSegment;
                    }
                    end = A;
                    start = (int64_t)A - (int64_t)kSegment * j;
                    if (start < 0) start = (start + kSegment) < end ? (start + kSegment) : end;
                }
                sm.start[t] = (uint32_t)start;
                nseg = (uint32_t)(end - start) >> 2;
            }
            uint32_t total;
            const uint32_t ex = warp_excl_scan(nseg, total);
            if (lane == 0) ctrl->warp_tot[warp] = total;
            __syncthreads();
            uint32_t prefix = running, tile_total = 0;
#pragma unroll
            for (int w = 0; w < kCtaWarps; w++) {
                uint32_t wt = ctrl->warp_tot[w];
                if ((uint32_t)w < warp) prefix += wt;
                tile_total += wt;
            }
            if (t < L) sm.segbase[t] = prefix + ex;
            running += tile_total;
            __syncthreads();
        }
        if (tid == 0) sm.segbase[L] = running;
        __syncthreads();
        const uint32_t S = running;
        if (S <= kRing) continue;  // collection.hpp:802-810

        auto locate = [&](uint32_t s) -> uint64_t {
            uint32_t lo = 0, len = L;
            while (len > 0) {
                uint32_t half = len >> 1, mid = lo + half;
                if (sm.segbase[mid] <= s) { lo = mid + 1; len -= half + 1; } else { len = half; }
            }
            uint32_t t = lo - 1;
            return (uint64_t)t * p.n + off + sm.start[t] + 4 * (s - sm.segbase[t]);
        };

        uint32_t base = 0;
        bool stop_now = false;
        do {
            uint32_t np = 0;
            // full ring sweeps (collection.hpp:813-866): warp w speculatively evaluates sweep number w of this round
            while (np < (uint32_t)kFilterBuffer && base + kRing <= S) {
                const uint32_t sb = base + kRing * warp;
                const bool valid = sb + kRing <= S;
                const uint32_t max_diff = ctrl->max_diff;
                uint32_t total = 0;
                if (valid) {
                    const uint32_t* seg = p.tbl_idx + locate(sb + lane);
                    uint32_t v0 = __ldg(seg), v1 = __ldg(seg + 1), v2 = __ldg(seg + 2), v3 = __ldg(seg + 3);
                    uint64_t s0 = __ldg(sk + ((uint64_t)v0 << 5 | lane)), s1 = __ldg(sk + ((uint64_t)v1 << 5 | lane));
                    uint64_t s2 = __ldg(sk + ((uint64_t)v2 << 5 | lane)), s3 = __ldg(sk + ((uint64_t)v3 << 5 | lane));
                    uint32_t p0 = (uint32_t)__popcll(s0 ^ my_sketch) <= max_diff, p1 = (uint32_t)__popcll(s1 ^ my_sketch) <= max_diff;
                    uint32_t p2 = (uint32_t)__popcll(s2 ^ my_sketch) <= max_diff, p3 = (uint32_t)__popcll(s3 ^ my_sketch) <= max_diff;
                    uint32_t cnt = p0 + p1 + p2 + p3;
                    uint32_t pos = warp_excl_scan(cnt, total);
                    uint32_t* out = sm.spec + warp * 128;
                    if (p0) out[pos++] = v0;
                    if (p1) out[pos++] = v1;
                    if (p2) out[pos++] = v2;
                    if (p3) out[pos++] = v3;
                }
                if (lane == 0) ctrl->spec_cnt[warp] = valid ? total : 0xffffffffu;
                __syncthreads();
                // in-order consumption: a sweep counts only while the buffer holds < 128 entries (collection.hpp:813)
                uint32_t consumed = 0, offs[kCtaWarps], cnts[kCtaWarps];
#pragma unroll
                for (int w = 0; w < kCtaWarps; w++) {
                    uint32_t sc = ctrl->spec_cnt[w];
                    bool take = (consumed == (uint32_t)w) && sc != 0xffffffffu && np < (uint32_t)kFilterBuffer;
                    offs[w] = np;
                    cnts[w] = take ? sc : 0;
                    if (take) {
                        np += sc;
                        consumed++;
                    }
                }
#pragma unroll
                for (int w = 0; w < kCtaWarps; w++)
                    for (uint32_t i = tid; i < cnts[w]; i += kCtaThreads) sm.pass_idx[offs[w] + i] = sm.spec[w * 128 + i];
                base += kRing * consumed;
                ctr.candidates += (unsigned long long)kRing * 4 * consumed;
                __syncthreads();
            }
            // tail (collection.hpp:869-903): ring slots not yet tested, descending slot order, index used as the sketch (:890-893)
            const uint32_t missing = (base + kRing > S) ? (base + kRing - S > (uint32_t)kRing ? (uint32_t)kRing : base + kRing - S) : 0;
            const uint32_t live = kRing - missing;
            if (warp == 0) {
                const uint32_t max_diff = ctrl->max_diff;
                uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0;
                if (lane < live) {
                    const uint32_t* seg = p.tbl_idx + locate(base + lane);
                    v0 = __ldg(seg); v1 = __ldg(seg + 1); v2 = __ldg(seg + 2); v3 = __ldg(seg + 3);
                    p0 = (uint32_t)__popcll((uint64_t)v0 ^ my_sketch) <= max_diff;
                    p1 = (uint32_t)__popcll((uint64_t)v1 ^ my_sketch) <= max_diff;
                    p2 = (uint32_t)__popcll((uint64_t)v2 ^ my_sketch) <= max_diff;
                    p3 = (uint32_t)__popcll((uint64_t)v3 ^ my_sketch) <= max_diff;
                }
                uint32_t cnt = p0 + p1 + p2 + p3, total;
                uint32_t ex = warp_excl_scan(cnt, total);
                uint32_t pos = np + (total - ex - cnt);
                if (p0) sm.pass_idx[pos++] = v0;
                if (p1) sm.pass_idx[pos++] = v1;
                if (p2) sm.pass_idx[pos++] = v2;
                if (p3) sm.pass_idx[pos++] = v3;
                if (lane == 0) ctrl->tail_cnt = total;
            }
            __syncthreads();
            np += ctrl->tail_cnt;
            ctr.candidates += 4ull * live;
            // empty the buffer (collection.hpp:909-925)
            if (tid == 0) ctrl->unk_cnt = 0;
            __syncthreads();
            rerank_cta(sm, np, rows, p.g.sl, memo, stage_rows, phase);
            ctr.distcomp += np;
            if (warp == 0) {
                uint32_t inserted = ctrl->inserted, minval16 = ctrl->minval16;
                maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, np);
                const uint32_t max_diff = p.msd[minval16 < 65536u ? minval16 : 65535u];  // filterer.hpp:108-111
                // stop rule (collection.hpp:927-943)
                uint32_t pulled = base + kRing, table_idx = L;
                if (pulled < S) {
                    uint32_t lo = 0, len = L;
                    while (len > 0) {
                        uint32_t half = len >> 1, mid = lo + half;
                        if (sm.segbase[mid] <= pulled) { lo = mid + 1; len -= half + 1; } else { len = half; }
                    }
                    table_idx = lo - 1;
                }
                float kth = __fdiv_rn((float)minval16, 65536.0f);
                float sim = kth < max_sim ? max_sim : kth;
                uint32_t bin = (uint32_t)__fdiv_rn(sim, 0.005f);
                bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
                uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (table_idx >> 5));
                if (lane == 0) {
                    ctrl->inserted = inserted;
                    ctrl->minval16 = minval16;
                    ctrl->max_diff = max_diff;
                    ctrl->stopped = (word >> (table_idx & 31)) & 1u;
                }
            }
            __syncthreads();
            stop_now = ctrl->stopped != 0;
        } while (!stop_now && base + kRing < S);
        __syncthreads();
    }
    if (warp == 0) {  // best_indices (collection.hpp:598, maxbuffer.hpp:79-96)
        uint32_t inserted = ctrl->inserted, minval16 = ctrl->minval16;
        maxbuffer_filter(sm.mb, P, k, inserted, minval16);
        if (lane == 0) ctrl->result_cnt = inserted;
    }
    __syncthreads();
    return ctrl->result_cnt;
}

template <int OCC>
__global__ void __launch_bounds__(kCtaThreads, OCC) k_probe_cta(SearchParams p, QueryBatch b, int stop_at_foreign,
                                                                uint16_t* __restrict__ memo_base, uint64_t memo_stride) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const CtaSmem sm = carve_cta(s_dyn, p.g.L, p.k, p.g.sl, kStageRows);
    CtaCtrl* ctrl = sm.ctrl;
    if (threadIdx.x == 0) mbar_init(&ctrl->mbar, 1);
    uint32_t phase = 0;
    __syncthreads();
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t state_bytes = sizeof(QueryStateHeader) + (uint64_t)p.k * 8;
    uint16_t* memo = memo_base ? memo_base + (uint64_t)blockIdx.x * memo_stride : nullptr;
    const uint32_t P = next_pow2(2 * p.k) < 32 ? 32 : next_pow2(2 * p.k);

    for (;;) {
        if (tid == 0) ctrl->work = atomicAdd(b.work_counter, 1u);
        __syncthreads();
        const uint32_t w = ctrl->work;
        __syncthreads();
        if (w >= b.nq) break;
        const uint32_t q = b.qperm[w];  // queries sorted by their nearest cluster
        QueryStateHeader* st = reinterpret_cast<QueryStateHeader*>(b.state + (uint64_t)q * state_bytes);
        if (st->done) continue;
        unsigned long long* st_heap = reinterpret_cast<unsigned long long*>(st + 1);
        uint32_t pos = st->next_pos;
        unsigned long long last_key = st->last_key;
        uint32_t visited = st->visited;
        ProbeCounters ctr{st->candidates, st->distcomp};
        const uint32_t heap_len0 = st->heap_len;
        for (uint32_t i = tid; i < heap_len0; i += kCtaThreads) sm.heap[i] = st_heap[i];
        for (uint32_t i = tid; i < p.g.sl; i += kCtaThreads) sm.qrow[i] = (int)b.q15[(uint64_t)q * p.g.sl + i];
        if (tid == 0) ctrl->heap_len = heap_len0;
        __syncthreads();
        const float* qv = b.queries + (uint64_t)q * p.g.d;
        const float qn = b.qnorm[q];
        const float* cd = b.cdist + (uint64_t)q * p.K;
        bool done = false;

        for (; pos < p.K; pos++) {
            // next cluster of the stable ascending centre-distance order (index.rs:592-616) and the current heap top
            if (warp == 0) {
                unsigned long long nk = ~0ull;
                for (uint32_t cc = lane; cc < p.K; cc += 32) {
                    unsigned long long key = ((unsigned long long)float_order_bits(cd[cc]) << 32) | cc;
                    if ((pos == 0 || key > last_key) && key < nk) nk = key;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    unsigned long long t = __shfl_xor_sync(0xffffffffu, nk, o);
                    nk = t < nk ? t : nk;
                }
                unsigned long long top = topk_peek(sm.heap, ctrl->heap_len);
                if (lane == 0) {
                    ctrl->nk = nk;
                    ctrl->top = top;
                }
            }
            __syncthreads();
            const unsigned long long nk = ctrl->nk;
            const uint32_t c = (uint32_t)nk;
            float max_dist = INFINITY;
            if (ctrl->heap_len > 0) {  // index.rs:342-361
                max_dist = float_from_order_bits((uint32_t)(ctrl->top >> 32));
                float cmin = __fsub_rn(float_from_order_bits((uint32_t)(nk >> 32)), p.radii[c]);
                if (cmin > max_dist) {
                    done = true;
                    break;
                }
            }
            if (stop_at_foreign && p.owner[c] != p.shard_rank) break;
            last_key = nk;
            visited++;
            const uint64_t off = p.offsets[c];
            const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
            __syncthreads();  // everyone has read ctrl->nk / top before warp 0 may overwrite them
            if (p.brute[c]) {
                // index.rs:364-378 with brute_force_search :666-685: members in assignment order into a local top-k, then merge
                float* s_dist = reinterpret_cast<float*>(sm.spec);
                uint32_t* s_pid = sm.spec + kCtaThreads;
                uint32_t loc_len = 0;
                for (uint32_t base = 0; base < nc; base += kCtaThreads) {
                    const uint32_t j = base + tid;
                    if (j < nc) {
                        uint32_t pid = p.perm[off + j];
                        s_pid[tid] = pid;
                        s_dist[tid] = distance_point(p.data + (uint64_t)pid * p.g.d, p.norms[pid], qv, qn, p.g.d);
                    }
                    __syncthreads();
                    if (warp == 0) {
                        const uint32_t lim = nc - base < (uint32_t)kCtaThreads ? nc - base : (uint32_t)kCtaThreads;
                        for (uint32_t l = 0; l < lim; l++) topk_add(sm.loc, loc_len, p.k, s_dist[l], s_pid[l]);
                    }
                    __syncthreads();
                }
                if (warp == 0) {
                    for (uint32_t i = lane; i < P; i += 32) sm.mb[i] = i < loc_len ? ~sm.loc[i] : 0ull;
                    __syncwarp();
                    warp_sort_desc(sm.mb, P);  // to_list (heap.rs:42-48): ascending by distance
                    uint32_t hl = ctrl->heap_len;
                    for (uint32_t i = 0; i < loc_len; i++) {
                        unsigned long long key = ~sm.mb[i];
                        topk_add(sm.heap, hl, p.k, float_from_order_bits((uint32_t)(key >> 32)), (uint32_t)key);
                    }
                    if (lane == 0) ctrl->heap_len = hl;
                }
            } else {
                const uint32_t fs = p.fset_of[c];
                const float max_sim = __fsub_rn(1.0f, __fdiv_rn(max_dist, 2.0f));  // puffinn_types.rs:77-79
                const uint32_t* codes = b.codes + (uint64_t)fs * p.g.L * b.nq + q;
                const uint64_t* qsk = b.sketches + ((uint64_t)fs * b.nq + q) * kNumSketches;
                const uint32_t* stop = p.stop + (uint64_t)fs * kMaxHashBits * kEstBins * p.stop_words;
                uint16_t* use_memo = (memo && nc <= memo_stride) ? memo : nullptr;
                const uint32_t cnt = probe_cluster_cta(p, sm, c, codes, b.nq, qsk, stop, max_sim, use_memo, kStageRows, phase, ctr);
                if (warp == 0) {  // map_candidates + fp32 distance + heap (index.rs:392-416), best first
                    uint32_t hl = ctrl->heap_len;
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
                            topk_add(sm.heap, hl, p.k, dl, il);
                        }
                    }
                    if (lane == 0) ctrl->heap_len = hl;
                }
            }
            __syncthreads();
        }
        if (pos >= p.K) done = true;
        __syncthreads();
        const uint32_t heap_len = ctrl->heap_len;
        for (uint32_t i = tid; i < heap_len; i += kCtaThreads) st_heap[i] = sm.heap[i];
        if (tid == 0) {
            st->heap_len = heap_len;
            st->next_pos = pos;
            st->last_key = last_key;
            st->visited = visited;
            st->done = done ? 1u : 0u;
            st->candidates = ctr.candidates;
            st->distcomp = ctr.distcomp;
        }
        __syncthreads();
    }
}

template <int OCC>
static void launch_probe_cta_occ(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev;
        CLANN_CUDA(cudaGetDevice(&dev));
        CLANN_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const size_t smem = cta_smem_bytes(p.g.L, p.k, p.g.sl, kStageRows);
    if (smem > 227 * 1024) throw std::invalid_argument("num_tables / k / dimension too large for the probe kernel's shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        CLANN_CUDA(cudaFuncSetAttribute(k_probe_cta<OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int ctas_per_sm = 0;
    CLANN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_probe_cta<OCC>, kCtaThreads, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    uint64_t grid = (uint64_t)sm_count * ctas_per_sm;  // persistent: a whole number of CTAs per SM
    if (b.nq < grid) grid = b.nq;
    // per-CTA similarity memo (u16 per local id of the cluster being probed); skipped when it would not fit in ~1 GiB
    static uint16_t* memo = nullptr;
    static size_t memo_cap = 0;
    uint64_t stride = ((uint64_t)p.max_cluster + 7) & ~7ull;
    size_t need = (size_t)grid * stride;
    uint16_t* use = nullptr;
    if (stride > 0 && need * sizeof(uint16_t) <= ((size_t)1 << 30)) {
        if (need > memo_cap) {
            if (memo) CLANN_CUDA(cudaFree(memo));
            CLANN_CUDA(cudaMalloc(&memo, need * sizeof(uint16_t)));
            memo_cap = need;
        }
        use = memo;
    }
    k_probe_cta<OCC><<<(unsigned)grid, kCtaThreads, smem, s>>>(p, b, stop_at_foreign ? 1 : 0, use, stride);
}

void launch_probe_cta(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    if (b.nq == 0) return;
    static int occ = 0;
    if (occ == 0) {
        const char* e = getenv("CLANN_PROBE_OCC");  // tuning knob: resident CTAs per SM the kernel is compiled for
        occ = e ? atoi(e) : 5;
        if (occ != 4 && occ != 5 && occ != 6 && occ != 8) occ = 5;
    }
    switch (occ) {
        case 4: launch_probe_cta_occ<4>(p, b, stop_at_foreign, s); break;
        case 6: launch_probe_cta_occ<6>(p, b, stop_at_foreign, s); break;
        case 8: launch_probe_cta_occ<8>(p, b, stop_at_foreign, s); break;
        default: launch_probe_cta_occ<5>(p, b, stop_at_foreign, s); break;
    }
}

}  // namespace clann
