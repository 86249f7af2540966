// k_probe_cta — the CLANN search loop with ONE CTA (4 warps) PER QUERY, organised so that no global-memory latency sits
// on the sequential part of puffinn::Index::search_maps.
//
// Same results, bit for bit, as the one-warp-per-query kernel in kernels_search.cu (which stays as the plain
// restatement and serves the legacy single-query ABI). What changes is the schedule of one (query, cluster) visit:
//
//   per depth, the segment stream is consumed in CHUNKS of 512 segments (16 ring sweeps, 2048 candidates):
//   phase A  all 128 threads: table indices of the chunk, the sketch word of every candidate (ring slot = segment
//            number mod 32) -> its Hamming distance to the query sketch, one byte per candidate in shared memory.
//            The distance is kept instead of a pass/fail bit because the filter threshold only ever tightens during a
//            visit (filterer.hpp:108-111 is monotone in the k-th similarity), so any later threshold can be applied
//            without touching global memory again.
//   phase B  Q15 similarity of every candidate that passes the threshold in force at the start of the chunk. A per-CTA
//            memo (one u16 per local id) returns similarities already computed during this visit — the reference
//            rescans nested ranges at every depth, so more than half of its distance computations are repeats (the
//            counter still counts them). Missing rows are gathered by the TMA unit, one bulk asynchronous copy per row
//            (cp.async.bulk global -> shared, completion on an mbarrier) into bank-conflict-free padded slots, and two
//            threads reduce each row against the query.
//   phase C  warp 0 replays the reference's sequential loop over the chunk out of shared memory only: ring sweeps in
//            order (lane = ring slot), the 128-entry passing buffer, the index-as-sketch tail (collection.hpp:890-893),
//            MaxBuffer inserts, threshold update, stop rule. A batch may straddle chunks; the stop rule may end the
//            visit in the middle of a chunk (the rest of the chunk was speculative work).
//
// Queries are scheduled in nearest-cluster order and only (SMs x CTAs/SM) of them are in flight, so the clusters being
// probed at any moment (rows + sketches + tables, ~3 MB each at the glove-100 shape) stay resident in the 126 MB L2.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <stdlib.h>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

constexpr int kCtaWarps = 4;
constexpr int kCtaThreads = kCtaWarps * 32;
constexpr uint32_t kSegPerThread = 4;                              // consecutive segments handled by one thread in phase A
constexpr uint32_t kChunkSegs = kCtaThreads * kSegPerThread;       // 512 segments = 16 ring sweeps per chunk
constexpr uint32_t kChunkCand = kChunkSegs * 4;                    // 2048 candidates
constexpr uint32_t kStageRows = 64;                                // Q15 rows gathered per bulk-copy round (2 threads per row)
constexpr uint32_t kNeedMore = 1u;

struct CtaCtrl {
    unsigned long long nk, top;
    unsigned long long mbar;  // mbarrier of the bulk row copies
    unsigned long long candidates, distcomp;  // performance.hpp:72-86, running totals of the query
    uint32_t work, heap_len, result_cnt;
    uint32_t inserted, minval16, max_diff, stopped;  // search_maps state carried between chunks
    uint32_t base, np, status;
    uint32_t unk_cnt, tail_unk;
    uint32_t warp_tot[kCtaWarps];
};
static_assert(sizeof(CtaCtrl) <= 128, "CtaCtrl must fit its 128-byte slot");

struct CtaSmem {
    CtaCtrl* ctrl;
    uint64_t* qsk;              // [32] query sketches of the function set in use
    int* qrow2;                 // [sl] query in Q15, doubled (see q15_mul_hi)
    uint8_t* stage;             // [kStageRows][stage_stride] gathered Q15 rows
    uint32_t* cid;              // [kChunkCand] local ids of the chunk's candidates, stream order
    uint16_t* csim;             // [kChunkCand] similarity (dot + 32768) of the candidates that pass the chunk threshold
    uint16_t* unk;              // [kChunkCand] chunk positions whose similarity is not memoised yet
    uint32_t* cpc;              // [kChunkCand / 4] Hamming distances, one byte per candidate
    uint2* lcp_up;              // [L]
    uint2* lcp_dn;              // [L]
    unsigned long long* mb;     // [P2K] MaxBuffer slots
    unsigned long long* heap;   // [k]
    unsigned long long* loc;    // [k]
    uint32_t* anchor;           // [L]
    uint32_t* code;             // [L]
    uint32_t* start;            // [L]
    uint32_t* segbase;          // [L+1]
    uint32_t* pass_idx;         // [kPassingCap]
    uint16_t* pass_sim;         // [kPassingCap]
};

// Row slots are padded to a number of 16-byte units that is 2 modulo 8: two threads per row reading alternate units
// then hit eight distinct bank groups per quarter warp.
__host__ __device__ inline uint32_t stage_stride_units(uint32_t sl) {
    uint32_t u = sl / 8;
    while ((u & 7u) != 2u) u++;
    return u;
}

__host__ __device__ inline uint32_t cta_smem_bytes(uint32_t L, uint32_t k, uint32_t sl) {
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint32_t b = 128 + 32 * 8;                            // ctrl, qsk
    b += sl * 4;                                          // qrow2
    b += kStageRows * stage_stride_units(sl) * 16;        // stage
    b += kChunkCand * 4;                                  // cid
    b += L * 8 * 2 + p2k * 8 + k * 8 * 2;                 // lcp_up, lcp_dn, mb, heap, loc
    b += L * 4 * 3 + (L + 1) * 4 + kPassingCap * 4;       // anchor, code, start, segbase, pass_idx
    b += kChunkCand;                                      // cpc
    b += kChunkCand * 2 * 2 + kPassingCap * 2;            // csim, unk, pass_sim
    return (b + 15) & ~15u;
}

__device__ __forceinline__ CtaSmem carve_cta(uint8_t* base, uint32_t L, uint32_t k, uint32_t sl) {
    CtaSmem s;
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint8_t* p = base;  // 16-byte aligned blocks first, then 8-, 4- and 2-byte arrays
    s.ctrl = reinterpret_cast<CtaCtrl*>(p); p += 128;
    s.qsk = reinterpret_cast<uint64_t*>(p); p += 32 * 8;
    s.qrow2 = reinterpret_cast<int*>(p); p += sl * 4;
    s.stage = p; p += kStageRows * stage_stride_units(sl) * 16;
    s.cid = reinterpret_cast<uint32_t*>(p); p += kChunkCand * 4;
    s.csim = reinterpret_cast<uint16_t*>(p); p += kChunkCand * 2;
    s.unk = reinterpret_cast<uint16_t*>(p); p += kChunkCand * 2;
    s.cpc = reinterpret_cast<uint32_t*>(p); p += kChunkCand;
    s.lcp_up = reinterpret_cast<uint2*>(p); p += L * 8;
    s.lcp_dn = reinterpret_cast<uint2*>(p); p += L * 8;
    s.mb = reinterpret_cast<unsigned long long*>(p); p += p2k * 8;
    s.heap = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    s.loc = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    s.anchor = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.code = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.start = reinterpret_cast<uint32_t*>(p); p += L * 4;
    s.segbase = reinterpret_cast<uint32_t*>(p); p += (L + 1) * 4;
    s.pass_idx = reinterpret_cast<uint32_t*>(p); p += kPassingCap * 4;
    s.pass_sim = reinterpret_cast<uint16_t*>(p);
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

__device__ __forceinline__ uint32_t ldg_nc_na_u32(const uint32_t* p) {
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint64_t ldg_nc_na_u64(const uint64_t* p) {
    uint64_t v;
    asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// One term of math.hpp:37-44, (a*b + 2^14) >> 15, for a given in the HIGH half of a 32-bit word and b doubled:
// (a*2^16) * (2b) + 2^31 = 2^17 (a*b + 2^14), whose upper word is the term. One IMAD.WIDE, no shift.
__device__ __forceinline__ int q15_mul_hi(int a_hi16, int b2) {
    return (int)(((long long)a_hi16 * (long long)b2 + 0x80000000ll) >> 32);
}

// Sum of the terms of one 16-byte unit (8 elements) against 8 doubled query elements.
__device__ __forceinline__ int q15_dot_unit(uint4 w, int4 a, int4 b) {
    int s = 0;
    s += q15_mul_hi((int)(w.x << 16), a.x); s += q15_mul_hi((int)(w.x & 0xffff0000u), a.y);
    s += q15_mul_hi((int)(w.y << 16), a.z); s += q15_mul_hi((int)(w.y & 0xffff0000u), a.w);
    s += q15_mul_hi((int)(w.z << 16), b.x); s += q15_mul_hi((int)(w.z & 0xffff0000u), b.y);
    s += q15_mul_hi((int)(w.w << 16), b.z); s += q15_mul_hi((int)(w.w & 0xffff0000u), b.w);
    return s;
}

// Q15 similarity (dot + 32768) of one row straight from global memory by one warp (rare path: tail entries whose
// similarity was not prefetched). cosine.hpp:19-23 / math.hpp:11-44.
__device__ __forceinline__ uint32_t warp_row_sim(const int16_t* __restrict__ row, const int* qrow2, uint32_t sl) {
    const uint32_t cpr = sl / 8;
    int s = 0;
    for (uint32_t ch = lane_id(); ch < cpr; ch += 32) {
        uint4 w = __ldg(reinterpret_cast<const uint4*>(row) + ch);
        const int4* qv = reinterpret_cast<const int4*>(qrow2) + 2 * ch;
        s += q15_dot_unit(w, qv[0], qv[1]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return (uint32_t)(uint16_t)(s + 32768);
}

// Exclusive prefix of a small per-lane count (0..7) with three ballots instead of a shuffle scan.
__device__ __forceinline__ uint32_t warp_excl_scan3(uint32_t cnt, uint32_t& total) {
    const uint32_t lt = (1u << lane_id()) - 1u;
    const uint32_t b0 = __ballot_sync(0xffffffffu, cnt & 1u), b1 = __ballot_sync(0xffffffffu, cnt & 2u),
                   b2 = __ballot_sync(0xffffffffu, cnt & 4u);
    total = __popc(b0) + 2 * __popc(b1) + 4 * __popc(b2);
    return __popc(b0 & lt) + 2 * __popc(b1 & lt) + 4 * __popc(b2 & lt);
}

// One PUFFINN query against cluster c by the whole CTA (collection.hpp:543-601 -> search_maps :768-948).
// Returns the number of results; sm.mb[0..cnt) holds them best first.
__device__ uint32_t probe_cluster_cta(const SearchParams& p, const CtaSmem& sm, uint32_t c, const uint32_t* __restrict__ codes,
                                      uint64_t code_stride, const uint32_t* __restrict__ stop, float max_sim, uint16_t* memo,
                                      uint32_t& phase) {
    const uint32_t L = p.g.L, k = p.k, sl = p.g.sl;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    const int16_t* rows = p.q15 + off * sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    CtaCtrl* ctrl = sm.ctrl;
    const uint32_t cpr = sl / 8;
    const uint32_t sunits = stage_stride_units(sl);

    if (memo) {
        uint4* mz = reinterpret_cast<uint4*>(memo);
        for (uint32_t i = tid; i < (nc + 7) / 8; i += kCtaThreads) mz[i] = make_uint4(0, 0, 0, 0);
    }
    // --- SearchBuffers ctor (collection.hpp:642-645): one thread per table
    for (uint32_t t = tid; t < L; t += kCtaThreads) {
        const uint32_t h = codes[(uint64_t)t * code_stride];
        const uint32_t* H = p.tbl_hash + (uint64_t)t * p.n + off;
        uint32_t lo = 0, len = nc;  // lower_bound == the reference's hinted halving search
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
            up[j] = pu < nc ? lcp24(__ldg(H + pu), h) : 0u;  // beyond the data lie the 0xffffffff sentinels (prefixmap.hpp:215-226)
            int64_t pd = (int64_t)lo - 1 - kSegment * j;
            dn[j] = pd >= 0 ? lcp24(__ldg(H + pd), h) : 0u;
        }
        sm.lcp_up[t] = make_uint2(up[0] | up[1] << 8 | up[2] << 16 | up[3] << 24, up[4] | up[5] << 8 | up[6] << 16 | up[7] << 24);
        sm.lcp_dn[t] = make_uint2(dn[0] | dn[1] << 8 | dn[2] << 16 | dn[3] << 24, dn[4] | dn[5] << 8 | dn[6] << 16 | dn[7] << 24);
    }
    if (tid == 0) {
        ctrl->inserted = 0;            // maxbuffer.hpp:53-55
        ctrl->minval16 = 0;
        ctrl->max_diff = kSketchBits;  // filterer.hpp:101
        ctrl->stopped = 0;
    }
    __syncthreads();
    const uint64_t my_sketch = sm.qsk[lane];  // ring slot == lane (used by warp 0 for the tail)

    for (uint32_t depth = kMaxHashBits; depth > 0; depth--) {
        if (ctrl->stopped) break;
        // --- fill_ranges (collection.hpp:650-667) with get_next_range (prefixmap.hpp:267-304) in closed form
        const uint32_t it = kMaxHashBits + 1 - depth;              // iteration 1..24
        const uint32_t dir_bit = 1u << (it >= 2 ? it - 2 : 0);     // removed bit (prefixmap.hpp:268-274)
        uint32_t running = 0;
        for (uint32_t t0 = 0; t0 < L; t0 += kCtaThreads) {
            const uint32_t t = t0 + tid;
            uint32_t nseg = 0;
            if (t < L) {
                const uint32_t h = sm.code[t];
                const uint32_t A = sm.anchor[t];
                const uint32_t* H = p.tbl_hash + (uint64_t)t * p.n + off;
                int64_t start, end;
                if ((h & dir_bit) == 0) {  // upward (prefixmap.hpp:277-290)
                    uint32_t j = lead_count(sm.lcp_up[t], depth);
                    if (j == 8) {
                        // run longer than the samples: first position >= A + 96 whose prefix differs, rounded up to the stride
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
                } else {  // downward (prefixmap.hpp:291-303)
                    uint32_t j = lead_count(sm.lcp_dn[t], depth);
                    if (j == 8) {
                        uint32_t hi = A >= 8 * kSegment ? A - 8 * kSegment : 0;  // positions [0, hi) undecided
                        uint32_t lo = 0, len = hi;
                        while (len > 0) {  // lower_bound of "prefix matches" (monotone: false ... false true ... true)
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
        const uint32_t S = running;
        if (S <= (uint32_t)kRing) continue;  // the initial ring fill swallows the whole stream (collection.hpp:802-810)
        if (tid == 0) {
            sm.segbase[L] = running;
            ctrl->base = 0;
            ctrl->np = 0;
        }
        __syncthreads();

        for (;;) {  // chunks of the segment stream of this depth
            const uint32_t cb = ctrl->base;                                         // first stream segment of the chunk
            const uint32_t ce = S - cb < kChunkSegs ? S : cb + kChunkSegs;          // one past the last
            const uint32_t md = ctrl->max_diff;                                     // threshold in force at the start of the chunk
            // ---------------------------------------------------------------- phase A: indices, Hamming distances, memo
            {
                const uint32_t s0 = cb + tid * kSegPerThread;
                uint32_t ids[kSegPerThread][4];
                uint32_t nmine = 0;
                if (s0 < ce) {
                    nmine = ce - s0 < kSegPerThread ? ce - s0 : kSegPerThread;
                    // segment number -> (table, position): upper_bound(segbase, s0) - 1, then walk forward
                    uint32_t lo = 0, len = L;
                    while (len > 0) {
                        uint32_t half = len >> 1, mid = lo + half;
                        if (sm.segbase[mid] <= s0) { lo = mid + 1; len -= half + 1; } else { len = half; }
                    }
                    uint32_t t = lo - 1;
#pragma unroll
                    for (uint32_t j = 0; j < kSegPerThread; j++) {
                        if (j < nmine) {
                            const uint32_t s = s0 + j;
                            while (t + 1 < L && sm.segbase[t + 1] <= s) t++;
                            const uint32_t* seg = p.tbl_idx + (uint64_t)t * p.n + off + sm.start[t] + 4 * (s - sm.segbase[t]);
                            ids[j][0] = ldg_nc_na_u32(seg); ids[j][1] = ldg_nc_na_u32(seg + 1);
                            ids[j][2] = ldg_nc_na_u32(seg + 2); ids[j][3] = ldg_nc_na_u32(seg + 3);
                        }
                    }
                }
                // every sketch word and memo entry of the thread's 16 candidates is requested before the first one is used
                uint64_t w[kSegPerThread][4];
                uint32_t mm[kSegPerThread][4];
#pragma unroll
                for (uint32_t j = 0; j < kSegPerThread; j++) {
                    if (j < nmine) {
                        const uint32_t slot = (s0 + j) & 31u;  // chunks start on a multiple of the ring size
#pragma unroll
                        for (int e = 0; e < 4; e++) w[j][e] = ldg_nc_na_u64(sk + ((uint64_t)ids[j][e] << 5 | slot));
#pragma unroll
                        for (int e = 0; e < 4; e++) mm[j][e] = memo ? (uint32_t)memo[ids[j][e]] : 0u;
                    }
                }
                uint32_t my_unk = 0;  // 4 bits per segment: candidates that pass md and are not memoised
                uint32_t n_unk = 0;
#pragma unroll
                for (uint32_t j = 0; j < kSegPerThread; j++) {
                    if (j < nmine) {
                        const uint32_t s = s0 + j;
                        const uint64_t qs = sm.qsk[s & 31u];
                        uint32_t pcs = 0;
                        const uint32_t ci = (s - cb) * 4;
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const uint32_t pc = (uint32_t)__popcll(w[j][e] ^ qs);
                            pcs |= pc << (8 * e);
                            if (pc <= md) {
                                if (mm[j][e]) {
                                    sm.csim[ci + e] = (uint16_t)mm[j][e];
                                } else {
                                    my_unk |= 1u << (4 * j + e);
                                    n_unk++;
                                }
                            }
                        }
                        sm.cpc[ci >> 2] = pcs;
                        *reinterpret_cast<uint4*>(sm.cid + ci) = make_uint4(ids[j][0], ids[j][1], ids[j][2], ids[j][3]);
                    }
                }
                if (n_unk) {
                    uint32_t pos = atomicAdd(&ctrl->unk_cnt, n_unk);
                    const uint32_t c0 = (s0 - cb) * 4;
                    while (my_unk) {
                        const uint32_t bit = __ffs(my_unk) - 1;
                        my_unk &= my_unk - 1;
                        sm.unk[pos++] = (uint16_t)(c0 + bit);
                    }
                }
            }
            __syncthreads();
            // ---------------------------------------------------------------- phase B: Q15 rerank of the missing rows
            const uint32_t nunk = ctrl->unk_cnt;
            for (uint32_t rb = 0; rb < nunk; rb += kStageRows) {
                const uint32_t nrows = nunk - rb < kStageRows ? nunk - rb : kStageRows;
                if (tid == 0) mbar_arrive_expect_tx(&ctrl->mbar, nrows * sl * 2);
                if (tid < nrows) {
                    const uint32_t id = sm.cid[sm.unk[rb + tid]];
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of the slot vs the async write
                    bulk_copy_g2s(sm.stage + tid * sunits * 16, rows + (uint64_t)id * sl, sl * 2, &ctrl->mbar);
                }
                mbar_wait(&ctrl->mbar, phase);
                phase ^= 1u;
                {
                    const uint32_t r = tid >> 1, half = tid & 1u;
                    int s = 0;
                    if (r < nrows) {
                        const uint4* row = reinterpret_cast<const uint4*>(sm.stage + r * sunits * 16);
                        const int4* qv = reinterpret_cast<const int4*>(sm.qrow2);
                        for (uint32_t ch = half; ch < cpr; ch += 2) s += q15_dot_unit(row[ch], qv[2 * ch], qv[2 * ch + 1]);
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    if (r < nrows && half == 0) {
                        const uint32_t pos = sm.unk[rb + r];
                        const uint16_t sim16 = (uint16_t)(s + 32768);
                        sm.csim[pos] = sim16;
                        if (memo) memo[sm.cid[pos]] = sim16;
                    }
                }
                __syncthreads();
            }
            // ---------------------------------------------------------------- phase C: the sequential loop, warp 0
            if (warp == 0) {
                uint32_t base = ctrl->base, np = ctrl->np;
                uint32_t inserted = ctrl->inserted, minval16 = ctrl->minval16, max_diff = md;
                unsigned long long cand = 0, dcomp = 0;
                uint32_t status = 0, stopped = 0;
                for (;;) {
                    bool need = false;
                    // full ring sweeps (collection.hpp:813-866): slot == lane, real sketches
                    while (np < (uint32_t)kFilterBuffer && base + kRing <= S) {
                        if (base + kRing > ce) { need = true; break; }
                        const uint32_t ci = (base - cb + lane) * 4;
                        const uint32_t pcs = sm.cpc[ci >> 2];
                        const uint32_t mask = __vcmpleu4(pcs, max_diff * 0x01010101u);
                        const uint32_t cnt = (uint32_t)__popc(mask) >> 3;
                        uint32_t total;
                        uint32_t pos = np + warp_excl_scan3(cnt, total);
                        if (cnt) {
                            const uint4 v = *reinterpret_cast<const uint4*>(sm.cid + ci);
                            const uint2 sv = *reinterpret_cast<const uint2*>(sm.csim + ci);
                            if (mask & 0x000000ffu) { sm.pass_idx[pos] = v.x; sm.pass_sim[pos] = (uint16_t)(sv.x & 0xffffu); pos++; }
                            if (mask & 0x0000ff00u) { sm.pass_idx[pos] = v.y; sm.pass_sim[pos] = (uint16_t)(sv.x >> 16); pos++; }
                            if (mask & 0x00ff0000u) { sm.pass_idx[pos] = v.z; sm.pass_sim[pos] = (uint16_t)(sv.y & 0xffffu); pos++; }
                            if (mask & 0xff000000u) { sm.pass_idx[pos] = v.w; sm.pass_sim[pos] = (uint16_t)(sv.y >> 16); pos++; }
                        }
                        np += total;
                        cand += kRing * 4;
                        base += kRing;
                    }
                    if (need) { status = kNeedMore; break; }
                    // tail (collection.hpp:869-903): the not-yet-tested ring slots, in descending slot order, tested with the
                    // point index itself in place of its sketch (:890-893)
                    const uint32_t live = base + kRing > S ? (S > base ? S - base : 0u) : (uint32_t)kRing;
                    if (base + live > ce) { status = kNeedMore; break; }
                    {
                        uint32_t cnt = 0, pm = 0, um = 0;
                        uint4 v = make_uint4(0, 0, 0, 0);
                        uint2 sv = make_uint2(0, 0);
                        if (lane < live) {
                            const uint32_t ci = (base - cb + lane) * 4;
                            v = *reinterpret_cast<const uint4*>(sm.cid + ci);
                            sv = *reinterpret_cast<const uint2*>(sm.csim + ci);
                            const uint32_t pcs = sm.cpc[ci >> 2];
                            const uint32_t known = __vcmpleu4(pcs, md * 0x01010101u);  // similarity was prefetched in phase B
                            pm |= ((uint32_t)__popcll((uint64_t)v.x ^ my_sketch) <= max_diff) ? 1u : 0u;
                            pm |= ((uint32_t)__popcll((uint64_t)v.y ^ my_sketch) <= max_diff) ? 2u : 0u;
                            pm |= ((uint32_t)__popcll((uint64_t)v.z ^ my_sketch) <= max_diff) ? 4u : 0u;
                            pm |= ((uint32_t)__popcll((uint64_t)v.w ^ my_sketch) <= max_diff) ? 8u : 0u;
                            cnt = __popc(pm);
                            um = pm & ~((known & 1u) | ((known >> 7) & 2u) | ((known >> 14) & 4u) | ((known >> 21) & 8u));
                        }
                        uint32_t total;
                        const uint32_t ex = warp_excl_scan3(cnt, total);
                        uint32_t pos = np + (total - ex - cnt);  // entries of higher slots come first
                        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
                        const uint32_t ss[4] = {sv.x & 0xffffu, sv.x >> 16, sv.y & 0xffffu, sv.y >> 16};
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            if (pm & (1u << e)) {
                                sm.pass_idx[pos] = vv[e];
                                sm.pass_sim[pos] = (uint16_t)ss[e];
                                if (um & (1u << e)) sm.unk[atomicAdd(&ctrl->tail_unk, 1u)] = (uint16_t)pos;
                                pos++;
                            }
                        }
                        np += total;
                        cand += 4ull * live;
                        __syncwarp();
                        // tail entries that failed the sketch filter of phase A but pass the index-as-sketch test: compute now
                        const uint32_t ntu = ctrl->tail_unk;
                        if (ntu) {
                            for (uint32_t i = 0; i < ntu; i++) {
                                const uint32_t pp = sm.unk[i];
                                const uint32_t sim16 = warp_row_sim(rows + (uint64_t)sm.pass_idx[pp] * sl, sm.qrow2, sl);
                                if (lane == 0) sm.pass_sim[pp] = (uint16_t)sim16;
                            }
                            __syncwarp();
                            if (lane == 0) ctrl->tail_unk = 0;
                            __syncwarp();
                        }
                    }
                    // empty the buffer (collection.hpp:909-925)
                    dcomp += np;
                    maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, np);
                    max_diff = p.msd[minval16 < 65536u ? minval16 : 65535u];  // filterer.hpp:108-111
                    np = 0;
                    // stop rule (collection.hpp:927-943)
                    const uint32_t pulled = base + kRing;
                    uint32_t table_idx = L;
                    if (pulled < S) {
                        uint32_t lo = 0, len = L;
                        while (len > 0) {
                            uint32_t half = len >> 1, mid = lo + half;
                            if (sm.segbase[mid] <= pulled) { lo = mid + 1; len -= half + 1; } else { len = half; }
                        }
                        table_idx = lo - 1;
                    }
                    const float kth = __fdiv_rn((float)minval16, 65536.0f);
                    const float sim = kth < max_sim ? max_sim : kth;  // std::max(kth, max_sim)
                    uint32_t bin = (uint32_t)__fdiv_rn(sim, 0.005f);  // crosspolytope.hpp:116-118
                    bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
                    const uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (table_idx >> 5));
                    if ((word >> (table_idx & 31)) & 1u) { stopped = 1; break; }
                    if (!(base + kRing < S)) break;  // status 0: depth finished
                }
                if (lane == 0) {
                    ctrl->base = base;
                    ctrl->np = np;
                    ctrl->inserted = inserted;
                    ctrl->minval16 = minval16;
                    ctrl->max_diff = max_diff;
                    ctrl->stopped = stopped;
                    ctrl->status = status;
                    ctrl->unk_cnt = 0;
                    ctrl->candidates += cand;
                    ctrl->distcomp += dcomp;
                }
            }
            __syncthreads();
            if (ctrl->status != kNeedMore) break;
        }
    }
    __syncthreads();
    if (warp == 0) {  // best_indices (collection.hpp:598, maxbuffer.hpp:79-96)
        uint32_t inserted = ctrl->inserted, minval16 = ctrl->minval16;
        maxbuffer_filter(sm.mb, P, k, inserted, minval16);
        if (lane == 0) ctrl->result_cnt = inserted;
    }
    __syncthreads();
    return ctrl->result_cnt;
}

template <int OCC>
__global__ void __launch_bounds__(kCtaThreads, OCC) k_probe_cta(SearchParams p, QueryBatch b, int stop_at_foreign, uint16_t* memo_base,
                                                                uint64_t memo_stride) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const CtaSmem sm = carve_cta(s_dyn, p.g.L, p.k, p.g.sl);
    CtaCtrl* ctrl = sm.ctrl;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t state_bytes = sizeof(QueryStateHeader) + (uint64_t)p.k * 8;
    uint16_t* memo = memo_base ? memo_base + (uint64_t)blockIdx.x * memo_stride : nullptr;
    const uint32_t P = next_pow2(2 * p.k) < 32 ? 32 : next_pow2(2 * p.k);
    if (tid == 0) {
        mbar_init(&ctrl->mbar, 1);
        ctrl->unk_cnt = 0;
        ctrl->tail_unk = 0;
    }
    uint32_t phase = 0;
    __syncthreads();

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
        const uint32_t heap_len0 = st->heap_len;
        for (uint32_t i = tid; i < heap_len0; i += kCtaThreads) sm.heap[i] = st_heap[i];
        for (uint32_t i = tid; i < p.g.sl; i += kCtaThreads) sm.qrow2[i] = 2 * (int)b.q15[(uint64_t)q * p.g.sl + i];
        if (tid == 0) {
            ctrl->heap_len = heap_len0;
            ctrl->candidates = st->candidates;
            ctrl->distcomp = st->distcomp;
        }
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
                float* s_dist = reinterpret_cast<float*>(sm.cid);
                uint32_t* s_pid = sm.cid + kCtaThreads;
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
                if (tid < kNumSketches) sm.qsk[tid] = b.sketches[((uint64_t)fs * b.nq + q) * kNumSketches + tid];
                const uint32_t* stop = p.stop + (uint64_t)fs * kMaxHashBits * kEstBins * p.stop_words;
                uint16_t* use_memo = (memo && nc <= memo_stride) ? memo : nullptr;
                const uint32_t cnt = probe_cluster_cta(p, sm, c, codes, b.nq, stop, max_sim, use_memo, phase);
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
            st->candidates = ctrl->candidates;
            st->distcomp = ctrl->distcomp;
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
    const size_t smem = cta_smem_bytes(p.g.L, p.k, p.g.sl);
    if (smem > 227 * 1024) throw std::invalid_argument("num_tables / k / dimension too large for the probe kernel's shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        CLANN_CUDA(cudaFuncSetAttribute(k_probe_cta<OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int ctas_per_sm = 0;
    CLANN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_probe_cta<OCC>, kCtaThreads, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    static int cap = -1;
    if (cap < 0) {
        const char* e = getenv("CLANN_PROBE_CTAS_PER_SM");  // tuning knob: fewer resident queries = smaller L2 working set
        cap = e ? atoi(e) : 0;
    }
    if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = cap;
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
        occ = e ? atoi(e) : 4;
        if (occ != 3 && occ != 4 && occ != 5 && occ != 6) occ = 4;
    }
    switch (occ) {
        case 3: launch_probe_cta_occ<3>(p, b, stop_at_foreign, s); break;
        case 5: launch_probe_cta_occ<5>(p, b, stop_at_foreign, s); break;
        case 6: launch_probe_cta_occ<6>(p, b, stop_at_foreign, s); break;
        default: launch_probe_cta_occ<4>(p, b, stop_at_foreign, s); break;
    }
}

}  // namespace clann
