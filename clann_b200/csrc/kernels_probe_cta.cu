// k_probe_cta — the CLANN search loop with ONE CTA (4 warps) PER QUERY, warp-specialised so that no global-memory latency
// sits on the sequential part of puffinn::Index::search_maps and the sequential part overlaps the memory-bound part.
//
// Same results, bit for bit, as the one-warp-per-query kernel in kernels_search.cu (which stays as the plain
// restatement and serves the legacy single-query ABI). What changes is the schedule of one (query, cluster) visit.
// Per depth the segment stream is cut into CHUNKS of 384 segments (12 ring sweeps, 1536 candidates); chunk boundaries
// do not depend on the data (every batch of the reference ends on a ring boundary), so chunks can be prepared ahead:
//
//   PRODUCERS (warps 1-3, 96 threads), one chunk ahead of the consumer through a two-slot ring in shared memory
//   ranges   per depth: get_next_range in closed form for every table, prefix sum -> segment stream
//   phase A  table indices of the chunk and the sketch word of every candidate (ring slot = segment number mod 32) ->
//            its Hamming distance to the query sketch, one byte per candidate. The distance is kept instead of a
//            pass/fail bit because the filter threshold only ever tightens during a visit (filterer.hpp:108-111 is
//            monotone in the k-th similarity): the consumer applies whatever threshold is in force when it gets there.
//   phase B  Q15 similarity of every candidate that passes the (possibly stale, hence looser) threshold the producers
//            last saw. A per-CTA memo (one u16 per local id) returns similarities already computed during this visit —
//            the reference rescans nested ranges at every depth, so more than half of its distance computations are
//            repeats (the counter still counts them). Missing rows are read with 128-bit loads, four lanes per row.
//   CONSUMER (warp 0) replays the reference's sequential loop out of shared memory only: ring sweeps in order (lane =
//            ring slot), the 128-entry passing buffer, the index-as-sketch tail (collection.hpp:890-893), MaxBuffer
//            inserts, threshold update, stop rule. A batch may straddle chunks; the stop rule may end the visit in the
//            middle of a chunk (what the producers prepared beyond that point was speculative).
//   Hand-over: one mbarrier pair (full / empty) per slot; an END record closes every visit so both sides always see
//            the same number of chunks and the barrier phases stay in step across visits.
//
// Queries are scheduled in nearest-cluster order and only (SMs x CTAs/SM) of them are in flight, so the clusters being
// probed at any moment (rows + sketches + tables, ~3 MB each at the glove-100 shape) stay resident in the 126 MB L2.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

#ifdef CLANN_TIMING
// Debug build only (-DCLANN_TIMING): clock64 totals per phase, summed over CTAs, printed by the launcher.
enum { T_SETUP, T_RANGES, T_WAIT_EMPTY, T_A, T_B, T_HAND, T_WAIT_FULL, T_C, T_VISIT, T_QUERY, T_SELECT, T_FINAL, T_CHUNKS, T_VISITS, T_N };
__device__ unsigned long long g_timing[T_N];
#define TSTART(var) long long var = clock64()
#define TADD(slot, var) do { long long _n = clock64(); t_acc[slot] += (unsigned long long)(_n - var); var = _n; } while (0)
#define TDECL unsigned long long t_acc[T_N] = {0}
#define TARG , unsigned long long* t_acc
#define TPASS , t_acc
#else
#define TSTART(var)
#define TADD(slot, var)
#define TDECL
#define TARG
#define TPASS
#endif

constexpr int kCtaWarps = 4;
constexpr int kCtaThreads = kCtaWarps * 32;
constexpr uint32_t kProducers = kCtaThreads - 32;                  // warps 1..3
constexpr uint32_t kSegPerThread = 4;                              // consecutive segments handled by one producer in phase A
constexpr uint32_t kChunkSegs = kProducers * kSegPerThread;        // 384 segments = 12 ring sweeps per chunk
constexpr uint32_t kChunkCand = kChunkSegs * 4;                    // 1536 candidates
constexpr uint32_t kRowsPerIter = kProducers / 4;                  // phase B: four lanes per row
constexpr uint32_t kSlots = 2;
constexpr uint32_t kNeedMore = 1u, kDepthDone = 0u, kStopped = 2u;

struct CtaCtrl {
    unsigned long long nk, top;
    unsigned long long full[kSlots], empty[kSlots];  // mbarriers of the chunk ring
    unsigned long long candidates, distcomp;         // performance.hpp:72-86, running totals of the query
    uint32_t work, heap_len, result_cnt;
    uint32_t inserted, minval16;                     // MaxBuffer state at the end of the visit (consumer -> everyone)
    volatile uint32_t max_diff, stopped;             // consumer -> producers (advisory threshold, stop flag)
    uint32_t p_md, p_stop;                           // the producers' uniform copy of the two, refreshed once per chunk
    uint32_t unk_cnt, tail_unk;
    uint32_t warp_tot[kCtaWarps];
};
static_assert(sizeof(CtaCtrl) <= 128, "CtaCtrl must fit its 128-byte slot");

// Chunk record: what the producers hand to the consumer.
struct SlotMeta {
    uint32_t depth;  // 0 = END of the visit
    uint32_t S;      // segments in the stream of this depth
    uint32_t cb, ce; // stream segments [cb, ce) are in the slot
    uint32_t md;     // every candidate with Hamming distance <= md has its similarity in csim
    uint32_t pad[3];
};

struct CtaSmem {
    CtaCtrl* ctrl;
    uint64_t* qsk;              // [32] query sketches of the function set in use
    int* qrow2;                 // [sl] query in Q15, doubled (see q15_mul_hi)
    // chunk ring
    uint8_t* ring;              // kSlots records of slot_bytes each:
    uint32_t slot_bytes;
    //   cid  [kChunkCand] u32  local ids of the chunk's candidates, stream order
    //   csim [kChunkCand] u16  similarity (dot + 32768) of the candidates with distance <= meta.md
    //   cpc  [kChunkCand] u8   Hamming distances
    //   segb [L+1] u32         copy of segbase for the stop rule's table lookup
    //   meta SlotMeta
    __device__ __forceinline__ uint32_t* cid(uint32_t slot) const { return reinterpret_cast<uint32_t*>(ring + slot * slot_bytes); }
    __device__ __forceinline__ uint16_t* csim(uint32_t slot) const { return reinterpret_cast<uint16_t*>(ring + slot * slot_bytes + kChunkCand * 4); }
    __device__ __forceinline__ uint32_t* cpc(uint32_t slot) const { return reinterpret_cast<uint32_t*>(ring + slot * slot_bytes + kChunkCand * 6); }
    __device__ __forceinline__ uint32_t* segb(uint32_t slot) const { return reinterpret_cast<uint32_t*>(ring + slot * slot_bytes + kChunkCand * 7); }
    __device__ __forceinline__ SlotMeta* meta(uint32_t slot) const { return reinterpret_cast<SlotMeta*>(ring + (slot + 1) * slot_bytes - 32); }
    // producers
    uint16_t* unk;              // [kChunkCand] chunk positions whose similarity is not memoised yet
    uint2* lcp_up;              // [L]
    uint2* lcp_dn;              // [L]
    uint32_t* anchor;           // [L]
    uint32_t* code;             // [L]
    uint32_t* start;            // [L]
    uint32_t* segbase;          // [L+1]
    // consumer
    unsigned long long* mb;     // [P2K] MaxBuffer slots
    unsigned long long* heap;   // [k]
    unsigned long long* loc;    // [k]
    uint32_t* pass_idx;         // [kPassingCap]
    uint16_t* pass_sim;         // [kPassingCap]
    uint16_t* tunk;             // [4 * kRing] passing-list positions of tail entries without a prefetched similarity
};

__host__ __device__ inline uint32_t align16(uint32_t v) { return (v + 15u) & ~15u; }

__host__ __device__ inline uint32_t cta_smem_bytes(uint32_t L, uint32_t k, uint32_t sl) {
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint32_t b = 128 + 32 * 8 + sl * 4;                                                        // ctrl, qsk, qrow2
    b += kSlots * (kChunkCand * 4 + kChunkCand * 2 + kChunkCand + align16((L + 1) * 4) + 32);  // ring
    b += kChunkCand * 2;                                                                       // unk
    b += L * 8 * 2 + p2k * 8 + k * 8 * 2;                                                      // lcp_up, lcp_dn, mb, heap, loc
    b += L * 4 * 3 + (L + 1) * 4 + kPassingCap * 4;                                            // anchor, code, start, segbase, pass_idx
    b += kPassingCap * 2 + 4 * kRing * 2;                                                      // pass_sim, tunk
    return align16(b);
}

__device__ __forceinline__ CtaSmem carve_cta(uint8_t* base, uint32_t L, uint32_t k, uint32_t sl) {
    CtaSmem s;
    uint32_t p2k = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    uint8_t* p = base;  // 16-byte aligned blocks first, then 8-, 4- and 2-byte arrays
    s.ctrl = reinterpret_cast<CtaCtrl*>(p); p += 128;
    s.qsk = reinterpret_cast<uint64_t*>(p); p += 32 * 8;
    s.qrow2 = reinterpret_cast<int*>(p); p += sl * 4;
    s.ring = p;
    s.slot_bytes = kChunkCand * 7 + align16((L + 1) * 4) + 32;
    p += kSlots * s.slot_bytes;
    s.unk = reinterpret_cast<uint16_t*>(p); p += kChunkCand * 2;
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
    s.pass_sim = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    s.tunk = reinterpret_cast<uint16_t*>(p);
    return s;
}

// --- mbarriers (SASS: SYNCS) and the producers' named barrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // watchdog: a broken hand-over protocol must fail loudly, not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory"); }

// L1::no_allocate loads are EVICT_FIRST in the L2 unless told otherwise (ncu: lts__t_sectors_*_evict_first_*, see
// kernels_probe2.cu); sketch words and rows are exactly what the other queries of the cluster re-read, so they carry an
// explicit evict_normal policy.
__device__ __forceinline__ uint64_t l2_keep_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ldg_nc_na_u32(const uint32_t* p) {
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(l2_keep_policy()));
    return v;
}
__device__ __forceinline__ uint64_t ldg_nc_na_u64(const uint64_t* p) {
    uint64_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(l2_keep_policy()));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_na_v4(const uint4* p) {
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
        : "l"(p), "l"(l2_keep_policy()));
    return v;
}

// Ask the TMA unit to stream [ptr, ptr+bytes) into L2 (SASS: UBLKPF). The range is widened to 16-byte boundaries.
__device__ __forceinline__ void l2_prefetch_range(const void* ptr, uint64_t bytes) {
    uint64_t a = reinterpret_cast<uint64_t>(ptr);
    const uint64_t end = (a + bytes + 15ull) & ~15ull;
    a &= ~15ull;
    const uint32_t sz = (uint32_t)(end - a);
    if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(sz) : "memory");
}

// Streams everything a visit of cluster c can touch (Q15 rows, sketches, the cluster's slice of every table) into L2:
// sequential DRAM traffic at full bandwidth instead of the random 32-byte reads the probe itself would issue.
__device__ __forceinline__ void l2_prefetch_cluster(const SearchParams& p, uint32_t c) {
    if (p.brute[c]) return;
    const uint64_t off = p.offsets[c];
    const uint64_t nc = p.offsets[c + 1] - off;
    constexpr uint64_t kPiece = 32 * 1024;
    const uint8_t* base[4] = {reinterpret_cast<const uint8_t*>(p.q15 + off * p.g.sl), reinterpret_cast<const uint8_t*>(p.sketches + off * kNumSketches),
                              reinterpret_cast<const uint8_t*>(p.tbl_hash + (uint64_t)p.g.L * off),
                              reinterpret_cast<const uint8_t*>(p.tbl_idx + (uint64_t)p.g.L * off)};
    const uint64_t len[4] = {nc * p.g.sl * 2, nc * kNumSketches * 8, nc * p.g.L * 4, nc * p.g.L * 4};
#pragma unroll
    for (int a = 0; a < 4; a++)
        for (uint64_t lo = (uint64_t)threadIdx.x * kPiece; lo < len[a]; lo += (uint64_t)blockDim.x * kPiece)
            l2_prefetch_range(base[a] + lo, lo + kPiece < len[a] ? kPiece : len[a] - lo);
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

struct VisitArgs {
    uint32_t c;                 // cluster
    const uint32_t* codes;      // this query's code of table 0 (stride code_stride)
    uint64_t code_stride;
    const uint32_t* stop;       // stop table of the function set
    float max_sim;
    uint16_t* memo;             // per-CTA similarity memo or null
};

// ------------------------------------------------------------------------------------------------ producers (warps 1-3)
__device__ __forceinline__ void produce_visit(const SearchParams& p, const CtaSmem& sm, const VisitArgs& a, uint32_t& it TARG) {
    const uint32_t L = p.g.L, sl = p.g.sl;
    const uint32_t ptid = threadIdx.x - 32, pwarp = ptid >> 5, lane = threadIdx.x & 31;
    const uint64_t off = p.offsets[a.c];
    const uint32_t nc = (uint32_t)(p.offsets[a.c + 1] - off);
    const int16_t* rows = p.q15 + off * sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;
    uint16_t* memo = a.memo;
    CtaCtrl* ctrl = sm.ctrl;
    const uint32_t cpr = sl / 8;
    bool stop_seen = false;
    TSTART(tp);

    for (uint32_t depth = kMaxHashBits; depth > 0 && !stop_seen; depth--) {
        // --- fill_ranges (collection.hpp:650-667) with get_next_range (prefixmap.hpp:267-304) in closed form
        uint32_t running = 0;
        for (uint32_t t0 = 0; t0 < L; t0 += kProducers) {
            const uint32_t t = t0 + ptid;
            uint32_t nseg = 0;
            if (t < L)
                sm.start[t] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)a.c * L + t) * kDirEntries, nc,
                                          sm.code[t], sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg);
            uint32_t total;
            const uint32_t ex = warp_excl_scan(nseg, total);
            if (lane == 0) ctrl->warp_tot[pwarp] = total;
            producers_sync();
            uint32_t prefix = running, tile_total = 0;
#pragma unroll
            for (uint32_t w = 0; w < kProducers / 32; w++) {
                uint32_t wt = ctrl->warp_tot[w];
                if (w < pwarp) prefix += wt;
                tile_total += wt;
            }
            if (t < L) sm.segbase[t] = prefix + ex;
            running += tile_total;
            producers_sync();
        }
        const uint32_t S = running;
        TADD(T_RANGES, tp);
        if (S <= (uint32_t)kRing) continue;  // the initial ring fill swallows the whole stream (collection.hpp:802-810)
        if (ptid == 0) sm.segbase[L] = S;

        for (uint32_t cb = 0; cb < S; cb += kChunkSegs) {  // chunk boundaries are data independent (see the header)
            const uint32_t slot = it & 1u;
            mbar_wait(&ctrl->empty[slot], ((it >> 1) & 1u) ^ 1u);
            TADD(T_WAIT_EMPTY, tp);
            if (ptid == 0) {
                ctrl->p_stop = ctrl->stopped;
                ctrl->p_md = ctrl->max_diff;
                ctrl->unk_cnt = 0;
            }
            producers_sync();  // also orders phase B of the previous chunk (memo writes) before this chunk's memo reads
            if (ctrl->p_stop) {
                stop_seen = true;
                break;
            }
            const uint32_t md = ctrl->p_md;
            const uint32_t ce = S - cb < kChunkSegs ? S : cb + kChunkSegs;
            uint32_t* cid = sm.cid(slot);
            uint16_t* csim = sm.csim(slot);
            uint32_t* cpc = sm.cpc(slot);
            // ---------------------------------------------------------------- phase A: indices, Hamming distances, memo
            {
                const uint32_t s0 = cb + ptid * kSegPerThread;
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
                            const uint32_t* seg = p.tbl_idx + table_base(off, nc, L, t) + sm.start[t] + 4 * (s - sm.segbase[t]);
                            ids[j][0] = __ldg(seg); ids[j][1] = __ldg(seg + 1);  // L1-allocating: the four words share a sector
                            ids[j][2] = __ldg(seg + 2); ids[j][3] = __ldg(seg + 3);
                        }
                    }
                }
                // every sketch word and memo entry of the thread's 16 candidates is requested before the first one is used
                uint64_t w[kSegPerThread][4];
                uint32_t mm[kSegPerThread][4];
#pragma unroll
                for (uint32_t j = 0; j < kSegPerThread; j++) {
                    if (j < nmine) {
                        const uint32_t slot_r = (s0 + j) & 31u;  // chunks start on a multiple of the ring size
#pragma unroll
                        for (int e = 0; e < 4; e++) w[j][e] = ldg_nc_na_u64(sk + ((uint64_t)ids[j][e] << 5 | slot_r));
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
                                    csim[ci + e] = (uint16_t)mm[j][e];
                                } else {
                                    my_unk |= 1u << (4 * j + e);
                                    n_unk++;
                                }
                            }
                        }
                        cpc[ci >> 2] = pcs;
                        *reinterpret_cast<uint4*>(cid + ci) = make_uint4(ids[j][0], ids[j][1], ids[j][2], ids[j][3]);
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
            producers_sync();
            TADD(T_A, tp);
            // ---------------------------------------------------------------- phase B: Q15 rerank of the missing rows
            {
                const uint32_t nunk = ctrl->unk_cnt;
                const uint32_t sub = ptid & 3u;
                const int4* qv = reinterpret_cast<const int4*>(sm.qrow2);
                for (uint32_t rb = 0; rb < nunk; rb += kRowsPerIter) {
                    const uint32_t r = rb + (ptid >> 2);
                    const bool valid = r < nunk;
                    int s = 0;
                    uint32_t pos = 0, id = 0;
                    if (valid) {
                        pos = sm.unk[r];
                        id = cid[pos];
                        const uint4* row = reinterpret_cast<const uint4*>(rows + (uint64_t)id * sl);
                        if (cpr <= 16) {  // d <= 128: at most four 16-byte units per lane, all requested before the first is used
                            uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0, u2 = u0, u3 = u0;
                            if (sub < cpr) u0 = ldg_nc_na_v4(row + sub);
                            if (sub + 4 < cpr) u1 = ldg_nc_na_v4(row + sub + 4);
                            if (sub + 8 < cpr) u2 = ldg_nc_na_v4(row + sub + 8);
                            if (sub + 12 < cpr) u3 = ldg_nc_na_v4(row + sub + 12);
                            if (sub < cpr) s += q15_dot_unit(u0, qv[2 * sub], qv[2 * sub + 1]);
                            if (sub + 4 < cpr) s += q15_dot_unit(u1, qv[2 * sub + 8], qv[2 * sub + 9]);
                            if (sub + 8 < cpr) s += q15_dot_unit(u2, qv[2 * sub + 16], qv[2 * sub + 17]);
                            if (sub + 12 < cpr) s += q15_dot_unit(u3, qv[2 * sub + 24], qv[2 * sub + 25]);
                        } else {
                            for (uint32_t ch = sub; ch < cpr; ch += 4) s += q15_dot_unit(ldg_nc_na_v4(row + ch), qv[2 * ch], qv[2 * ch + 1]);
                        }
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    if (valid && sub == 0) {
                        const uint16_t sim16 = (uint16_t)(s + 32768);
                        csim[pos] = sim16;
                        if (memo) memo[id] = sim16;
                    }
                }
            }
            TADD(T_B, tp);
            // ---------------------------------------------------------------- hand the slot over
            for (uint32_t i = ptid; i <= L; i += kProducers) sm.segb(slot)[i] = sm.segbase[i];
            if (ptid == 0) {
                SlotMeta* m = sm.meta(slot);
                m->depth = depth;
                m->S = S;
                m->cb = cb;
                m->ce = ce;
                m->md = md;
            }
            mbar_arrive(&ctrl->full[slot]);  // release: this thread's writes to the slot are visible to whoever sees the phase complete
            it++;
            TADD(T_HAND, tp);
#ifdef CLANN_TIMING
            t_acc[T_CHUNKS]++;
#endif
        }
    }
    // END record: closes the visit for the consumer whatever happened above
    {
        const uint32_t slot = it & 1u;
        mbar_wait(&ctrl->empty[slot], ((it >> 1) & 1u) ^ 1u);
        if (ptid == 0) sm.meta(slot)->depth = 0;
        mbar_arrive(&ctrl->full[slot]);
        it++;
    }
}

// ------------------------------------------------------------------------------------------------ consumer (warp 0)
__device__ __forceinline__ void consume_visit(const SearchParams& p, const CtaSmem& sm, const VisitArgs& a, uint32_t thr_lo,
                                              uint32_t thr_hi, uint32_t& it TARG) {
    const uint32_t L = p.g.L, k = p.k, sl = p.g.sl;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t off = p.offsets[a.c];
    const int16_t* rows = p.q15 + off * sl;
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    CtaCtrl* ctrl = sm.ctrl;
    const uint64_t my_sketch = sm.qsk[lane];  // ring slot == lane
    const float max_sim = a.max_sim;

    uint32_t inserted = 0, minval16 = 0, max_diff = kSketchBits;  // maxbuffer.hpp:53-55, filterer.hpp:101
    uint32_t base = 0, np = 0, cur_depth = 0, status = kDepthDone;
    unsigned long long cand = 0, dcomp = 0;
    uint32_t stop_key = 0xffffffffu, stop_word = 0;  // lane w < stop_words caches word w of the stop row (depth, bin)
    TSTART(tc);

    for (;;) {
        const uint32_t slot = it & 1u;
        mbar_wait(&ctrl->full[slot], (it >> 1) & 1u);
        TADD(T_WAIT_FULL, tc);
        const SlotMeta m = *sm.meta(slot);
        if (m.depth == 0 || status == kStopped) {  // END, or draining what the producers prepared past the stop
            __syncwarp();
            if (lane == 0) mbar_arrive(&ctrl->empty[slot]);
            it++;
            if (m.depth == 0) break;
            continue;
        }
        const uint32_t S = m.S, cb = m.cb, ce = m.ce, md = m.md, depth = m.depth;
        if ((depth != cur_depth) == (status == kNeedMore)) __trap();  // a depth ends exactly when the replay stops asking for more
        if (depth != cur_depth) {
            cur_depth = depth;
            base = 0;
            np = 0;
        }
        const uint32_t* cid = sm.cid(slot);
        const uint16_t* csim = sm.csim(slot);
        const uint32_t* cpc = sm.cpc(slot);
        const uint32_t* segb = sm.segb(slot);
        for (;;) {
            bool need = false;
            // full ring sweeps (collection.hpp:813-866): slot == lane, real sketches
            while (np < (uint32_t)kFilterBuffer && base + kRing <= S) {
                if (base + kRing > ce) { need = true; break; }
                const uint32_t ci = (base - cb + lane) * 4;
                const uint32_t pcs = cpc[ci >> 2];
                const uint32_t mask = __vcmpleu4(pcs, max_diff * 0x01010101u);
                const uint32_t cnt = (uint32_t)__popc(mask) >> 3;
                uint32_t total;
                uint32_t pos = np + warp_excl_scan3(cnt, total);
                if (cnt) {
                    const uint4 v = *reinterpret_cast<const uint4*>(cid + ci);
                    const uint2 sv = *reinterpret_cast<const uint2*>(csim + ci);
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
                    v = *reinterpret_cast<const uint4*>(cid + ci);
                    sv = *reinterpret_cast<const uint2*>(csim + ci);
                    const uint32_t pcs = cpc[ci >> 2];
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
                        if (um & (1u << e)) sm.tunk[atomicAdd(&ctrl->tail_unk, 1u)] = (uint16_t)pos;
                        pos++;
                    }
                }
                np += total;
                cand += 4ull * live;
                __syncwarp();
                // tail entries that failed the sketch filter of phase A but pass the index-as-sketch test: compute now
                const uint32_t ntu = __any_sync(0xffffffffu, um != 0) ? ctrl->tail_unk : 0u;
                if (ntu) {
                    for (uint32_t i = 0; i < ntu; i++) {
                        const uint32_t pp = sm.tunk[i];
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
            const uint32_t minval_before = minval16;
            maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, np);
            np = 0;
            if (minval16 != minval_before) {
                // filterer.hpp:108-111 through the threshold form of the host table: #{m : thr[m] > minval}
                max_diff = __popc(__ballot_sync(0xffffffffu, thr_lo > minval16)) + __popc(__ballot_sync(0xffffffffu, thr_hi > minval16));
                if (lane == 0) ctrl->max_diff = max_diff;  // advisory, for the producers
            }
            // stop rule (collection.hpp:927-943)
            const uint32_t pulled = base + kRing;
            uint32_t table_idx = L;
            if (pulled < S) {
                // upper_bound(segbase[0..L), pulled) - 1 by counting, 32 tables per ballot
                uint32_t cnt_le = 0;
                for (uint32_t t0 = 0; t0 < L; t0 += 32) {
                    const uint32_t t = t0 + lane;
                    cnt_le += __popc(__ballot_sync(0xffffffffu, t < L && segb[t] <= pulled));
                }
                table_idx = cnt_le - 1;
            }
            const float kth = __fdiv_rn((float)minval16, 65536.0f);
            const float sim = kth < max_sim ? max_sim : kth;  // std::max(kth, max_sim)
            uint32_t bin = (uint32_t)__fdiv_rn(sim, 0.005f);  // crosspolytope.hpp:116-118
            bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
            const uint32_t key = (depth - 1) * kEstBins + bin;
            if (key != stop_key) {
                stop_key = key;
                stop_word = lane < p.stop_words ? __ldg(a.stop + (uint64_t)key * p.stop_words + lane) : 0u;
            }
            const uint32_t word = __shfl_sync(0xffffffffu, stop_word, (table_idx >> 5) & 31u);
            if ((word >> (table_idx & 31)) & 1u) { status = kStopped; break; }
            if (!(base + kRing < S)) { status = kDepthDone; break; }
        }
        __syncwarp();
        if (lane == 0) {
            if (status == kStopped) ctrl->stopped = 1;
            mbar_arrive(&ctrl->empty[slot]);
        }
        it++;
        TADD(T_C, tc);
    }
    if (lane == 0) {
        ctrl->inserted = inserted;
        ctrl->minval16 = minval16;
        ctrl->candidates += cand;
        ctrl->distcomp += dcomp;
    }
}

// One PUFFINN query against cluster c by the whole CTA (collection.hpp:543-601 -> search_maps :768-948).
// Returns the number of results; sm.mb[0..cnt) holds them best first.
__device__ uint32_t probe_cluster_cta(const SearchParams& p, const CtaSmem& sm, const VisitArgs& a, uint32_t thr_lo, uint32_t thr_hi,
                                      uint32_t& it TARG) {
    const uint32_t L = p.g.L, k = p.k;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t off = p.offsets[a.c];
    const uint32_t nc = (uint32_t)(p.offsets[a.c + 1] - off);
    const uint32_t P = next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k);
    CtaCtrl* ctrl = sm.ctrl;
    TSTART(t0);

    if (a.memo) {
        uint4* mz = reinterpret_cast<uint4*>(a.memo);
        for (uint32_t i = tid; i < (nc + 7) / 8; i += kCtaThreads) mz[i] = make_uint4(0, 0, 0, 0);
    }
    // --- SearchBuffers ctor (collection.hpp:642-645): one thread per table
    for (uint32_t t = tid; t < L; t += kCtaThreads) {
        const uint32_t h = a.codes[(uint64_t)t * a.code_stride];
        uint32_t A;
        uint2 up, dn;
        table_anchor(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)a.c * L + t) * kDirEntries, nc, h, A, up, dn);
        sm.code[t] = h;
        sm.anchor[t] = A;
        sm.lcp_up[t] = up;
        sm.lcp_dn[t] = dn;
    }
    if (tid == 0) {
        ctrl->max_diff = kSketchBits;  // filterer.hpp:101
        ctrl->stopped = 0;
    }
    __syncthreads();
    TADD(T_SETUP, t0);
    if (warp == 0) consume_visit(p, sm, a, thr_lo, thr_hi, it TPASS);
    else produce_visit(p, sm, a, it TPASS);
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
                                                                uint64_t memo_stride, uint32_t prefetch_ahead) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const CtaSmem sm = carve_cta(s_dyn, p.g.L, p.k, p.g.sl);
    CtaCtrl* ctrl = sm.ctrl;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint64_t state_bytes = sizeof(QueryStateHeader) + (uint64_t)p.k * 8;
    uint16_t* memo = memo_base ? memo_base + (uint64_t)blockIdx.x * memo_stride : nullptr;
    const uint32_t P = next_pow2(2 * p.k) < 32 ? 32 : next_pow2(2 * p.k);
    if (tid == 0) {
        for (uint32_t i = 0; i < kSlots; i++) {
            mbar_init(&ctrl->full[i], kProducers);
            mbar_init(&ctrl->empty[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ctrl->unk_cnt = 0;
        ctrl->tail_unk = 0;
    }
    const uint32_t thr_lo = p.msd_thr[lane], thr_hi = p.msd_thr[lane + 32];
    uint32_t it = 0;  // chunks handed over so far (identical in producers and consumer at every visit boundary)
    TDECL;
    __syncthreads();

    for (;;) {
        if (tid == 0) ctrl->work = atomicAdd(b.work_counter, 1u);
        __syncthreads();
        const uint32_t w = ctrl->work;
        __syncthreads();
        if (w >= b.nq) break;
        const uint32_t q = (stop_at_foreign & 2) ? w : b.qperm[w];  // queries sorted by their nearest cluster (bit 1: debug, unsorted)
        if (prefetch_ahead) {
            // b.first holds the nearest cluster of every work item, in work order: whoever takes the first query of a
            // cluster prefetches it (cold start), and the cluster that comes up `prefetch_ahead` work items later
            if (w < prefetch_ahead && (w == 0 || b.first[w] != b.first[w - 1])) l2_prefetch_cluster(p, b.first[w] & 0xfffffu);
            const uint64_t wa = (uint64_t)w + prefetch_ahead;
            if (wa < b.nq && b.first[wa] != b.first[wa - 1]) l2_prefetch_cluster(p, b.first[wa] & 0xfffffu);
        }
        QueryStateHeader* st = reinterpret_cast<QueryStateHeader*>(b.state + (uint64_t)q * state_bytes);
        if (st->done) continue;
        TSTART(tq);
        TSTART(tm);
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
            if ((stop_at_foreign & 1) && p.owner[c] != p.shard_rank) break;
            last_key = nk;
            visited++;
            const uint64_t off = p.offsets[c];
            const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
            __syncthreads();  // everyone has read ctrl->nk / top before warp 0 may overwrite them
            if (p.brute[c]) {
                // index.rs:364-378 with brute_force_search :666-685: members in assignment order into a local top-k, then merge
                float* s_dist = reinterpret_cast<float*>(sm.cid(0));
                uint32_t* s_pid = sm.cid(0) + kCtaThreads;
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
                VisitArgs a;
                a.c = c;
                a.codes = b.codes + (uint64_t)fs * p.g.L * b.nq + q;
                a.code_stride = b.nq;
                a.stop = p.stop + (uint64_t)fs * kMaxHashBits * kEstBins * p.stop_words;
                a.max_sim = __fsub_rn(1.0f, __fdiv_rn(max_dist, 2.0f));  // puffinn_types.rs:77-79
                a.memo = (memo && nc <= memo_stride) ? memo : nullptr;
                if (tid < kNumSketches) sm.qsk[tid] = b.sketches[((uint64_t)fs * b.nq + q) * kNumSketches + tid];
                TADD(T_SELECT, tm);
                const uint32_t cnt = probe_cluster_cta(p, sm, a, thr_lo, thr_hi, it TPASS);
                TADD(T_VISIT, tm);
#ifdef CLANN_TIMING
                t_acc[T_VISITS]++;
#endif
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
                TADD(T_FINAL, tm);
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
        TADD(T_QUERY, tq);
    }
#ifdef CLANN_TIMING
    // representative threads: tid 0 (consumer lane 0 + CTA-level), tid 32 (producer 0)
    if (tid == 0) {
        const int slots[] = {T_SETUP, T_WAIT_FULL, T_C, T_VISIT, T_QUERY, T_SELECT, T_FINAL, T_VISITS};
        for (int i : slots) atomicAdd(&g_timing[i], t_acc[i]);
    }
    if (tid == 32) {
        const int slots[] = {T_RANGES, T_WAIT_EMPTY, T_A, T_B, T_HAND, T_CHUNKS};
        for (int i : slots) atomicAdd(&g_timing[i], t_acc[i]);
    }
#endif
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
    const int cap = (int)tune_get("probe_ctas", 0);  // knob: fewer resident queries = smaller L2 working set
    if (cap > 0 && cap < ctas_per_sm) ctas_per_sm = cap;
    uint64_t grid = (uint64_t)sm_count * ctas_per_sm;  // persistent: a whole number of CTAs per SM
    if (b.nq < grid) grid = b.nq;
    // per-CTA similarity memo (u16 per local id of the cluster being probed), from the index workspace
    uint16_t* use = (b.memo && grid <= b.memo_slots) ? b.memo : nullptr;
    const uint64_t stride = b.memo_stride;
    const int nosort = tune_get("probe_nosort", 0) ? 2 : 0;  // debug: measure what the nearest-cluster work order buys
    const int pf = (int)tune_get("probe_prefetch", -2);  // work items of lookahead for the L2 cluster prefetch; 0 = off, -2 = one grid
    // the prefetch follows the nearest-cluster work order, which only exists for a fresh batch (not for multi-GPU re-entry)
    const uint32_t ahead = stop_at_foreign ? 0u : (pf == -2 ? (uint32_t)grid : (uint32_t)pf);
    k_probe_cta<OCC><<<(unsigned)grid, kCtaThreads, smem, s>>>(p, b, (stop_at_foreign ? 1 : 0) | nosort, use, stride, ahead);
#ifdef CLANN_TIMING
    {
        CLANN_CUDA(cudaStreamSynchronize(s));
        unsigned long long h[T_N], z[T_N] = {0};
        CLANN_CUDA(cudaMemcpyFromSymbol(h, g_timing, sizeof(h)));
        CLANN_CUDA(cudaMemcpyToSymbol(g_timing, z, sizeof(z)));
        const char* names[T_N] = {"setup", "ranges", "wait_empty", "A", "B", "hand", "wait_full", "C", "visit", "query", "select", "final", "chunks", "visits"};
        const double nqd = (double)b.nq;
        fprintf(stderr, "[clann timing] grid=%llu per-query kcycles:", (unsigned long long)grid);
        for (int i = 0; i < T_N; i++) fprintf(stderr, " %s=%.2f", names[i], i >= T_CHUNKS ? h[i] / nqd : h[i] / nqd / 1e3);
        fprintf(stderr, "\n");
    }
#endif
}

void launch_probe_cta(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    if (b.nq == 0) return;
    int occ = (int)tune_get("probe_cta_occ", 4);  // knob: resident CTAs per SM the kernel is compiled for
    if (occ != 3 && occ != 4 && occ != 5 && occ != 6) occ = 4;
    switch (occ) {
        case 3: launch_probe_cta_occ<3>(p, b, stop_at_foreign, s); break;
        case 5: launch_probe_cta_occ<5>(p, b, stop_at_foreign, s); break;
        case 6: launch_probe_cta_occ<6>(p, b, stop_at_foreign, s); break;
        default: launch_probe_cta_occ<4>(p, b, stop_at_foreign, s); break;
    }
}

}  // namespace clann
