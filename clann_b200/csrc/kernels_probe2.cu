// k_probe2 — the CLANN search loop with one warp per query (as k_probe in kernels_search.cu, same results bit for bit),
// rescheduled so that FEW queries are in flight and each of them keeps MANY memory requests in flight.
//
// Why (measured on B200, glove-100 shape, 10 000 planted queries; profiles/exp/probe_sweep.py): k_probe needs 24 warps
// per SM to hide its dependent-load chains (about 500 us per query unloaded), and with 3 552 queries in flight the
// clusters being probed (3 MB each) do not stay in the L2: 10.3 GB of DRAM reads per launch. With 8 warps per SM the
// same kernel reads 4.9 GB, with 4 warps 3.8 GB — but then runs at the latency of its chains. k_probe2 shortens the
// chains instead of multiplying the warps:
//   * ring sweeps are loaded FOUR at a time (table indices, then the sketch word of every candidate) and kept in
//     registers as Hamming distances; the filter threshold only tightens during a visit (filterer.hpp:108-111), so the
//     sweeps are consumed in order with whatever threshold is in force, and what a batch does not consume is carried to
//     the next batch. The next group is requested before the rerank of the current batch and lands behind it;
//   * the Q15 rows a batch needs are fetched by TMA bulk copies (cp.async.bulk, one 2*SL-byte row per lane per issue,
//     mbarrier complete_tx) into a per-warp staging area — all rows of a batch in flight at once instead of eight per
//     round trip — and the similarity is computed out of shared memory;
//   * the per-visit similarity memo (one u16 per local id) lives in shared memory;
//   * anchors and per-depth ranges are evaluated for three tables per lane in lockstep (independent loads overlap).
// One CTA of 8 warps per SM by default (knobs probe2_warps / probe2_ctas); queries in nearest-cluster order.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <stdlib.h>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

namespace {


__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a copy that never lands must fail loudly, not hang the GPU
    } while (!ok);
}
// L2 eviction policy of the random reads of a visit (sketch words, Q15 rows). Measured (ncu, lts__t_sectors_*_evict_first_*):
// ld.global.nc.L1::no_allocate and cp.async.bulk default to EVICT_FIRST in the L2, which throws away exactly the sectors the
// other queries of the same cluster are about to touch; an explicit policy keeps them. kind: 0 evict_first, 1 evict_normal,
// 2 evict_last.
__device__ __forceinline__ uint64_t make_l2_policy(int kind) {
    uint64_t pol;
    if (kind == 0) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// TMA bulk copy global -> shared (SASS: UBLKCP); bytes a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(pol)
                 : "memory");
}
// Loads whose position in the instruction stream matters (they are issued before a wait and consumed after it).
__device__ __forceinline__ uint64_t ld_nc_u64_pinned(const uint64_t* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_nc_u32_pinned(const uint32_t* p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// An opaque zero: keeps arithmetic that consumes early-issued loads from being scheduled before the point it is produced.
__device__ __forceinline__ uint64_t opaque_zero64() {
    uint64_t z;
    asm volatile("mov.b64 %0, 0;" : "=l"(z));
    return z;
}

// Per-warp scratch carved out of dynamic shared memory.
struct W2 {
    uint8_t* stage;            // [stage_rows][2*sl] Q15 rows of the batch being reranked (TMA destination)
    int16_t* qrow;             // [sl] this warp's query in Q15
    uint16_t* memo;            // [memo_cap] similarity memo of the visit (dot + 32768, 0 = unknown)
    unsigned long long* mb;    // [P2K] MaxBuffer slots (maxbuffer.hpp:20): (sim16 << 32) | local id
    unsigned long long* heap;  // [k] TopKClosestHeap (src/core/heap.rs): (order_bits(dist) << 32) | point id
    unsigned long long* loc;   // [k] local heap of a brute-force cluster (index.rs:671)
    unsigned long long* mbar;  // [1] completion barrier of the row copies
    uint2* lcp_up;             // [L] 8 bytes: common-prefix length with the code at anchor + 12 j, j = 0..7
    uint2* lcp_dn;             // [L] same at anchor - 1 - 12 j
    uint32_t* pass_idx;        // [kPassingCap] passing-filter ids (collection.hpp:783)
    uint32_t* anchor;          // [L] lower bound of the query code in each table (prefixmap.hpp:36-57), unpadded position
    uint32_t* code;            // [L] query code per table
    uint32_t* start;           // [L] range start of the current depth
    uint32_t* segbase;         // [L+1] exclusive prefix of 4-entry segment counts of the current depth
    uint16_t* pass_sim;        // [kPassingCap] Q15 similarity of the passing ids as dot + 32768
    uint16_t* unk;             // [kPassingCap] positions of the passing list whose similarity is not memoised yet
};

struct Layout2 {
    uint32_t stage_rows, memo_cap, per_warp;
    int l2_policy;       // make_l2_policy kind for sketch words
    int l2_policy_rows;  // ... for Q15 rows
    int l2_policy_idx;   // ... for table indices
};

__host__ __device__ inline uint32_t p2k_of(uint32_t k) { return next_pow2(2 * k) < 32 ? 32 : next_pow2(2 * k); }

// bytes of everything but the staging area and the memo
__host__ __device__ inline uint32_t fixed_bytes2(uint32_t L, uint32_t k, uint32_t sl) {
    uint32_t b = sl * 2;                    // qrow (sl is a multiple of 16 -> 32-byte multiple)
    b += p2k_of(k) * 8 + k * 8 * 2 + 8;     // mb, heap, loc, mbar
    b += L * 8 * 2;                         // lcp_up, lcp_dn
    b += kPassingCap * 4 + L * 4 * 3 + (L + 1) * 4;
    b += kPassingCap * 2 * 2;
    return (b + 15) & ~15u;
}

__device__ __forceinline__ W2 carve2(uint8_t* base, uint32_t L, uint32_t k, uint32_t sl, const Layout2& lay) {
    W2 w;
    uint8_t* p = base;
    w.stage = p; p += (size_t)lay.stage_rows * sl * 2;
    w.qrow = reinterpret_cast<int16_t*>(p); p += sl * 2;
    w.memo = reinterpret_cast<uint16_t*>(p); p += (size_t)lay.memo_cap * 2;
    w.mb = reinterpret_cast<unsigned long long*>(p); p += p2k_of(k) * 8;
    w.heap = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    w.loc = reinterpret_cast<unsigned long long*>(p); p += k * 8;
    w.mbar = reinterpret_cast<unsigned long long*>(p); p += 8;
    w.lcp_up = reinterpret_cast<uint2*>(p); p += L * 8;
    w.lcp_dn = reinterpret_cast<uint2*>(p); p += L * 8;
    w.pass_idx = reinterpret_cast<uint32_t*>(p); p += kPassingCap * 4;
    w.anchor = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.code = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.start = reinterpret_cast<uint32_t*>(p); p += L * 4;
    w.segbase = reinterpret_cast<uint32_t*>(p); p += (L + 1) * 4;
    w.pass_sim = reinterpret_cast<uint16_t*>(p); p += kPassingCap * 2;
    w.unk = reinterpret_cast<uint16_t*>(p);
    return w;
}

// ------------------------------------------------------------------------------------------------ anchors and ranges

// fill_ranges (collection.hpp:650-667) with get_next_range (prefixmap.hpp:267-304) in closed form (table_range,
// probe_common.cuh), kNT tables per lane: start[] and the exclusive prefix of segment counts; returns the stream length.
__device__ __forceinline__ uint32_t ranges_lockstep(const SearchParams& p, const W2& sm, uint32_t c, uint64_t off, uint32_t nc,
                                                    uint32_t depth) {
    const uint32_t L = p.g.L, lane = lane_id();
    uint32_t running = 0;
    for (uint32_t t0 = 0; t0 < L; t0 += 32 * kNT) {
        uint32_t nseg[kNT], st[kNT];
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            const uint32_t t = t0 + 32 * j + lane;
            nseg[j] = 0;
            st[j] = 0;
            if (t < L)
                st[j] = table_range(p.tbl_hash + table_base(off, nc, L, t), p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries, nc, sm.code[t],
                                    sm.anchor[t], sm.lcp_up[t], sm.lcp_dn[t], depth, nseg[j]);
        }
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            const uint32_t t = t0 + 32 * j + lane;
            uint32_t total;
            const uint32_t ex = warp_excl_scan(nseg[j], total);
            if (t < L) {
                sm.start[t] = st[j];
                sm.segbase[t] = running + ex;
            }
            running += total;
        }
    }
    if (lane == 0) sm.segbase[L] = running;
    __syncwarp();
    return running;
}

// ------------------------------------------------------------------------------------------------ rerank

// Memo pass: similarities already known go straight to pass_sim; the positions still unknown are listed in unk.
__device__ __forceinline__ uint32_t rerank_lookup(const W2& sm, uint32_t count, const uint16_t* memo) {
    const uint32_t lane = lane_id();
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
    return nunk;
}

// Requests the Q15 rows of unk[c0 .. c0+cnt) into the staging area (one TMA bulk copy per row).
__device__ __forceinline__ void rows_request(const W2& sm, const int16_t* __restrict__ rows, uint32_t sl, uint32_t c0, uint32_t cnt,
                                             uint64_t pol) {
    const uint32_t lane = lane_id(), row_bytes = sl * 2;
    if (lane == 0) mbar_expect_tx(sm.mbar, cnt * row_bytes);
    __syncwarp();
    for (uint32_t r = lane; r < cnt; r += 32)
        bulk_g2s(sm.stage + (size_t)r * row_bytes, rows + (uint64_t)sm.pass_idx[sm.unk[c0 + r]] * sl, row_bytes, sm.mbar, pol);
}

// Q15 similarity (cosine.hpp:19-23 / math.hpp:11-44) of the cnt staged rows, G lanes per row -> pass_sim and the memo.
template <int G>
__device__ __forceinline__ void rows_compute(const W2& sm, uint32_t sl, uint32_t c0, uint32_t cnt, const int (&qreg)[8], bool qreg_valid,
                                             uint16_t* memo) {
    constexpr int CPI = 32 / G;
    const uint32_t lane = lane_id(), sub = lane % G, grp = lane / G, cpr = sl / 8, row_bytes = sl * 2;
    for (uint32_t r0 = 0; r0 < cnt; r0 += CPI * 4) {
        int part[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t r = r0 + u * CPI + grp;
            int s = 0;
            if (r < cnt) {
                const uint4* src = reinterpret_cast<const uint4*>(sm.stage + (size_t)r * row_bytes);
                if (qreg_valid) {
                    if (sub < cpr) {
                        const uint4 w = src[sub];
                        s += q15_mul(unpack_lo(w.x), qreg[0]); s += q15_mul(unpack_hi(w.x), qreg[1]);
                        s += q15_mul(unpack_lo(w.y), qreg[2]); s += q15_mul(unpack_hi(w.y), qreg[3]);
                        s += q15_mul(unpack_lo(w.z), qreg[4]); s += q15_mul(unpack_hi(w.z), qreg[5]);
                        s += q15_mul(unpack_lo(w.w), qreg[6]); s += q15_mul(unpack_hi(w.w), qreg[7]);
                    }
                } else {  // rows wider than 32 chunks (d > 256): the query comes from shared memory
                    for (uint32_t ch = sub; ch < cpr; ch += G) {
                        const uint4 a = src[ch];
                        const uint4 b = *reinterpret_cast<const uint4*>(sm.qrow + ch * 8);
                        s += q15_mul(unpack_lo(a.x), unpack_lo(b.x)); s += q15_mul(unpack_hi(a.x), unpack_hi(b.x));
                        s += q15_mul(unpack_lo(a.y), unpack_lo(b.y)); s += q15_mul(unpack_hi(a.y), unpack_hi(b.y));
                        s += q15_mul(unpack_lo(a.z), unpack_lo(b.z)); s += q15_mul(unpack_hi(a.z), unpack_hi(b.z));
                        s += q15_mul(unpack_lo(a.w), unpack_lo(b.w)); s += q15_mul(unpack_hi(a.w), unpack_hi(b.w));
                    }
                }
            }
            part[u] = s;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int s = part[u];
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            const uint32_t r = r0 + u * CPI + grp;
            if (sub == 0 && r < cnt) {
                const uint32_t pos = sm.unk[c0 + r];
                const uint16_t sim16 = (uint16_t)(s + 32768);
                sm.pass_sim[pos] = sim16;
                if (memo) memo[sm.pass_idx[pos]] = sim16;
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------ probe of one cluster

// One PUFFINN query against cluster c (collection.hpp:543-601 -> search_maps :768-948). On return sm.mb[0..cnt) holds the
// best entries, best first (maxbuffer.hpp:79-96). `phase` is the parity of the warp's mbarrier.
template <int G, int kAhead>
__device__ uint32_t probe_cluster2(const SearchParams& p, const W2& sm, const Layout2& lay, uint32_t c, const uint32_t* __restrict__ codes,
                                   uint64_t code_stride, uint64_t my_sketch, const uint32_t* __restrict__ stop, float max_sim,
                                   const int (&qreg)[8], bool qreg_valid, uint16_t* gmemo, uint64_t gmemo_stride, uint32_t& phase,
                                   ProbeCounters& ctr, uint16_t* prefilled) {
    const uint32_t L = p.g.L, k = p.k, sl = p.g.sl;
    const uint32_t lane = lane_id();
    const uint64_t off = p.offsets[c];
    const uint32_t nc = (uint32_t)(p.offsets[c + 1] - off);
    const uint32_t P = p2k_of(k);
    const int16_t* rows = p.q15 + off * sl;
    const uint64_t* sk = p.sketches + off * kNumSketches;

    const uint64_t pol = make_l2_policy(lay.l2_policy), pol_rows = make_l2_policy(lay.l2_policy_rows),
                   pol_idx = make_l2_policy(lay.l2_policy_idx);
    uint32_t inserted = 0, minval16 = 0, max_diff = kSketchBits;  // maxbuffer.hpp:53-55, filterer.hpp:101
    // similarity memo of this visit: shared memory when the cluster fits, else the global scratch, else none
    uint16_t* memo = nc <= lay.memo_cap ? sm.memo : ((gmemo && nc <= gmemo_stride) ? gmemo : nullptr);
    if (prefilled) memo = prefilled;  // first visit: similarities to the whole cluster computed in advance (launch_dense_sims)
    else if (memo) {
        uint4* mz = reinterpret_cast<uint4*>(memo);
        for (uint32_t i = lane; i < (nc + 7) / 8; i += 32) mz[i] = make_uint4(0, 0, 0, 0);
    }

    anchors_lockstep(p, sm, c, off, nc, codes, code_stride);

    bool stopped = false;
    for (uint32_t depth = kMaxHashBits; depth > 0 && !stopped; depth--) {
        const uint32_t S = ranges_lockstep(p, sm, c, off, nc, depth);
        if (S <= kRing) continue;  // the initial ring fill swallows the whole stream (collection.hpp:802-810)

        // stream segment number -> position of its first entry in tbl_idx
        auto locate = [&](uint32_t s) -> uint64_t {
            uint32_t lo = 0, len = L;  // upper_bound(segbase, s) - 1 over segbase[0..L)
            while (len > 0) {
                uint32_t half = len >> 1, mid = lo + half;
                if (sm.segbase[mid] <= s) { lo = mid + 1; len -= half + 1; } else { len = half; }
            }
            const uint32_t t = lo - 1;
            return table_base(off, nc, L, t) + sm.start[t] + 4 * (s - sm.segbase[t]);
        };

        // Sweeps ahead of the consumer: slot j holds ring sweep [base + 32 j, base + 32 j + 32) — this lane's segment
        // (ring slot == lane): its four table indices and the Hamming distance of each to the query sketch.
        uint32_t v[kAhead][4], ham[kAhead];
        uint32_t ahead = 0;  // valid slots
        uint32_t base = 0;   // first stream segment held by the ring

        auto load_idx = [&](uint32_t from, uint32_t to) {
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                if ((uint32_t)u >= from && (uint32_t)u < to) {
                    const uint32_t* seg = p.tbl_idx + locate(base + kRing * u + lane);
                    v[u][0] = ld_nc_u32_pinned(seg, pol_idx); v[u][1] = ld_nc_u32_pinned(seg + 1, pol_idx);
                    v[u][2] = ld_nc_u32_pinned(seg + 2, pol_idx); v[u][3] = ld_nc_u32_pinned(seg + 3, pol_idx);
                }
            }
        };
        uint64_t sw[kAhead][4];
        auto load_sketch = [&](uint32_t from, uint32_t to) {
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                if ((uint32_t)u >= from && (uint32_t)u < to) {
#pragma unroll
                    for (int e = 0; e < 4; e++) sw[u][e] = ld_nc_u64_pinned(sk + ((uint64_t)v[u][e] << 5 | lane), pol);
                }
            }
        };
        auto finish_ham = [&](uint32_t from, uint32_t to, uint64_t mine) {
#pragma unroll
            for (int u = 0; u < kAhead; u++) {
                if ((uint32_t)u >= from && (uint32_t)u < to) {
                    ham[u] = (uint32_t)__popcll(sw[u][0] ^ mine) | (uint32_t)__popcll(sw[u][1] ^ mine) << 8 |
                             (uint32_t)__popcll(sw[u][2] ^ mine) << 16 | (uint32_t)__popcll(sw[u][3] ^ mine) << 24;
                }
            }
        };

        do {
            uint32_t np = 0;
            while (np < kFilterBuffer && base + kRing <= S) {  // collection.hpp:813-866: a full ring sweep
                if (ahead == 0) {
                    const uint32_t avail = (S - base) / kRing;
                    ahead = avail < (uint32_t)kAhead ? avail : (uint32_t)kAhead;
                    load_idx(0, ahead);
                    load_sketch(0, ahead);
                    finish_ham(0, ahead, my_sketch);
                }
                const uint32_t hm = ham[0];
                const uint32_t p0 = (hm & 0xffu) <= max_diff, p1 = ((hm >> 8) & 0xffu) <= max_diff;
                const uint32_t p2 = ((hm >> 16) & 0xffu) <= max_diff, p3 = (hm >> 24) <= max_diff;
                uint32_t cnt = p0 + p1 + p2 + p3, total;
                uint32_t pos = np + warp_excl_scan(cnt, total);
                if (p0) sm.pass_idx[pos++] = v[0][0];
                if (p1) sm.pass_idx[pos++] = v[0][1];
                if (p2) sm.pass_idx[pos++] = v[0][2];
                if (p3) sm.pass_idx[pos++] = v[0][3];
                np += total;
                ctr.candidates += kRing * 4;
                base += kRing;
#pragma unroll
                for (int u = 0; u + 1 < kAhead; u++) {
                    v[u][0] = v[u + 1][0]; v[u][1] = v[u + 1][1]; v[u][2] = v[u + 1][2]; v[u][3] = v[u + 1][3];
                    ham[u] = ham[u + 1];
                }
                ahead--;
            }
            // top up the sweeps ahead: indices now, sketch words after the memo pass, Hamming distances after the rerank
            const uint32_t avail = (S - base) / kRing;
            const uint32_t from = ahead, to = avail < (uint32_t)kAhead ? avail : (uint32_t)kAhead;
            if (to > from) load_idx(from, to);
            // tail (collection.hpp:869-903): the not-yet-tested ring slots, in descending slot order, tested with the point
            // index itself in place of its sketch (:890-893)
            {
                const uint32_t live = (S - base) < (uint32_t)kRing ? (S - base) : (uint32_t)kRing;  // slots 0..live-1 hold segments base+slot
                uint32_t v0 = 0, v1 = 0, v2 = 0, v3 = 0, p0 = 0, p1 = 0, p2 = 0, p3 = 0;
                if (to > 0) {  // a full ring: slot 0 of the sweeps ahead
                    v0 = v[0][0]; v1 = v[0][1]; v2 = v[0][2]; v3 = v[0][3];
                } else if (lane < live) {
                    const uint32_t* seg = p.tbl_idx + locate(base + lane);
                    v0 = __ldg(seg); v1 = __ldg(seg + 1); v2 = __ldg(seg + 2); v3 = __ldg(seg + 3);
                }
                if (lane < live) {
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
            const uint32_t nunk = rerank_lookup(sm, np, memo);
            if (to > from) load_sketch(from, to);
            for (uint32_t c0 = 0; c0 < nunk; c0 += lay.stage_rows) {
                const uint32_t cnt = nunk - c0 < lay.stage_rows ? nunk - c0 : lay.stage_rows;
                rows_request(sm, rows, sl, c0, cnt, pol_rows);
                mbar_wait(sm.mbar, phase);
                phase ^= 1u;
                rows_compute<G>(sm, sl, c0, cnt, qreg, qreg_valid, memo);
            }
            if (to > from) {
                finish_ham(from, to, my_sketch | opaque_zero64());
                ahead = to;
            }
            maxbuffer_insert_list(sm.mb, P, k, inserted, minval16, sm.pass_idx, sm.pass_sim, np);
            ctr.distcomp += np;
            max_diff = p.msd[minval16 < 65536u ? minval16 : 65535u];  // filterer.hpp:108-111
            // stop rule (collection.hpp:927-943)
            uint32_t pulled = base + kRing;
            uint32_t table_idx = L;
            if (pulled < S) {
                uint32_t lo = 0, len = L;
                while (len > 0) {
                    uint32_t half = len >> 1, mid = lo + half;
                    if (sm.segbase[mid] <= pulled) { lo = mid + 1; len -= half + 1; } else { len = half; }
                }
                table_idx = lo - 1;
            }
            float kth = __fdiv_rn((float)minval16, 65536.0f);
            float sim = kth < max_sim ? max_sim : kth;  // std::max(kth, max_sim)
            uint32_t bin = (uint32_t)__fdiv_rn(sim, 0.005f);  // crosspolytope.hpp:116-118
            bin = bin > (uint32_t)(kEstBins - 1) ? (uint32_t)(kEstBins - 1) : bin;
            uint32_t word = __ldg(stop + ((uint64_t)(depth - 1) * kEstBins + bin) * p.stop_words + (table_idx >> 5));
            if ((word >> (table_idx & 31)) & 1u) {
                stopped = true;
                break;
            }
        } while (base + kRing < S);
    }
    // best_indices (collection.hpp:598, maxbuffer.hpp:79-96)
    maxbuffer_filter(sm.mb, P, k, inserted, minval16);
    return inserted;
}

// ------------------------------------------------------------------------------------------------ CLANN search loop

// src/core/index.rs:311-439 — one warp per query, pulled from a global counter in nearest-cluster order; persistent warps.
template <int G, int AH, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_probe2(SearchParams p, QueryBatch b, Layout2 lay, int stop_at_foreign, uint16_t* gmemo_base,
                                                   uint64_t gmemo_stride) {
    extern __shared__ __align__(16) uint8_t s_dyn[];
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const W2 sm = carve2(s_dyn + (size_t)warp * lay.per_warp, p.g.L, p.k, p.g.sl, lay);
    const uint64_t state_bytes = sizeof(QueryStateHeader) + (uint64_t)p.k * 8;
    const uint32_t cpr = p.g.sl / 8;
    const bool qreg_valid = cpr <= 32;
    uint16_t* gmemo = gmemo_base ? gmemo_base + ((uint64_t)blockIdx.x * (blockDim.x >> 5) + warp) * gmemo_stride : nullptr;
    if (lane == 0) {
        mbar_init(sm.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;

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
        ProbeCounters ctr{st->candidates, st->distcomp};
        for (uint32_t i = lane; i < heap_len; i += 32) sm.heap[i] = st_heap[i];
        // query row -> shared memory and this lane's 16-byte chunk -> registers
        for (uint32_t i = lane; i < p.g.sl / 2; i += 32)
            reinterpret_cast<uint32_t*>(sm.qrow)[i] = reinterpret_cast<const uint32_t*>(b.q15 + (uint64_t)q * p.g.sl)[i];
        __syncwarp();
        int qreg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (qreg_valid && (lane % G) < cpr) {
            uint4 w = *reinterpret_cast<const uint4*>(sm.qrow + (lane % G) * 8);
            qreg[0] = unpack_lo(w.x); qreg[1] = unpack_hi(w.x); qreg[2] = unpack_lo(w.y); qreg[3] = unpack_hi(w.y);
            qreg[4] = unpack_lo(w.z); qreg[5] = unpack_hi(w.z); qreg[6] = unpack_lo(w.w); qreg[7] = unpack_hi(w.w);
        }
        const float* qv = b.queries + (uint64_t)q * p.g.d;
        const float qn = b.qnorm[q];
        bool done = false;

        const float* cd = b.cdist + (uint64_t)q * p.K;
        for (; pos < p.K; pos++) {
            // next cluster of the stable ascending centre-distance order (index.rs:592-616): smallest key above last_key
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
            const uint32_t c = (uint32_t)nk;
            float max_dist = INFINITY;
            if (heap_len > 0) {  // index.rs:342-361
                unsigned long long top = topk_peek(sm.heap, heap_len);
                max_dist = float_from_order_bits((uint32_t)(top >> 32));
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
                const uint32_t P = p2k_of(p.k);
                for (uint32_t i = lane; i < P; i += 32) sm.mb[i] = i < loc_len ? ~sm.loc[i] : 0ull;  // ~ turns asc into desc
                __syncwarp();
                warp_sort_desc(sm.mb, P);
                for (uint32_t i = 0; i < loc_len; i++) {
                    unsigned long long key = ~sm.mb[i];
                    topk_add(sm.heap, heap_len, p.k, float_from_order_bits((uint32_t)(key >> 32)), (uint32_t)key);
                }
            } else {
                const uint32_t fs = p.fset_of[c];
                const float max_sim = __fsub_rn(1.0f, __fdiv_rn(max_dist, 2.0f));  // puffinn_types.rs:77-79
                const uint32_t* codes = b.codes + (uint64_t)fs * p.g.L * b.nq + q;
                const uint64_t my_sketch = b.sketches[((uint64_t)fs * b.nq + q) * kNumSketches + lane];
                const uint32_t* stop = p.stop + (uint64_t)fs * kMaxHashBits * kEstBins * p.stop_words;
                uint16_t* prefilled = (b.dense != nullptr && pos == 0 && !stop_at_foreign) ? b.dense + (uint64_t)q * b.dense_stride : nullptr;
                uint32_t cnt = probe_cluster2<G, AH>(p, sm, lay, c, codes, b.nq, my_sketch, stop, max_sim, qreg, qreg_valid, gmemo, gmemo_stride,
                                                 phase, ctr, prefilled);
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
                        topk_add(sm.heap, heap_len, p.k, dl, il);
                    }
                }
            }
        }
        if (pos >= p.K) done = true;
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
        }
        __syncwarp();
    }
}

int rerank_group2(uint32_t sl) {
    uint32_t cpr = sl / 8;
    int g = 2;
    while ((uint32_t)g < cpr && g < 32) g <<= 1;
    return g;
}

template <int G, int AH, int MAXT>
void launch_probe2_go(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev;
        CLANN_CUDA(cudaGetDevice(&dev));
        CLANN_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    uint32_t warps = (uint32_t)tune_get("probe2_warps", 8);  // knob: resident queries per CTA
    if (warps < 1 || warps > MAXT / 32) warps = MAXT / 32;
    uint32_t ctas = (uint32_t)tune_get("probe2_ctas", 1);    // knob: CTAs per SM
    if (ctas < 1 || ctas > 4) ctas = 1;
    // shared memory per warp: fixed scratch, then the memo (the largest cluster, if it fits), then the row staging area
    const uint32_t budget = (uint32_t)((size_t)226 * 1024 / ctas / warps) & ~15u;
    const uint32_t fixed = fixed_bytes2(p.g.L, p.k, p.g.sl);
    const uint32_t row_bytes = p.g.sl * 2;
    if (fixed + row_bytes > budget) throw std::invalid_argument("num_tables / k / dimension too large for the probe kernel's shared memory");
    Layout2 lay;
    uint32_t rest = budget - fixed;
    const uint32_t min_stage = row_bytes * 16 < rest ? row_bytes * 16 : row_bytes;
    const uint32_t memo_want = (p.max_cluster + 7u) & ~7u;
    lay.memo_cap = (memo_want * 2 + min_stage <= rest) ? memo_want : 0u;
    rest -= lay.memo_cap * 2;
    uint32_t want_rows = (uint32_t)tune_get("probe2_stage_rows", 64);
    if (want_rows < 1) want_rows = 1;
    lay.stage_rows = rest / row_bytes < want_rows ? rest / row_bytes : want_rows;
    lay.per_warp = (fixed + lay.memo_cap * 2 + lay.stage_rows * row_bytes + 15u) & ~15u;
    const size_t smem = (size_t)warps * lay.per_warp;
    static size_t configured = 0;
    if (smem > configured) {
        CLANN_CUDA(cudaFuncSetAttribute(k_probe2<G, AH, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int ctas_per_sm = 0;
    CLANN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_probe2<G, AH, MAXT>, warps * 32, smem));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if ((uint32_t)ctas_per_sm > ctas) ctas_per_sm = (int)ctas;
    uint64_t want = (b.nq + warps - 1) / warps;
    uint64_t grid = (uint64_t)sm_count * ctas_per_sm;  // persistent: a whole number of CTAs per SM
    if (want < grid) grid = want ? want : 1;
    // clusters beyond the shared-memory memo fall back to the global scratch of the index workspace
    uint16_t* gmemo = (b.memo && (uint64_t)grid * warps <= b.memo_slots && tune_get("probe_nomemo", 0) == 0) ? b.memo : nullptr;
    if (tune_get("probe_nomemo", 0) != 0) lay.memo_cap = 0;
    lay.l2_policy = (int)tune_get("probe2_l2", 1);  // knob: 0 evict_first (what the unhinted loads get), 1 evict_normal, 2 evict_last
    lay.l2_policy_rows = (int)tune_get("probe2_l2rows", 1);
    lay.l2_policy_idx = (int)tune_get("probe2_l2idx", 1);
    k_probe2<G, AH, MAXT><<<(unsigned)grid, warps * 32, smem, s>>>(p, b, lay, stop_at_foreign ? 1 : 0, gmemo, b.memo_stride);
}

// Register budget follows the CTA size: up to 8 warps may use 255 registers (four sweeps ahead), 12 warps 168, 16 warps 128
// (two sweeps ahead).
template <int G>
void launch_probe2_g(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    const uint32_t warps = (uint32_t)tune_get("probe2_warps", 8);
    if (warps > 12) launch_probe2_go<G, 2, 512>(p, b, stop_at_foreign, s);
    else if (warps > 8) launch_probe2_go<G, 4, 384>(p, b, stop_at_foreign, s);
    else launch_probe2_go<G, 4, 256>(p, b, stop_at_foreign, s);
}

}  // namespace

void launch_probe_pipelined(const SearchParams& p, const QueryBatch& b, bool stop_at_foreign, cudaStream_t s) {
    if (b.nq == 0) return;
    switch (rerank_group2(p.g.sl)) {
        case 2: launch_probe2_g<2>(p, b, stop_at_foreign, s); break;
        case 4: launch_probe2_g<4>(p, b, stop_at_foreign, s); break;
        case 8: launch_probe2_g<8>(p, b, stop_at_foreign, s); break;
        case 16: launch_probe2_g<16>(p, b, stop_at_foreign, s); break;
        default: launch_probe2_g<32>(p, b, stop_at_foreign, s); break;
    }
}

}  // namespace clann
