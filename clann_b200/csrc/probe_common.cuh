// Device helpers of the probe kernel (kernels_search.cu: one warp per query): per-query state, fp32 distance, warp primitives, TopKClosestHeap and MaxBuffer emulation, range helpers.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#pragma once

#include "kernels.h"

namespace clann {

// ------------------------------------------------------------------------------------------------ query state

// Per-query running state; also the unit exchanged between ranks in the multi-GPU stepping mode.
struct QueryStateHeader {
    uint32_t next_pos;   // number of clusters of the visiting order already consumed
    uint32_t done;       // 1 = search finished (early exit or all clusters seen)
    uint32_t heap_len;   // entries in the TopKClosestHeap
    uint32_t visited;    // clusters probed
    unsigned long long candidates;  // performance.hpp:82-86 summed over visits
    unsigned long long distcomp;    // performance.hpp:72-76 summed over visits
    unsigned long long last_key;    // (order_bits(centre distance) << 32 | cluster) of the last consumed cluster, 0 = none
    unsigned long long stop_point;  // last PUFFINN visit: (stop depth << 32) | table index at the stop (collection.hpp:927-943), 0 = no stop
    // followed by k x u64 heap keys: (order_bits(distance) << 32) | point id
};



// angulardata.rs:29-35
__device__ __forceinline__ float distance_point(const float* __restrict__ row, float row_norm, const float* __restrict__ q, float qn,
                                                uint32_t d) {
    float dot = ndarray_dot_thread(row, q, d);
    float cs = __fdiv_rn(dot, __fmul_rn(row_norm, qn));
    return __fsub_rn(1.0f, cs);
}


__host__ __device__ inline uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}


// ------------------------------------------------------------------------------------------------ warp helpers

__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, uint32_t& total) {
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((int)lane_id() >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// Bitonic sort (descending) of 32 u64 keys held one per lane: 15 shuffle steps, no shared memory.
__device__ __forceinline__ unsigned long long warp_sort_desc32(unsigned long long key) {
    unsigned long long x = ~key;  // ascending sort of the complement
    const uint32_t lane = lane_id();
#pragma unroll
    for (uint32_t kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, x, j);
            const bool up = (lane & kk) == 0;
            const bool keep_min = ((lane & j) == 0) == up;
            const unsigned long long lo = x < other ? x : other, hi = x < other ? other : x;
            x = keep_min ? lo : hi;
        }
    }
    return ~x;
}

// Bitonic sort (descending) of P (power of two) u64 keys in shared memory by one warp.
__device__ __forceinline__ void warp_sort_desc(unsigned long long* keys, uint32_t P) {
    if (P == 32) {
        unsigned long long k = keys[lane_id()];
        __syncwarp();
        keys[lane_id()] = warp_sort_desc32(k);
        __syncwarp();
        return;
    }
    for (uint32_t kk = 2; kk <= P; kk <<= 1) {
        for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane_id(); i < P; i += 32) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool desc = (i & kk) == 0;
                    if ((a < b) == desc) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}


// heap.rs:23-36 — bounded max-heap on (distance, index): push while not full, else replace the maximum iff the new
// distance is strictly smaller. Executed by the whole warp; `len` is warp-uniform.
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// returns what TopKClosestHeap::add returns (heap.rs:23-36): true when the element entered the heap (warp-uniform)
__device__ __forceinline__ bool topk_add(unsigned long long* heap, uint32_t& len, uint32_t cap, float dist, uint32_t id) {
    unsigned long long key = ((unsigned long long)float_order_bits(dist) << 32) | id;
    if (len < cap) {
        if (lane_id() == 0) heap[len] = key;
        len++;
        __syncwarp();
        return true;
    }
    unsigned long long best = 0;
    uint32_t where = 0;
    for (uint32_t i = lane_id(); i < len; i += 32) {
        unsigned long long v = heap[i];
        if (v >= best) {
            best = v;
            where = i;
        }
    }
    unsigned long long mx = warp_max_u64(best);
    // strict comparison on the distance only (heap.rs:27)
    const bool enters = (uint32_t)(key >> 32) < (uint32_t)(mx >> 32);
    if (enters) {
        uint32_t owner = __ffs(__ballot_sync(0xffffffffu, best == mx && lane_id() < len)) - 1;
        if (lane_id() == owner) heap[where] = key;
    }
    __syncwarp();
    return enters || len == 0;  // k == 0: `else if let Some(max)` finds nothing and add() returns true
}

__device__ __forceinline__ unsigned long long topk_peek(const unsigned long long* heap, uint32_t len) {
    unsigned long long best = 0;
    for (uint32_t i = lane_id(); i < len; i += 32) {
        unsigned long long v = heap[i];
        best = v > best ? v : best;
    }
    return warp_max_u64(best);
}

// maxbuffer.hpp:25-46 — filter(): sort by (value desc, id desc), drop adjacent duplicate ids, keep k,
// minval = k-th value iff k entries survive. mb holds `inserted` keys; slots up to P are scratch.
__device__ __forceinline__ void maxbuffer_filter(unsigned long long* mb, uint32_t P, uint32_t k, uint32_t& inserted, uint32_t& minval16) {
    for (uint32_t i = inserted + lane_id(); i < P; i += 32) mb[i] = 0;  // pad sorts last
    __syncwarp();
    warp_sort_desc(mb, P);
    uint32_t kept = 0;
    uint32_t prev_id = 0xffffffffu;
    bool have_prev = false;
    for (uint32_t base = 0; base < inserted; base += 32) {
        uint32_t i = base + lane_id();
        unsigned long long key = (i < inserted) ? mb[i] : 0;
        uint32_t id = (uint32_t)key;
        uint32_t left = __shfl_up_sync(0xffffffffu, id, 1);
        bool keep = i < inserted;
        if (lane_id() == 0) {
            if (have_prev && id == prev_id) keep = false;
        } else if (id == left) {
            keep = false;
        }
        uint32_t bal = __ballot_sync(0xffffffffu, keep);
        uint32_t pos = kept + __popc(bal & ((1u << lane_id()) - 1));
        __syncwarp();
        if (keep) mb[pos] = key;
        kept += __popc(bal);
        uint32_t last_valid = (inserted - base) < 32 ? (inserted - base - 1) : 31;
        prev_id = __shfl_sync(0xffffffffu, id, last_valid);
        have_prev = true;
        __syncwarp();
    }
    inserted = kept < k ? kept : k;
    if (inserted == k && k != 0) minval16 = (uint32_t)(mb[k - 1] >> 32);
    __syncwarp();
}

// maxbuffer.hpp:64-76 for a list of candidates in order: reject sim <= minval; an accepted entry that finds all 2k slots
// taken first runs filter() and is then stored WITHOUT being re-checked against the new minval (:68-75).
__device__ __forceinline__ void maxbuffer_insert_list(unsigned long long* mb, uint32_t P, uint32_t k, uint32_t& inserted,
                                                      uint32_t& minval16, const uint32_t* ids, const uint16_t* sims, uint32_t count) {
    const uint32_t lane = lane_id();
    for (uint32_t base = 0; base < count; base += 32) {
        const uint32_t i = base + lane;
        const uint32_t v = (i < count) ? sims[i] : 0;
        const uint32_t id = (i < count) ? ids[i] : 0;
        const unsigned long long key = ((unsigned long long)v << 32) | id;
        uint32_t todo = __ballot_sync(0xffffffffu, i < count);
        while (todo) {
            const bool accept = ((todo >> lane) & 1u) && v > minval16;
            const uint32_t acc = __ballot_sync(0xffffffffu, accept);
            if (!acc) break;
            const uint32_t space = 2 * k - inserted;
            const uint32_t nacc = __popc(acc);
            const uint32_t rank = __popc(acc & ((1u << lane) - 1));
            if (nacc <= space) {
                if (accept) mb[inserted + rank] = key;
                inserted += nacc;
                __syncwarp();
                break;
            }
            if (accept && rank < space) mb[inserted + rank] = key;
            inserted += space;
            __syncwarp();
            const uint32_t trig = __fns(acc, 0, space + 1);  // lane of the accepted entry that finds the buffer full
            maxbuffer_filter(mb, P, k, inserted, minval16);
            if (lane == trig) mb[inserted] = key;
            inserted += 1;
            __syncwarp();
            todo &= ~((2u << trig) - 1);
        }
    }
}


struct ProbeCounters {
    unsigned long long candidates, distcomp;
    uint32_t stop_point;  // of the last visit: (depth << 16) | table_idx where the stop rule fired, 0 = ran out of depths
};

__device__ __forceinline__ uint32_t lcp24(uint32_t a, uint32_t b) {
    uint32_t x = (a ^ b) & 0xffffffu;
    return x ? (uint32_t)(__clz(x) - 8) : 24u;
}

// number of leading samples (of 8 byte-packed lcp values) that are >= depth
__device__ __forceinline__ uint32_t lead_count(uint2 packed, uint32_t depth) {
    uint32_t rep = depth * 0x01010101u;
    uint32_t m0 = __vcmpgeu4(packed.x, rep), m1 = __vcmpgeu4(packed.y, rep);
    return (uint32_t)(__popc(m0) + __popc(m1)) >> 3;
}



// ------------------------------------------------------------------------------------------------ per-table probe state

// SearchBuffers ctor for one table (collection.hpp:642-645 -> prefixmap.hpp:36-57,250-260): the anchor A = lower bound of
// the query code h in the cluster's sorted codes H[0..nc), found inside bucket [dir[b], dir[b+1]) of its top 12 bits b
// (lower_bound == the reference's hinted halving search), plus the common-prefix length with the codes at
// A + 12 j and A - 1 - 12 j, j = 0..7, packed one byte each (beyond the data lie the 0xffffffff sentinels,
// prefixmap.hpp:215-226, which match nothing).
__device__ __forceinline__ void table_anchor(const uint32_t* __restrict__ H, const uint32_t* __restrict__ dir, uint32_t nc, uint32_t h,
                                             uint32_t& anchor, uint2& lcp_up, uint2& lcp_dn) {
    const uint32_t b = h >> (kMaxHashBits - kDirBits);
    uint32_t lo = __ldg(dir + b), len = __ldg(dir + b + 1) - lo;
    while (len > 8) {
        uint32_t half = len >> 1, mid = lo + half;
        if (__ldg(H + mid) < h) { lo = mid + 1; len -= half + 1; } else { len = half; }
    }
    uint32_t below = 0;  // the codes are sorted: the lower bound is the number of entries below h
#pragma unroll
    for (uint32_t j = 0; j < 8; j++)
        if (j < len) below += __ldg(H + lo + j) < h ? 1u : 0u;
    lo += below;
    anchor = lo;
    uint32_t up[8], dn[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t pu = lo + kSegment * j;
        up[j] = pu < nc ? lcp24(__ldg(H + pu), h) : 0u;
        int64_t pd = (int64_t)lo - 1 - kSegment * j;
        dn[j] = pd >= 0 ? lcp24(__ldg(H + pd), h) : 0u;
    }
    lcp_up = make_uint2(up[0] | up[1] << 8 | up[2] << 16 | up[3] << 24, up[4] | up[5] << 8 | up[6] << 16 | up[7] << 24);
    lcp_dn = make_uint2(dn[0] | dn[1] << 8 | dn[2] << 16 | dn[3] << 24, dn[4] | dn[5] << 8 | dn[6] << 16 | dn[7] << 24);
}

// PrefixMap::get_next_range (prefixmap.hpp:267-304) in closed form for one table at `depth`: with [lo, hi)
// the block of codes sharing the query's `depth`-bit prefix, upward [A, A + 12 ceil((hi - A)/12)) then the `>= len-12`
// clamp, downward [A - 12 ceil((A - lo)/12), A) then the `< 12` clamp; the anchors are never advanced (:278,292).
// The block comes from the directory when depth <= 12 bits (kDirBits), from the stride-12 samples otherwise (a block longer than the
// samples is resolved by a binary search confined to the code's bucket). Returns the range start; nseg = 4-entry segments.
__device__ __forceinline__ uint32_t table_range(const uint32_t* __restrict__ H, const uint32_t* __restrict__ dir, uint32_t nc, uint32_t h,
                                                uint32_t A, uint2 lcp_up, uint2 lcp_dn, uint32_t depth, uint32_t& nseg) {
    const uint32_t it = kMaxHashBits + 1 - depth;              // iteration 1..24
    const uint32_t dir_bit = 1u << (it >= 2 ? it - 2 : 0);     // removed bit (prefixmap.hpp:268-274)
    const bool upward = (h & dir_bit) == 0;
    const uint32_t b = h >> (kMaxHashBits - kDirBits);
    // j = 12-entry strides covered by the matching block on the probed side of the anchor. The common cases are written
    // without a branch on the direction (the tables of one query split about evenly between the two): one directory word for
    // depth <= kDirBits, the stride-12 samples otherwise.
    uint32_t j;
    if (depth <= kDirBits) {
        const uint32_t sh = kDirBits - depth, pb = b >> sh;
        const uint32_t edge = __ldg(dir + ((upward ? pb + 1 : pb) << sh));  // end of the block above / start of the block below
        const uint32_t gap = upward ? (edge > A ? edge - A : 0u) : (A > edge ? A - edge : 0u);
        j = (gap + kSegment - 1) / kSegment;
    } else {
        j = lead_count(upward ? lcp_up : lcp_dn, depth);
        if (j == 8) {
            if (upward) {  // first position >= A + 96 whose prefix differs; it cannot lie beyond the bucket of the top 12 bits
                uint32_t lo = A + 8 * kSegment, end = __ldg(dir + b + 1);
                uint32_t len = end > lo ? end - lo : 0;
                while (len > 0) {
                    uint32_t half = len >> 1, mid = lo + half;
                    if (lcp24(__ldg(H + mid), h) >= depth) { lo = mid + 1; len -= half + 1; } else { len = half; }
                }
                j = (lo - A + kSegment - 1) / kSegment;
            } else {       // first position of the matching block below A - 96, not before the bucket start
                const uint32_t beg = __ldg(dir + b);
                const uint32_t top = A >= 8 * kSegment ? A - 8 * kSegment : 0;  // positions [beg, top) undecided
                uint32_t lo = beg, len = top > beg ? top - beg : 0;
                while (len > 0) {  // lower_bound of "prefix matches" (monotone: false ... false true ... true)
                    uint32_t half = len >> 1, mid = lo + half;
                    if (lcp24(__ldg(H + mid), h) < depth) { lo = mid + 1; len -= half + 1; } else { len = half; }
                }
                j = (A - lo + kSegment - 1) / kSegment;
            }
        }
    }
    // prefixmap.hpp:277-290 (upward: start = A, end = A + 12 j, `end >= len - 12` clamp) and :291-303 (downward: end = A,
    // start = A - 12 j, `start < 12` clamp): either clamp takes one stride off the far end, never past the anchor.
    uint64_t span = (uint64_t)kSegment * j;
    const bool over = upward ? ((uint64_t)A + span >= (uint64_t)nc) : (span > (uint64_t)A);
    if (over) span = span > kSegment ? span - kSegment : 0;
    nseg = (uint32_t)(span >> 2);
    return upward ? A : (uint32_t)((uint64_t)A - span);
}

// ------------------------------------------------------------------------------------------------ TMA bulk copy + mbarrier

__device__ __forceinline__ uint32_t pc_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pc_mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pc_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pc_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pc_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pc_mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(pc_smem_addr(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();  // a copy that never lands must fail loudly, not hang the GPU
    } while (!ok);
}
// global -> shared bulk copy (SASS: UBLKCP), bytes a multiple of 16, both addresses 16-byte aligned. The destination may have
// been written through the generic proxy before (zeroing, memo updates): the proxy fence orders those writes before the copy.
__device__ __forceinline__ void pc_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pc_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(pc_smem_addr(bar))
                 : "memory");
}

constexpr int kNT = 3;  // tables per lane evaluated in lockstep (anchors, ranges)

// SearchBuffers ctor (collection.hpp:642-645 -> prefixmap.hpp:36-57,250-260) for kNT tables per lane in lockstep: the same
// anchors and stride-12 common-prefix samples as table_anchor (probe_common.cuh); the searches of the kNT tables advance
// together so that their loads overlap.
template <typename Smem>
__device__ __forceinline__ void anchors_lockstep(const SearchParams& p, const Smem& sm, uint32_t c, uint64_t off, uint32_t nc,
                                                 const uint32_t* __restrict__ codes, uint64_t code_stride) {
    const uint32_t L = p.g.L, lane = lane_id();
    for (uint32_t t0 = 0; t0 < L; t0 += 32 * kNT) {
        uint32_t h[kNT], lo[kNT], len[kNT];
        const uint32_t* H[kNT];
        bool valid[kNT];
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            const uint32_t t = t0 + 32 * j + lane;
            valid[j] = t < L;
            const uint32_t tt = valid[j] ? t : 0;
            h[j] = valid[j] ? __ldg(codes + (uint64_t)tt * code_stride) : 0u;
            H[j] = p.tbl_hash + table_base(off, nc, L, tt);
        }
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            const uint32_t t = valid[j] ? t0 + 32 * j + lane : 0;
            const uint32_t* dir = p.tbl_dir + ((uint64_t)c * L + t) * kDirEntries;
            const uint32_t b = h[j] >> (kMaxHashBits - kDirBits);
            lo[j] = __ldg(dir + b);
            len[j] = valid[j] ? __ldg(dir + b + 1) - lo[j] : 0u;
        }
        for (;;) {
            bool more = false;
#pragma unroll
            for (int j = 0; j < kNT; j++) more |= len[j] > 8;
            if (!more) break;
            uint32_t probe[kNT];
#pragma unroll
            for (int j = 0; j < kNT; j++) probe[j] = len[j] > 8 ? __ldg(H[j] + lo[j] + (len[j] >> 1)) : 0u;
#pragma unroll
            for (int j = 0; j < kNT; j++) {
                if (len[j] > 8) {
                    const uint32_t half = len[j] >> 1;
                    if (probe[j] < h[j]) { lo[j] += half + 1; len[j] -= half + 1; } else { len[j] = half; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            uint32_t below = 0;  // the codes are sorted: the lower bound is the number of entries below h
#pragma unroll
            for (uint32_t i = 0; i < 8; i++)
                if (i < len[j]) below += __ldg(H[j] + lo[j] + i) < h[j] ? 1u : 0u;
            lo[j] += below;
        }
#pragma unroll
        for (int j = 0; j < kNT; j++) {
            uint32_t up[8], dn[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const uint32_t pu = lo[j] + kSegment * i;
                up[i] = (valid[j] && pu < nc) ? lcp24(__ldg(H[j] + pu), h[j]) : 0u;
                const int64_t pd = (int64_t)lo[j] - 1 - kSegment * i;
                dn[i] = (valid[j] && pd >= 0) ? lcp24(__ldg(H[j] + pd), h[j]) : 0u;
            }
            if (valid[j]) {
                const uint32_t t = t0 + 32 * j + lane;
                sm.code[t] = h[j];
                sm.anchor[t] = lo[j];
                sm.lcp_up[t] = make_uint2(up[0] | up[1] << 8 | up[2] << 16 | up[3] << 24, up[4] | up[5] << 8 | up[6] << 16 | up[7] << 24);
                sm.lcp_dn[t] = make_uint2(dn[0] | dn[1] << 8 | dn[2] << 16 | dn[3] << 24, dn[4] | dn[5] << 8 | dn[6] << 16 | dn[7] << 24);
            }
        }
    }
    __syncwarp();
}

}  // namespace clann
