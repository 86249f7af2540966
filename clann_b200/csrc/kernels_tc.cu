// Tensor-core kernels (sm_100a: tcgen05.mma with TMEM accumulators, operands staged by TMA tensor copies).
//
// k_sketch_tc — the SimHash sketch projection (filterer.hpp:76-102: 2048 hyperplanes x SL Q15 multiply-round-accumulates per
// vector, math.hpp:37-44) with the contraction on the tensor pipe and the reference's result bit for bit.
//
// The reference's dot is R = sum_i ((a_i b_i + 2^14) >> 15) in a wrapping int16, and the sketch bit is R >= 0
// (simhash.hpp:41-44). The per-element rounding is not linear, so R itself is not a GEMM. But the EXACT integer
// S = sum_i a_i b_i is, and every rounded term differs from a_i b_i / 2^15 by at most one half:
//        2^15 R  in  (S - SL 2^14,  S + SL 2^14].
// Hence S >= SL 2^14 decides R > 0, S < -SL 2^14 decides R < 0 (with |S| < 2^30 - SL 2^14 ruling out int16 wrap-around),
// and only the band |S| < SL 2^14 — about 1.4 % of the (vector, hyperplane) pairs for d = 100 — has to be evaluated
// the reference's way, which is done right here on the CUDA cores from the same Q15 data. S is computed exactly in int32 from
// int8 slices a = 256 ah + al (ah signed, al unsigned; likewise b):
//        S = 2^16 sum ah bh + 2^8 (sum ah bl + sum al bh) + sum al bl
// = four `tcgen05.mma.kind::i8` products per K step into three TMEM accumulators (P1, P2, P0).
//
// One CTA = up to 128 consecutive vectors (MMA M = 128, TMEM lane = vector) against all 2048 hyperplanes in 16 tiles of N = 128.
//   warp 0     TMA producer: the vectors' slices once, then hyperplane slices through a 2-stage ring (cp.async.bulk.tensor, 128B swizzle)
//   warp 1     allocates TMEM, issues the MMAs (one elected thread), commits to mbarriers
//   warps 2-9  epilogue: tcgen05.ld the three accumulators, decide the bits, queue the undecided pairs, resolve them exactly,
//              write the 32 sketch words of every vector
#include <cuda.h>

#include "kernels.h"

namespace clann {

namespace tc {

constexpr uint32_t kTileM = 128;        // vectors per CTA
constexpr uint32_t kTileN = 128;        // hyperplanes per MMA tile
constexpr uint32_t kKBlock = 128;       // int8 elements per shared-memory K block (one 128-byte swizzle atom)
constexpr uint32_t kSliceBytes = kTileM * kKBlock;     // 16 KiB: one int8 slice tile
constexpr uint32_t kStages = 2;
constexpr uint32_t kEpiWarps = 8;
constexpr uint32_t kThreads = (2 + kEpiWarps) * 32;
constexpr uint32_t kListCap = 6144;     // undecided (vector, hyperplane) pairs queued per CTA (~3 700 expected at d = 100); beyond that a thread resolves its own
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a barrier that never completes must fail loudly (trap -> launch error), not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();
    } while (!ok);
}
// 2-D tiled TMA load (SASS: UTMALDG): box {128 bytes, 128 rows} at element coordinates (x, y) -> 128B-swizzled shared tile.
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* holder) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(kTmemCols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier receives one arrival when every MMA issued so far by this thread has completed (implies
// fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32 (SASS: UTCIMMA). M = 128, N = 128, K = 32 per instruction.
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// Shared-memory matrix descriptor of a K-major tile of 128-byte rows with the 128B swizzle (what the TMA box above writes):
// start address and offsets in 16-byte units; stride between 8-row groups = 1024 bytes; descriptor version 1 (sm_100);
// layout type 2 = SWIZZLE_128B. Advancing K by 32 int8 inside the swizzle atom = +32 bytes on the start address.
__device__ __forceinline__ uint64_t smem_desc(const void* tile, uint32_t k_byte_offset) {
    const uint32_t addr = smem_u32(tile) + k_byte_offset;
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor of kind::i8: D = s32 (bits 4-5 = 2), A / B signedness (bits 7-9 / 10-12: 1 = signed, 0 = unsigned),
// both operands K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24.
__device__ __forceinline__ uint32_t idesc_i8(uint32_t a_signed, uint32_t b_signed) {
    return (2u << 4) | (a_signed << 7) | (b_signed << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);
}
// 32 lanes x 32 consecutive 32-bit columns of TMEM -> 32 registers per thread (SASS: LDTM); lane = quarter * 32 + threadIdx % 32.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tc

// Q15 rows -> int8 slices for the tensor pipe: out[row][0 .. KP) = high bytes (signed), out[row][KP .. 2 KP) = low bytes
// (unsigned), zero-padded from SL to KP. One thread per (row, 8 elements).
__global__ void __launch_bounds__(256) k_q15_slices(const int16_t* __restrict__ q15, uint64_t rows, uint32_t sl, uint32_t kp,
                                                    uint8_t* __restrict__ out) {
    const uint32_t cpr = kp / 8;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cpr) return;
    const uint64_t row = i / cpr;
    const uint32_t ch = (uint32_t)(i % cpr);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ch * 8 < sl) v = __ldg(reinterpret_cast<const uint4*>(q15 + row * sl) + ch);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t hi[2], lo[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t a = w[2 * h], b = w[2 * h + 1];  // elements 4h .. 4h+3 (two per word)
        lo[h] = (a & 0xffu) | ((a >> 16) & 0xffu) << 8 | (b & 0xffu) << 16 | ((b >> 16) & 0xffu) << 24;
        hi[h] = ((a >> 8) & 0xffu) | ((a >> 24) & 0xffu) << 8 | ((b >> 8) & 0xffu) << 16 | ((b >> 24) & 0xffu) << 24;
    }
    uint8_t* dst = out + row * 2 * kp + ch * 8;
    *reinterpret_cast<uint2*>(dst) = make_uint2(hi[0], hi[1]);
    *reinterpret_cast<uint2*>(dst + kp) = make_uint2(lo[0], lo[1]);
}

// One 128-vector tile of the sketch projection. rows_map: 2-D uint8 tensor [rows][2 KP] of vector slices; planes_map: the same for
// the hyperplanes of all function sets ([n_fsets * 2048][2 KP]). q15 / planes: the Q15 originals for the exact evaluation of the
// undecided pairs. sketches[(out_row0 + r) * 32 + s] receives the words (function 64 s + b -> bit 63 - b, independent.hpp:77-84).

__global__ void __launch_bounds__(tc::kThreads, 1)
k_sketch_tc(const __grid_constant__ CUtensorMap rows_map, const __grid_constant__ CUtensorMap planes_map,
            const SketchTcTile* __restrict__ tiles, uint32_t slice_row_base, const int16_t* __restrict__ q15,
            const int16_t* __restrict__ planes, uint32_t sl, uint32_t kp, uint64_t* __restrict__ sketches) {
    using namespace tc;
    extern __shared__ uint8_t s_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t KB = kp / kKBlock;  // K blocks (1 for SL <= 128)
    // carve: A slices [KB][2] x 16 KiB, B ring [kStages][2] x 16 KiB, sketch words 32 KiB, pair list, barriers
    uint8_t* s_a = base;
    uint8_t* s_b = s_a + (size_t)KB * 2 * kSliceBytes;
    unsigned long long* s_out = reinterpret_cast<unsigned long long*>(s_b + (size_t)kStages * 2 * kSliceBytes);  // [128][32]
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_out + kTileM * kNumSketches);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_list + kListCap);
    uint64_t* a_full = bars;
    uint64_t* b_full = bars + 1;            // [kStages]
    uint64_t* b_empty = bars + 1 + kStages;  // [kStages]
    uint64_t* t_full = bars + 1 + 2 * kStages;
    uint64_t* t_empty = t_full + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(t_empty + 1);
    uint32_t* s_count = s_tmem + 1;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const SketchTcTile tile = tiles[blockIdx.x];
    constexpr uint32_t kTilesN = kNumPlanes / kTileN;  // 16
    // blockIdx.y splits the 16 plane tiles between CTAs (small batches: more CTAs than row tiles, shorter per-CTA latency)
    const uint32_t j_begin = blockIdx.y * (kTilesN / gridDim.y), j_end = j_begin + kTilesN / gridDim.y;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (uint32_t s = 0; s < kStages; s++) {
            mbar_init(b_full + s, 1);
            mbar_init(b_empty + s, 1);
        }
        mbar_init(t_full, 1);
        mbar_init(t_empty, kEpiWarps * 32);
        *s_count = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(s_tmem);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // vectors: both slices of every K block, once
            mbar_expect_tx(a_full, KB * 2 * kSliceBytes);
            for (uint32_t kb = 0; kb < KB; kb++) {
                tma_load_2d(s_a + (size_t)(kb * 2 + 0) * kSliceBytes, &rows_map, a_full, (int32_t)(kb * kKBlock), (int32_t)(tile.in_row0 - slice_row_base));
                tma_load_2d(s_a + (size_t)(kb * 2 + 1) * kSliceBytes, &rows_map, a_full, (int32_t)(kp + kb * kKBlock), (int32_t)(tile.in_row0 - slice_row_base));
            }
            // hyperplanes: ring over (tile of 128 planes, K block)
            uint32_t it = 0;
            for (uint32_t j = j_begin; j < j_end; j++) {
                for (uint32_t kb = 0; kb < KB; kb++, it++) {
                    const uint32_t s = it % kStages;
                    mbar_wait(b_empty + s, ((it / kStages) & 1u) ^ 1u);
                    mbar_expect_tx(b_full + s, 2 * kSliceBytes);
                    const int32_t y = (int32_t)(tile.fset * kNumPlanes + j * kTileN);
                    tma_load_2d(s_b + (size_t)(s * 2 + 0) * kSliceBytes, &planes_map, b_full + s, (int32_t)(kb * kKBlock), y);
                    tma_load_2d(s_b + (size_t)(s * 2 + 1) * kSliceBytes, &planes_map, b_full + s, (int32_t)(kp + kb * kKBlock), y);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id_ss = idesc_i8(1, 1), id_su = idesc_i8(1, 0), id_us = idesc_i8(0, 1), id_uu = idesc_i8(0, 0);
            mbar_wait(a_full, 0);
            uint32_t it = 0;
            for (uint32_t j = j_begin; j < j_end; j++) {
                mbar_wait(t_empty, ((j - j_begin) & 1u) ^ 1u);  // the epilogue has drained the accumulators of the previous tile
                fence_after_sync();
                for (uint32_t kb = 0; kb < KB; kb++, it++) {
                    const uint32_t s = it % kStages;
                    mbar_wait(b_full + s, (it / kStages) & 1u);
                    fence_after_sync();
                    const uint8_t* a_hi = s_a + (size_t)(kb * 2 + 0) * kSliceBytes;
                    const uint8_t* a_lo = s_a + (size_t)(kb * 2 + 1) * kSliceBytes;
                    const uint8_t* b_hi = s_b + (size_t)(s * 2 + 0) * kSliceBytes;
                    const uint8_t* b_lo = s_b + (size_t)(s * 2 + 1) * kSliceBytes;
#pragma unroll
                    for (uint32_t k = 0; k < kKBlock / 32; k++) {
                        const uint32_t ko = k * 32, first = (kb | k) == 0 ? 0u : 1u;
                        const uint64_t dah = smem_desc(a_hi, ko), dal = smem_desc(a_lo, ko);
                        const uint64_t dbh = smem_desc(b_hi, ko), dbl = smem_desc(b_lo, ko);
                        mma_i8(tmem + 0 * kTileN, dah, dbh, id_ss, first);   // P1 = sum ah bh
                        mma_i8(tmem + 1 * kTileN, dah, dbl, id_su, first);   // P2 = sum ah bl ...
                        mma_i8(tmem + 1 * kTileN, dal, dbh, id_us, 1u);      //      + sum al bh
                        mma_i8(tmem + 2 * kTileN, dal, dbl, id_uu, first);   // P0 = sum al bl
                    }
                    mma_commit(b_empty + s);  // the stage may be refilled once these MMAs have read it
                }
                mma_commit(t_full);
            }
        }
    } else {
        // ---- epilogue: warp e handles TMEM lanes of quarter (warp % 4) and the 64-plane half (e / 4) of every tile = one sketch word
        const uint32_t e = warp - 2;
        const uint32_t quarter = warp & 3u, half = e >> 2;
        const uint32_t r = quarter * 32 + lane;               // vector of the tile (TMEM lane)
        const bool live = r < tile.count;
        const int T = (int)(sl * 64u);                        // SL 2^14 / 2^8
        const int wrap_guard = (1 << 22) - T - 1;
        const uint32_t cpr = sl / 8;
        const uint32_t et = threadIdx.x - 64;                 // 0..255 among the epilogue threads
        for (uint32_t j = j_begin; j < j_end; j++) {
            mbar_wait(t_full, (j - j_begin) & 1u);
            fence_after_sync();
            unsigned long long word = 0, undecided = 0;
#pragma unroll
            for (uint32_t c0 = 0; c0 < 64; c0 += 32) {
                int p1[32], p2[32], p0[32];
                const uint32_t col = half * 64 + c0;
                const uint32_t taddr = tmem + ((quarter * 32u) << 16) + col;
                tmem_ld32(taddr + 0 * kTileN, p1);
                tmem_ld32(taddr + 1 * kTileN, p2);
                tmem_ld32(taddr + 2 * kTileN, p0);
                tmem_ld_wait();
                if (c0 == 32) {
                    // the accumulators are in registers: hand TMEM back to the MMA warp for the next tile
                    fence_before_sync();
                    mbar_arrive(t_empty);
                }
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    // Z = floor(S / 256): S = 2^16 P1 + 2^8 P2 + P0 lies in [256 Z, 256 Z + 255]
                    const int Z = p1[c] * 256 + p2[c] + (p0[c] >> 8);
                    const bool one = Z >= T && Z < wrap_guard;
                    const bool zero = Z < -T && Z > -wrap_guard;
                    const unsigned long long bit = 1ull << (63 - (c0 + c));
                    if (one) word |= bit;
                    if (!one && !zero) undecided |= bit;
                }
            }
            const uint32_t sk = j * 2 + half;  // sketch word of this half tile: planes 64 sk .. 64 sk + 63
            if (!live) undecided = 0;
            // queue the undecided pairs; when the list is full the thread evaluates its own (any input stays correct)
            while (undecided) {
                const uint32_t b = 63 - (uint32_t)__clzll(undecided);  // highest bit first: plane 64 sk + (63 - b)
                undecided &= ~(1ull << b);
                const uint32_t slot = atomicAdd(s_count, 1u);
                if (slot < kListCap) {
                    s_list[slot] = r << 16 | sk << 6 | (63 - b);
                } else {
                    const int16_t* x = q15 + (uint64_t)(tile.in_row0 + r) * sl;
                    const int16_t* y = planes + ((uint64_t)tile.fset * kNumPlanes + sk * 64 + (63 - b)) * sl;
                    int acc = 0;
                    for (uint32_t i = 0; i < sl; i++) acc += q15_mul((int)x[i], (int)y[i]);
                    if ((int16_t)acc >= 0) word |= 1ull << b;
                }
            }
            s_out[r * kNumSketches + sk] = word;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        // ---- resolve the queued pairs exactly (math.hpp:37-44 on the Q15 originals): 8 lanes per pair, four pairs in flight per
        // group, so that the L2 round trips of 128 pairs overlap
        {
            const uint32_t npairs = min(*s_count, kListCap);
            const uint32_t sub = lane & 7u, grp = et >> 3;   // 32 groups of 8 lanes
            for (uint32_t base0 = 0; base0 < npairs; base0 += 128) {  // trip count uniform over the warp: the loop shuffles
                int part[4];
                uint32_t ent[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = base0 + grp * 4 + u;
                    ent[u] = i < npairs ? s_list[i] : 0xffffffffu;
                    part[u] = 0;
                    if (ent[u] == 0xffffffffu) continue;
                    const uint32_t rr = ent[u] >> 16, f = ent[u] & 0xffffu;
                    const uint4* x = reinterpret_cast<const uint4*>(q15 + (uint64_t)(tile.in_row0 + rr) * sl);
                    const uint4* y = reinterpret_cast<const uint4*>(planes + ((uint64_t)tile.fset * kNumPlanes + f) * sl);
                    int s = 0;
                    for (uint32_t ch = sub; ch < cpr; ch += 8) {
                        const uint4 a = __ldg(x + ch), bq = __ldg(y + ch);
                        s += q15_mul(unpack_lo(a.x), unpack_lo(bq.x)); s += q15_mul(unpack_hi(a.x), unpack_hi(bq.x));
                        s += q15_mul(unpack_lo(a.y), unpack_lo(bq.y)); s += q15_mul(unpack_hi(a.y), unpack_hi(bq.y));
                        s += q15_mul(unpack_lo(a.z), unpack_lo(bq.z)); s += q15_mul(unpack_hi(a.z), unpack_hi(bq.z));
                        s += q15_mul(unpack_lo(a.w), unpack_lo(bq.w)); s += q15_mul(unpack_hi(a.w), unpack_hi(bq.w));
                    }
                    part[u] = s;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    int s = part[u];
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (sub == 0 && ent[u] != 0xffffffffu && (int16_t)s >= 0) {
                        const uint32_t rr = ent[u] >> 16, f = ent[u] & 0xffffu;
                        atomicOr(&s_out[rr * kNumSketches + (f >> 6)], 1ull << (63 - (f & 63u)));
                    }
                }
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        // ---- this CTA's words of every live vector
        const uint32_t w0 = j_begin * 2, nw = (j_end - j_begin) * 2;
        for (uint32_t i = et; i < tile.count * nw; i += kEpiWarps * 32) {
            const uint32_t rr = i / nw, wq = w0 + i % nw;
            sketches[((uint64_t)tile.out_row0 + rr) * kNumSketches + wq] = s_out[rr * kNumSketches + wq];
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        tc::tmem_dealloc(tmem);
    }
}

// ------------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CLANN_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess) throw CudaError("cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows][2 KP] uint8, box = 128 bytes x 128 rows, 128B swizzle; rows beyond the tensor read as zero
static CUtensorMap slice_map(const uint8_t* base, uint64_t rows, uint32_t kp) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)2 * kp, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)2 * kp};
    const cuuint32_t box[2] = {tc::kKBlock, tc::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}

uint32_t sketch_tc_kp(uint32_t sl) { return (sl + tc::kKBlock - 1) / tc::kKBlock * tc::kKBlock; }

bool sketch_tc_supported(uint32_t sl) { return sl % 8 == 0 && sketch_tc_kp(sl) <= 2 * tc::kKBlock; }

void launch_q15_slices(const int16_t* q15, uint64_t rows, uint32_t sl, uint8_t* out, cudaStream_t s) {
    if (rows == 0) return;
    const uint32_t kp = sketch_tc_kp(sl);
    const uint64_t threads = rows * (kp / 8);
    k_q15_slices<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(q15, rows, sl, kp, out);
}

// Host tiles are RowTile-like (in_row0, out_row0, count <= 128, fset); `tiles` is a device array of SketchTcTile.
// row_slices covers vector rows [slice_row_base, slice_row_base + slice_rows) of q15; plane_slices all function sets.
void launch_sketch_tc(const void* tiles, uint32_t n_tiles, const uint8_t* row_slices, uint32_t slice_row_base, uint64_t slice_rows,
                      const uint8_t* plane_slices, uint32_t n_fsets, const int16_t* q15, const int16_t* planes, uint32_t sl,
                      uint64_t* sketches, cudaStream_t s) {
    if (n_tiles == 0) return;
    const uint32_t kp = sketch_tc_kp(sl);
    const CUtensorMap rows_map = slice_map(row_slices, slice_rows, kp);
    const CUtensorMap planes_map = slice_map(plane_slices, (uint64_t)n_fsets * kNumPlanes, kp);
    const uint32_t KB = kp / tc::kKBlock;
    const size_t smem = 1024 + (size_t)KB * 2 * tc::kSliceBytes + (size_t)tc::kStages * 2 * tc::kSliceBytes +
                        (size_t)tc::kTileM * kNumSketches * 8 + (size_t)tc::kListCap * 4 + 128;
    ensure_dynamic_smem((const void*)(k_sketch_tc), smem);
    // few row tiles (a query batch): split the plane tiles over blockIdx.y until the grid covers the GPU about twice
    int sms = 0, dev = 0;
    CLANN_CUDA(cudaGetDevice(&dev));
    CLANN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t split = 1;
    while (split < 4 && (uint64_t)n_tiles * split < 2ull * (uint32_t)sms) split *= 2;
    k_sketch_tc<<<dim3(n_tiles, split), tc::kThreads, smem, s>>>(rows_map, planes_map, static_cast<const SketchTcTile*>(tiles), slice_row_base, q15,
                                                    planes, sl, kp, sketches);
}


// ================================================================================================ centre scoring GEMM
//
// src/core/index.rs:592-600 scores a query against every centre: K dot products of length d. For a batch that is the
// nq x K x d GEMM north_star asks to put on the tensor cores. The reference's values are fp32 in ndarray's summation order,
// and they decide the visiting order and the prune test bit for bit, so the tensor pipe cannot simply replace them. It
// produces a SCREEN instead: tf32 products (kind::tf32, fp32 accumulate in TMEM) give every distance to within
// kCentreEps; k_center_refine then evaluates exactly (the reference's arithmetic) the 32 nearest candidates of each query —
// the only ones a search ever touches unless it walks more than 32 clusters — and publishes, per query, the bound below which
// every stored value is exact. The probe kernel re-evaluates a query's whole row exactly the moment its walk reaches that bound.

namespace tc {
constexpr uint32_t kGemmKBlock = 32;                    // fp32 elements per 128-byte swizzle row
constexpr uint32_t kGemmTile = kTileM * 128;            // bytes of one 128 x 32 fp32 tile
constexpr uint32_t kGemmMaxKB = 4;                      // K blocks resident at once: d <= 128 in one pass, more in rounds

// D[tmem] (+)= A[smem] * B[smem]^T with tf32 operands (SASS: UTCHMMA.tf32 family), M = 128, N = 128, K = 8 per instruction.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32: D = f32 (bits 4-5 = 1), A and B formats = 2 (TF32), K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((kTileN >> 3) << 17) | ((kTileM >> 4) << 24);
}
}  // namespace tc

// One CTA = 128 queries x 128 centres: approx[q * K + c] = 1 - dot_tf32(query q, centre c) / (|q| |c|).
// q_map: 2-D fp32 tensor [nq][d], c_map: [K][d]; box = 32 floats x 128 rows, 128B swizzle; out-of-range rows and columns read 0.
__global__ void __launch_bounds__(192, 1)
k_center_gemm_tc(const __grid_constant__ CUtensorMap q_map, const __grid_constant__ CUtensorMap c_map, uint64_t nq, uint32_t K,
                 uint32_t d, const float* __restrict__ qnorm, const float* __restrict__ cnorm, float* __restrict__ approx) {
    using namespace tc;
    extern __shared__ uint8_t s_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(s_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = base;                                   // [kGemmMaxKB] query tiles
    uint8_t* s_b = base + (size_t)kGemmMaxKB * kGemmTile;  // [kGemmMaxKB] centre tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + (size_t)kGemmMaxKB * kGemmTile);
    uint64_t* full = bars;       // operands of the current round landed
    uint64_t* drained = bars + 1;  // the MMAs of the current round have read them
    uint64_t* done = bars + 2;   // accumulator complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q0 = blockIdx.x * kTileM, c0 = blockIdx.y * kTileN;
    const uint32_t nkb = (d + kGemmKBlock - 1) / kGemmKBlock;
    const uint32_t rounds = (nkb + kGemmMaxKB - 1) / kGemmMaxKB;
    if (threadIdx.x == 0) {
        mbar_init(full, 1);
        mbar_init(drained, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        // 128 columns of TMEM for the fp32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "n"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem = *s_tmem;
    if (warp == 0) {
        if (lane == 0) {
            for (uint32_t r = 0; r < rounds; r++) {
                if (r > 0) mbar_wait(drained, (r - 1) & 1u);
                const uint32_t kb0 = r * kGemmMaxKB, kbn = min(kGemmMaxKB, nkb - kb0);
                mbar_expect_tx(full, kbn * 2 * kGemmTile);
                for (uint32_t i = 0; i < kbn; i++) {
                    tma_load_2d(s_a + (size_t)i * kGemmTile, &q_map, full, (int32_t)((kb0 + i) * kGemmKBlock), (int32_t)q0);
                    tma_load_2d(s_b + (size_t)i * kGemmTile, &c_map, full, (int32_t)((kb0 + i) * kGemmKBlock), (int32_t)c0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t id = idesc_tf32();
            for (uint32_t r = 0; r < rounds; r++) {
                mbar_wait(full, r & 1u);
                fence_after_sync();
                const uint32_t kb0 = r * kGemmMaxKB, kbn = min(kGemmMaxKB, nkb - kb0);
                for (uint32_t i = 0; i < kbn; i++) {
#pragma unroll
                    for (uint32_t k = 0; k < kGemmKBlock / 8; k++)
                        mma_tf32(tmem, smem_desc(s_a + (size_t)i * kGemmTile, k * 32), smem_desc(s_b + (size_t)i * kGemmTile, k * 32), id,
                                 (r | i | k) == 0 ? 0u : 1u);
                }
                mma_commit(drained);
            }
            mma_commit(done);
        }
    } else {
        // epilogue warps 2..5: TMEM lane quarter = warp % 4, thread = query row
        const uint32_t quarter = warp & 3u;
        const uint32_t q = q0 + quarter * 32 + lane;
        mbar_wait(done, 0);
        fence_after_sync();
        const float qn = q < nq ? qnorm[q] : 1.0f;
#pragma unroll 1
        for (uint32_t cc = 0; cc < kTileN; cc += 32) {
            int v[32];
            tmem_ld32(tmem + ((quarter * 32u) << 16) + cc, v);
            tmem_ld_wait();
            if (q < nq) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const uint32_t c = c0 + cc + j;
                    if (c < K) approx[(uint64_t)q * K + c] = __fsub_rn(1.0f, __fdiv_rn(__int_as_float(v[j]), __fmul_rn(cnorm[c], qn)));
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        fence_after_sync();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(128) : "memory");
    }
}

// [rows][d] fp32, box = 32 floats x 128 rows, 128B swizzle
static CUtensorMap f32_map(const float* base, uint64_t rows, uint32_t d) {
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
    const cuuint32_t box[2] = {tc::kGemmKBlock, tc::kTileM};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled (fp32) failed (" + std::to_string((int)r) + ")");
    return m;
}

bool center_gemm_tc_supported(uint32_t d) { return d % 4 == 0 && d >= 4; }

void launch_center_gemm_tc(const float* queries, const float* qnorm, uint64_t nq, const float* center_rows, const float* center_norms,
                           uint32_t K, uint32_t d, float* approx, cudaStream_t s) {
    if (nq == 0 || K == 0) return;
    const CUtensorMap qm = f32_map(queries, nq, d), cm = f32_map(center_rows, K, d);
    const size_t smem = 1024 + (size_t)2 * tc::kGemmMaxKB * tc::kGemmTile + 64;
    ensure_dynamic_smem((const void*)(k_center_gemm_tc), smem);
    dim3 grid((unsigned)((nq + tc::kTileM - 1) / tc::kTileM), (K + tc::kTileN - 1) / tc::kTileN);
    k_center_gemm_tc<<<grid, 192, smem, s>>>(qm, cm, nq, K, d, qnorm, center_norms, approx);
}

}  // namespace clann
