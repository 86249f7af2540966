// Build-side kernels: greedy k-center (gmm.rs), Q15 store, SimHash sketches, FHT cross-polytope table codes,
// segmented stable radix sort of the hash tables, Monte-Carlo collision estimates.
// Citations are file:line into /root/reference (libpuffinn/include/puffinn unless a src/ path is given).
#include <cuda_fp16.h>

#include "kernels.h"

namespace clann {

// ------------------------------------------------------------------------------------------------ CLANN layer

// src/metricdata/angulardata.rs:12-19 — norms[i] = sqrt(row.dot(row)); 8 lanes per row.
__global__ void __launch_bounds__(256) k_row_norms(const float* __restrict__ data, uint64_t n, uint32_t d, float* __restrict__ norms) {
    uint64_t row = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    uint64_t r = row < n ? row : n - 1;
    const float* x = data + r * d;
    float dot = ndarray_dot_group8(x, x, d);
    if (row < n && (threadIdx.x & 7) == 0) norms[row] = __fsqrt_rn(dot);
}

__device__ __forceinline__ uint64_t gmm_key(float dist, uint64_t i) {
    // arg-max with first-max tie rule (src/core/gmm.rs:5-15): larger distance wins, then the smaller index.
    return ((uint64_t)float_order_bits(dist) << 32) | (uint64_t)(0xffffffffu - (uint32_t)i);
}
__device__ __forceinline__ uint32_t gmm_key_index(uint64_t key) { return 0xffffffffu - (uint32_t)(key & 0xffffffffu); }

// src/core/gmm.rs:40-53 — one pass: distances of all points to centre c, strict-< reassignment, arg-max of the updated
// distances (which selects centre c+1). Memory-bound: streams n*d floats once.
__global__ void __launch_bounds__(256) k_gmm_pass(const float* __restrict__ data, const float* __restrict__ norms, uint64_t row0, uint64_t n,
                                                  uint32_t d, uint32_t c, uint64_t* __restrict__ keys, float* __restrict__ dist,
                                                  uint32_t* __restrict__ assign) {
    __shared__ uint64_t s_key[8];
    const uint32_t ci = (c == 0) ? 0u : gmm_key_index(keys[c - 1]);
    uint64_t row = row0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3);  // rows [row0, n) of the full arrays
    const bool valid = row < n;
    uint64_t r = valid ? row : n - 1;
    const float* x = data + r * d;
    const float* y = data + (uint64_t)ci * d;
    float dot = ndarray_dot_group8(x, y, d);
    // angulardata.rs:25-27: 1.0 - (dot / (norms[i] * norms[j]))
    float nd = __fsub_rn(1.0f, __fdiv_rn(dot, __fmul_rn(norms[r], norms[ci])));
    uint64_t key = 0;
    if (valid && (threadIdx.x & 7) == 0) {
        float cur;
        if (c == 0) {
            cur = nd;
            dist[row] = nd;
            assign[row] = 0;
        } else {
            cur = dist[row];
            if (nd < cur) {  // gmm.rs:49
                cur = nd;
                dist[row] = nd;
                assign[row] = c;
            }
        }
        key = gmm_key(cur, row);
    }
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane_id() == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t best = s_key[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = s_key[w] > best ? s_key[w] : best;
        atomicMax((unsigned long long*)&keys[c], (unsigned long long)best);
    }
}

// Distances from the centre chosen for pass c to every earlier centre j < c (angulardata.rs:25-27), for the triangle-inequality
// filter of k_gmm_pass_v. One thread per earlier centre.
__global__ void __launch_bounds__(128) k_gmm_centre_dists(const float* __restrict__ data, const float* __restrict__ norms, uint32_t d,
                                                          uint32_t c, const uint64_t* __restrict__ keys, float* __restrict__ cc) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= c) return;
    const uint32_t ci = gmm_key_index(keys[c - 1]);
    const uint32_t cj = j == 0 ? 0u : gmm_key_index(keys[j - 1]);
    const float dot = ndarray_dot_thread(data + (uint64_t)cj * d, data + (uint64_t)ci * d, d);
    cc[j] = __fsub_rn(1.0f, __fdiv_rn(dot, __fmul_rn(norms[cj], norms[ci])));
}

// gmm.rs:40-53, vectorised (d % 4 == 0): two lanes per row, each streaming its half of every 8-element chunk as one 128-bit load
// (a full 32-byte sector per row and load instruction, 16 rows per warp in flight), the centre row in shared memory. The
// arithmetic is ndarray's unrolled_dot exactly: lane s owns the partial sums p[4s .. 4s+3], the fold (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7)
// and the tail run on the even lane.
// PRUNE: a row is not read at all when the triangle inequality already rules out a reassignment. With chord lengths
// e(x,y) = sqrt(2 dist(x,y)) (a metric on directions), e(row, new) >= e(centre_of_row, new) - e(row, centre_of_row); if
// e(centre_of_row, new) >= 2 e(row, centre_of_row) + 0.03 the new distance exceeds the current one by more than any rounding of the
// fp32 evaluations involved (each within ~1e-5 of the exact value, i.e. within 4.5e-3 in chord length), so the reference's
// strict `new < old` (gmm.rs:49) is false and nothing changes for that row. Every comparison with a NaN (zero vectors) is
// false, which keeps such rows on the evaluated path.
template <bool PRUNE>
__global__ void __launch_bounds__(256) k_gmm_pass_v(const float* __restrict__ data, const float* __restrict__ norms, uint64_t row0, uint64_t n,
                                                    uint32_t d, uint32_t c, uint64_t* __restrict__ keys, float* __restrict__ dist,
                                                    uint32_t* __restrict__ assign, const float* __restrict__ cc) {
    extern __shared__ __align__(16) float s_y[];  // centre row
    __shared__ uint64_t s_key[8];
    const uint32_t ci = (c == 0) ? 0u : gmm_key_index(keys[c - 1]);
    for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) s_y[i] = data[(uint64_t)ci * d + i];
    __syncthreads();
    const uint64_t row = row0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1);  // rows [row0, n) of the full arrays
    const uint32_t s = threadIdx.x & 1u;
    const bool valid = row < n;
    const uint64_t r = valid ? row : n - 1;
    float cur = INFINITY;
    bool need = valid;
    if (c > 0 && valid) {
        cur = dist[r];
        if (PRUNE) {
            const float dcc = cc[assign[r]];
            const bool skip = sqrtf(2.0f * dcc) >= 2.0f * sqrtf(2.0f * fmaxf(cur, 0.0f)) + 0.03f;
            need = !skip;
        }
    }
    const uint32_t mask = __ballot_sync(0xffffffffu, need);
    if (need) {
        const float4* x = reinterpret_cast<const float4*>(data + r * d) + s;
        const float4* y = reinterpret_cast<const float4*>(s_y) + s;
        float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f, p3 = 0.0f;
        const uint32_t chunks = d / 8;
#pragma unroll 4
        for (uint32_t j = 0; j < chunks; j++) {
            const float4 a = __ldg(x + 2 * j), b = y[2 * j];
            p0 = __fadd_rn(p0, __fmul_rn(a.x, b.x));
            p1 = __fadd_rn(p1, __fmul_rn(a.y, b.y));
            p2 = __fadd_rn(p2, __fmul_rn(a.z, b.z));
            p3 = __fadd_rn(p3, __fmul_rn(a.w, b.w));
        }
        const float q0 = __shfl_xor_sync(mask, p0, 1), q1 = __shfl_xor_sync(mask, p1, 1);
        const float q2 = __shfl_xor_sync(mask, p2, 1), q3 = __shfl_xor_sync(mask, p3, 1);
        if (s == 0) {
            float sum = 0.0f;
            sum = __fadd_rn(sum, __fadd_rn(p0, q0));
            sum = __fadd_rn(sum, __fadd_rn(p1, q1));
            sum = __fadd_rn(sum, __fadd_rn(p2, q2));
            sum = __fadd_rn(sum, __fadd_rn(p3, q3));
            const float* xs = data + r * d;
            for (uint32_t i = chunks * 8; i < d; i++) sum = __fadd_rn(sum, __fmul_rn(xs[i], s_y[i]));
            const float nd = __fsub_rn(1.0f, __fdiv_rn(sum, __fmul_rn(norms[r], norms[ci])));  // angulardata.rs:25-27
            if (c == 0 || nd < cur) {  // gmm.rs:49
                cur = nd;
                dist[r] = nd;
                assign[r] = c;
            }
        }
    }
    uint64_t key = (valid && s == 0) ? gmm_key(cur, row) : 0;
    for (int o = 16; o > 0; o >>= 1) {
        const uint64_t other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane_id() == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t best = s_key[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = s_key[w] > best ? s_key[w] : best;
        atomicMax((unsigned long long*)&keys[c], (unsigned long long)best);
    }
}

// gmm.rs:56-60 radii, cluster sizes, and decoding of the centre list.
__global__ void k_gmm_finish(const uint64_t* __restrict__ keys, uint32_t K, uint64_t n, const float* __restrict__ dist,
                             const uint32_t* __restrict__ assign, uint32_t* __restrict__ centers, float* __restrict__ radii,
                             uint32_t* __restrict__ sizes) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) centers[i] = (i == 0) ? 0u : gmm_key_index(keys[i - 1]);
    if (i < n) {
        uint32_t a = assign[i];
        // radii start at 0.0 and take f32::max: as signed ints, negative floats order below 0 and positives order naturally.
        // f32::max ignores a NaN operand (gmm.rs:58-60; the distance of a zero vector is NaN)
        const float di = dist[i];
        if (di == di) atomicMax((int*)&radii[a], __float_as_int(di));
        atomicAdd(&sizes[a], 1u);
    }
}

// index.rs:188-192 — member lists in ascending point order = a stable partition of the rows by cluster id. One warp per chunk of
// consecutive rows keeps a private row of K counters (counts[chunk * K + c], L2-resident): count, exclusive scan per cluster over
// the chunks (starting at the cluster's offset), then the same walk again turns the counters into running write positions.
// Within a 32-row tile the rows of one cluster are ranked by lane (match_any), so the order inside a cluster is the row order.
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_assign_partition(const uint32_t* __restrict__ assign, uint64_t n, uint32_t K, uint64_t chunk_rows,
                                                          uint32_t n_chunks, uint32_t* __restrict__ counts, uint32_t* __restrict__ perm) {
    const uint32_t w = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31u;
    if (w >= n_chunks) return;
    const uint64_t lo = (uint64_t)w * chunk_rows, hi = lo + chunk_rows < n ? lo + chunk_rows : n;
    uint32_t* mine = counts + (uint64_t)w * K;
    for (uint64_t base = lo; base < hi; base += 32) {
        const uint64_t i = base + lane;
        const bool valid = i < hi;
        const uint32_t c = valid ? assign[i] : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, c);
        const uint32_t leader = __ffs(peers) - 1u;
        uint32_t start = 0;
        if (valid && lane == leader) {
            start = mine[c];
            mine[c] = start + __popc(peers);
        }
        if (SCATTER) {
            start = __shfl_sync(0xffffffffu, start, leader);
            if (valid) perm[start + __popc(peers & ((1u << lane) - 1u))] = (uint32_t)i;
        }
        __syncwarp();  // the next tile's leader of the same cluster may be another lane
    }
}

// one warp per cluster: counts[b * K + c] -> offsets[c] + (rows of cluster c in the chunks before b)
__global__ void __launch_bounds__(256) k_assign_scan(uint32_t* __restrict__ counts, uint32_t K, uint32_t n_chunks,
                                                     const uint64_t* __restrict__ offsets) {
    const uint32_t c = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31u;
    if (c >= K) return;
    uint32_t running = (uint32_t)offsets[c];
    for (uint32_t b0 = 0; b0 < n_chunks; b0 += 32) {
        const uint32_t b = b0 + lane;
        const uint32_t v = b < n_chunks ? counts[(uint64_t)b * K + c] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((int)lane >= o) incl += t;
        }
        if (b < n_chunks) counts[(uint64_t)b * K + c] = running + incl - v;
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
}

void launch_assign_partition(const uint32_t* assign, uint64_t n, uint32_t K, const uint64_t* offsets, uint32_t* counts, uint32_t n_chunks,
                             uint32_t* perm, cudaStream_t s) {
    if (n == 0 || K == 0 || n_chunks == 0) return;
    uint64_t chunk_rows = (n + n_chunks - 1) / n_chunks;
    chunk_rows = (chunk_rows + 31) & ~31ull;
    CLANN_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_chunks * K * sizeof(uint32_t), s));
    const unsigned grid = (unsigned)(((uint64_t)n_chunks * 32 + 255) / 256);
    k_assign_partition<false><<<grid, 256, 0, s>>>(assign, n, K, chunk_rows, n_chunks, counts, perm);
    k_assign_scan<<<(unsigned)(((uint64_t)K * 32 + 255) / 256), 256, 0, s>>>(counts, K, n_chunks, offsets);
    k_assign_partition<true><<<grid, 256, 0, s>>>(assign, n, K, chunk_rows, n_chunks, counts, perm);
}

__global__ void k_gather_rows(const float* __restrict__ data, const uint32_t* __restrict__ rows, uint32_t count, uint32_t d,
                              float* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (uint64_t)count * d) {
        uint32_t r = (uint32_t)(i / d), j = (uint32_t)(i % d);
        out[i] = data[(uint64_t)rows[r] * d + j];
    }
}

// fp16 rows -> fp32 (exact): the ingest path of BASELINE.json's 100M x 96 fp16 configuration.
__global__ void __launch_bounds__(256) k_widen_f16(const uint16_t* __restrict__ in, uint64_t count, float* __restrict__ out) {
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (i + 1 < count) {
        const __half2 h = *reinterpret_cast<const __half2*>(in + i);
        const float2 f = __half22float2(h);
        out[i] = f.x;
        out[i + 1] = f.y;
    } else if (i < count) {
        out[i] = __half2float(*reinterpret_cast<const __half*>(in + i));
    }
}

void launch_widen_f16(const uint16_t* in, uint64_t count, float* out, cudaStream_t s) {
    if (count) k_widen_f16<<<(unsigned)(((count + 1) / 2 + 255) / 256), 256, 0, s>>>(in, count, out);
}

// ------------------------------------------------------------------------------------------------ Q15 store

// format/unit_vector.hpp:61-89. The sum of squares follows the shape g++ -O3 gives the reference loop (:71-74):
// products rounded then added in order for the first d - d%4 elements, FMAs for the last d%4 (see oracle/clann_oracle.c).
__global__ void __launch_bounds__(256) k_store_q15(const float* __restrict__ data, const uint32_t* __restrict__ perm, uint64_t rows,
                                                   uint32_t d, uint32_t sl, int16_t* __restrict__ q15) {
    uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const float* v = data + (uint64_t)(perm ? perm[row] : row) * d;
    float acc = 0.0f;
    const uint32_t body = d & ~3u;
    for (uint32_t i = 0; i < body; i++) acc = __fadd_rn(acc, __fmul_rn(v[i], v[i]));
    for (uint32_t i = body; i < d; i++) acc = __fmaf_rn(v[i], v[i], acc);
    const float len = __fsqrt_rn(acc);
    int16_t* out = q15 + row * sl;
    for (uint32_t i = 0; i < d; i++) {
        float x = v[i];
        if (len != 0.0f) x = __fdiv_rn(x, len);
        out[i] = to_q15(x);
    }
    for (uint32_t i = d; i < sl; i++) out[i] = 0;
}

// ------------------------------------------------------------------------------------------------ SimHash sketches

// filterer.hpp:76-102 + simhash.hpp:41-44 + independent.hpp:70-86 (function 64*s+b -> bit 63-b of sketch s).
// One CTA = one tile of <=32 rows x 256 hyperplanes (4 sketches). Thread = hyperplane; rows live in shared memory as
// int32 so that each 16-element unit costs 4 broadcast LDS.128 + 48 integer ops per row.
__global__ void __launch_bounds__(256, 3) k_sketch(const int16_t* __restrict__ q15, const RowTile* __restrict__ tiles,
                                                const int16_t* __restrict__ planes, uint32_t sl, uint64_t* __restrict__ sketches) {
    extern __shared__ int s_rows[];  // [32][sl]
    const RowTile tile = tiles[blockIdx.x];
    for (uint32_t e = threadIdx.x; e < 32 * sl; e += blockDim.x) {
        uint32_t p = e / sl, i = e % sl;
        s_rows[e] = (p < tile.count) ? (int)q15[(uint64_t)(tile.in_row0 + p) * sl + i] : 0;
    }
    __syncthreads();
    const uint32_t f = blockIdx.y * 256 + threadIdx.x;  // hyperplane 0..2047
    const uint4* prow = reinterpret_cast<const uint4*>(planes + ((uint64_t)tile.fset * kNumPlanes + f) * sl);
    int acc[32];
#pragma unroll
    for (int p = 0; p < 32; p++) acc[p] = 0;
    for (uint32_t u = 0; u < sl / 16; u++) {
        uint4 w0 = __ldg(prow + 2 * u), w1 = __ldg(prow + 2 * u + 1);
        int a[16];
        a[0] = unpack_lo(w0.x); a[1] = unpack_hi(w0.x); a[2] = unpack_lo(w0.y); a[3] = unpack_hi(w0.y);
        a[4] = unpack_lo(w0.z); a[5] = unpack_hi(w0.z); a[6] = unpack_lo(w0.w); a[7] = unpack_hi(w0.w);
        a[8] = unpack_lo(w1.x); a[9] = unpack_hi(w1.x); a[10] = unpack_lo(w1.y); a[11] = unpack_hi(w1.y);
        a[12] = unpack_lo(w1.z); a[13] = unpack_hi(w1.z); a[14] = unpack_lo(w1.w); a[15] = unpack_hi(w1.w);
#pragma unroll
        for (int p = 0; p < 32; p++) {
            const int4* pv = reinterpret_cast<const int4*>(s_rows + p * sl + u * 16);
            int4 v0 = pv[0], v1 = pv[1], v2 = pv[2], v3 = pv[3];
            int s = acc[p];
            s += q15_mul(a[0], v0.x); s += q15_mul(a[1], v0.y); s += q15_mul(a[2], v0.z); s += q15_mul(a[3], v0.w);
            s += q15_mul(a[4], v1.x); s += q15_mul(a[5], v1.y); s += q15_mul(a[6], v1.z); s += q15_mul(a[7], v1.w);
            s += q15_mul(a[8], v2.x); s += q15_mul(a[9], v2.y); s += q15_mul(a[10], v2.z); s += q15_mul(a[11], v2.w);
            s += q15_mul(a[12], v3.x); s += q15_mul(a[13], v3.y); s += q15_mul(a[14], v3.z); s += q15_mul(a[15], v3.w);
            acc[p] = s;
        }
    }
    // bit = dot >= 0 on the wrapping int16 accumulator (math.hpp:37-44, simhash.hpp:43)
    uint32_t mine = 0;
#pragma unroll
    for (int p = 0; p < 32; p++) {
        uint32_t bal = __ballot_sync(0xffffffffu, (int16_t)acc[p] >= 0);
        if ((int)lane_id() == p) mine = __brev(bal);  // lane (plane b) -> bit 31-b of this 32-bit half
    }
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t sk = blockIdx.y * 4 + (warp >> 1);
    const uint32_t half = warp & 1;  // planes 0..31 of the sketch are its high word
    if (lane_id() < tile.count) {
        uint32_t* out32 = reinterpret_cast<uint32_t*>(sketches);
        out32[((uint64_t)(tile.out_row0 + lane_id()) * kNumSketches + sk) * 2 + (half ? 0 : 1)] = mine;
    }
}

// ------------------------------------------------------------------------------------------------ FHT cross-polytope codes

// external/ffht/fht_avx.c:107-195,375-567: unnormalised in-place WHT, strides 1,2,4,... in that order, fl(u+v), fl(u-v).
template <int N, bool UNROLL>
__device__ __forceinline__ void fht_inplace(float* x) {
    if constexpr (UNROLL) {
#pragma unroll
        for (int h = 1; h < N; h <<= 1) {
#pragma unroll
            for (int i = 0; i < N; i += 2 * h) {
#pragma unroll
                for (int j = i; j < i + h; j++) {
                    float u = x[j], v = x[j + h];
                    x[j] = __fadd_rn(u, v);
                    x[j + h] = __fsub_rn(u, v);
                }
            }
        }
    } else {
#pragma unroll 1
        for (int h = 1; h < N; h <<= 1) {
#pragma unroll 1
            for (int i = 0; i < N; i += 2 * h) {
#pragma unroll 4
                for (int j = i; j < i + h; j++) {
                    float u = x[j], v = x[j + h];
                    x[j] = __fadd_rn(u, v);
                    x[j + h] = __fsub_rn(u, v);
                }
            }
        }
    }
}

// One FHT-CP function value (crosspolytope.hpp:187-209, encode_closest_axis :131-144) for the row held by this lane.
// sb: sign bits of this function, [3][W] words, bit i set = sign -1 (random_signs, :160-165).
template <int M>
__device__ __forceinline__ uint32_t fht_cp_hash(const int16_t* row, uint32_t d, const uint32_t* __restrict__ sb) {
    constexpr int N = 1 << M;
    constexpr int W = (N + 31) / 32;
    constexpr bool UNROLL = (M <= 7);
    float x[N];
    if constexpr (UNROLL) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = (i < (int)d) ? __fmul_rn((float)row[i], 1.0f / 32768.0f) : 0.0f;
    } else {
        for (int i = 0; i < N; i++) x[i] = (i < (int)d) ? __fmul_rn((float)row[i], 1.0f / 32768.0f) : 0.0f;
    }
#pragma unroll 1
    for (int r = 0; r < kRotations; r++) {
        if constexpr (UNROLL) {
#pragma unroll
            for (int i = 0; i < N; i++) {
                uint32_t w = __ldg(sb + r * W + (i >> 5));
                x[i] = __uint_as_float(__float_as_uint(x[i]) ^ ((w << (31 - (i & 31))) & 0x80000000u));
            }
        } else {
            for (int i = 0; i < N; i++) {
                uint32_t w = __ldg(sb + r * W + (i >> 5));
                x[i] = __uint_as_float(__float_as_uint(x[i]) ^ ((w << (31 - (i & 31))) & 0x80000000u));
            }
        }
        fht_inplace<N, UNROLL>(x);
    }
    int res = 0;
    float best = 0.0f;
    if constexpr (UNROLL) {
#pragma unroll
        for (int i = 0; i < N; i++) {
            float v = x[i];
            if (v > best) { res = i; best = v; }
            else if (-v > best) { res = i + N; best = -v; }
        }
    } else {
        for (int i = 0; i < N; i++) {
            float v = x[i];
            if (v > best) { res = i; best = v; }
            else if (-v > best) { res = i + N; best = -v; }
        }
    }
    return (uint32_t)res;
}

// independent.hpp:70-86: table code = fph function values concatenated MSB-first, >> bits_to_cut.
// CTA = one tile of <=32 rows; lane = row, warp = table (so sign bits are warp-uniform loads).
template <int M>
__global__ void __launch_bounds__(256) k_codes(const int16_t* __restrict__ q15, const RowTile* __restrict__ tiles,
                                               const uint32_t* __restrict__ signbits, HashGeom g, uint32_t* __restrict__ codes,
                                               uint64_t code_stride, uint64_t fset_stride) {
    extern __shared__ int16_t s_q[];  // [32][sl + 2] (odd word stride: conflict-free column reads)
    constexpr int N = 1 << M;
    constexpr int W = (N + 31) / 32;
    const RowTile tile = tiles[blockIdx.x];
    const uint32_t stride = g.sl + 2;
    for (uint32_t e = threadIdx.x; e < 32 * g.sl; e += blockDim.x) {
        uint32_t p = e / g.sl, i = e % g.sl;
        s_q[p * stride + i] = (p < tile.count) ? q15[(uint64_t)(tile.in_row0 + p) * g.sl + i] : (int16_t)0;
    }
    __syncthreads();
    const uint32_t t = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= g.L) return;
    const int16_t* row = s_q + lane_id() * stride;
    uint64_t code = 0;
    for (uint32_t f = 0; f < g.fph; f++) {
        const uint32_t func = t * g.fph + f;
        const uint32_t* sb = signbits + ((uint64_t)tile.fset * (g.L * g.fph) + func) * (kRotations * W);
        uint32_t h = fht_cp_hash<M>(row, g.d, sb);
        code = (code << g.bpf) | h;
    }
    code >>= g.cut;
    if (lane_id() < tile.count) {
        if (tile.code_stride) codes[tile.code_base + (uint64_t)t * tile.code_stride + lane_id()] = (uint32_t)code;
        else codes[(uint64_t)tile.fset * fset_stride + (uint64_t)t * code_stride + tile.out_row0 + lane_id()] = (uint32_t)code;
    }
}

// The same codes for 2^M = 32..128 points with the per-function overheads taken out of the instruction stream (the FHT itself
// is 3 x N log2 N / ... adds and cannot shrink): the 32 rows of the tile are converted to float ONCE per CTA (the plain
// kernel re-reads and converts the int16 row for each of the L*fph functions), the +-1 diagonals of a table are expanded to
// sign-bit masks in shared memory once per warp (one XOR per element instead of shift + and + xor), and the signed arg-max
// (crosspolytope.hpp:131-144: strict comparisons, the first maximum of |v| wins, code = index + N for a negative winner) is a
// left-biased tournament on (value, index) — 3 instructions per comparison instead of 6 per element.
template <int M>
__global__ void __launch_bounds__(256) k_codes_fast(const int16_t* __restrict__ q15, const RowTile* __restrict__ tiles,
                                                    const uint32_t* __restrict__ signbits, HashGeom g, uint32_t* __restrict__ codes,
                                                    uint64_t code_stride, uint64_t fset_stride) {
    constexpr int N = 1 << M;
    constexpr int W = (N + 31) / 32;
    constexpr int PITCH = N + 4;  // words per row: 16-byte aligned, conflict-free LDS.128 with lane = row
    extern __shared__ __align__(16) uint8_t s_raw[];
    float* s_f = reinterpret_cast<float*>(s_raw);                                      // [32][PITCH]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_raw + (size_t)32 * PITCH * 4);     // [8 warps][fph][3][N]
    const RowTile tile = tiles[blockIdx.x];
    for (uint32_t e = threadIdx.x; e < 32 * N; e += blockDim.x) {
        const uint32_t r = e / N, i = e % N;
        const float v = (r < tile.count && i < g.d) ? __fmul_rn((float)q15[(uint64_t)(tile.in_row0 + r) * g.sl + i], 1.0f / 32768.0f) : 0.0f;
        s_f[r * PITCH + i] = v;
    }
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t t = blockIdx.y * (blockDim.x >> 5) + warp;
    uint32_t* mymask = s_mask + (size_t)warp * g.fph * kRotations * N;
    if (t < g.L) {
        const uint32_t* sb = signbits + ((uint64_t)tile.fset * (g.L * g.fph) + (uint64_t)t * g.fph) * (kRotations * W);
        for (uint32_t e = lane; e < g.fph * kRotations * N; e += 32) {
            const uint32_t fr = e / N, i = e % N;  // fr = function * 3 + rotation
            mymask[e] = ((__ldg(sb + fr * W + (i >> 5)) >> (i & 31)) & 1u) << 31;
        }
    }
    __syncthreads();
    if (t >= g.L) return;
    const float4* row4 = reinterpret_cast<const float4*>(s_f + lane * PITCH);
    uint64_t code = 0;
    for (uint32_t f = 0; f < g.fph; f++) {
        float x[N];
        const uint4* mk = reinterpret_cast<const uint4*>(mymask + (size_t)f * kRotations * N);
#pragma unroll
        for (int i = 0; i < N / 4; i++) {
            if ((i & 3) == 0) asm volatile("" ::: "memory");  // keep the loads from being hoisted en bloc (255 registers + spills)
            const float4 v = row4[i];
            const uint4 m = mk[i];
            x[4 * i + 0] = __uint_as_float(__float_as_uint(v.x) ^ m.x);
            x[4 * i + 1] = __uint_as_float(__float_as_uint(v.y) ^ m.y);
            x[4 * i + 2] = __uint_as_float(__float_as_uint(v.z) ^ m.z);
            x[4 * i + 3] = __uint_as_float(__float_as_uint(v.w) ^ m.w);
        }
        fht_inplace<N, true>(x);
#pragma unroll 1
        for (int r = 1; r < kRotations; r++) {
            const uint4* mr = mk + r * (N / 4);
#pragma unroll
            for (int i = 0; i < N / 4; i++) {
                if ((i & 3) == 0) asm volatile("" ::: "memory");
                const uint4 m = mr[i];
                x[4 * i + 0] = __uint_as_float(__float_as_uint(x[4 * i + 0]) ^ m.x);
                x[4 * i + 1] = __uint_as_float(__float_as_uint(x[4 * i + 1]) ^ m.y);
                x[4 * i + 2] = __uint_as_float(__float_as_uint(x[4 * i + 2]) ^ m.z);
                x[4 * i + 3] = __uint_as_float(__float_as_uint(x[4 * i + 3]) ^ m.w);
            }
            fht_inplace<N, true>(x);
        }
        // tournament in blocks of 16 (keeps the index registers few): the left entry wins ties, so the lowest index among equal
        // |v| survives, as in the sequential scan
        float bv = 0.0f;
        int bi = 0;
#pragma unroll
        for (int c = 0; c < N / 16; c++) {
            float v[8];
            int id[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const bool right = fabsf(x[16 * c + 2 * i + 1]) > fabsf(x[16 * c + 2 * i]);
                v[i] = right ? x[16 * c + 2 * i + 1] : x[16 * c + 2 * i];
                id[i] = 16 * c + 2 * i + (right ? 1 : 0);
            }
#pragma unroll
            for (int len = 4; len >= 1; len >>= 1) {
#pragma unroll
                for (int i = 0; i < len; i++) {
                    const bool right = fabsf(v[2 * i + 1]) > fabsf(v[2 * i]);
                    v[i] = right ? v[2 * i + 1] : v[2 * i];
                    id[i] = right ? id[2 * i + 1] : id[2 * i];
                }
            }
            const bool right = c > 0 && fabsf(v[0]) > fabsf(bv);
            if (c == 0 || right) {
                bv = v[0];
                bi = id[0];
            }
        }
        const uint32_t hval = (uint32_t)bi + (bv < 0.0f ? (uint32_t)N : 0u);
        code = (code << g.bpf) | hval;
    }
    code >>= g.cut;
    if (lane < tile.count) {
        if (tile.code_stride) codes[tile.code_base + (uint64_t)t * tile.code_stride + lane] = (uint32_t)code;
        else codes[(uint64_t)tile.fset * fset_stride + (uint64_t)t * code_stride + tile.out_row0 + lane] = (uint32_t)code;
    }
}

// ------------------------------------------------------------------------------------------------ segmented radix sort

constexpr uint32_t kSortThreads = 512;
constexpr uint32_t kSortWarps = kSortThreads / 32;
constexpr uint32_t kSortSmemCapSmall = 4096;   // 16 B/entry -> 64 KB + 36 KB bookkeeping: 2 CTAs / SM
constexpr uint32_t kSortSmemCapLarge = 11264;  // 176 KB + 36 KB: 1 CTA / SM
constexpr uint32_t kSortBookWords = 3 * 256 + 2 * kSortWarps * 256 + 256;

uint32_t segment_sort_smem_capacity() { return kSortSmemCapLarge; }

// sorthash.hpp:133-194 — stable LSD radix sort with three byte passes over (hash, payload) pairs; one CTA per segment.
// Stability inside a pass: elements are taken 512 at a time in order; __match_any gives each element its rank among
// the equal digits of its warp, per-warp digit counts are prefix-summed across warps on top of the running bin offset.
// Segments longer than `cap` ping-pong through a global scratch pair instead of shared memory.
__global__ void __launch_bounds__(kSortThreads) k_segment_sort(const SortSegment* __restrict__ segs, uint32_t cap, uint32_t* keys,
                                                               uint32_t* idx, uint32_t* scratch_keys, uint32_t* scratch_idx) {
    extern __shared__ uint32_t s_mem[];
    uint32_t* s_hist = s_mem;                           // [3][256]
    uint32_t* s_cnt = s_hist + 3 * 256;                 // [warps][256]
    uint32_t* s_off = s_cnt + kSortWarps * 256;         // [warps][256]
    uint32_t* s_bin = s_off + kSortWarps * 256;         // [256]
    uint32_t* s_buf = s_bin + 256;                      // [4][cap]

    const SortSegment seg = segs[blockIdx.x];
    const uint32_t len = seg.len;
    if (len == 0) return;
    uint32_t* gk = keys + seg.base;
    uint32_t* gi = idx + seg.base;
    const bool in_smem = len <= cap;
    uint32_t *b0k, *b0i, *b1k, *b1i;
    if (in_smem) {
        b0k = s_buf; b0i = s_buf + cap; b1k = s_buf + 2 * cap; b1i = s_buf + 3 * cap;
    } else {
        // one global scratch pair: pass 1 -> scratch, pass 2 -> (keys, idx) in place, pass 3 -> scratch, then copy back
        b0k = scratch_keys + seg.scratch_base; b0i = scratch_idx + seg.scratch_base; b1k = gk; b1i = gi;
    }
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (uint32_t i = tid; i < 3 * 256 + kSortWarps * 256; i += kSortThreads) s_mem[i] = 0;  // s_hist and s_cnt
    __syncthreads();
    for (uint32_t i = tid; i < len; i += kSortThreads) {
        uint32_t h = gk[i];
        atomicAdd(&s_hist[h & 0xff], 1u);
        atomicAdd(&s_hist[256 + ((h >> 8) & 0xff)], 1u);
        atomicAdd(&s_hist[512 + ((h >> 16) & 0xff)], 1u);
    }
    __syncthreads();

    for (int pass = 0; pass < 3; pass++) {
        const uint32_t *sk, *si;
        uint32_t *dk, *di;
        if (pass == 0) { sk = gk; si = nullptr; dk = b0k; di = b0i; }
        else if (pass == 1) { sk = b0k; si = b0i; dk = b1k; di = b1i; }
        else { sk = b1k; si = b1i; dk = in_smem ? gk : b0k; di = in_smem ? gi : b0i; }
        const int shift = 8 * pass;
        // exclusive prefix of this pass's histogram -> running bin offsets (warp 0)
        if (warp == 0) {
            uint32_t run = 0;
            for (int c = 0; c < 8; c++) {
                uint32_t v = s_hist[pass * 256 + c * 32 + lane];
                uint32_t incl = v;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if ((int)lane >= o) incl += t;
                }
                s_bin[c * 32 + lane] = run + incl - v;
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        __syncthreads();
        for (uint32_t base = 0; base < len; base += kSortThreads) {
            const uint32_t i = base + tid;
            const bool valid = i < len;
            uint32_t key = 0, payload = 0, digit = 0xffffu;
            if (valid) {
                key = sk[i];
                payload = si ? si[i] : i;
                digit = (key >> shift) & 0xff;
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1));
            if (valid && rank == 0) s_cnt[warp * 256 + digit] = __popc(peers);
            __syncthreads();
            if (tid < 256) {
                uint32_t run = s_bin[tid];
#pragma unroll 8
                for (uint32_t w = 0; w < kSortWarps; w++) {
                    uint32_t c = s_cnt[w * 256 + tid];
                    s_cnt[w * 256 + tid] = 0;
                    s_off[w * 256 + tid] = run;
                    run += c;
                }
                s_bin[tid] = run;
            }
            __syncthreads();
            if (valid) {
                uint32_t pos = s_off[warp * 256 + digit] + rank;
                dk[pos] = key;
                di[pos] = payload;
            }
        }
        __syncthreads();
    }
    if (!in_smem) {
        for (uint32_t i = tid; i < len; i += kSortThreads) {
            gk[i] = b0k[i];
            gi[i] = b0i[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ collision estimates

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// crosspolytope.hpp:16-88 — Monte-Carlo estimate of the probability that a random rotation maps x = e1 and
// y = (alpha, sqrt(1-alpha^2), 0, ...) to cross-polytope hashes that agree in their top `used_bits` bits.
// One thread per (alpha bin, repetition); any good RNG is statistically equivalent to the reference's clock-seeded one.
__global__ void k_cp_trials(uint32_t m, uint32_t reps, uint64_t seed, float eps, uint32_t* __restrict__ counts) {
    const uint32_t bin = blockIdx.x;
    const uint32_t rep = blockIdx.y * blockDim.x + threadIdx.x;
    if (rep >= reps) return;
    // alpha advances by 2*eps in double starting from -1 (crosspolytope.hpp:31-34,86)
    double alpha = -1.0;
    const double step = (double)(2 * eps);
    for (uint32_t i = 0; i < bin; i++) alpha += step;
    const double beta = sqrt(1.0 - alpha * alpha);
    const uint32_t dims = 1u << m;
    uint32_t hx = 0, hy = 0;
    double vx = 0.0, vy = 0.0;
    uint64_t ctr = splitmix64(seed ^ (((uint64_t)bin << 40) | ((uint64_t)rep << 16)));
    for (uint32_t j = 0; j < dims; j++) {
        uint64_t r1 = splitmix64(ctr + 2 * j), r2 = splitmix64(ctr + 2 * j + 1);
        double u1 = ((double)(r1 >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0,1]
        double u2 = (double)(r2 >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
        double rad = sqrt(-2.0 * log(u1));
        double sn, cs;
        sincospi(2.0 * u2, &sn, &cs);
        double z1 = rad * cs, z2 = rad * sn;
        if (fabs(z1) > vx) {
            vx = fabs(z1);
            hx = j;
            if (z1 < 0) hx |= (1u << m);
        }
        double h = alpha * z1 + beta * z2;
        if (fabs(h) > vy) {
            vy = fabs(h);
            hy = j;
            if (h < 0) hy |= (1u << m);
        }
    }
    for (uint32_t used = 0; used <= m + 1; used++) {
        uint32_t shift = m + 1 - used;
        if ((hx >> shift) == (hy >> shift)) atomicAdd(&counts[used * kEstBins + bin], 1u);
    }
}

__global__ void k_cp_finalize(uint32_t m, uint32_t reps, const uint32_t* __restrict__ counts, float* __restrict__ est) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (m + 2) * kEstBins) est[i] = (float)counts[i] / (float)reps;
}

// ------------------------------------------------------------------------------------------------ launchers

void launch_row_norms(const float* data, uint64_t n, uint32_t d, float* norms, cudaStream_t s) {
    uint64_t threads = n * 8;
    k_row_norms<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(data, n, d, norms);
}

void launch_gmm_pass(const float* data, const float* norms, uint64_t row0, uint64_t row1, uint32_t d, uint32_t c, uint64_t* keys, float* dist,
                     uint32_t* assign, float* cc, cudaStream_t s) {
    // rows [row0, row1) of the full arrays (a rank's share when the clustering is sharded; everything otherwise)
    const uint64_t n = row1, rows = row1 > row0 ? row1 - row0 : 0;
    if (rows == 0) {
        if (c > 0 && cc && d % 4 == 0 && tune_get("gmm_vec", 1) != 0 && tune_get("gmm_prune", 1) != 0) return;
        return;
    }
    // knobs (never change a result): gmm_vec 0 = the 8-lanes-per-row kernel; gmm_prune 0 = evaluate every row in every pass
    if (d % 4 == 0 && d * sizeof(float) <= 40 * 1024 && tune_get("gmm_vec", 1) != 0) {
        const uint64_t threads = rows * 2;
        const unsigned grid = (unsigned)((threads + 255) / 256);
        const size_t smem = d * sizeof(float);
        if (c > 0 && cc && tune_get("gmm_prune", 1) != 0) {
            k_gmm_centre_dists<<<(c + 127) / 128, 128, 0, s>>>(data, norms, d, c, keys, cc);
            k_gmm_pass_v<true><<<grid, 256, smem, s>>>(data, norms, row0, n, d, c, keys, dist, assign, cc);
        } else {
            k_gmm_pass_v<false><<<grid, 256, smem, s>>>(data, norms, row0, n, d, c, keys, dist, assign, cc);
        }
        return;
    }
    uint64_t threads = rows * 8;
    k_gmm_pass<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(data, norms, row0, n, d, c, keys, dist, assign);
}

void launch_gmm_finish(const uint64_t* keys, uint32_t K, uint64_t n, const float* dist, const uint32_t* assign, uint32_t* centers,
                       float* radii, uint32_t* sizes, cudaStream_t s) {
    uint64_t threads = n > K ? n : K;
    k_gmm_finish<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(keys, K, n, dist, assign, centers, radii, sizes);
}

void launch_gather_rows(const float* data, const uint32_t* rows, uint32_t count, uint32_t d, float* out, cudaStream_t s) {
    uint64_t threads = (uint64_t)count * d;
    if (threads == 0) return;
    k_gather_rows<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(data, rows, count, d, out);
}

void launch_store_q15(const float* data, const uint32_t* perm, uint64_t rows, uint32_t d, uint32_t sl, int16_t* q15, cudaStream_t s) {
    if (rows == 0) return;
    k_store_q15<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(data, perm, rows, d, sl, q15);
}

void launch_sketch(const int16_t* q15, const RowTile* tiles, uint32_t n_tiles, const int16_t* planes, uint32_t sl, uint64_t* sketches,
                   cudaStream_t s) {
    if (n_tiles == 0) return;
    size_t smem = (size_t)32 * sl * sizeof(int);
    if (smem > 48 * 1024) ensure_dynamic_smem((const void*)(k_sketch), smem);
    dim3 grid(n_tiles, kNumPlanes / 256);
    k_sketch<<<grid, 256, smem, s>>>(q15, tiles, planes, sl, sketches);
}

template <int M>
static void launch_codes_m(const int16_t* q15, const RowTile* tiles, uint32_t n_tiles, const uint32_t* signbits, HashGeom g,
                           uint32_t* codes, uint64_t code_stride, uint64_t fset_stride, cudaStream_t s) {
    dim3 grid(n_tiles, (g.L + 7) / 8);
    if constexpr (M >= 5 && M <= 7) {
        if (tune_get("codes_fast", 1) != 0) {  // knob: 0 = the plain kernel (A/B)
            constexpr int N = 1 << M;
            const size_t fsmem = (size_t)32 * (N + 4) * 4 + (size_t)8 * g.fph * kRotations * N * 4;
            if (fsmem > 48 * 1024) ensure_dynamic_smem((const void*)(k_codes_fast<M>), fsmem);
            k_codes_fast<M><<<grid, 256, fsmem, s>>>(q15, tiles, signbits, g, codes, code_stride, fset_stride);
            return;
        }
    }
    size_t smem = (size_t)32 * (g.sl + 2) * sizeof(int16_t);
    if (smem > 48 * 1024) CLANN_CUDA(cudaFuncSetAttribute(k_codes<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_codes<M><<<grid, 256, smem, s>>>(q15, tiles, signbits, g, codes, code_stride, fset_stride);
}

void launch_codes(const int16_t* q15, const RowTile* tiles, uint32_t n_tiles, const uint32_t* signbits, HashGeom g, uint32_t* codes,
                  uint64_t code_stride, uint64_t fset_stride, cudaStream_t s) {
    if (n_tiles == 0) return;
    switch (g.m) {
        case 0: launch_codes_m<0>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 1: launch_codes_m<1>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 2: launch_codes_m<2>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 3: launch_codes_m<3>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 4: launch_codes_m<4>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 5: launch_codes_m<5>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 6: launch_codes_m<6>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 7: launch_codes_m<7>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 8: launch_codes_m<8>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 9: launch_codes_m<9>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        case 10: launch_codes_m<10>(q15, tiles, n_tiles, signbits, g, codes, code_stride, fset_stride, s); break;
        default: throw std::invalid_argument("dimension above 1024 is not supported");
    }
}

void launch_segment_sort(const SortSegment* segs, uint32_t n_segs, uint32_t max_len, uint32_t* keys, uint32_t* idx,
                         uint32_t* scratch_keys, uint32_t* scratch_idx, cudaStream_t s) {
    if (n_segs == 0) return;
    const uint32_t cap = (max_len <= kSortSmemCapSmall) ? kSortSmemCapSmall : kSortSmemCapLarge;
    size_t smem = ((size_t)kSortBookWords + (size_t)4 * cap) * sizeof(uint32_t);
    ensure_dynamic_smem((const void*)(k_segment_sort), smem);
    k_segment_sort<<<n_segs, kSortThreads, smem, s>>>(segs, cap, keys, idx, scratch_keys, scratch_idx);
}

// One CTA per (table, cluster): one pass over the sorted 24-bit codes; wherever the top kDirBits bits step from `lo` to `hi`
// at position i, entries lo+1 .. hi of the directory are i (the lower bound of b << 12 for each of those b); position nc closes
// the directory. O(nc + 2^kDirBits) per table instead of 2^kDirBits binary searches.
__global__ void __launch_bounds__(256) k_build_dir(const uint32_t* __restrict__ tbl_hash, uint64_t n, const uint64_t* __restrict__ offsets,
                                                   const uint8_t* __restrict__ skip, uint32_t K, uint32_t* __restrict__ dir) {
    const uint32_t c = blockIdx.x, t = blockIdx.y;
    if (skip[c]) return;
    const uint64_t off = offsets[c];
    const uint32_t nc = (uint32_t)(offsets[c + 1] - off);
    const uint32_t* H = tbl_hash + table_base(off, nc, gridDim.y, t);
    uint32_t* out = dir + ((uint64_t)c * gridDim.y + t) * kDirEntries;
    constexpr uint32_t shift = kMaxHashBits - kDirBits;
    for (uint32_t i = threadIdx.x; i <= nc; i += blockDim.x) {
        const int hi = i < nc ? (int)(__ldg(H + i) >> shift) : (int)(1u << kDirBits);
        const int lo = i > 0 ? (int)(__ldg(H + i - 1) >> shift) : -1;
        for (int b = lo + 1; b <= hi; b++) out[b] = i;
    }
}

void launch_build_dir(const uint32_t* tbl_hash, uint64_t n, const uint64_t* offsets, const uint8_t* skip, uint32_t K, uint32_t L,
                      uint32_t* dir, cudaStream_t s) {
    if (K == 0 || L == 0) return;
    dim3 grid(K, L);
    k_build_dir<<<grid, 256, 0, s>>>(tbl_hash, n, offsets, skip, K, dir);
}

void launch_cp_estimates(uint32_t m, uint32_t reps, uint64_t seed, float* est, uint32_t* scratch_counts, cudaStream_t s) {
    CLANN_CUDA(cudaMemsetAsync(scratch_counts, 0, sizeof(uint32_t) * (m + 2) * kEstBins, s));
    dim3 grid(kEstBins, (reps + 127) / 128);
    k_cp_trials<<<grid, 128, 0, s>>>(m, reps, seed, 5e-3f, scratch_counts);
    k_cp_finalize<<<((m + 2) * kEstBins + 255) / 256, 256, 0, s>>>(m, reps, scratch_counts, est);
}

}  // namespace clann
