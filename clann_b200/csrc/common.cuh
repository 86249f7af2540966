// Shared constants, error plumbing and small device helpers for libclann_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>

namespace clann {

// PUFFINN constants (citations into /root/reference/libpuffinn/include/puffinn)
constexpr int kNumSketches = 32;     // filterer.hpp:16
constexpr int kSketchBits = 64;      // typedefs.hpp:9
constexpr int kNumPlanes = kNumSketches * kSketchBits;
constexpr int kMaxHashBits = 24;     // typedefs.hpp:13
constexpr int kSegment = 12;         // prefixmap.hpp:60
constexpr int kEstBins = 201;        // crosspolytope.hpp:37-86 with eps = 5e-3
constexpr int kRotations = 3;        // crosspolytope.hpp:221-226
constexpr int kRing = 32;            // collection.hpp:604
constexpr int kFilterBuffer = 128;   // collection.hpp:777
constexpr int kPassingCap = kFilterBuffer + 8 * kRing;  // collection.hpp:783

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

#define CLANN_CUDA(expr)                                                                                   \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess)                                                                             \
            throw ::clann::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                     ":" + std::to_string(__LINE__) + ")");                               \
    } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to a (function, device) pair: the largest value set so far is remembered
// per pair and under a lock, so that host threads driving the library side by side (two in-process ranks in the tests) or a
// process that holds indices on several devices never launch with more shared memory than the attribute allows.
inline void ensure_dynamic_smem(const void* kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> largest;
    int dev = 0;
    CLANN_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur = largest[std::make_pair(kernel, dev)];
    if (bytes > cur) {
        CLANN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cur = bytes;
    }
}

// One work tile for the row-parallel hashing kernels: up to 32 consecutive Q15 rows that share a function set.
struct RowTile {
    uint32_t in_row0;   // first row in the Q15 input array
    uint32_t out_row0;  // first row in the output array
    uint32_t count;     // 1..32
    uint32_t fset;      // function set id
    // table codes only: when code_stride != 0 the code of table t for row r of the tile goes to codes[code_base + t*code_stride + r]
    // (index build: cluster-major tables), otherwise to the launcher's [fset][table][row] layout (query batches)
    uint32_t code_stride;
    uint32_t pad;
    uint64_t code_base;
};

// Cluster-major table layout: the L tables of a cluster are adjacent, so that everything one (query, cluster) visit touches
// lies in a handful of 2 MB pages (measured on B200: random reads over an L2-resident working set drop from 288 to 37 G
// sectors/s once it is spread over more pages than the TLB covers — profiles/micro/tlb_spread.cu). Entry i of table t of the
// cluster whose rows start at `off` (nc rows) is at L*off + t*nc + i; its directory at (c*L + t)*kDirEntries.
__host__ __device__ inline uint64_t table_base(uint64_t off, uint32_t nc, uint32_t L, uint32_t t) {
    return (uint64_t)L * off + (uint64_t)t * nc;
}

// Geometry of the hash family for dimension d (independent.hpp:29-33, crosspolytope.hpp:301-303, generic.hpp:34-40).
struct HashGeom {
    uint32_t d, sl, m, npts, bpf, fph, cut, L;
};

inline uint32_t ceil_log(uint32_t v) {  // math.hpp:105-113
    uint32_t lg = 0, p = 1;
    while (p < v) {
        lg++;
        p *= 2;
    }
    return lg;
}

inline HashGeom make_geom(uint32_t d, uint32_t L) {
    HashGeom g;
    g.d = d;
    g.sl = (d + 15) / 16 * 16;
    g.m = ceil_log(d);
    g.npts = 1u << g.m;
    g.bpf = g.m + 1;
    g.fph = (kMaxHashBits + g.bpf - 1) / g.bpf;
    g.cut = g.bpf * g.fph - kMaxHashBits;
    g.L = L;
    return g;
}

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// format/unit_vector.hpp:40-45: min(v * 2^15, 32767), truncation toward zero.
__device__ __forceinline__ int16_t to_q15(float v) {
    float s = __fmul_rn(v, 32768.0f);
    s = fminf(s, 32767.0f);
    return (int16_t)__float2int_rz(s);
}

// One term of math.hpp:37-44: ((a*b >> 14) + 1) >> 1 == (a*b + 2^14) >> 15 with an arithmetic shift.
__device__ __forceinline__ int q15_mul(int a, int b) { return (a * b + 16384) >> 15; }

// The same term for a given in the HIGH half of a 32-bit word and b doubled: (a*2^16) * (2b) + 2^31 = 2^17 (a*b + 2^14),
// whose upper word is the term — one IMAD.HI with a 64-bit addend and no shift. Measured on B200: IMAD.HI issues at a
// lower rate than IMAD + SHF, so the sketch projection and the warp kernel's rerank keep q15_mul (A/B: 3.09 vs 3.55 ms).
__device__ __forceinline__ int q15_mul_hi(int a_hi16, int b2) {
    return (int)(((long long)a_hi16 * (long long)b2 + 0x80000000ll) >> 32);
}

// Sign-extending unpack of a packed pair of int16 (PRMT with the sign-replicate selector bit).
__device__ __forceinline__ int unpack_lo(uint32_t w) {
    int r;
    asm("prmt.b32 %0, %1, 0, 0x9910;" : "=r"(r) : "r"(w));
    return r;
}
__device__ __forceinline__ int unpack_hi(uint32_t w) { return (int)w >> 16; }

// ndarray 0.16.1 numeric_util::unrolled_dot (the f32 `dot` behind angulardata.rs:13,26,30): eight partial sums over
// chunks of eight, folded (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7) into `sum`, then the tail in order. Rust emits no FMA.
__device__ __forceinline__ float ndarray_dot_thread(const float* __restrict__ x, const float* __restrict__ y, uint32_t d) {
    float p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    uint32_t i = 0;
    for (; i + 8 <= d; i += 8) {
        p0 = __fadd_rn(p0, __fmul_rn(x[i + 0], y[i + 0]));
        p1 = __fadd_rn(p1, __fmul_rn(x[i + 1], y[i + 1]));
        p2 = __fadd_rn(p2, __fmul_rn(x[i + 2], y[i + 2]));
        p3 = __fadd_rn(p3, __fmul_rn(x[i + 3], y[i + 3]));
        p4 = __fadd_rn(p4, __fmul_rn(x[i + 4], y[i + 4]));
        p5 = __fadd_rn(p5, __fmul_rn(x[i + 5], y[i + 5]));
        p6 = __fadd_rn(p6, __fmul_rn(x[i + 6], y[i + 6]));
        p7 = __fadd_rn(p7, __fmul_rn(x[i + 7], y[i + 7]));
    }
    float sum = 0.0f;
    sum = __fadd_rn(sum, __fadd_rn(p0, p4));
    sum = __fadd_rn(sum, __fadd_rn(p1, p5));
    sum = __fadd_rn(sum, __fadd_rn(p2, p6));
    sum = __fadd_rn(sum, __fadd_rn(p3, p7));
    for (; i < d; i++) sum = __fadd_rn(sum, __fmul_rn(x[i], y[i]));
    return sum;
}

// The same dot computed by 8 consecutive lanes (sub = lane & 7 owns partial sum p_sub); every lane of the group
// returns the result. Row reads are coalesced 32-byte sectors across the group.
__device__ __forceinline__ float ndarray_dot_group8(const float* __restrict__ x, const float* __restrict__ y, uint32_t d) {
    const uint32_t sub = threadIdx.x & 7;
    const uint32_t full = d & ~7u;
    float p = 0.0f;
    for (uint32_t i = sub; i < full; i += 8) p = __fadd_rn(p, __fmul_rn(x[i], y[i]));
    float pr = __fadd_rn(p, __shfl_xor_sync(0xffffffffu, p, 4));  // lanes 0..3 of the group: p0+p4, p1+p5, p2+p6, p3+p7
    const int base = (threadIdx.x & 31) & ~7;
    float s0 = __shfl_sync(0xffffffffu, pr, base + 0);
    float s1 = __shfl_sync(0xffffffffu, pr, base + 1);
    float s2 = __shfl_sync(0xffffffffu, pr, base + 2);
    float s3 = __shfl_sync(0xffffffffu, pr, base + 3);
    float sum = 0.0f;
    sum = __fadd_rn(sum, s0);
    sum = __fadd_rn(sum, s1);
    sum = __fadd_rn(sum, s2);
    sum = __fadd_rn(sum, s3);
    for (uint32_t i = full; i < d; i++) sum = __fadd_rn(sum, __fmul_rn(x[i], y[i]));
    return sum;
}

// Order-preserving map float -> uint32 (works for negatives too).
__device__ __forceinline__ uint32_t float_order_bits(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_order_bits(uint32_t u) {
    uint32_t b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(b);
}

#endif  // __CUDACC__

}  // namespace clann
