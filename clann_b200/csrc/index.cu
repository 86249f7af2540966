// Host side of libclann_b200: the device-resident clustered index, its build / search orchestration and the C ABI
// declared in include/clann_b200.h. Citations are file:line into /root/reference.
#include "../../include/clann_b200.h"

#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <random>
#include <thread>
#include <vector>

#include "kernels.h"
#include "probe_common.cuh"

namespace clann {

static thread_local std::string g_last_error;

struct StatusError : std::runtime_error {
    int code;
    StatusError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) CLANN_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void ensure(size_t count) {
        if (count > n) alloc(count);
    }
    void zero(cudaStream_t s = 0) {
        if (n) CLANN_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    void upload(const T* src, size_t count, cudaStream_t s = 0) {
        ensure(count);
        if (count) CLANN_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& v, cudaStream_t s = 0) { upload(v.data(), v.size(), s); }
    std::vector<T> download(size_t count, size_t offset = 0) const {
        std::vector<T> v(count);
        if (count) CLANN_CUDA(cudaMemcpy(v.data(), p + offset, count * sizeof(T), cudaMemcpyDeviceToHost));
        return v;
    }
};

// ------------------------------------------------------------------------------------------------ host-side function sets

// format/unit_vector.hpp:61-89 on the host (used for the SimHash hyperplanes, simhash.hpp:17-23); same arithmetic as
// k_store_q15 (see the note there on the summation shape).
static void store_q15_host(const float* v, uint32_t d, uint32_t sl, int16_t* out) {
    float acc = 0.0f;
    uint32_t body = d & ~3u;
    for (uint32_t i = 0; i < body; i++) {
        volatile float prod = v[i] * v[i];  // keep the product rounded (no contraction)
        acc = acc + prod;
    }
    for (uint32_t i = body; i < d; i++) acc = fmaf(v[i], v[i], acc);
    float len = sqrtf(acc);
    for (uint32_t i = 0; i < d; i++) {
        float x = v[i];
        if (len != 0.0f) x = x / len;
        float s = x * 32768.0f;
        if (s > 32767.0f) s = 32767.0f;
        out[i] = (int16_t)s;
    }
    for (uint32_t i = d; i < sl; i++) out[i] = 0;
}

struct FunctionSet {
    std::vector<int16_t> planes;     // [2048][sl]
    std::vector<uint32_t> signbits;  // [L*fph][3][W]
    std::vector<float> est;          // [(m+2)][201]
    bool have_est = false, have_fn = false;
};

static uint32_t sign_words(const HashGeom& g) { return (g.npts + 31) / 32; }

static void pack_signs(const HashGeom& g, const int8_t* signs, std::vector<uint32_t>& out) {
    const uint32_t W = sign_words(g);
    const size_t nfunc = (size_t)g.L * g.fph;
    out.assign(nfunc * kRotations * W, 0u);
    for (size_t f = 0; f < nfunc; f++)
        for (uint32_t r = 0; r < (uint32_t)kRotations; r++)
            for (uint32_t i = 0; i < g.npts; i++)
                if (signs[(f * kRotations + r) * g.npts + i] < 0) out[(f * kRotations + r) * W + i / 32] |= 1u << (i % 32);
}

// Fresh functions as the reference draws them: 2048 normalised Gaussian hyperplanes stored as Q15 (simhash.hpp:17-23,
// unit_vector.hpp:91-100) and L*fph x 3 x 2^m uniform signs (crosspolytope.hpp:160-165).
static void generate_functions(const HashGeom& g, uint64_t seed, FunctionSet& fs) {
    std::mt19937_64 rng(seed);
    std::normal_distribution<float> normal(0.0f, 1.0f);
    fs.planes.resize((size_t)kNumPlanes * g.sl);
    std::vector<float> v(g.d);
    for (int f = 0; f < kNumPlanes; f++) {
        for (uint32_t i = 0; i < g.d; i++) v[i] = normal(rng);
        store_q15_host(v.data(), g.d, g.sl, fs.planes.data() + (size_t)f * g.sl);
    }
    const size_t nsign = (size_t)g.L * g.fph * kRotations * g.npts;
    std::vector<int8_t> signs(nsign);
    uint64_t bits = 0;
    int left = 0;
    for (size_t i = 0; i < nsign; i++) {
        if (left == 0) {
            bits = rng();
            left = 64;
        }
        signs[i] = (bits & 1) ? 1 : -1;
        bits >>= 1;
        left--;
    }
    pack_signs(g, signs.data(), fs.signbits);
    fs.have_fn = true;
}

// hash_source/hash_source.hpp:49-57 with crosspolytope.hpp:116-118, evaluated per similarity bin.
static float concat_prob(const HashGeom& g, const float* est, uint32_t num_bits, uint32_t bin) {
    uint32_t whole = num_bits / g.bpf, rem = num_bits % g.bpf;
    float wp = est[(size_t)g.bpf * kEstBins + bin];
    float rp = est[(size_t)rem * kEstBins + bin];
    return (float)(pow((double)wp, (double)(int)whole) * (double)rp);
}

// The stop rule (collection.hpp:927-943) tabulated with the host's libm so that the device decision is bit-identical to
// hash_source/independent.hpp:108-119: bit t of row (depth, bin) = failure_probability(depth, t, depth==24 ? t : L, sim) <= 1-recall.
static void build_stop_table(const HashGeom& g, const float* est, float recall, uint32_t stop_words, uint32_t* out) {
    const float thr = 1 - recall;
    for (uint32_t depth = 1; depth <= (uint32_t)kMaxHashBits; depth++) {
        for (uint32_t bin = 0; bin < (uint32_t)kEstBins; bin++) {
            uint32_t* row = out + ((size_t)(depth - 1) * kEstBins + bin) * stop_words;
            for (uint32_t w = 0; w < stop_words; w++) row[w] = 0;
            float col = concat_prob(g, est, depth, bin);
            float last = concat_prob(g, est, depth + 1, bin);
            double one_minus_col = 1.0 - (double)col;
            float one_minus_last = 1.0f - last;  // `1-last_prob` is evaluated in float (independent.hpp:118)
            for (uint32_t t = 0; t <= g.L; t++) {
                uint64_t max_tables = (depth == (uint32_t)kMaxHashBits) ? t : g.L;
                double a = pow(one_minus_col, (double)(uint64_t)t);
                double b = pow((double)one_minus_last, (double)(max_tables - t));
                float fp = (float)(a * b);
                if (fp <= thr) row[t >> 5] |= 1u << (t & 31);
            }
        }
    }
}

// filterer.hpp:108-111 + simhash.hpp:96-102 tabulated over every value MaxBuffer::smallest_value() can take.
static void build_msd_table(std::vector<uint8_t>& msd) {
    msd.resize(65536);
    for (uint32_t v = 0; v < 65536; v++) {
        float sim = (float)v / 65536.0f;
        float arg = 2.0f * sim - 1.0f;
        float cp = (float)(1.0 - (double)acosf(arg) / M_PI);
        float r = roundf((float)(64.0 * (1.0 - (double)cp)));
        msd[v] = (uint8_t)r;
    }
}

// --- Index::serialize reader (collection.hpp:185-203; field list in SURVEY.md section 8c): extracts the function set.
struct Reader {
    const uint8_t* p;
    uint64_t len, off = 0;
    void bytes(void* dst, uint64_t n) {
        if (n > len - off) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream truncated");  // overflow-safe: off <= len always
        if (dst) memcpy(dst, p + off, n);
        off += n;
    }
    template <typename T>
    T get() {
        T v;
        bytes(&v, sizeof(T));
        return v;
    }
};

static void parse_reference_stream(const uint8_t* blob, uint64_t len, const HashGeom& g, FunctionSet& fs, uint32_t* n_points) {
    Reader r{blob, len};
    uint32_t d = r.get<uint32_t>(), sl = r.get<uint32_t>(), n = r.get<uint32_t>();
    if (d != g.d || sl != g.sl) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream has a different dimension");
    r.bytes(nullptr, (uint64_t)n * sl * 2);                    // Dataset rows (dataset.hpp:79-86)
    if (r.get<int32_t>() != 0) throw StatusError(CLANN_ERR_SERIALIZE, "sketch source is not IndependentHashSource");
    r.get<uint32_t>(); r.get<uint32_t>();                      // SimHash dataset description
    if (r.get<uint64_t>() != (uint64_t)kNumPlanes) throw StatusError(CLANN_ERR_SERIALIZE, "expected 2048 sketch functions");
    fs.planes.resize((size_t)kNumPlanes * sl);
    for (int f = 0; f < kNumPlanes; f++) {                     // simhash.hpp:33-38
        if (r.get<uint32_t>() != sl) throw StatusError(CLANN_ERR_SERIALIZE, "hyperplane length mismatch");
        r.bytes(fs.planes.data() + (size_t)f * sl, (uint64_t)sl * 2);
    }
    r.get<uint32_t>(); r.get<uint32_t>(); r.get<uint8_t>(); r.get<uint32_t>(); r.get<uint32_t>();  // independent.hpp:64-68
    uint64_t n_sk = r.get<uint64_t>();
    if (n_sk > len / 8) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream truncated");
    r.bytes(nullptr, n_sk * 8);                                // stored sketches (filterer.hpp:62-68)
    if (r.get<int32_t>() != 0) throw StatusError(CLANN_ERR_SERIALIZE, "hash source is not IndependentHashSource");
    uint32_t rot = r.get<uint32_t>();
    r.get<uint32_t>(); r.get<float>();                         // estimation_repetitions, estimation_eps
    if (rot != (uint32_t)kRotations) throw StatusError(CLANN_ERR_SERIALIZE, "expected 3 FHT rotations");
    if (!r.get<uint8_t>()) throw StatusError(CLANN_ERR_SERIALIZE, "reference index was never rebuilt (no hash source)");
    r.get<uint32_t>(); r.get<uint32_t>();                      // family: dataset description
    r.get<uint32_t>(); r.get<uint32_t>(); r.get<float>();      // family: args
    uint64_t rows = r.get<uint64_t>();                         // crosspolytope.hpp:104-114
    if (rows != g.m + 2) throw StatusError(CLANN_ERR_SERIALIZE, "collision estimate table has the wrong height");
    fs.est.resize((size_t)(g.m + 2) * kEstBins);
    for (uint64_t b = 0; b < rows; b++) {
        if (r.get<uint64_t>() != (uint64_t)kEstBins) throw StatusError(CLANN_ERR_SERIALIZE, "collision estimate table has the wrong width");
        r.bytes(fs.est.data() + b * kEstBins, sizeof(float) * kEstBins);
    }
    for (float v : fs.est)
        if (!(v >= 0.0f && v <= 1.0f)) throw StatusError(CLANN_ERR_SERIALIZE, "collision estimate outside [0, 1]");
    r.get<float>();                                            // eps
    uint64_t n_fn = r.get<uint64_t>();
    if (n_fn != (uint64_t)g.L * g.fph) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream has a different number of tables");
    std::vector<int8_t> signs((size_t)n_fn * kRotations * g.npts);
    for (uint64_t f = 0; f < n_fn; f++) {                      // crosspolytope.hpp:178-184
        int32_t dd = r.get<int32_t>(), mm = r.get<int32_t>();
        uint32_t rr = r.get<uint32_t>();
        if ((uint32_t)dd != g.d || (uint32_t)mm != g.m || rr != (uint32_t)kRotations) throw StatusError(CLANN_ERR_SERIALIZE, "FHT function header mismatch");
        r.bytes(signs.data() + f * kRotations * g.npts, (uint64_t)kRotations * g.npts);
    }
    pack_signs(g, signs.data(), fs.signbits);
    fs.have_fn = fs.have_est = true;
    if (n_points) *n_points = n;
}

// --- Index::serialize writer (collection.hpp:185-203): the byte stream a puffinn::Index<CosineSimilarity> built over the
// same rows, with the same functions, would write — so that an index built here can be loaded by the CPU reference
// (Index(std::istream&), collection.hpp:147-170) and vice versa (parse_reference_stream). Field list in SURVEY.md 8c.
struct Writer {
    std::vector<uint8_t> out;
    template <typename T>
    void put(T v) {
        const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
        out.insert(out.end(), b, b + sizeof(T));
    }
    void bytes(const void* src, uint64_t n) {
        const uint8_t* b = static_cast<const uint8_t*>(src);
        out.insert(out.end(), b, b + n);
    }
};

static std::vector<uint8_t> write_reference_stream(const HashGeom& g, const FunctionSet& fs, uint32_t nc, const int16_t* q15,
                                                   const uint64_t* sketches, const uint32_t* tbl_hash, const uint32_t* tbl_idx) {
    Writer w;
    const uint32_t d = g.d, sl = g.sl, L = g.L;
    w.out.reserve((size_t)nc * (sl * 2 + kNumSketches * 8 + (size_t)L * 8) + (size_t)kNumPlanes * (sl * 2 + 4) + (size_t)L * 8200 * 4 + 65536);
    // Dataset (dataset.hpp:79-86): description {args = d, storage_len}, inserted_vectors, rows
    w.put<uint32_t>(d); w.put<uint32_t>(sl); w.put<uint32_t>(nc);
    w.bytes(q15, (uint64_t)nc * sl * 2);
    // Filterer (filterer.hpp:62-68): sketch args, SimHash source, sketches
    w.put<int32_t>(0);                                   // HashSourceType::Independent (independent.hpp:135-139); SimHash args are empty
    w.put<uint32_t>(d); w.put<uint32_t>(sl);             // family: dataset description
    w.put<uint64_t>((uint64_t)kNumPlanes);
    for (int f = 0; f < kNumPlanes; f++) {               // simhash.hpp:33-38
        w.put<uint32_t>(sl);
        w.bytes(fs.planes.data() + (size_t)f * sl, (uint64_t)sl * 2);
    }
    w.put<uint32_t>(kNumSketches); w.put<uint32_t>(kSketchBits); w.put<uint8_t>(1);  // num_hashers, functions_per_hasher, bits_per_function
    w.put<uint32_t>(0); w.put<uint32_t>(0);                                          // next_function, bits_to_cut (independent.hpp:64-68)
    w.put<uint64_t>((uint64_t)nc * kNumSketches);
    w.bytes(sketches, (uint64_t)nc * kNumSketches * 8);
    // hash args (independent.hpp:135-139 + crosspolytope.hpp:234-238)
    w.put<int32_t>(0); w.put<int32_t>(kRotations); w.put<uint32_t>(1000); w.put<float>(0.005f);
    w.put<uint8_t>(1);                                   // has_hash_source
    // IndependentHashSource<FHTCrossPolytopeHash> (independent.hpp:56-68): family (crosspolytope.hpp:291-295) ...
    w.put<uint32_t>(d); w.put<uint32_t>(sl);
    w.put<int32_t>(kRotations); w.put<uint32_t>(1000); w.put<float>(0.005f);
    w.put<uint64_t>((uint64_t)g.m + 2);                  // estimates (crosspolytope.hpp:104-114)
    for (uint32_t b = 0; b < g.m + 2; b++) {
        w.put<uint64_t>((uint64_t)kEstBins);
        w.bytes(fs.est.data() + (size_t)b * kEstBins, sizeof(float) * kEstBins);
    }
    w.put<float>(0.005f);
    // ... functions (crosspolytope.hpp:178-184): ±1 signs per rotation
    const uint64_t n_fn = (uint64_t)L * g.fph;
    const uint32_t W = sign_words(g);
    w.put<uint64_t>(n_fn);
    std::vector<int8_t> signs((size_t)kRotations * g.npts);
    for (uint64_t f = 0; f < n_fn; f++) {
        w.put<int32_t>((int32_t)d); w.put<int32_t>((int32_t)g.m); w.put<uint32_t>(kRotations);
        for (uint32_t r = 0; r < (uint32_t)kRotations; r++)
            for (uint32_t i = 0; i < g.npts; i++)
                signs[(size_t)r * g.npts + i] = (fs.signbits[(f * kRotations + r) * W + i / 32] >> (i % 32)) & 1u ? -1 : 1;
        w.bytes(signs.data(), signs.size());
    }
    w.put<uint32_t>(L); w.put<uint32_t>(g.fph); w.put<uint8_t>((uint8_t)g.bpf); w.put<uint32_t>(0); w.put<uint32_t>(g.cut);
    // tables (collection.hpp:196-201 -> prefixmap.hpp:128-154): 12 + n_c + 12 entries, sentinels (0, 0xffffffff), prefix_index
    w.put<uint64_t>((uint64_t)L);
    w.put<uint8_t>(0);                                   // use_chunks
    const uint32_t len = nc + 2 * kSegment;
    std::vector<uint32_t> padded(len), pidx((1u << 13) + 1);
    for (uint32_t t = 0; t < L; t++) {
        const uint32_t* H = tbl_hash + (size_t)t * nc;
        const uint32_t* I = tbl_idx + (size_t)t * nc;
        w.put<uint64_t>((uint64_t)len);
        for (uint32_t i = 0; i < (uint32_t)kSegment; i++) padded[i] = padded[len - 1 - i] = 0u;
        memcpy(padded.data() + kSegment, I, (size_t)nc * 4);
        w.bytes(padded.data(), (uint64_t)len * 4);
        for (uint32_t i = 0; i < (uint32_t)kSegment; i++) padded[i] = padded[len - 1 - i] = 0xffffffffu;  // IMPOSSIBLE_PREFIX
        memcpy(padded.data() + kSegment, H, (size_t)nc * 4);
        w.bytes(padded.data(), (uint64_t)len * 4);
        w.put<uint64_t>(0);                              // rebuilding data
        w.put<uint32_t>(kMaxHashBits);                   // hash_length
        uint32_t idx = 0;                                // prefixmap.hpp:231-240: first position per 13-bit prefix
        for (uint32_t prefix = 0; prefix < (1u << 13); prefix++) {
            while (idx < nc && (H[idx] >> (kMaxHashBits - 13)) < prefix) idx++;
            pidx[prefix] = kSegment + idx;
        }
        pidx[1u << 13] = kSegment + nc;
        w.bytes(pidx.data(), pidx.size() * 4);
    }
    w.put<uint32_t>(nc);                                 // last_rebuild = points present at the last rebuild (collection.hpp:304)
    return std::move(w.out);
}

// Full reader of the same stream: everything a PUFFINN index needs to answer queries without being rebuilt.
struct LoadedStream {
    uint32_t d = 0, sl = 0, n = 0, L = 0;
    std::vector<int16_t> rows;        // [n][sl]
    std::vector<uint64_t> sketches;   // [n][32]
    std::vector<uint32_t> hashes;     // [L][n] unpadded
    std::vector<uint32_t> indices;    // [L][n]
    FunctionSet fs;
};

static void read_reference_stream(const uint8_t* blob, uint64_t len, LoadedStream& out) {
    Reader r{blob, len};
    out.d = r.get<uint32_t>();
    out.sl = r.get<uint32_t>();
    out.n = r.get<uint32_t>();
    if (out.d == 0 || out.d > 1024 || out.sl != (out.d + 15) / 16 * 16) throw StatusError(CLANN_ERR_SERIALIZE, "not a cosine PUFFINN index stream");
    if ((uint64_t)out.n * out.sl * 2 > len) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream truncated");
    out.rows.resize((size_t)out.n * out.sl);
    r.bytes(out.rows.data(), out.rows.size() * 2);
    // the function set: reuse the importer on the same bytes once the number of tables is known (it sits behind the functions)
    const HashGeom g1 = make_geom(out.d, 1);
    {
        Reader t{blob, len};
        t.bytes(nullptr, 12 + out.rows.size() * 2);
        t.bytes(nullptr, 4 + 8);                                            // sketch source type, SimHash description
        if (t.get<uint64_t>() != (uint64_t)kNumPlanes) throw StatusError(CLANN_ERR_SERIALIZE, "expected 2048 sketch functions");
        t.bytes(nullptr, (uint64_t)kNumPlanes * (4 + (uint64_t)out.sl * 2));
        t.bytes(nullptr, 4 + 4 + 1 + 4 + 4);
        const uint64_t n_sk = t.get<uint64_t>();
        if (n_sk != (uint64_t)out.n * kNumSketches) throw StatusError(CLANN_ERR_SERIALIZE, "sketch count does not match the dataset");
        out.sketches.resize(n_sk);
        t.bytes(out.sketches.data(), n_sk * 8);
        t.bytes(nullptr, 4 + 4 + 4 + 4);                                    // hash args
        if (!t.get<uint8_t>()) throw StatusError(CLANN_ERR_SERIALIZE, "the index was never rebuilt (no hash source)");
        t.bytes(nullptr, 8 + 12);                                           // family description + args
        const uint64_t rows = t.get<uint64_t>();
        if (rows > 64) throw StatusError(CLANN_ERR_SERIALIZE, "collision estimate table has the wrong height");
        for (uint64_t b = 0; b < rows; b++) {
            const uint64_t w = t.get<uint64_t>();
            if (w > len / 4) throw StatusError(CLANN_ERR_SERIALIZE, "reference stream truncated");
            t.bytes(nullptr, w * 4);
        }
        t.bytes(nullptr, 4);                                                // eps
        const uint64_t n_fn = t.get<uint64_t>();
        if (n_fn == 0 || n_fn % g1.fph || n_fn > len / 12) throw StatusError(CLANN_ERR_SERIALIZE, "unexpected number of hash functions");
        out.L = (uint32_t)(n_fn / g1.fph);
        t.bytes(nullptr, n_fn * (12 + (uint64_t)kRotations * g1.npts));
        if (t.get<uint32_t>() != out.L) throw StatusError(CLANN_ERR_SERIALIZE, "table count mismatch in the hash source");
        t.bytes(nullptr, 4 + 1 + 4 + 4);
        if (t.get<uint64_t>() != (uint64_t)out.L) throw StatusError(CLANN_ERR_SERIALIZE, "table count mismatch");
        if (t.get<uint8_t>() != 0) throw StatusError(CLANN_ERR_SERIALIZE, "chunked streams are not supported");
        out.hashes.resize((size_t)out.L * out.n);
        out.indices.resize((size_t)out.L * out.n);
        std::vector<uint32_t> padded;
        for (uint32_t tb = 0; tb < out.L; tb++) {                           // prefixmap.hpp:99-126
            const uint64_t plen = t.get<uint64_t>();
            if (plen != (uint64_t)out.n + 2 * kSegment) throw StatusError(CLANN_ERR_SERIALIZE, "table length does not match the dataset");
            padded.resize(plen);
            t.bytes(padded.data(), plen * 4);
            memcpy(out.indices.data() + (size_t)tb * out.n, padded.data() + kSegment, (size_t)out.n * 4);
            t.bytes(padded.data(), plen * 4);
            memcpy(out.hashes.data() + (size_t)tb * out.n, padded.data() + kSegment, (size_t)out.n * 4);
            {
                // the probe trusts the tables: every index must name a row, the codes must be 24-bit and sorted
                const uint32_t* I = out.indices.data() + (size_t)tb * out.n;
                const uint32_t* H = out.hashes.data() + (size_t)tb * out.n;
                for (uint32_t i = 0; i < out.n; i++) {
                    if (I[i] >= out.n) throw StatusError(CLANN_ERR_SERIALIZE, "table index out of range in the stream");
                    if (H[i] >> kMaxHashBits || (i && H[i] < H[i - 1])) throw StatusError(CLANN_ERR_SERIALIZE, "table codes are not sorted 24-bit values");
                }
            }
            if (t.get<uint64_t>() != 0) throw StatusError(CLANN_ERR_SERIALIZE, "the stream holds pending (unsorted) insertions");
            t.bytes(nullptr, 4 + ((1u << 13) + 1) * 4ull);                  // hash_length, prefix_index (rebuilt here as the 12-bit bucket directory)
        }
    }
    parse_reference_stream(blob, len, make_geom(out.d, out.L), out.fs, nullptr);
}

// ------------------------------------------------------------------------------------------------ the index

}  // namespace clann

using namespace clann;

// ------------------------------------------------------------------------------------------------ collectives

// NCCL is resolved at run time (dlopen): the library has no link-time dependency on it, a single-GPU user never loads it, and
// inside a process that already holds a copy (torch's bundled libnccl) that copy is the one found.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    static bool loaded = false;
    if (!loaded) {
        void* h = nullptr;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) throw StatusError(CLANN_ERR_CONFIG, "NCCL (libnccl.so.2) not found: clann_comm_init needs it; use clann_set_collectives for another transport");
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
        if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.AllReduce)
            throw StatusError(CLANN_ERR_CONFIG, "libnccl.so.2 lacks the expected entry points");
        loaded = true;
    }
    return api;
}

#define CLANN_NCCL(expr)                                                                                                  \
    do {                                                                                                                  \
        ncclResult_t _r = (expr);                                                                                         \
        if (_r != ncclSuccess)                                                                                            \
            throw CudaError(std::string(#expr) + ": " + (nccl_api().GetErrorString ? nccl_api().GetErrorString(_r) : "NCCL error")); \
    } while (0)

struct clann_index {
    // configuration
    clann_config cfg{};
    uint64_t n = 0;
    uint64_t n_rows = 0;  // rows of the PUFFINN-layer arrays on this rank: n, or the rows of the clusters this shard owns
    HashGeom g{};
    uint32_t K = 0;
    uint64_t seed = 0x5eedc1a7ull;
    bool per_cluster_functions = false;
    bool puffinn_mode = false;  // legacy single-index handle: one cluster, no CLANN layer
    uint32_t shard_rank = 0, shard_count = 1;
    bool clustering_imposed = false, built = false;

    // cluster-sharded search (search_sharded): the transport — NCCL (clann_comm_init) or caller-supplied collectives
    // (clann_set_collectives) — and its buffers, all sized for the global batch
    ncclComm_t comm = nullptr;
    clann_allgather_fn user_allgather = nullptr;
    clann_allreduce_min_u64_fn user_allreduce_min = nullptr;
    void* user_ctx = nullptr;


    // host mirrors
    std::vector<uint32_t> h_centers, h_sizes, h_assign;
    std::vector<uint64_t> h_offsets;
    std::vector<float> h_radii;
    std::vector<uint8_t> h_brute, h_owner;
    std::vector<uint32_t> h_fset_of;
    std::vector<FunctionSet> fsets;  // 1 (shared) or K
    std::vector<uint8_t> h_msd;
    double build_ms[5] = {0, 0, 0, 0, 0};  // gmm phase, hashing, sort, total, the k-center passes alone
    uint32_t visit_log_cap = 0;  // option "visit_log"

    // device: dataset + CLANN layer
    DevBuf<float> d_data, d_norms, d_dist, d_radii, d_center_rows, d_center_norms;
    DevBuf<uint32_t> d_assign, d_centers, d_sizes, d_perm, d_fset_of;
    DevBuf<uint64_t> d_keys, d_offsets;
    DevBuf<uint8_t> d_brute, d_owner, d_msd;
    DevBuf<uint32_t> d_msd_thr;
    // device: PUFFINN layer
    DevBuf<int16_t> d_q15, d_planes;
    DevBuf<uint8_t> d_plane_slices;  // int8 slices of the hyperplanes for the tensor-pipe sketch projection (kernels_tc.cu)
    DevBuf<uint64_t> d_sketches;
    DevBuf<uint32_t> d_tbl_hash, d_tbl_idx, d_tbl_dir, d_signbits, d_stop;
    uint32_t stop_words = 0;
    float stop_recall = -1.0f;

    // search workspace: everything one batch of queries needs between query preparation and the result copy. Set 0 serves
    // the stream-ordered calls; the others rotate under clann_search_device_async so that consecutive batches overlap.
    struct SearchWs {
        uint64_t ws_nq = 0;
        // side stream (high priority) for the first-visit anchors, which are latency-bound and independent of the dense similarities
        cudaStream_t aux = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        ~SearchWs() {
            if (aux) cudaStreamDestroy(aux);
            if (ev_fork) cudaEventDestroy(ev_fork);
            if (ev_join) cudaEventDestroy(ev_join);
        }
        bool ws_tc_center = false;  // the batch in this workspace was scored by the tensor-pipe screen (cdist exact only below exact_limit)
        bool ws_fs = false;  // the workspace was sized with the first-visit stream buffers (knob first_stream)
        uint32_t ws_vlog = 0;  // rows per query of w_visit_log
        DevBuf<uint32_t> w_visit_log;
        DevBuf<float> w_qnorm, w_cdist, w_exact_limit;
        DevBuf<int16_t> w_q15;
        DevBuf<uint32_t> w_codes, w_first, w_qperm, w_counter, w_vis, w_sort_k, w_sort_i;
        DevBuf<SortSegment> w_sort_seg;
        DevBuf<uint64_t> w_sketches;
        DevBuf<unsigned long long> w_cand, w_dc;
        DevBuf<uint8_t> w_state;
        DevBuf<uint16_t> w_memo;  // similarity memo scratch of the probe kernels
        uint64_t w_memo_stride = 0;
        uint32_t w_memo_slots = 0;
        DevBuf<uint16_t> w_dense;  // dense first-visit similarities [nq][w_dense_stride]
        uint64_t w_dense_stride = 0;
        DevBuf<unsigned long long> w_stats;           // {running sum of clusters visited, finished-block ticket} of k_finish
        DevBuf<uint32_t> w_pre_anchor, w_pre_range;  // first-visit anchors [nq][L] and ranges [nq][24][L]
        DevBuf<uint4> w_pre_lcp;                     // first-visit common-prefix samples [nq][L]
        // first-visit candidate stream (QueryBatch::fs_*): [nq][w_fs_cap] segments
        DevBuf<uint16_t> w_fs_idx;
        DevBuf<uint32_t> w_fs_hd, w_fs_meta;
        DevBuf<uint8_t> w_fs_tab;
        uint32_t w_fs_cap = 0;
        DevBuf<RowTile> w_tiles, w_tiles_codes;
        DevBuf<uint8_t> w_qslices;        // int8 slices of the query batch (tensor-pipe sketches)
        DevBuf<SketchTcTile> w_tc_tiles;  // 128-query tiles per function set
        uint32_t w_n_tc_tiles = 0;
        uint32_t w_ntiles = 0;
        uint64_t w_tiles_codes_nq = 0;
        // device staging of the host-buffer entry points
        DevBuf<float> h_queries, h_dists;
        DevBuf<uint32_t> h_ids, h_counts;
    };
    static constexpr int kPipeMax = 4;
    SearchWs wsv[1 + kPipeMax];
    SearchWs* W = &wsv[0];
    struct ShardLane {  // one sub-batch in flight: its buffers (sized for the sub-batch), its stream, its three workspaces
        DevBuf<uint32_t> first_all, list0, list1, list2, counts;
        DevBuf<unsigned long long> packed, packed1, packed2, top_local, top_all, counters;
        DevBuf<float> q0, q1;
        uint32_t* h_counts = nullptr;  // pinned: {routed to this rank in round one, still open after it, of those: served by this rank}
        uint64_t nq = 0;               // size the buffers were made for
        uint64_t q_lo = 0, q_n = 0;    // the sub-batch of the current call
        uint32_t n0 = 0, n1 = 0, n2 = 0;
        // streaming mode (sharded_submit): the batch this lane carries and the phase it runs at the next tick (0 route, 1 round one,
        // 2 scoring of the open queries, 3 round two + merge); phase < 0 = idle
        int phase = -1;
        const float* sq = nullptr;
        uint32_t* s_ids = nullptr;
        float* s_dists = nullptr;
        uint32_t* s_counts = nullptr;
        cudaEvent_t fork = nullptr;
        cudaStream_t stream = nullptr;
        cudaEvent_t ready = nullptr, done = nullptr;
        cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // phase boundaries of the last call
        SearchWs ws[3];                // route / open-query scoring, round one, round two
    };
    static constexpr int kShardLanes = 4;
    ShardLane lanes[kShardLanes];
    struct {
        bool last_was_sharded = false;
        int lanes_used = 0;
        cudaEvent_t fork = nullptr;
        uint64_t tick = 0;  // streaming mode: submissions so far
    } sh;
    // {clusters visited, queries} of the last finished batch, written by k_finish into mapped host memory and read without any
    // synchronisation: when queries walk many clusters (avg > 3: overlapping or unclustered data) the dense first-visit
    // precompute buys nothing and the 24-warp schedule of the gather path is the faster one
    unsigned long long* h_stats = nullptr;
    unsigned long long* h_stats_dev = nullptr;
    cudaStream_t pipe_stream[kPipeMax] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_done[kPipeMax] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t pipe_calls = 0;
    DevBuf<float> w_queries, w_out_dists;
    DevBuf<uint32_t> w_out_ids, w_out_counts;
    uint64_t last_nq = 0;
    const float* cur_queries = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t last_launches = 0, last_pre_launches = 0;
    bool profile_valid = false;

    ~clann_index() {
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : pipe_done)
            if (e) cudaEventDestroy(e);
        for (auto& st : pipe_stream)
            if (st) cudaStreamDestroy(st);
        if (h_stats) cudaFreeHost(h_stats);
        for (auto& ln : lanes) {
            if (ln.h_counts) cudaFreeHost(ln.h_counts);
            if (ln.stream) cudaStreamDestroy(ln.stream);
            for (cudaEvent_t e : {ln.ready, ln.done})
                if (e) cudaEventDestroy(e);
            for (auto& e : ln.ev)
                if (e) cudaEventDestroy(e);
        }
        if (sh.fork) cudaEventDestroy(sh.fork);
        if (comm) nccl_api().CommDestroy(comm);
    }

    void reset_workspaces() {
        if (pipe_calls) cudaDeviceSynchronize();  // batches still in flight on the internal streams
        for (auto& w : wsv) {
            w.ws_nq = 0;
            w.w_tiles_codes_nq = 0;
        }
        for (auto& ln : lanes)
            for (auto& w : ln.ws) {
                w.ws_nq = 0;
                w.w_tiles_codes_nq = 0;
            }
        if (h_stats) h_stats[0] = h_stats[1] = 0;
    }

    uint32_t n_fsets() const { return (uint32_t)fsets.size(); }

    // index.rs:78-80
    static uint32_t num_clusters(float factor, uint64_t n) {
        double v = floor((double)factor * sqrt((double)n));
        uint64_t k = v < 1.0 ? 1 : (uint64_t)v;
        return (uint32_t)k;
    }

    // dtype: 0 = f32, 1 = f16 (IEEE half; converted to f32 exactly on the device — the index is the one the reference would build
    // from the same rows widened to f32). on_device: `data` is a device pointer (copied / converted device to device).
    void init(const void* data, uint64_t n_, uint32_t d, const clann_config& c, int dtype = 0, bool on_device = false) {
        if (n_ == 0) throw StatusError(CLANN_ERR_DATA, "empty dataset");  // index.rs:72-74
        if (!data || d == 0) throw StatusError(CLANN_ERR_ARG, "null data or zero dimension");
        if (dtype != 0 && dtype != 1) throw StatusError(CLANN_ERR_ARG, "dtype must be 0 (f32) or 1 (f16)");
        if (d > 1024) throw StatusError(CLANN_ERR_CONFIG, "dimension above 1024 is not supported");
        if (c.num_tables == 0) throw StatusError(CLANN_ERR_CONFIG, "num tables should be >0");  // collection.hpp:242-244
        if (c.k == 0) throw StatusError(CLANN_ERR_CONFIG, "k must be at least 1");
        if (n_ > 0xfffffff0ull) throw StatusError(CLANN_ERR_CONFIG, "more than 2^32 points are not supported");
        int dev_count = 0;
        if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
            throw StatusError(CLANN_ERR_CUDA, "no CUDA device: libclann_b200 has no CPU fallback");
        cfg = c;
        n = n_;
        g = make_geom(d, (uint32_t)c.num_tables);
        K = puffinn_mode ? 1u : num_clusters(c.num_clusters_factor, n);
        const size_t count = (size_t)n * d;
        if (dtype == 0) {
            d_data.alloc(count);
            CLANN_CUDA(cudaMemcpy(d_data.p, data, count * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
        } else {
            d_data.alloc(count);
            // widen in chunks through a bounded staging buffer (host input) or straight from the caller's device buffer
            const size_t chunk = std::min<size_t>(count, (size_t)256 << 20);
            DevBuf<uint16_t> stage;
            if (!on_device) stage.alloc(chunk);
            for (size_t off = 0; off < count; off += chunk) {
                const size_t m = std::min(chunk, count - off);
                const uint16_t* src = static_cast<const uint16_t*>(data) + off;
                if (!on_device) {
                    CLANN_CUDA(cudaMemcpy(stage.p, src, m * sizeof(uint16_t), cudaMemcpyHostToDevice));
                    src = stage.p;
                }
                launch_widen_f16(src, m, d_data.p + off, 0);
            }
            CLANN_CUDA(cudaDeviceSynchronize());
        }
        for (auto& e : ev)
            if (!e) CLANN_CUDA(cudaEventCreate(&e));
    }

    void set_clustering(uint64_t K_, const uint64_t* centers, const uint64_t* assignment, const float* radii) {
        if (K_ == 0 || !centers || !assignment || !radii) throw StatusError(CLANN_ERR_ARG, "bad clustering");
        K = (uint32_t)K_;
        h_centers.resize(K);
        h_radii.assign(radii, radii + K);
        for (uint32_t c = 0; c < K; c++) {
            if (centers[c] >= n) throw StatusError(CLANN_ERR_BOUNDS, "centre id out of range");
            h_centers[c] = (uint32_t)centers[c];
        }
        h_assign.resize(n);
        for (uint64_t i = 0; i < n; i++) {
            if (assignment[i] >= K) throw StatusError(CLANN_ERR_BOUNDS, "assignment out of range");
            h_assign[i] = (uint32_t)assignment[i];
        }
        clustering_imposed = true;
        built = false;
    }

    void ensure_fsets() {
        size_t want = per_cluster_functions ? K : 1;
        if (fsets.size() != want) fsets.resize(want);
    }

    void import_reference(uint64_t cluster, const void* blob, uint64_t len) {
        if (cluster >= K) throw StatusError(CLANN_ERR_BOUNDS, "cluster out of range");
        per_cluster_functions = true;
        ensure_fsets();
        parse_reference_stream(static_cast<const uint8_t*>(blob), len, g, fsets[cluster], nullptr);
        built = false;
    }

    void set_functions(uint64_t cluster, const int16_t* planes, const int8_t* signs, const float* est) {
        FunctionSet* fs;
        if (cluster == UINT64_MAX) {
            per_cluster_functions = false;
            ensure_fsets();
            fs = &fsets[0];
        } else {
            if (cluster >= K) throw StatusError(CLANN_ERR_BOUNDS, "cluster out of range");
            per_cluster_functions = true;
            ensure_fsets();
            fs = &fsets[cluster];
        }
        if (planes && signs) {
            fs->planes.assign(planes, planes + (size_t)kNumPlanes * g.sl);
            pack_signs(g, signs, fs->signbits);
            fs->have_fn = true;
        }
        if (est) {
            fs->est.assign(est, est + (size_t)(g.m + 2) * kEstBins);
            fs->have_est = true;
        }
        built = false;
    }

    // A PUFFINN index straight from its serialized stream (collection.hpp:147-170): rows, sketches, tables and functions are
    // taken as they are, nothing is rehashed; only the bucket directory and the host-built stop / threshold tables are derived.
    void load_stream(const uint8_t* blob, uint64_t len) {
        int dev_count = 0;
        if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
            throw StatusError(CLANN_ERR_CUDA, "no CUDA device: libclann_b200 has no CPU fallback");
        LoadedStream ls;
        read_reference_stream(blob, len, ls);
        cudaStream_t s = 0;
        puffinn_mode = true;
        cfg = clann_config{ls.L, 1.0f, 10, 0.9f};
        n = ls.n;
        n_rows = n;
        g = make_geom(ls.d, ls.L);
        K = 1;
        for (auto& e : ev)
            if (!e) CLANN_CUDA(cudaEventCreate(&e));
        h_centers.assign(1, 0u);
        h_radii.assign(1, 0.0f);
        h_assign.assign(n, 0u);
        h_sizes.assign(1, (uint32_t)n);
        h_offsets = {0, n};
        h_brute.assign(1, 0);
        h_owner.assign(1, 0);
        h_fset_of.assign(1, 0u);
        per_cluster_functions = false;
        fsets.clear();
        fsets.push_back(std::move(ls.fs));
        std::vector<uint32_t> ident(n);
        std::iota(ident.begin(), ident.end(), 0u);
        d_perm.upload(ident, s);
        d_offsets.upload(h_offsets, s);
        d_brute.upload(h_brute, s);
        d_owner.upload(h_owner, s);
        d_fset_of.upload(h_fset_of, s);
        d_radii.upload(h_radii, s);
        d_centers.upload(h_centers, s);
        d_q15.upload(ls.rows, s);
        d_sketches.upload(ls.sketches, s);
        d_tbl_hash.upload(ls.hashes, s);   // one cluster: cluster-major == table-major
        d_tbl_idx.upload(ls.indices, s);
        d_tbl_dir.alloc((size_t)g.L * kDirEntries);
        launch_build_dir(d_tbl_hash.p, n, d_offsets.p, d_brute.p, 1, g.L, d_tbl_dir.p, s);
        prepare_functions(s);
        CLANN_CUDA(cudaStreamSynchronize(s));
        CLANN_CUDA(cudaGetLastError());
        reset_workspaces();
        built = true;
    }

    // ---- build -----------------------------------------------------------------------------------------------

    void run_gmm(cudaStream_t s) {
        // angulardata.rs:12-19
        d_norms.alloc(n);
        launch_row_norms(d_data.p, n, g.d, d_norms.p, s);
        build_ms[4] = 0.0;
        if (clustering_imposed) return;
        if (n <= K) {  // gmm.rs:26-31: every point its own centre
            K = (uint32_t)n;
            h_centers.resize(K);
            std::iota(h_centers.begin(), h_centers.end(), 0u);
            h_assign.resize(n);
            std::iota(h_assign.begin(), h_assign.end(), 0u);
            h_radii.assign(K, 0.0f);
            return;
        }
        // Sharded clustering, when a communicator exists at build time: every rank runs the K passes over ITS rows only
        // (rows are replicated, so the newest centre's row is at hand everywhere) and the packed arg-max key of each pass is
        // all-reduced (max, 8 bytes, in stream); the distances and assignments of the shares are all-gathered once at the end. The
        // arg-max of the union is the max of the shares' arg-maxes under the same key, so the clustering is the single-GPU one.
        const bool shard_gmm = shard_count > 1 && comm && !user_allgather && tune_get("shard_gmm", 1) != 0;
        const uint64_t chunk = shard_gmm ? (n + shard_count - 1) / shard_count : n;
        const uint64_t row0 = shard_gmm ? std::min<uint64_t>(n, (uint64_t)shard_rank * chunk) : 0;
        const uint64_t row1 = shard_gmm ? std::min<uint64_t>(n, row0 + chunk) : n;
        d_dist.alloc(shard_gmm ? chunk * shard_count : n);
        d_assign.alloc(shard_gmm ? chunk * shard_count : n);
        d_keys.alloc(K);
        d_keys.zero(s);
        DevBuf<float> d_cc;
        d_cc.alloc(K);
        cudaEvent_t p0, p1;  // the K passes alone (build_ms[4]): what the build roofline is about
        CLANN_CUDA(cudaEventCreate(&p0)); CLANN_CUDA(cudaEventCreate(&p1));
        CLANN_CUDA(cudaEventRecord(p0, s));
        for (uint32_t c = 0; c < K; c++) {
            launch_gmm_pass(d_data.p, d_norms.p, row0, row1, g.d, c, d_keys.p, d_dist.p, d_assign.p, d_cc.p, s);
            if (shard_gmm) CLANN_NCCL(nccl_api().AllReduce(d_keys.p + c, d_keys.p + c, 1, ncclUint64, ncclMax, comm, s));
        }
        CLANN_CUDA(cudaEventRecord(p1, s));
        if (shard_gmm) {
            CLANN_NCCL(nccl_api().AllGather(d_dist.p + shard_rank * chunk, d_dist.p, chunk, ncclFloat32, comm, s));
            CLANN_NCCL(nccl_api().AllGather(d_assign.p + shard_rank * chunk, d_assign.p, chunk, ncclUint32, comm, s));
        }
        d_centers.alloc(K);
        d_radii.alloc(K);
        d_radii.zero(s);
        d_sizes.alloc(K);
        d_sizes.zero(s);
        launch_gmm_finish(d_keys.p, K, n, d_dist.p, d_assign.p, d_centers.p, d_radii.p, d_sizes.p, s);
        CLANN_CUDA(cudaStreamSynchronize(s));
        {
            float ms = 0.0f;
            CLANN_CUDA(cudaEventElapsedTime(&ms, p0, p1));
            build_ms[4] = ms;
            cudaEventDestroy(p0); cudaEventDestroy(p1);
        }
        h_centers = d_centers.download(K);
        h_radii = d_radii.download(K);
        h_assign = d_assign.download(n);
        d_dist.release();
        d_keys.release();
    }

    // index.rs:188-192 — per-cluster member lists in ascending point order, laid out back to back.
    void invert_assignment(cudaStream_t s) {
        h_sizes.assign(K, 0);
        for (uint64_t i = 0; i < n; i++) h_sizes[h_assign[i]]++;
        h_offsets.assign(K + 1, 0);
        for (uint32_t c = 0; c < K; c++) h_offsets[c + 1] = h_offsets[c] + h_sizes[c];
        DevBuf<uint32_t> keys, scratch_k, scratch_i;
        keys.upload(h_assign, s);
        d_perm.alloc(n);
        // stable partition by cluster id: counting sort over warp-sized chunks of the rows (knob assign_partition, default 1) ...
        if (tune_get("assign_partition", 1) != 0 && n < (1ull << 32)) {
            uint64_t chunks = std::min<uint64_t>((n + 31) / 32, 4096);
            while (chunks > 1 && chunks * K * sizeof(uint32_t) > (256ull << 20)) chunks >>= 1;
            DevBuf<uint32_t> counts;
            counts.alloc(chunks * K);
            d_offsets.upload(h_offsets, s);
            launch_assign_partition(keys.p, n, K, d_offsets.p, counts.p, (uint32_t)chunks, d_perm.p, s);
            CLANN_CUDA(cudaStreamSynchronize(s));
            return;
        }
        // ... or the same radix sort the tables use, one segment of n keys (one CTA)
        DevBuf<SortSegment> seg;
        std::vector<SortSegment> hs(1);
        hs[0] = SortSegment{0, 0, (uint32_t)n, 0};
        seg.upload(hs, s);
        if (n > segment_sort_smem_capacity()) {
            scratch_k.alloc(n);
            scratch_i.alloc(n);
        }
        launch_segment_sort(seg.p, 1, (uint32_t)n, keys.p, d_perm.p, scratch_k.p, scratch_i.p, s);
        d_offsets.upload(h_offsets, s);
        CLANN_CUDA(cudaStreamSynchronize(s));
    }

    void assign_owners() {
        // longest-processing-time on cluster sizes; deterministic, so every rank derives the same map
        h_owner.assign(K, 0);
        if (shard_count <= 1) return;
        std::vector<uint32_t> order(K);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return h_sizes[a] > h_sizes[b]; });
        std::vector<uint64_t> load(shard_count, 0);
        for (uint32_t c : order) {
            uint32_t best = 0;
            for (uint32_t r = 1; r < shard_count; r++)
                if (load[r] < load[best]) best = r;
            h_owner[c] = (uint8_t)best;
            load[best] += h_sizes[c];
        }
    }

    // Per-cluster sets that were all imported and turn out identical (an index this library saved with its shared set and
    // loads again through clann_import_reference) collapse into one shared set: a query is then hashed once, not per cluster.
    void collapse_identical_function_sets() {
        if (!per_cluster_functions || fsets.size() <= 1 || fsets.size() != K) return;
        const FunctionSet* first = nullptr;
        for (size_t f = 0; f < fsets.size(); f++) {
            if (h_brute[f]) continue;
            const FunctionSet& fs = fsets[f];
            if (!fs.have_fn || !fs.have_est) return;
            if (!first) first = &fs;
            else if (fs.planes != first->planes || fs.signbits != first->signbits || fs.est != first->est) return;
        }
        if (!first) return;
        FunctionSet keep = *first;
        fsets.assign(1, keep);
        per_cluster_functions = false;
    }

    void prepare_functions(cudaStream_t s) {
        ensure_fsets();
        // collision estimates: one Monte-Carlo table per dimension, shared by every set that did not import its own
        std::vector<float> shared_est;
        auto need_est = [&]() -> const std::vector<float>& {
            if (shared_est.empty()) {
                DevBuf<float> d_est;
                DevBuf<uint32_t> d_counts;
                d_est.alloc((size_t)(g.m + 2) * kEstBins);
                d_counts.alloc((size_t)(g.m + 2) * kEstBins);
                launch_cp_estimates(g.m, 1000, seed ^ 0xC0111510ull, d_est.p, d_counts.p, s);  // crosspolytope.hpp:221-226
                CLANN_CUDA(cudaStreamSynchronize(s));
                shared_est = d_est.download((size_t)(g.m + 2) * kEstBins);
            }
            return shared_est;
        };
        for (size_t f = 0; f < fsets.size(); f++) {
            FunctionSet& fs = fsets[f];
            bool used = !per_cluster_functions || !h_brute[f];
            if (!used) {
                // keep array shapes uniform for unused slots
                if (!fs.have_fn) {
                    fs.planes.assign((size_t)kNumPlanes * g.sl, 0);
                    fs.signbits.assign((size_t)g.L * g.fph * kRotations * sign_words(g), 0u);
                }
                if (!fs.have_est) fs.est.assign((size_t)(g.m + 2) * kEstBins, 1.0f);
                continue;
            }
            if (!fs.have_fn) generate_functions(g, seed + 0x9E3779B97F4A7C15ull * (f + 1), fs);
            if (!fs.have_est) {
                fs.est = need_est();
                fs.have_est = true;
            }
        }
        // upload
        const size_t plane_sz = (size_t)kNumPlanes * g.sl, sign_sz = (size_t)g.L * g.fph * kRotations * sign_words(g);
        std::vector<int16_t> all_planes(plane_sz * fsets.size());
        std::vector<uint32_t> all_signs(sign_sz * fsets.size());
        for (size_t f = 0; f < fsets.size(); f++) {
            memcpy(all_planes.data() + f * plane_sz, fsets[f].planes.data(), plane_sz * sizeof(int16_t));
            memcpy(all_signs.data() + f * sign_sz, fsets[f].signbits.data(), sign_sz * sizeof(uint32_t));
        }
        d_planes.upload(all_planes, s);
        d_signbits.upload(all_signs, s);
        if (sketch_tc_supported(g.sl)) {
            d_plane_slices.alloc((size_t)fsets.size() * kNumPlanes * 2 * sketch_tc_kp(g.sl));
            launch_q15_slices(d_planes.p, (uint64_t)fsets.size() * kNumPlanes, g.sl, d_plane_slices.p, s);
        }
        build_msd_table(h_msd);
        d_msd.upload(h_msd, s);
        {
            // threshold form of the same table for the probe kernel: the table is non-increasing in v, so
            // msd[v] = #{m in [0,64) : thr[m] > v} with thr[m] = smallest v whose entry is <= m
            std::vector<uint32_t> thr(64, 65536u);
            for (uint32_t v = 65536; v-- > 0;) {
                if (v + 1 < 65536 && h_msd[v] < h_msd[v + 1]) throw StatusError(CLANN_ERR_CUDA, "max_sketch_diff table is not monotone");
                for (uint32_t m = h_msd[v]; m < 64; m++) thr[m] = v;
            }
            d_msd_thr.upload(thr, s);
        }
        CLANN_CUDA(cudaStreamSynchronize(s));
        stop_recall = -1.0f;
    }

    // Stop tables for a recall value (Config.delta for the batched ABI; per call for the legacy ABI).
    void ensure_stop_table(float recall, cudaStream_t s) {
        if (recall == stop_recall && d_stop.p) return;
        stop_words = (g.L + 1 + 31) / 32;
        const size_t per = (size_t)kMaxHashBits * kEstBins * stop_words;
        std::vector<uint32_t> all(per * fsets.size());
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (;;) {
                size_t f = next.fetch_add(1);
                if (f >= fsets.size()) break;
                build_stop_table(g, fsets[f].est.data(), recall, stop_words, all.data() + f * per);
            }
        };
        unsigned nt = std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), fsets.size());
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        d_stop.upload(all, s);
        CLANN_CUDA(cudaStreamSynchronize(s));
        stop_recall = recall;
    }

    // knob tc_center (default 1): query x centre scoring as a tf32 tensor-core GEMM screen + exact evaluation of each query's 32
    // nearest candidates (k_center_gemm_tc, k_center_refine) instead of the all-exact CUDA-core kernel. Queries that walk past
    // their 32 candidates re-evaluate their row inside the probe, so when the last finished batch averaged more than 8 clusters
    // per query (overlapping or unclustered data) the all-exact kernel is used from the start. Never changes a result.
    bool use_tc_center() const {
        if (!center_gemm_tc_supported(g.d) || puffinn_mode || tune_get("tc_center", 1) == 0) return false;
        if (h_stats && tune_get("dense_adaptive", 1) != 0) {
            const unsigned long long visited = reinterpret_cast<volatile unsigned long long*>(h_stats)[0];
            const unsigned long long queries = reinterpret_cast<volatile unsigned long long*>(h_stats)[1];
            if (queries > 0 && visited > 8 * queries) return false;
        }
        return true;
    }

    // knob tc_sketch (default 1): the sketch projection on the tensor pipe (k_sketch_tc) instead of the CUDA-core kernel
    bool use_tc_sketch() const { return sketch_tc_supported(g.sl) && d_plane_slices.p && tune_get("tc_sketch", 1) != 0; }

    // filterer.hpp:76-97 for every PUFFINN cluster this rank builds, on the tensor pipe: tiles of up to 128 consecutive rows of
    // one cluster; the int8 slices of the rows go through a bounded scratch buffer (at most ~4 M rows at a time).
    void build_sketches_tc(cudaStream_t s) {
        const uint32_t kp = sketch_tc_kp(g.sl);
        const uint64_t cap_rows = (uint64_t)std::max<int64_t>(1 << 16, tune_get("tc_sketch_chunk_rows", 4 << 20));
        DevBuf<uint8_t> slices;
        DevBuf<SketchTcTile> d_tc;
        std::vector<SketchTcTile> tc_tiles;
        uint64_t lo = 0, hi = 0;  // row span [lo, hi) covered by the pending tiles
        auto flush = [&]() {
            if (tc_tiles.empty()) return;
            slices.ensure((hi - lo) * 2 * kp);
            launch_q15_slices(d_q15.p + lo * g.sl, hi - lo, g.sl, slices.p, s);
            d_tc.upload(tc_tiles, s);
            launch_sketch_tc(d_tc.p, (uint32_t)tc_tiles.size(), slices.p, (uint32_t)lo, hi - lo, d_plane_slices.p, n_fsets(), d_q15.p,
                             d_planes.p, g.sl, d_sketches.p, s);
            CLANN_CUDA(cudaStreamSynchronize(s));  // the host vectors and the scratch are reused by the next chunk
            tc_tiles.clear();
        };
        for (uint32_t c = 0; c < K; c++) {
            if (h_brute[c] || h_sizes[c] == 0) continue;
            if (shard_count > 1 && h_owner[c] != shard_rank) continue;
            const uint64_t c_lo = h_offsets[c], c_hi = h_offsets[c] + h_sizes[c];
            if (!tc_tiles.empty() && c_hi - lo > cap_rows) flush();
            if (tc_tiles.empty()) lo = c_lo;
            hi = c_hi;
            for (uint32_t r = 0; r < h_sizes[c]; r += 128) {
                const uint32_t row0 = (uint32_t)c_lo + r;
                tc_tiles.push_back(SketchTcTile{row0, row0, std::min<uint32_t>(128, h_sizes[c] - r), h_fset_of[c]});
            }
        }
        flush();
    }

    void build() {
        cudaStream_t s = 0;
        if (shard_rank >= shard_count) throw StatusError(CLANN_ERR_CONFIG, "shard_rank must be below shard_count");
        reset_workspaces();  // the search workspace (memo stride, tiles) depends on the clustering
        cudaEvent_t e0, e1, e2, e3;
        CLANN_CUDA(cudaEventCreate(&e0)); CLANN_CUDA(cudaEventCreate(&e1)); CLANN_CUDA(cudaEventCreate(&e2)); CLANN_CUDA(cudaEventCreate(&e3));
        CLANN_CUDA(cudaEventRecord(e0, s));
        run_gmm(s);
        invert_assignment(s);
        // ClusterCenter records (index.rs:194-220)
        h_brute.resize(K);
        for (uint32_t c = 0; c < K; c++)
            h_brute[c] = puffinn_mode ? 0 : (h_sizes[c] < 100 || h_sizes[c] < cfg.k);  // index.rs:204-205
        collapse_identical_function_sets();
        h_fset_of.resize(K);
        for (uint32_t c = 0; c < K; c++) h_fset_of[c] = per_cluster_functions ? c : 0;
        assign_owners();
        n_rows = n;
        if (shard_count > 1) {
            // A shard keeps the PUFFINN layer (Q15 rows, sketches, tables) of ITS clusters only, packed back to back: the per-GPU
            // footprint of that layer is 1/shard_count of the index (the fp32 rows, centres and radii stay replicated — greedy
            // k-center and the final distances read them). Foreign clusters become empty in this rank's layout; they are never
            // probed here (the sharded entry points stop at / skip them).
            std::vector<uint32_t> perm = d_perm.download(n), local;
            local.reserve(n / shard_count + 1024);
            std::vector<uint64_t> off(K + 1, 0);
            for (uint32_t c = 0; c < K; c++) {
                if (h_owner[c] == shard_rank) local.insert(local.end(), perm.begin() + h_offsets[c], perm.begin() + h_offsets[c] + h_sizes[c]);
                else h_sizes[c] = 0;
                off[c + 1] = local.size();
            }
            h_offsets = off;
            n_rows = local.size();
            if (local.empty()) local.push_back(0);
            d_perm.upload(local, s);
            d_offsets.upload(h_offsets, s);
            CLANN_CUDA(cudaStreamSynchronize(s));
        }
        d_brute.upload(h_brute, s);
        d_fset_of.upload(h_fset_of, s);
        d_owner.upload(h_owner, s);
        d_radii.upload(h_radii, s);
        d_centers.upload(h_centers, s);
        d_center_rows.alloc((size_t)K * g.d);
        launch_gather_rows(d_data.p, d_centers.p, K, g.d, d_center_rows.p, s);
        d_center_norms.alloc(K);
        {
            std::vector<float> norms = d_norms.download(n);
            std::vector<float> cn(K);
            for (uint32_t c = 0; c < K; c++) cn[c] = norms[h_centers[c]];
            d_center_norms.upload(cn, s);
            CLANN_CUDA(cudaStreamSynchronize(s));
        }
        CLANN_CUDA(cudaEventRecord(e1, s));

        prepare_functions(s);

        // PUFFINN layer: Q15 rows in cluster order (puffinn.rs:41-46 -> dataset.hpp:109-124)
        d_q15.alloc((size_t)std::max<uint64_t>(n_rows, 1) * g.sl);
        launch_store_q15(d_data.p, d_perm.p, n_rows, g.d, g.sl, d_q15.p, s);
        std::vector<RowTile> tiles;
        std::vector<SortSegment> segs;
        uint32_t max_len = 0;
        uint64_t scratch_total = 0;
        const uint32_t cap = segment_sort_smem_capacity();
        for (uint32_t c = 0; c < K; c++) {
            if (h_brute[c] || h_sizes[c] == 0) continue;
            if (shard_count > 1 && h_owner[c] != shard_rank) continue;  // other ranks build their own clusters
            for (uint32_t r = 0; r < h_sizes[c]; r += 32) {
                uint32_t row0 = (uint32_t)h_offsets[c] + r;
                tiles.push_back(RowTile{row0, row0, std::min<uint32_t>(32, h_sizes[c] - r), h_fset_of[c], h_sizes[c], 0,
                                        (uint64_t)g.L * h_offsets[c] + r});
            }
            max_len = std::max(max_len, h_sizes[c]);
        }
        DevBuf<RowTile> d_tiles;
        d_tiles.upload(tiles, s);
        d_sketches.alloc((size_t)std::max<uint64_t>(n_rows, 1) * kNumSketches);
        if (use_tc_sketch()) build_sketches_tc(s);
        else launch_sketch(d_q15.p, d_tiles.p, (uint32_t)tiles.size(), d_planes.p, g.sl, d_sketches.p, s);
        d_tbl_hash.alloc((size_t)g.L * std::max<uint64_t>(n_rows, 1));
        d_tbl_idx.alloc((size_t)g.L * std::max<uint64_t>(n_rows, 1));
        launch_codes(d_q15.p, d_tiles.p, (uint32_t)tiles.size(), d_signbits.p, g, d_tbl_hash.p, n_rows, 0, s);
        CLANN_CUDA(cudaEventRecord(e2, s));
        // collection.hpp:299-302 -> prefixmap.hpp:169-247: one stable radix sort per (cluster, table)
        for (uint32_t c = 0; c < K; c++) {
            if (h_brute[c] || h_sizes[c] == 0) continue;
            if (shard_count > 1 && h_owner[c] != shard_rank) continue;
            for (uint32_t t = 0; t < g.L; t++) {
                SortSegment sg{table_base(h_offsets[c], h_sizes[c], g.L, t), 0, h_sizes[c], 0};
                if (h_sizes[c] > cap) {
                    sg.scratch_base = scratch_total;
                    scratch_total += h_sizes[c];
                }
                segs.push_back(sg);
            }
        }
        DevBuf<SortSegment> d_segs;
        DevBuf<uint32_t> scratch_k, scratch_i;
        d_segs.upload(segs, s);
        if (scratch_total) {
            scratch_k.alloc(scratch_total);
            scratch_i.alloc(scratch_total);
        }
        launch_segment_sort(d_segs.p, (uint32_t)segs.size(), max_len, d_tbl_hash.p, d_tbl_idx.p, scratch_k.p, scratch_i.p, s);
        {
            // bucket directory over the top 12 code bits of every table this rank built
            std::vector<uint8_t> skip(K);
            for (uint32_t c = 0; c < K; c++)
                skip[c] = h_brute[c] || h_sizes[c] == 0 || (shard_count > 1 && h_owner[c] != shard_rank);
            DevBuf<uint8_t> d_skip;
            d_skip.upload(skip, s);
            d_tbl_dir.alloc((size_t)g.L * K * kDirEntries);
            launch_build_dir(d_tbl_hash.p, n_rows, d_offsets.p, d_skip.p, K, g.L, d_tbl_dir.p, s);
            CLANN_CUDA(cudaStreamSynchronize(s));
        }
        CLANN_CUDA(cudaEventRecord(e3, s));
        CLANN_CUDA(cudaStreamSynchronize(s));
        CLANN_CUDA(cudaGetLastError());
        float ms;
        CLANN_CUDA(cudaEventElapsedTime(&ms, e0, e1)); build_ms[0] = ms;
        CLANN_CUDA(cudaEventElapsedTime(&ms, e1, e2)); build_ms[1] = ms;
        CLANN_CUDA(cudaEventElapsedTime(&ms, e2, e3)); build_ms[2] = ms;
        CLANN_CUDA(cudaEventElapsedTime(&ms, e0, e3)); build_ms[3] = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
        ensure_stop_table(cfg.delta, s);
        built = true;
    }

    // ---- search ----------------------------------------------------------------------------------------------

    SearchParams params() const {
        SearchParams p{};
        p.g = g;
        p.k = (uint32_t)cfg.k;
        p.K = K;
        p.n_fsets = n_fsets();
        p.n = n;
        p.stop_words = stop_words;
        p.data = d_data.p;
        p.norms = d_norms.p;
        p.perm = d_perm.p;
        p.offsets = d_offsets.p;
        p.q15 = d_q15.p;
        p.sketches = d_sketches.p;
        p.tbl_hash = d_tbl_hash.p;
        p.tbl_idx = d_tbl_idx.p;
        p.tbl_dir = d_tbl_dir.p;
        p.center_rows = d_center_rows.p;
        p.center_norms = d_center_norms.p;
        p.radii = d_radii.p;
        p.brute = d_brute.p;
        p.fset_of = d_fset_of.p;
        p.owner = d_owner.p;
        p.stop = d_stop.p;
        p.msd = d_msd.p;
        p.msd_thr = d_msd_thr.p;
        p.shard_rank = shard_rank;
        p.max_cluster = h_sizes.empty() ? 0u : *std::max_element(h_sizes.begin(), h_sizes.end());
        p.prefetch_rows = 0;
        p.reserve_sms = 0;
        return p;
    }

    void ensure_workspace(uint64_t nq, cudaStream_t s) {
        const bool want_fs = tune_get("first_stream", 0) != 0;
        if (nq == W->ws_nq && want_fs == W->ws_fs && visit_log_cap == W->ws_vlog) return;
        W->ws_fs = want_fs;
        W->ws_vlog = visit_log_cap;
        const uint32_t F = n_fsets();
        const uint32_t k = (uint32_t)cfg.k;
        // capacity in steps of 2048 queries: batches of slightly different sizes (the sharded search) reuse the allocations
        const uint64_t cq = (nq + 2047) / 2048 * 2048;
        W->w_qnorm.ensure(cq);
        W->w_q15.ensure(cq * g.sl);
        W->w_codes.ensure((size_t)F * g.L * cq);
        W->w_sketches.ensure((size_t)F * cq * kNumSketches);
        W->w_cdist.ensure(cq * K);
        W->w_exact_limit.ensure(cq);
        W->w_first.ensure(cq);
        W->w_qperm.ensure(cq);
        if (nq > segment_sort_smem_capacity()) {
            W->w_sort_k.ensure(cq);
            W->w_sort_i.ensure(cq);
        }
        W->w_sort_seg.ensure(1);
        W->w_state.ensure(cq * query_state_bytes(k));
        {
            // one memo region (u16 per local id of the largest cluster) per resident probe warp; skipped beyond 1 GiB
            const uint32_t max_cluster = h_sizes.empty() ? 0u : *std::max_element(h_sizes.begin(), h_sizes.end());
            W->w_memo_stride = ((uint64_t)max_cluster + 7) & ~7ull;
            W->w_memo_slots = probe_memo_slots();
            const uint64_t need = W->w_memo_stride * W->w_memo_slots;
            if (need > 0 && need * sizeof(uint16_t) <= (1ull << 30)) W->w_memo.ensure(need);
            else W->w_memo_slots = 0;
        }
        {
            // dense first-visit similarities: one u16 per (query, row of the largest cluster); skipped beyond 2 GiB
            const uint32_t max_cluster = h_sizes.empty() ? 0u : *std::max_element(h_sizes.begin(), h_sizes.end());
            W->w_dense_stride = ((uint64_t)max_cluster + 63) & ~63ull;
            const uint64_t need = W->w_dense_stride * cq;
            if (need > 0 && need * sizeof(uint16_t) <= (2ull << 30) && !puffinn_mode) W->w_dense.ensure(need);
            else W->w_dense_stride = 0;
            const uint64_t words = cq * kMaxHashBits * g.L;
            if (W->w_dense_stride && words * 4 <= (2ull << 30)) {
                W->w_pre_anchor.ensure(cq * g.L);
                W->w_pre_range.ensure(words);
                W->w_pre_lcp.ensure(cq * g.L);
            }
            // first-visit candidate stream: room for 2 x the largest cluster in segments per query (a visit scans ~4 candidates
            // per cluster row on the planted shapes, p99 ~2x that), 12 bytes per segment; only what a visit needs is written
            W->w_fs_cap = 0;
            if (W->w_dense_stride && max_cluster <= 65536u && g.L <= 255 && want_fs) {
                uint64_t cap = std::min<uint64_t>(std::max<uint64_t>(2ull * max_cluster, 1024), 16384);
                const int64_t knob = tune_get("first_stream_cap", 0);
                if (knob > 0) cap = (uint64_t)knob;
                cap = (cap + 31) & ~31ull;
                if (cq * cap * 12 <= (6ull << 30)) {
                    W->w_fs_idx.ensure(cq * cap * 4);
                    W->w_fs_hd.ensure(cq * cap);
                    W->w_fs_tab.ensure(cq * cap / 32);
                    W->w_fs_meta.ensure(cq * kFsMeta);
                    W->w_fs_cap = (uint32_t)cap;
                }
            }
        }
        W->w_counter.ensure(2);
        if (!W->w_stats.p) {
            W->w_stats.ensure(2);
            CLANN_CUDA(cudaMemsetAsync(W->w_stats.p, 0, 2 * sizeof(unsigned long long), s));
        }
        if (!h_stats) {
            CLANN_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_stats), 2 * sizeof(unsigned long long), cudaHostAllocMapped));
            h_stats[0] = h_stats[1] = 0;
            CLANN_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&h_stats_dev), h_stats, 0));
        }
        if (visit_log_cap) W->w_visit_log.ensure(cq * visit_log_cap * 4);
        W->w_cand.ensure(cq);
        W->w_dc.ensure(cq);
        W->w_vis.ensure(cq);
        // The tile lists of the hashing kernels and the sort segment are written by a kernel: a batch of a new size (the sub-batches of
        // the sharded search change size with every call) costs no host-device synchronisation.
        const uint32_t per_f = (uint32_t)((nq + 31) / 32), per_f_tc = (uint32_t)((nq + 127) / 128);
        W->w_ntiles = per_f * F;
        W->w_tiles.ensure(W->w_ntiles);
        W->w_tiles_codes.ensure(W->w_ntiles);
        W->w_n_tc_tiles = 0;
        if (sketch_tc_supported(g.sl) && nq < (1ull << 31)) {
            W->w_n_tc_tiles = per_f_tc * F;
            W->w_tc_tiles.ensure(W->w_n_tc_tiles);
            W->w_qslices.ensure(cq * 2 * sketch_tc_kp(g.sl));
        }
        launch_query_tiles(nq, F, W->w_tiles.p, W->w_tiles_codes.p, W->w_n_tc_tiles ? W->w_tc_tiles.p : nullptr, W->w_sort_seg.p, s);
        W->w_tiles_codes_nq = nq;
        W->ws_nq = nq;
    }

    QueryBatch batch(const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts) {
        QueryBatch b{};
        b.nq = nq;
        b.queries = d_queries;
        b.qnorm = W->w_qnorm.p;
        b.q15 = W->w_q15.p;
        b.codes = W->w_codes.p;
        b.sketches = W->w_sketches.p;
        b.cdist = W->w_cdist.p;
        b.exact_limit = W->ws_tc_center ? W->w_exact_limit.p : nullptr;
        b.first = W->w_first.p;
        b.qperm = W->w_qperm.p;
        b.state = W->w_state.p;
        b.work_counter = W->w_counter.p;
        b.memo = W->w_memo_slots ? W->w_memo.p : nullptr;
        b.memo_stride = W->w_memo_stride;
        b.memo_slots = W->w_memo_slots;
        b.dense = nullptr;  // set by the single-pass search paths (use_dense_sims)
        b.dense_stride = W->w_dense_stride;
        b.pre_anchor = nullptr;
        b.pre_range = nullptr;
        b.pre_lcp = nullptr;
        b.fs_idx = nullptr;
        b.fs_hd = nullptr;
        b.fs_tab = nullptr;
        b.fs_meta = nullptr;
        b.fs_cap = W->w_fs_cap;
        b.first_is_own = false;
        b.shard_packed = nullptr;
        b.visit_log = W->ws_vlog ? W->w_visit_log.p : nullptr;
        b.visit_cap = W->ws_vlog;
        b.stats_dev = (d_ids && h_stats_dev) ? W->w_stats.p : nullptr;
        b.stats_host = h_stats_dev;
        b.out_ids = d_ids;
        b.out_dists = d_dists;
        b.out_counts = d_counts;
        b.cnt_candidates = W->w_cand.p;
        b.cnt_distcomp = W->w_dc.p;
        b.cnt_visited = W->w_vis.p;
        return b;
    }

    void require_built() const {
        if (!built) throw StatusError(CLANN_ERR_NOT_BUILT, "index has not been built");
    }
    // A cluster-sharded index holds tables for its own clusters only: the single-pass entry points would walk foreign clusters
    // through tables that were never built. Only the stepping / sharded calls may search it.
    void require_unsharded(const char* what) const {
        if (shard_count > 1)
            throw StatusError(CLANN_ERR_CONFIG, std::string(what) + " needs the whole index on this GPU; a cluster-sharded index "
                                                "(shard_count > 1) is searched with clann_search_begin/step/merge/end or clann_search_sharded");
    }
    void require_owned(uint64_t cluster) const {
        if (shard_count > 1 && h_owner[cluster] != shard_rank)
            throw StatusError(CLANN_ERR_CONFIG, "cluster is owned by another shard: its rows and tables are not on this rank");
    }

    // hashing of the query batch with every function set in use, centre ordering, state reset
    void search_begin(const float* d_queries, uint64_t nq, cudaStream_t s) {
        require_built();
        ensure_workspace(nq, s);
        W->ws_tc_center = use_tc_center();  // decided once per batch: every later view of this workspace must agree with it
        sh.last_was_sharded = false;
        SearchParams p = params();
        QueryBatch b = batch(d_queries, nq, nullptr, nullptr, nullptr);
        launch_prep_queries(p, b, s);
        // sketches are indexed by out_row = fset*nq + q (W->w_tiles); codes by fset*L*nq + t*nq + q (w_code_tiles)
        if (use_tc_sketch() && W->w_n_tc_tiles && nq >= 64) {
            launch_q15_slices(b.q15, nq, g.sl, W->w_qslices.p, s);
            launch_sketch_tc(W->w_tc_tiles.p, W->w_n_tc_tiles, W->w_qslices.p, 0, nq, d_plane_slices.p, n_fsets(), b.q15, d_planes.p, g.sl,
                             b.sketches, s);
        } else {
            launch_sketch(b.q15, W->w_tiles.p, W->w_ntiles, d_planes.p, g.sl, b.sketches, s);
        }
        launch_codes(b.q15, w_code_tiles(nq, s), W->w_ntiles, d_signbits.p, g, b.codes, nq, (uint64_t)g.L * nq, s);
        launch_center_order(p, b, s);
        // work order: stable sort of the queries by nearest cluster (same radix sort as the tables)
        launch_segment_sort(W->w_sort_seg.p, 1, (uint32_t)nq, b.first, b.qperm, W->w_sort_k.p, W->w_sort_i.p, s);
        launch_init_state(p, b, s);
        if (b.visit_log) CLANN_CUDA(cudaMemsetAsync(b.visit_log, 0, nq * (uint64_t)b.visit_cap * 16, s));
        cur_queries = d_queries;
        last_nq = nq;
        last_launches = (use_tc_sketch() && W->w_n_tc_tiles && nq >= 64) ? 8 : 7;
    }

    const RowTile* w_code_tiles(uint64_t nq, cudaStream_t s) {
        if (W->w_tiles_codes_nq != nq || !W->w_tiles_codes.p) {  // (ensure_workspace already wrote them for its nq)
            std::vector<RowTile> tiles;
            for (uint32_t f = 0; f < n_fsets(); f++)
                for (uint64_t q0 = 0; q0 < nq; q0 += 32)
                    tiles.push_back(RowTile{(uint32_t)q0, (uint32_t)q0, (uint32_t)std::min<uint64_t>(32, nq - q0), f, 0, 0, 0});
            W->w_tiles_codes.upload(tiles, s);
            CLANN_CUDA(cudaStreamSynchronize(s));
            W->w_tiles_codes_nq = nq;
        }
        return W->w_tiles_codes.p;
    }

    void search_device(const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, cudaStream_t s) {
        require_built();
        require_unsharded("clann_search / clann_search_device");
        if (nq == 0) return;
        CLANN_CUDA(cudaEventRecord(ev[0], s));
        search_begin(d_queries, nq, s);
        SearchParams p = params();
        QueryBatch b = batch(d_queries, nq, d_ids, d_dists, d_counts);
        CLANN_CUDA(cudaEventRecord(ev[1], s));
        const bool dense = use_dense_sims(p, b, s);
        launch_probe(p, b, false, s);
        CLANN_CUDA(cudaEventRecord(ev[2], s));
        launch_finish(p, b, s);
        CLANN_CUDA(cudaEventRecord(ev[3], s));
        last_launches = dense ? 9 + last_pre_launches : 9;
        profile_valid = true;
    }

    // The same search on one of the internal streams (rotating, pipeline_depth of them) with its own workspace set, not ordered after anything
    // the caller has in flight: consecutive batches overlap — the next batch's hashing runs beside the probe of the current
    // one and its probe fills the SMs the current probe's last wave leaves idle. The caller guarantees that the query buffer
    // is complete when the call is made; results are complete once search_flush() has been waited on.
    // Fills the memo of every query's first visit in advance (launch_dense_sims) when the default probe kernel will use it.
    bool use_dense_sims(const SearchParams& p, QueryBatch& b, cudaStream_t s) {
        if (!W->w_dense.p || W->w_dense_stride == 0) return false;
        if (h_stats && tune_get("dense_adaptive", 1) != 0) {
            const unsigned long long visited = reinterpret_cast<volatile unsigned long long*>(h_stats)[0];
            const unsigned long long queries = reinterpret_cast<volatile unsigned long long*>(h_stats)[1];
            if (queries > 0 && visited > 3 * queries) return false;  // the last finished batch walked > 3 clusters per query
        }
        b.dense = W->w_dense.p;
        if (!dense_sims_supported(p) || tune_get("order_longest_first", 0) != 0 || b.nq == 0 || p.max_cluster == 0) {
            b.dense = nullptr;
            return false;
        }
        // knob first_stream (default 0): the whole first visit's candidate stream ahead of the probe (launch_first_stream); it
        // includes the anchors, so first_ranges is not launched then. Measured on B200 (glove-100 shape, 10 000 queries,
        // profiles/r2d_*): the probe falls from 1.89 to 1.35 ms (DRAM reads 2.85 -> 0.81 GB) but k_first_stream itself takes
        // 1.22 ms (0.82 ms as one CTA per query), so the step is slower and the stream stays opt-in.
        if (W->w_fs_cap && tune_get("first_stream", 0) != 0) {
            launch_dense_sims(p, b, s);  // the stream's depth prediction reads the dense similarities
            b.fs_idx = W->w_fs_idx.p;
            b.fs_hd = W->w_fs_hd.p;
            b.fs_tab = W->w_fs_tab.p;
            b.fs_meta = W->w_fs_meta.p;
            if (launch_first_stream(p, b, s)) {
                last_pre_launches = 2;
                return true;
            }
            b.fs_idx = nullptr; b.fs_hd = nullptr; b.fs_tab = nullptr; b.fs_meta = nullptr;
            last_pre_launches = 1;
            return true;
        }
        // knob first_ranges: 0 off, 1 anchors + every depth's range, 2 anchors + samples only (default: same total time as 1 on the
        // glove-100 shape — 0.10 + 1.94 ms against 0.32 + 1.72 ms — with 13 MB instead of 94 MB of workspace per 10 000 queries)
        const int64_t fr = tune_get("first_ranges", 2);
        if (W->w_pre_range.p && fr != 0) {
            b.pre_anchor = W->w_pre_anchor.p;
            if (fr == 2) b.pre_lcp = W->w_pre_lcp.p;
            else b.pre_range = W->w_pre_range.p;
            // k_first_ranges (840 000 independent searches, 13 % SM throughput, pure latency) does not depend on the dense similarities
            // (ALU-bound): it runs beside them on a high-priority side stream (knob first_ranges_overlap)
            // Measured on B200 (glove-100 shape): the rerank kernels take 2.55 ms with the side stream against 2.48 ms without (the
            // join costs more than the 0.09 ms it hides; three batches in flight gain 1 %), so it is off by default.
            if (tune_get("first_ranges_overlap", 0) != 0) {
                if (!W->aux) {
                    int lo = 0, hi = 0;
                    CLANN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    CLANN_CUDA(cudaStreamCreateWithPriority(&W->aux, cudaStreamNonBlocking, hi));
                    CLANN_CUDA(cudaEventCreateWithFlags(&W->ev_fork, cudaEventDisableTiming));
                    CLANN_CUDA(cudaEventCreateWithFlags(&W->ev_join, cudaEventDisableTiming));
                }
                CLANN_CUDA(cudaEventRecord(W->ev_fork, s));          // query codes and work order are ready
                CLANN_CUDA(cudaStreamWaitEvent(W->aux, W->ev_fork, 0));
                launch_first_ranges(p, b, W->aux);
                CLANN_CUDA(cudaEventRecord(W->ev_join, W->aux));
                launch_dense_sims(p, b, s);
                CLANN_CUDA(cudaStreamWaitEvent(s, W->ev_join, 0));   // the probe needs both
            } else {
                launch_dense_sims(p, b, s);
                launch_first_ranges(p, b, s);
            }
            last_pre_launches = 2;
        } else {
            launch_dense_sims(p, b, s);
            last_pre_launches = 1;
        }
        return true;
    }

    // ---- cluster-sharded search: one process per GPU, every rank calls this with the same global batch ------------
    void all_gather(const void* send, void* recv, uint64_t bytes, cudaStream_t s) {
        if (user_allgather) {
            if (user_allgather(user_ctx, send, recv, bytes, s) != 0) throw StatusError(CLANN_ERR_SEARCH, "the caller's all-gather failed");
        } else if (comm) {
            CLANN_NCCL(nccl_api().AllGather(send, recv, bytes, ncclUint8, comm, s));
        } else throw StatusError(CLANN_ERR_CONFIG, "no transport: call clann_comm_init or clann_set_collectives first");
    }
    void all_reduce_min_u64(void* buf, uint64_t count, cudaStream_t s) {
        if (user_allreduce_min) {
            if (user_allreduce_min(user_ctx, buf, count, s) != 0) throw StatusError(CLANN_ERR_SEARCH, "the caller's all-reduce failed");
        } else if (comm) {
            CLANN_NCCL(nccl_api().AllReduce(buf, buf, count, ncclUint64, ncclMin, comm, s));
        } else throw StatusError(CLANN_ERR_CONFIG, "no transport: call clann_comm_init or clann_set_collectives first");
    }

    // Each rank owns the clusters assign_owners gave it and holds tables for those only. Per batch:
    //   route   every rank scores its 1/world slice of the queries against the (replicated) centres; one all-gather of the nearest
    //           cluster ids tells every rank which queries start in its own clusters
    //   round 1 a rank runs the reference's loop (index.rs:311-439) on those queries, unchanged, for as long as the visiting order
    //           stays inside its own clusters: prune test, radius early exit, max_sim, heap — first visits are 83 % of all visits
    //           on the planted shape and 4 of 5 queries end here
    //   bound   one all-reduce(min) of 8 bytes per query: finished, or the k-th distance reached and the clusters consumed
    //   round 2 every rank walks the order of the queries still open, prunes with the agreed bound (tightened by its own k-th
    //           distance once its heap is full) and visits its own clusters among those — a superset of the reference's visits
    //   merge   one all-gather of nq x k x (distance, id) and a k-way merge
    // No step moves more than a few megabytes; results have recall >= the single-GPU search (identical whenever the walk of a query
    // stays on one rank).
    // The call can split the batch into two halves ("lanes", knob shard_lanes) that run the phases above on their own streams,
    // interleaved by the host, so that one half's waits (count read-backs, collectives, the latency-bound round two) overlap the
    // other half's round one. Collectives are issued in the same order on every rank. Off by default (see search_sharded).
    void lane_prepare(ShardLane& ln, uint64_t nq) {
        const uint32_t world = shard_count, k = (uint32_t)cfg.k, d = g.d;
        const uint64_t chunk = (nq + world - 1) / world;
        if (ln.nq != nq) {
            ln.first_all.ensure(chunk * world);
            ln.list0.ensure(nq);
            ln.list1.ensure(nq);
            ln.list2.ensure(nq);
            ln.packed2.ensure(nq);
            ln.counts.ensure(4);
            ln.packed.ensure(nq);
            ln.packed1.ensure(nq);
            ln.top_local.ensure(nq * k);
            ln.top_all.ensure(nq * k * world);
            ln.counters.ensure(nq * 3);
            ln.q0.ensure(nq * d);
            ln.q1.ensure(nq * d);
            ln.nq = nq;
        }
        if (!ln.h_counts) CLANN_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ln.h_counts), 4 * sizeof(uint32_t), cudaHostAllocDefault));
        if (!ln.stream) CLANN_CUDA(cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking));
        if (!ln.ready) CLANN_CUDA(cudaEventCreateWithFlags(&ln.ready, cudaEventDisableTiming));
        if (!ln.done) CLANN_CUDA(cudaEventCreateWithFlags(&ln.done, cudaEventDisableTiming));
        for (auto& e : ln.ev)
            if (!e) CLANN_CUDA(cudaEventCreate(&e));
    }

    // route: nearest centre of this rank's slice of the sub-batch, all-gather, selection of the queries that start here
    void lane_route(ShardLane& ln, const float* d_queries) {
        cudaStream_t s = ln.stream;
        const uint32_t world = shard_count, rank = shard_rank, d = g.d;
        const uint64_t nq = ln.q_n, chunk = (nq + world - 1) / world;
        CLANN_CUDA(cudaEventRecord(ln.ev[0], s));
        CLANN_CUDA(cudaMemsetAsync(ln.counts.p, 0, 4 * sizeof(uint32_t), s));
        const uint64_t lo = std::min<uint64_t>(nq, rank * chunk), hi = std::min<uint64_t>(nq, lo + chunk);
        W = &ln.ws[0];
        if (hi > lo) {
            ensure_workspace(hi - lo, s);
            W->ws_tc_center = use_tc_center();
            SearchParams p = params();
            QueryBatch b = batch(d_queries + lo * d, hi - lo, nullptr, nullptr, nullptr);
            launch_prep_queries(p, b, s);
            launch_center_order(p, b, s);
            CLANN_CUDA(cudaMemcpyAsync(ln.first_all.p + lo, b.first, (hi - lo) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        }
        CLANN_CUDA(cudaEventRecord(ln.ev[1], s));
        // in-place all-gather: every rank's slice sits at rank * chunk of the same buffer
        all_gather(ln.first_all.p + rank * chunk, ln.first_all.p, chunk * sizeof(uint32_t), s);
        launch_shard_select_owned(ln.first_all.p, d_owner.p, rank, nq, ln.list0.p, ln.counts.p, s);
        CLANN_CUDA(cudaMemcpyAsync(ln.h_counts, ln.counts.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaEventRecord(ln.ready, s));
    }

    // round one on the queries routed here, then the bound exchange and the selection of the queries still open
    void lane_round_one(ShardLane& ln, const float* d_queries) {
        cudaStream_t s = ln.stream;
        const uint32_t k = (uint32_t)cfg.k, d = g.d;
        const uint64_t nq = ln.q_n;
        CLANN_CUDA(cudaEventSynchronize(ln.ready));
        const uint32_t n0 = ln.n0 = ln.h_counts[0];
        CLANN_CUDA(cudaEventRecord(ln.ev[2], s));
        launch_fill_u64(ln.packed.p, nq, 0xff800000ffffffffull, s);  // {+inf, nothing consumed}
        launch_fill_u64(ln.top_local.p, nq * k, ~0ull, s);
        CLANN_CUDA(cudaMemsetAsync(ln.counters.p, 0, nq * 3 * sizeof(unsigned long long), s));
        W = &ln.ws[1];
        if (n0) {
            launch_shard_gather_rows(d_queries, ln.list0.p, n0, d, ln.q0.p, s);
            search_begin(ln.q0.p, n0, s);
            SearchParams p = params();
            QueryBatch b = batch(ln.q0.p, n0, nullptr, nullptr, nullptr);
            b.first_is_own = true;  // by construction of list0
            use_dense_sims(p, b, s);
            launch_probe(p, b, 1, s);
            launch_shard_pack_bounds(W->w_state.p, k, ln.list0.p, n0, ln.packed.p, s);
            launch_shard_collect(W->w_state.p, k, ln.list0.p, n0, ln.top_local.p, false, ln.counters.p, s);
        }
        CLANN_CUDA(cudaEventRecord(ln.ev[3], s));
        all_reduce_min_u64(ln.packed.p, nq, s);
        launch_shard_select_open(ln.packed.p, nq, ln.list1.p, ln.packed1.p, ln.counts.p + 1, s);
        CLANN_CUDA(cudaMemcpyAsync(ln.h_counts + 1, ln.counts.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaEventRecord(ln.ready, s));
    }

    // every rank scores the open queries against the centres (cheap), but will hash and probe only those for which it owns a
    // cluster the agreed bound does not prune
    void lane_score_open(ShardLane& ln, const float* d_queries) {
        cudaStream_t s = ln.stream;
        const uint32_t rank = shard_rank, d = g.d;
        CLANN_CUDA(cudaEventSynchronize(ln.ready));
        const uint32_t n1 = ln.n1 = ln.h_counts[1];
        CLANN_CUDA(cudaEventRecord(ln.ev[4], s));
        ln.h_counts[2] = 0;
        if (n1) {
            W = &ln.ws[0];
            launch_shard_gather_rows(d_queries, ln.list1.p, n1, d, ln.q1.p, s);
            ensure_workspace(n1, s);
            W->ws_tc_center = use_tc_center();
            SearchParams p = params();
            QueryBatch b = batch(ln.q1.p, n1, nullptr, nullptr, nullptr);
            launch_prep_queries(p, b, s);
            launch_center_order(p, b, s);
            launch_shard_select_mine(b.cdist, b.exact_limit, d_radii.p, d_owner.p, rank, K, ln.list1.p, ln.packed1.p, n1, ln.list2.p,
                                     ln.packed2.p, ln.counts.p + 2, s);
            CLANN_CUDA(cudaMemcpyAsync(ln.h_counts + 2, ln.counts.p + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        }
        CLANN_CUDA(cudaEventRecord(ln.ready, s));
    }

    // round two, the all-gather of the candidate lists and the k-way merge into the caller's output rows of this sub-batch
    void lane_round_two(ShardLane& ln, const float* d_queries, uint32_t* d_ids, float* d_dists, uint32_t* d_counts) {
        cudaStream_t s = ln.stream;
        const uint32_t world = shard_count, k = (uint32_t)cfg.k, d = g.d;
        const uint64_t nq = ln.q_n;
        CLANN_CUDA(cudaEventSynchronize(ln.ready));
        const uint32_t n2 = ln.n2 = ln.n1 ? ln.h_counts[2] : 0;
        W = &ln.ws[2];
        if (n2) {
            launch_shard_gather_rows(d_queries, ln.list2.p, n2, d, ln.q1.p, s);
            search_begin(ln.q1.p, n2, s);
            SearchParams p = params();
            QueryBatch b = batch(ln.q1.p, n2, nullptr, nullptr, nullptr);
            b.shard_packed = ln.packed2.p;
            launch_probe(p, b, 2, s);
            launch_shard_collect(W->w_state.p, k, ln.list2.p, n2, ln.top_local.p, true, ln.counters.p, s);
        }
        CLANN_CUDA(cudaEventRecord(ln.ev[5], s));
        all_gather(ln.top_local.p, ln.top_all.p, nq * k * sizeof(unsigned long long), s);
        launch_shard_final_merge(ln.top_all.p, world, nq, k, d_ids, d_dists, d_counts, s);
        CLANN_CUDA(cudaEventRecord(ln.ev[6], s));
        CLANN_CUDA(cudaEventRecord(ln.done, s));
    }

    // Two WHOLE batches in flight (clann_search_sharded_pair): the same interleaving with one lane per batch — the next batch's
    // round one fills the SMs while the current batch sits in its count read-backs, collectives and latency-bound round two. This is
    // the batch pipelining of clann_search_device_async for the sharded search.
    void search_sharded_multi(int nb, const float* const* d_queries, uint64_t nq, uint32_t* const* d_ids, float* const* d_dists,
                              uint32_t* const* d_counts, cudaStream_t s) {
        require_built();
        if (shard_count < 2) throw StatusError(CLANN_ERR_CONFIG, "clann_search_sharded_multi needs an index built with shard_count > 1");
        if (nb < 1 || nb > kShardLanes) throw StatusError(CLANN_ERR_ARG, "1 to 4 batches in flight");
        require_no_stream_in_flight();
        if (nq == 0) return;
        if (nq >= (1ull << 32)) throw StatusError(CLANN_ERR_ARG, "batch too large");
        SearchWs* saved = W;
        try {
            if (!sh.fork) CLANN_CUDA(cudaEventCreateWithFlags(&sh.fork, cudaEventDisableTiming));
            CLANN_CUDA(cudaEventRecord(sh.fork, s));
            for (int i = 0; i < nb; i++) {
                ShardLane& ln = lanes[i];
                ln.q_lo = (uint64_t)i * nq;  // position in the concatenated counters
                ln.q_n = nq;
                lane_prepare(ln, nq);
                CLANN_CUDA(cudaStreamWaitEvent(ln.stream, sh.fork, 0));
            }
            for (int i = 0; i < nb; i++) lane_route(lanes[i], d_queries[i]);
            for (int i = 0; i < nb; i++) lane_round_one(lanes[i], d_queries[i]);
            for (int i = 0; i < nb; i++) lane_score_open(lanes[i], d_queries[i]);
            for (int i = 0; i < nb; i++) lane_round_two(lanes[i], d_queries[i], d_ids[i], d_dists[i], d_counts[i]);
            for (int i = 0; i < nb; i++) CLANN_CUDA(cudaStreamWaitEvent(s, lanes[i].done, 0));
            sh.lanes_used = nb;
            sh.last_was_sharded = true;
            last_nq = (uint64_t)nb * nq;
            last_launches = 0;
        } catch (...) {
            W = saved;
            throw;
        }
        W = saved;
    }

    // Streaming form (clann_search_sharded_submit / _flush): a software pipeline over consecutive calls. Every call ("tick") takes
    // one new batch and moves each of the up to four batches in flight forward by ONE phase, newest first:
    //     tick t:  route(t)  round one(t-1)  scoring of the open queries(t-2)  round two + merge(t-3)
    // so that (i) the count a phase reads back was produced a whole tick earlier — the host no longer waits inside a batch, only for
    // the previous tick's round one, after the next round one is already queued —, (ii) a collective that waits for a slower rank
    // stalls only its own lane's stream, and (iii) the latency-bound round two of one batch runs in the shadow of the round one of a
    // later batch instead of next to the round two of its twin (search_sharded_multi keeps its batches in the same phase). The
    // collectives are issued in the same order on every rank as long as every rank makes the same sequence of calls. Results of a
    // batch are complete (in stream order on the stream of the call that finishes it) after three more submissions or a flush.
    void require_no_stream_in_flight() const {
        for (const auto& ln : lanes)
            if (ln.phase >= 0) throw StatusError(CLANN_ERR_SEARCH, "batches submitted with clann_search_sharded_submit are in flight: call clann_search_sharded_flush first");
    }

    void sharded_tick(cudaStream_t s) {
        for (int age = 0; age < kShardLanes; age++) {
            if ((uint64_t)age > sh.tick) break;
            ShardLane& ln = lanes[(sh.tick - age) % kShardLanes];
            switch (ln.phase) {
                case 0: lane_route(ln, ln.sq); ln.phase = 1; break;
                case 1: lane_round_one(ln, ln.sq); ln.phase = 2; break;
                case 2: lane_score_open(ln, ln.sq); ln.phase = 3; break;
                case 3:
                    lane_round_two(ln, ln.sq, ln.s_ids, ln.s_dists, ln.s_counts);
                    CLANN_CUDA(cudaStreamWaitEvent(s, ln.done, 0));  // the finishing call's stream sees the results
                    ln.phase = -1;
                    break;
                default: break;
            }
        }
        sh.tick++;
    }

    void sharded_submit(const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, cudaStream_t s) {
        require_built();
        if (shard_count < 2) throw StatusError(CLANN_ERR_CONFIG, "clann_search_sharded_submit needs an index built with shard_count > 1");
        if (nq == 0 || nq >= (1ull << 32)) throw StatusError(CLANN_ERR_ARG, "batch of 1 .. 2^32 - 1 queries");
        SearchWs* saved = W;
        try {
            ShardLane& ln = lanes[sh.tick % kShardLanes];
            if (ln.phase >= 0) throw StatusError(CLANN_ERR_SEARCH, "sharded pipeline out of step");  // cannot happen: a batch lives four ticks
            lane_prepare(ln, nq);
            if (!ln.fork) CLANN_CUDA(cudaEventCreateWithFlags(&ln.fork, cudaEventDisableTiming));
            CLANN_CUDA(cudaEventRecord(ln.fork, s));
            CLANN_CUDA(cudaStreamWaitEvent(ln.stream, ln.fork, 0));  // the caller's queries are ready
            ln.q_lo = 0;
            ln.q_n = nq;
            ln.sq = d_queries;
            ln.s_ids = d_ids;
            ln.s_dists = d_dists;
            ln.s_counts = d_counts;
            ln.phase = 0;
            sharded_tick(s);
            sh.lanes_used = 0;  // per-batch counters / phase times are those of the blocking calls
            sh.last_was_sharded = true;
            last_nq = 0;
        } catch (...) {
            W = saved;
            throw;
        }
        W = saved;
    }

    void sharded_flush(cudaStream_t s) {
        SearchWs* saved = W;
        try {
            for (int i = 0; i < kShardLanes - 1; i++) {
                bool any = false;
                for (auto& ln : lanes) any = any || ln.phase >= 0;
                if (!any) break;
                sharded_tick(s);  // a tick without a new batch (its lane stays idle)
            }
        } catch (...) {
            W = saved;
            throw;
        }
        W = saved;
    }

    void search_sharded(const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, cudaStream_t s) {
        require_built();
        if (shard_count < 2) throw StatusError(CLANN_ERR_CONFIG, "clann_search_sharded needs an index built with shard_count > 1");
        require_no_stream_in_flight();
        if (nq == 0) return;
        if (nq >= (1ull << 32)) throw StatusError(CLANN_ERR_ARG, "batch too large");
        const uint32_t k = (uint32_t)cfg.k, d = g.d;
        // knob shard_lanes (default 1): sub-batches in flight. Measured on 2 B200 (glove-100 shape, 20 000 queries per step): two
        // lanes 4.58 ms vs one 4.24 ms — the persistent probe kernel of one half fills every SM, so the halves serialise anyway and
        // each pays its own last wave; small batches are never split
        int nl = (int)tune_get("shard_lanes", 1);
        nl = nl < 1 ? 1 : (nl > kShardLanes ? kShardLanes : nl);
        if (nq < 2048ull * shard_count) nl = 1;
        SearchWs* saved = W;
        try {
            if (!sh.fork) CLANN_CUDA(cudaEventCreateWithFlags(&sh.fork, cudaEventDisableTiming));
            CLANN_CUDA(cudaEventRecord(sh.fork, s));
            const uint64_t per = (nq + nl - 1) / nl;
            for (int i = 0; i < nl; i++) {
                ShardLane& ln = lanes[i];
                ln.q_lo = std::min<uint64_t>(nq, (uint64_t)i * per);
                ln.q_n = std::min<uint64_t>(per, nq - ln.q_lo);
                lane_prepare(ln, per);
                CLANN_CUDA(cudaStreamWaitEvent(ln.stream, sh.fork, 0));  // the caller's queries are ready
            }
            for (int i = 0; i < nl; i++) lane_route(lanes[i], d_queries + lanes[i].q_lo * d);
            for (int i = 0; i < nl; i++) lane_round_one(lanes[i], d_queries + lanes[i].q_lo * d);
            for (int i = 0; i < nl; i++) lane_score_open(lanes[i], d_queries + lanes[i].q_lo * d);
            for (int i = 0; i < nl; i++)
                lane_round_two(lanes[i], d_queries + lanes[i].q_lo * d, d_ids + lanes[i].q_lo * k, d_dists + lanes[i].q_lo * k,
                               d_counts + lanes[i].q_lo);
            for (int i = 0; i < nl; i++) CLANN_CUDA(cudaStreamWaitEvent(s, lanes[i].done, 0));  // the caller's stream sees the results
            sh.lanes_used = nl;
            sh.last_was_sharded = true;
            last_nq = nq;
            last_launches = 0;
        } catch (...) {
            W = saved;
            throw;
        }
        W = saved;
    }

    int next_pipe_slot() {
        int depth = (int)tune_get("pipeline_depth", 3);  // batches in flight (knob; 3 measured 1.5 % above 2, 4 slower)
        depth = depth < 1 ? 1 : (depth > kPipeMax ? kPipeMax : depth);
        const int slot = (int)(pipe_calls++ % (uint64_t)depth);
        if (!pipe_stream[slot]) {
            CLANN_CUDA(cudaStreamCreateWithFlags(&pipe_stream[slot], cudaStreamNonBlocking));
            CLANN_CUDA(cudaEventCreateWithFlags(&pipe_done[slot], cudaEventDisableTiming));
        }
        return slot;
    }

    void search_on_slot(int slot, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts) {
        cudaStream_t s = pipe_stream[slot];
        SearchWs* saved = W;
        W = &wsv[1 + slot];
        try {
            search_begin(d_queries, nq, s);
            SearchParams p = params();
            QueryBatch b = batch(d_queries, nq, d_ids, d_dists, d_counts);
            const bool dense = use_dense_sims(p, b, s);
            launch_probe(p, b, false, s);
            launch_finish(p, b, s);
            last_launches = dense ? 9 + last_pre_launches : 9;
        } catch (...) {
            W = saved;
            throw;
        }
        W = saved;
    }

    void search_device_async(const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts) {
        require_built();
        require_unsharded("clann_search_device_async");
        if (nq == 0) return;
        const int slot = next_pipe_slot();
        search_on_slot(slot, d_queries, nq, d_ids, d_dists, d_counts);
        CLANN_CUDA(cudaEventRecord(pipe_done[slot], pipe_stream[slot]));
    }

    // Host buffers (pinned, for the copies to be asynchronous): H2D, search and D2H of one batch on the slot's stream.
    void search_host_async(const float* queries, uint64_t nq, uint32_t* ids, float* dists, uint32_t* counts) {
        require_built();
        require_unsharded("clann_search_async");
        if (nq == 0) return;
        const int slot = next_pipe_slot();
        cudaStream_t s = pipe_stream[slot];
        SearchWs& w = wsv[1 + slot];
        const uint32_t k = (uint32_t)cfg.k;
        w.h_queries.ensure(nq * g.d);
        w.h_ids.ensure(nq * k);
        w.h_dists.ensure(nq * k);
        w.h_counts.ensure(nq);
        CLANN_CUDA(cudaMemcpyAsync(w.h_queries.p, queries, nq * g.d * sizeof(float), cudaMemcpyHostToDevice, s));
        search_on_slot(slot, w.h_queries.p, nq, w.h_ids.p, w.h_dists.p, w.h_counts.p);
        CLANN_CUDA(cudaMemcpyAsync(ids, w.h_ids.p, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaMemcpyAsync(dists, w.h_dists.p, nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaMemcpyAsync(counts, w.h_counts.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaEventRecord(pipe_done[slot], s));
    }

    void search_wait() {
        for (auto& st : pipe_stream)
            if (st) CLANN_CUDA(cudaStreamSynchronize(st));
        CLANN_CUDA(cudaGetLastError());
    }

    void search_flush(cudaStream_t s) {
        for (auto& e : pipe_done)
            if (e) CLANN_CUDA(cudaStreamWaitEvent(s, e, 0));
    }
};

// ------------------------------------------------------------------------------------------------ C ABI (batched)

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return CLANN_OK;
    } catch (const StatusError& e) {
        g_last_error = e.what();
        return e.code;
    } catch (const CudaError& e) {
        g_last_error = e.what();
        return CLANN_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return CLANN_ERR_CREATION;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return CLANN_ERR_ARG;
    } catch (...) {
        g_last_error = "unknown error";
        return CLANN_ERR_ARG;
    }
}

extern "C" {

const char* clann_last_error(void) { return g_last_error.c_str(); }

int clann_init_with_config(const float* data, uint64_t n, uint32_t d, const clann_config* config, clann_index** out) {
    if (out) *out = nullptr;
    return guarded([&] {
        if (!config || !out) throw StatusError(CLANN_ERR_ARG, "null config or output pointer");
        auto* ix = new clann_index();
        try {
            ix->init(data, n, d, *config);
        } catch (...) {
            delete ix;
            throw;
        }
        *out = ix;
    });
}

int clann_init_with_config_ex(const void* data, uint64_t n, uint32_t d, const clann_config* config, int dtype, int on_device,
                              clann_index** out) {
    if (out) *out = nullptr;
    return guarded([&] {
        if (!config || !out) throw StatusError(CLANN_ERR_ARG, "null config or output pointer");
        auto* ix = new clann_index();
        try {
            ix->init(data, n, d, *config, dtype, on_device != 0);
        } catch (...) {
            delete ix;
            throw;
        }
        *out = ix;
    });
}

int clann_set_option(clann_index* index, const char* key, int64_t value) {
    return guarded([&] {
        if (!index || !key) throw StatusError(CLANN_ERR_ARG, "null index or key");
        std::string k(key);
        if (k == "seed") index->seed = (uint64_t)value;
        else if (k == "function_sets") index->per_cluster_functions = value != 0;
        else if (k == "strict") { if (value != 1) throw StatusError(CLANN_ERR_ARG, "only strict mode is implemented"); }
        else if (k == "shard_rank") {
            if (value < 0 || value > 254) throw StatusError(CLANN_ERR_ARG, "shard_rank must be in 0..254");
            index->shard_rank = (uint32_t)value;
        } else if (k == "shard_count") {
            if (value < 1 || value > 255) throw StatusError(CLANN_ERR_ARG, "shard_count must be in 1..255");
            index->shard_count = (uint32_t)value;
        } else if (k == "visit_log") {
            // rows of the per-visit log kept per query (0 = off): the cluster granularity of the reference's metrics. Not part of the index.
            if (value < 0 || value > 65535) throw StatusError(CLANN_ERR_ARG, "visit_log must be in 0..65535");
            index->visit_log_cap = (uint32_t)value;
            return;
        } else throw StatusError(CLANN_ERR_ARG, "unknown option '" + k + "'");
        index->built = false;
    });
}

int clann_set_delta(clann_index* index, float delta) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        if (!(delta > 0.0f && delta < 1.0f)) throw StatusError(CLANN_ERR_CONFIG, "delta must be in (0, 1)");
        index->cfg.delta = delta;
        if (index->built) index->ensure_stop_table(delta, 0);  // the tables and functions do not depend on delta (collection.hpp:927-943)
    });
}

int clann_tune(const char* key, int64_t value) {
    return guarded([&] {
        if (!key) throw StatusError(CLANN_ERR_ARG, "null key");
        clann::tune_set(key, value);
    });
}

int clann_set_clustering(clann_index* index, uint64_t K, const uint64_t* centers, const uint64_t* assignment, const float* radii) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->set_clustering(K, centers, assignment, radii);
    });
}

int clann_import_reference(clann_index* index, uint64_t cluster, const void* blob, uint64_t len) {
    return guarded([&] {
        if (!index || !blob) throw StatusError(CLANN_ERR_ARG, "null index or blob");
        index->import_reference(cluster, blob, len);
    });
}

int clann_set_functions(clann_index* index, uint64_t cluster, const int16_t* planes, const int8_t* signs, const float* est) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->set_functions(cluster, planes, signs, est);
    });
}

int clann_build(clann_index* index) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        try {
            index->build();
        } catch (const CudaError& e) {
            throw StatusError(CLANN_ERR_CREATION, e.what());  // index.rs:267-273
        }
    });
}

int clann_search_device(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts,
                        void* stream) {
    return guarded([&] {
        if (!index || (nq && (!d_queries || !d_ids || !d_dists || !d_counts))) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_device(d_queries, nq, d_ids, d_dists, d_counts, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_device_async(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts) {
    return guarded([&] {
        if (!index || (nq && (!d_queries || !d_ids || !d_dists || !d_counts))) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_device_async(d_queries, nq, d_ids, d_dists, d_counts);
    });
}

int clann_search_async(clann_index* index, const float* queries, uint64_t nq, uint32_t* ids, float* dists, uint32_t* counts) {
    return guarded([&] {
        if (!index || (nq && (!queries || !ids || !dists || !counts))) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_host_async(queries, nq, ids, dists, counts);
    });
}

int clann_search_wait(clann_index* index) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->search_wait();
    });
}

int clann_search_flush(clann_index* index, void* stream) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->search_flush(static_cast<cudaStream_t>(stream));
    });
}

int clann_search(clann_index* index, const float* queries, uint64_t nq, uint32_t* ids, float* dists, uint32_t* counts) {
    return guarded([&] {
        if (!index || (nq && (!queries || !ids || !dists || !counts))) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->require_built();
        if (nq == 0) return;
        const uint32_t k = (uint32_t)index->cfg.k;
        cudaStream_t s = 0;
        index->w_queries.ensure(nq * index->g.d);
        index->w_out_ids.ensure(nq * k);
        index->w_out_dists.ensure(nq * k);
        index->w_out_counts.ensure(nq);
        CLANN_CUDA(cudaMemcpyAsync(index->w_queries.p, queries, nq * index->g.d * sizeof(float), cudaMemcpyHostToDevice, s));
        index->search_device(index->w_queries.p, nq, index->w_out_ids.p, index->w_out_dists.p, index->w_out_counts.p, s);
        CLANN_CUDA(cudaMemcpyAsync(ids, index->w_out_ids.p, nq * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaMemcpyAsync(dists, index->w_out_dists.p, nq * k * sizeof(float), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaMemcpyAsync(counts, index->w_out_counts.p, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CLANN_CUDA(cudaStreamSynchronize(s));
        CLANN_CUDA(cudaGetLastError());
    });
}

int clann_search_begin(clann_index* index, const float* d_queries, uint64_t nq, void* stream) {
    return guarded([&] {
        if (!index || !d_queries) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_begin(d_queries, nq, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_step(clann_index* index, void* stream) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->require_built();
        SearchParams p = index->params();
        QueryBatch b = index->batch(index->cur_queries, index->last_nq, nullptr, nullptr, nullptr);
        launch_probe(p, b, index->shard_count > 1, static_cast<cudaStream_t>(stream));
        index->last_launches++;
    });
}

uint64_t clann_state_bytes(const clann_index* index) { return index ? query_state_bytes((uint32_t)index->cfg.k) : 0; }
void* clann_state_ptr(clann_index* index) { return index ? index->W->w_state.p : nullptr; }

int clann_search_merge(clann_index* index, const void* d_all_states, int world, uint64_t* active_out, void* stream) {
    return guarded([&] {
        if (!index || !d_all_states || world < 1) throw StatusError(CLANN_ERR_ARG, "bad merge arguments");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        SearchParams p = index->params();
        QueryBatch b = index->batch(index->cur_queries, index->last_nq, nullptr, nullptr, nullptr);
        launch_merge_states(p, b, static_cast<const uint8_t*>(d_all_states), world, index->W->w_counter.p + 1, s);
        index->last_launches++;
        if (active_out) {
            uint32_t a = 0;
            CLANN_CUDA(cudaMemcpyAsync(&a, index->W->w_counter.p + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            CLANN_CUDA(cudaStreamSynchronize(s));
            *active_out = a;
        }
    });
}

int clann_search_end(clann_index* index, uint32_t* d_ids, float* d_dists, uint32_t* d_counts, void* stream) {
    return guarded([&] {
        if (!index || !d_ids || !d_dists || !d_counts) throw StatusError(CLANN_ERR_ARG, "null pointer");
        SearchParams p = index->params();
        QueryBatch b = index->batch(index->cur_queries, index->last_nq, d_ids, d_dists, d_counts);
        launch_finish(p, b, static_cast<cudaStream_t>(stream));
        index->last_launches++;
    });
}

int clann_comm_unique_id(uint8_t* out, uint64_t cap) {
    return guarded([&] {
        if (!out || cap < sizeof(ncclUniqueId)) throw StatusError(CLANN_ERR_ARG, "unique id buffer must hold 128 bytes");
        ncclUniqueId id;
        CLANN_NCCL(nccl_api().GetUniqueId(&id));
        memcpy(out, &id, sizeof(id));
    });
}

int clann_comm_init(clann_index* index, int rank, int world, const uint8_t* unique_id) {
    return guarded([&] {
        if (!index || !unique_id || world < 1 || rank < 0 || rank >= world) throw StatusError(CLANN_ERR_ARG, "bad communicator arguments");
        if ((uint32_t)world != index->shard_count || (uint32_t)rank != index->shard_rank)
            throw StatusError(CLANN_ERR_CONFIG, "rank / world must equal the index's shard_rank / shard_count options");
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        if (index->comm) {
            nccl_api().CommDestroy(index->comm);
            index->comm = nullptr;
        }
        CLANN_NCCL(nccl_api().CommInitRank(&index->comm, world, id, rank));
        // first collective now: NCCL sets up its channels lazily (~1 s), which must not land inside a build or a search
        DevBuf<unsigned long long> warm;
        warm.alloc(1);
        warm.zero(0);
        CLANN_NCCL(nccl_api().AllReduce(warm.p, warm.p, 1, ncclUint64, ncclMax, index->comm, 0));
        CLANN_CUDA(cudaStreamSynchronize(0));
    });
}

int clann_set_collectives(clann_index* index, clann_allgather_fn allgather, clann_allreduce_min_u64_fn allreduce_min, void* ctx) {
    return guarded([&] {
        if (!index || !allgather || !allreduce_min) throw StatusError(CLANN_ERR_ARG, "null index or collective");
        index->user_allgather = allgather;
        index->user_allreduce_min = allreduce_min;
        index->user_ctx = ctx;
    });
}

int clann_search_sharded(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists, uint32_t* d_counts,
                         void* stream) {
    return guarded([&] {
        if (!index || (nq && (!d_queries || !d_ids || !d_dists || !d_counts))) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_sharded(d_queries, nq, d_ids, d_dists, d_counts, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_sharded_pair(clann_index* index, const float* d_queries_a, const float* d_queries_b, uint64_t nq, uint32_t* d_ids_a,
                              float* d_dists_a, uint32_t* d_counts_a, uint32_t* d_ids_b, float* d_dists_b, uint32_t* d_counts_b, void* stream) {
    return guarded([&] {
        if (!index || (nq && (!d_queries_a || !d_queries_b || !d_ids_a || !d_dists_a || !d_counts_a || !d_ids_b || !d_dists_b || !d_counts_b)))
            throw StatusError(CLANN_ERR_ARG, "null pointer");
        const float* q[2] = {d_queries_a, d_queries_b};
        uint32_t* ids[2] = {d_ids_a, d_ids_b};
        float* dd[2] = {d_dists_a, d_dists_b};
        uint32_t* cc[2] = {d_counts_a, d_counts_b};
        index->search_sharded_multi(2, q, nq, ids, dd, cc, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_sharded_multi(clann_index* index, int n_batches, const float* const* d_queries, uint64_t nq, uint32_t* const* d_ids,
                               float* const* d_dists, uint32_t* const* d_counts, void* stream) {
    return guarded([&] {
        if (!index || !d_queries || !d_ids || !d_dists || !d_counts) throw StatusError(CLANN_ERR_ARG, "null pointer");
        for (int i = 0; i < n_batches && i < 4; i++)
            if (nq && (!d_queries[i] || !d_ids[i] || !d_dists[i] || !d_counts[i])) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->search_sharded_multi(n_batches, d_queries, nq, d_ids, d_dists, d_counts, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_sharded_submit(clann_index* index, const float* d_queries, uint64_t nq, uint32_t* d_ids, float* d_dists,
                                uint32_t* d_counts, void* stream) {
    return guarded([&] {
        if (!index || !d_queries || !d_ids || !d_dists || !d_counts) throw StatusError(CLANN_ERR_ARG, "null pointer");
        index->sharded_submit(d_queries, nq, d_ids, d_dists, d_counts, static_cast<cudaStream_t>(stream));
    });
}

int clann_search_sharded_flush(clann_index* index, void* stream) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        index->sharded_flush(static_cast<cudaStream_t>(stream));
    });
}

int clann_shard_stats(clann_index* index, uint64_t* routed_round_one, uint64_t* open_round_two, float* phase_ms) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        uint64_t n0 = 0, n1 = 0;
        for (int l = 0; l < index->sh.lanes_used; l++) {
            n0 += index->lanes[l].n0;
            n1 += index->lanes[l].n1;
        }
        if (routed_round_one) *routed_round_one = n0;
        if (open_round_two) *open_round_two = n1;
        if (phase_ms) {  // device time of the phases, summed over the sub-batches in flight (they overlap on the device)
            if (index->sh.lanes_used == 0) throw StatusError(CLANN_ERR_ARG, "no sharded search yet");
            for (int i = 0; i < 6; i++) phase_ms[i] = 0.0f;
            for (int l = 0; l < index->sh.lanes_used; l++) {
                CLANN_CUDA(cudaEventSynchronize(index->lanes[l].ev[6]));
                for (int i = 0; i < 6; i++) {
                    float ms = 0.0f;
                    CLANN_CUDA(cudaEventElapsedTime(&ms, index->lanes[l].ev[i], index->lanes[l].ev[i + 1]));
                    phase_ms[i] += ms;
                }
            }
        }
    });
}

int clann_get_counters(clann_index* index, uint64_t nq, uint64_t* candidates, uint64_t* distance_computations, uint32_t* clusters_visited) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        if (nq > index->last_nq) throw StatusError(CLANN_ERR_BOUNDS, "more counters requested than queries searched");
        CLANN_CUDA(cudaDeviceSynchronize());
        if (index->sh.last_was_sharded) {
            // this rank's share of every query's counters (sum them over the ranks for the totals)
            for (int l = 0; l < index->sh.lanes_used; l++) {
                const auto& ln = index->lanes[l];
                if (ln.q_lo >= nq) break;
                const uint64_t m = std::min<uint64_t>(ln.q_n, nq - ln.q_lo);
                std::vector<unsigned long long> c = ln.counters.download(m * 3);
                for (uint64_t q = 0; q < m; q++) {
                    if (candidates) candidates[ln.q_lo + q] = c[q * 3];
                    if (distance_computations) distance_computations[ln.q_lo + q] = c[q * 3 + 1];
                    if (clusters_visited) clusters_visited[ln.q_lo + q] = (uint32_t)c[q * 3 + 2];
                }
            }
            return;
        }
        if (candidates) CLANN_CUDA(cudaMemcpy(candidates, index->W->w_cand.p, nq * 8, cudaMemcpyDeviceToHost));
        if (distance_computations) CLANN_CUDA(cudaMemcpy(distance_computations, index->W->w_dc.p, nq * 8, cudaMemcpyDeviceToHost));
        if (clusters_visited) CLANN_CUDA(cudaMemcpy(clusters_visited, index->W->w_vis.p, nq * 4, cudaMemcpyDeviceToHost));
    });
}

int clann_last_search_profile(clann_index* index, float* ms, uint32_t* launches) {
    return guarded([&] {
        if (!index || !index->profile_valid) throw StatusError(CLANN_ERR_ARG, "no profiled search yet");
        if (ms) {
            CLANN_CUDA(cudaEventSynchronize(index->ev[3]));
            CLANN_CUDA(cudaEventElapsedTime(&ms[0], index->ev[0], index->ev[1]));
            CLANN_CUDA(cudaEventElapsedTime(&ms[1], index->ev[1], index->ev[2]));
            CLANN_CUDA(cudaEventElapsedTime(&ms[2], index->ev[2], index->ev[3]));
        }
        if (launches) *launches = index->last_launches;
    });
}

int clann_export(clann_index* index, int what, uint64_t arg, void* dst, uint64_t cap, uint64_t* size) {
    return guarded([&] {
        if (!index) throw StatusError(CLANN_ERR_ARG, "null index");
        CLANN_CUDA(cudaDeviceSynchronize());
        auto emit = [&](const void* src, uint64_t bytes) {
            if (size) *size = bytes;
            if (dst) {
                if (cap < bytes) throw StatusError(CLANN_ERR_BOUNDS, "export buffer too small");
                memcpy(dst, src, bytes);
            }
        };
        auto emit_dev = [&](const void* src, uint64_t bytes) {
            if (size) *size = bytes;
            if (dst) {
                if (cap < bytes) throw StatusError(CLANN_ERR_BOUNDS, "export buffer too small");
                CLANN_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
            }
        };
        const uint32_t K = index->K;
        auto need_cluster = [&]() {
            index->require_built();
            if (arg >= K) throw StatusError(CLANN_ERR_BOUNDS, "cluster out of range");
        };
        auto need_owned_cluster = [&]() {  // per-cluster arrays that only the owning shard builds
            need_cluster();
            index->require_owned(arg);
        };
        switch (what) {
            case CLANN_X_NUM_CLUSTERS: {
                uint64_t k64 = K;
                emit(&k64, 8);
                break;
            }
            case CLANN_X_CENTERS: {
                index->require_built();
                std::vector<uint64_t> v(index->h_centers.begin(), index->h_centers.end());
                emit(v.data(), v.size() * 8);
                break;
            }
            case CLANN_X_ASSIGNMENT: {
                index->require_built();
                std::vector<uint64_t> v(index->h_assign.begin(), index->h_assign.end());
                emit(v.data(), v.size() * 8);
                break;
            }
            case CLANN_X_RADII: index->require_built(); emit(index->h_radii.data(), (uint64_t)K * 4); break;
            case CLANN_X_OFFSETS: index->require_built(); emit(index->h_offsets.data(), (uint64_t)(K + 1) * 8); break;
            case CLANN_X_PERM: index->require_built(); emit_dev(index->d_perm.p, index->n_rows * 4); break;
            case CLANN_X_Q15:
                need_cluster();
                emit_dev(index->d_q15.p + index->h_offsets[arg] * index->g.sl, (uint64_t)index->h_sizes[arg] * index->g.sl * 2);
                break;
            case CLANN_X_SKETCHES:
                need_owned_cluster();
                emit_dev(index->d_sketches.p + index->h_offsets[arg] * kNumSketches, (uint64_t)index->h_sizes[arg] * kNumSketches * 8);
                break;
            case CLANN_X_TABLE_HASHES:
            case CLANN_X_TABLE_INDICES: {
                need_owned_cluster();
                const uint32_t nc = index->h_sizes[arg], L = index->g.L;
                const uint64_t bytes = (uint64_t)L * nc * 4;
                if (size) *size = bytes;
                if (dst) {
                    if (cap < bytes) throw StatusError(CLANN_ERR_BOUNDS, "export buffer too small");
                    const uint32_t* src = what == CLANN_X_TABLE_HASHES ? index->d_tbl_hash.p : index->d_tbl_idx.p;
                    // cluster-major layout: the L tables of the cluster are adjacent
                    CLANN_CUDA(cudaMemcpy(dst, src + table_base(index->h_offsets[arg], nc, L, 0), bytes, cudaMemcpyDeviceToHost));
                }
                break;
            }
            case CLANN_X_BRUTE: index->require_built(); emit(index->h_brute.data(), K); break;
            case CLANN_X_NORMS: index->require_built(); emit_dev(index->d_norms.p, index->n * 4); break;
            case CLANN_X_EST: {
                need_cluster();
                const FunctionSet& fs = index->fsets[index->h_fset_of[arg]];
                emit(fs.est.data(), fs.est.size() * 4);
                break;
            }
            case CLANN_X_QUERY_CODES: {
                need_cluster();
                const uint64_t nq = index->last_nq, L = index->g.L;
                const uint32_t f = index->h_fset_of[arg];
                // device layout [L][nq] -> host layout [nq][L]
                std::vector<uint32_t> tmp = index->W->w_codes.download(L * nq, (size_t)f * L * nq);
                std::vector<uint32_t> outv(nq * L);
                for (uint64_t t = 0; t < L; t++)
                    for (uint64_t q = 0; q < nq; q++) outv[q * L + t] = tmp[t * nq + q];
                emit(outv.data(), outv.size() * 4);
                break;
            }
            case CLANN_X_QUERY_SKETCHES: {
                need_cluster();
                const uint64_t nq = index->last_nq;
                const uint32_t f = index->h_fset_of[arg];
                emit_dev(index->W->w_sketches.p + (size_t)f * nq * kNumSketches, nq * kNumSketches * 8);
                break;
            }
            case CLANN_X_CLUSTER_ORDER: {
                // index.rs:592-616: stable ascending order of the centre distances, derived on the host for the tests
                index->require_built();
                const uint64_t nq = index->last_nq;
                // the workspace may hold the tensor-pipe screen (exact only below each query's limit): evaluate every distance
                // exactly into a scratch copy. Needs the query buffer of the last search to be still alive.
                DevBuf<float> d_cd;
                DevBuf<uint32_t> d_first;
                d_cd.alloc(nq * K);
                d_first.alloc(nq);
                {
                    SearchParams p = index->params();
                    QueryBatch b = index->batch(index->cur_queries, nq, nullptr, nullptr, nullptr);
                    b.cdist = d_cd.p;
                    b.first = d_first.p;
                    b.exact_limit = nullptr;
                    launch_center_order(p, b, 0);
                    CLANN_CUDA(cudaDeviceSynchronize());
                }
                std::vector<float> cd = d_cd.download(nq * K);
                std::vector<uint32_t> order(nq * K);
                for (uint64_t q = 0; q < nq; q++) {
                    uint32_t* o = order.data() + q * K;
                    std::iota(o, o + K, 0u);
                    const float* dq = cd.data() + q * K;
                    std::stable_sort(o, o + K, [&](uint32_t a, uint32_t b) { return dq[a] < dq[b]; });
                }
                emit(order.data(), order.size() * 4);
                break;
            }
            case CLANN_X_BUILD_MS: emit(index->build_ms, sizeof(index->build_ms)); break;
            case CLANN_X_REFERENCE_STREAM: {
                need_owned_cluster();
                if (index->h_brute[arg] || index->h_sizes[arg] == 0)
                    throw StatusError(CLANN_ERR_SERIALIZE, "brute-force clusters have no PUFFINN index (index.rs:204-205)");
                const uint32_t nc = index->h_sizes[arg], L = index->g.L;
                const uint64_t off = index->h_offsets[arg];
                std::vector<int16_t> rows = index->d_q15.download((size_t)nc * index->g.sl, off * index->g.sl);
                std::vector<uint64_t> sks = index->d_sketches.download((size_t)nc * kNumSketches, off * kNumSketches);
                std::vector<uint32_t> th = index->d_tbl_hash.download((size_t)L * nc, table_base(off, nc, L, 0));
                std::vector<uint32_t> ti = index->d_tbl_idx.download((size_t)L * nc, table_base(off, nc, L, 0));
                std::vector<uint8_t> blob = write_reference_stream(index->g, index->fsets[index->h_fset_of[arg]], nc, rows.data(), sks.data(),
                                                                   th.data(), ti.data());
                emit(blob.data(), blob.size());
                break;
            }
            case CLANN_X_TABLE_DIR: {
                need_owned_cluster();
                const uint64_t L = index->g.L;
                emit_dev(index->d_tbl_dir.p + (uint64_t)arg * L * kDirEntries, L * kDirEntries * 4);
                break;
            }
            case CLANN_X_QUERY_ANCHORS:
            case CLANN_X_QUERY_RANGES: {
                need_cluster();
                if (index->last_nq == 0 || !index->W->w_codes.p) throw StatusError(CLANN_ERR_NOT_BUILT, "no search batch to trace yet");
                if (index->h_brute[arg] || index->h_sizes[arg] == 0 || (index->shard_count > 1 && index->h_owner[arg] != index->shard_rank))
                    throw StatusError(CLANN_ERR_ARG, "cluster has no tables on this rank (brute-force, empty or foreign)");
                const uint64_t nq = index->last_nq, L = index->g.L;
                const uint64_t na = nq * L, nr = nq * kMaxHashBits * L * 2;
                const uint64_t bytes = (what == CLANN_X_QUERY_ANCHORS ? na : nr) * 4;
                if (size) *size = bytes;
                if (dst) {
                    if (cap < bytes) throw StatusError(CLANN_ERR_BOUNDS, "export buffer too small");
                    DevBuf<uint32_t> d_a, d_r;
                    d_a.alloc(na);
                    d_r.alloc(nr);
                    SearchParams p = index->params();
                    QueryBatch b = index->batch(index->cur_queries, nq, nullptr, nullptr, nullptr);
                    launch_export_ranges(p, b, (uint32_t)arg, d_a.p, d_r.p, 0);
                    CLANN_CUDA(cudaMemcpy(dst, what == CLANN_X_QUERY_ANCHORS ? d_a.p : d_r.p, bytes, cudaMemcpyDeviceToHost));
                }
                break;
            }
            case CLANN_X_STOP_POINTS: {
                index->require_built();
                const uint64_t nq = index->last_nq;
                const uint64_t sb = query_state_bytes((uint32_t)index->cfg.k);
                std::vector<uint8_t> st = index->W->w_state.download(nq * sb);
                std::vector<uint32_t> outv(nq * 2);
                for (uint64_t q = 0; q < nq; q++) {
                    const QueryStateHeader* h = reinterpret_cast<const QueryStateHeader*>(st.data() + q * sb);
                    outv[2 * q] = (uint32_t)(h->stop_point >> 32);
                    outv[2 * q + 1] = (uint32_t)h->stop_point;
                }
                emit(outv.data(), outv.size() * 4);
                break;
            }
            case CLANN_X_VISIT_LOG: {
                index->require_built();
                if (!index->W->ws_vlog || !index->W->w_visit_log.p)
                    throw StatusError(CLANN_ERR_ARG, "no visit log: set option visit_log before the search");
                std::vector<uint32_t> log = index->W->w_visit_log.download(index->last_nq * index->W->ws_vlog * 4);
                emit(log.data(), log.size() * 4);
                break;
            }
            default: throw StatusError(CLANN_ERR_ARG, "unknown export selector");
        }
    });
}

void clann_destroy(clann_index* index) { delete index; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ C ABI (legacy CPUFFINN_*)

// One reference `puffinn::Index<CosineSimilarity>` handle: rows staged on the host until rebuild (dataset.hpp:109-124),
// then a single-cluster device index searched by one warp per call.
struct CPUFFINN {
    int dim = 0;
    std::vector<float> rows;
    uint32_t count = 0;
    clann_index* ix = nullptr;
    uint32_t built_k = 0;
    unsigned num_maps = 0;
    bool loaded = false;  // came from CPUFFINN_load_from_file: the fp32 rows are not known, so it cannot be rebuilt here
    DevBuf<float> d_query;
    DevBuf<uint32_t> d_out;  // [k] ids, then count, distance computations, stop point
    ~CPUFFINN() { delete ix; }
};

static std::atomic<unsigned> g_legacy_distcomp{0};  // performance.hpp:65-80: counter of the last query, process-global

// One query of a legacy handle: hash + sketch the query, one warp runs Index::search (collection.hpp:543-601) with the given
// FilterType. Writes up to k ids (best first) and returns their number; *stop = (depth << 16 | table) where the stop rule fired.
static uint32_t legacy_search(CPUFFINN* index, const float* query, uint32_t k, float recall, float max_sim, int filter_type,
                              uint32_t* result, uint32_t* stop) {
    clann_index* ix = index->ix;
    ix->cfg.k = k;
    cudaStream_t s = 0;
    ix->ensure_workspace(1, s);
    ix->ensure_stop_table(recall, s);
    index->d_query.upload(query, (size_t)index->dim, s);
    index->d_out.ensure((size_t)k + 3);
    SearchParams p = ix->params();
    QueryBatch b = ix->batch(index->d_query.p, 1, nullptr, nullptr, nullptr);
    launch_prep_queries(p, b, s);
    launch_sketch(b.q15, ix->W->w_tiles.p, ix->W->w_ntiles, ix->d_planes.p, ix->g.sl, b.sketches, s);
    launch_codes(b.q15, ix->w_code_tiles(1, s), ix->W->w_ntiles, ix->d_signbits.p, ix->g, b.codes, 1, ix->g.L, s);
    launch_puffinn_search(p, b, ix->d_stop.p, max_sim, filter_type, index->d_out.p, index->d_out.p + k, index->d_out.p + k + 1,
                          index->d_out.p + k + 2, s);
    std::vector<uint32_t> out(k + 3);
    CLANN_CUDA(cudaMemcpyAsync(out.data(), index->d_out.p, sizeof(uint32_t) * (k + 3), cudaMemcpyDeviceToHost, s));
    CLANN_CUDA(cudaStreamSynchronize(s));
    CLANN_CUDA(cudaGetLastError());
    const uint32_t cnt = out[k] < k ? out[k] : k;
    for (uint32_t i = 0; i < cnt; i++) result[i] = out[i];
    g_legacy_distcomp.store(out[k + 1]);
    if (stop) *stop = out[k + 2];
    return cnt;
}

extern "C" {

CPUFFINN* CPUFFINN_index_create(const char* dataset_type, int dataset_args) {
    if (!dataset_type || strcmp(dataset_type, "angular") != 0 || dataset_args <= 0) {
        // c_binder.cpp:39-50 also accepts "jaccard" but every later call treats the handle as cosine; refuse it instead
        fprintf(stderr, "Error: Unsupported dataset type '%s'. Only 'angular' is supported.\n", dataset_type ? dataset_type : "(null)");
        return nullptr;
    }
    try {
        auto* h = new CPUFFINN();
        h->dim = dataset_args;
        return h;
    } catch (...) {
        return nullptr;
    }
}

void CPUFFINN_index_insert_cosine(CPUFFINN* index, float* point, int dimension) {
    if (!index || !point) return;
    if (dimension != index->dim) {  // unit_vector.hpp:66-68 throws std::invalid_argument through extern "C"
        fprintf(stderr, "CPUFFINN_index_insert_cosine: dimension %d != %d, point ignored\n", dimension, index->dim);
        return;
    }
    try {
        index->rows.insert(index->rows.end(), point, point + dimension);
        index->count++;
    } catch (...) {
    }
}

static void legacy_build(CPUFFINN* h, uint32_t k) {
    clann_config cfg{h->num_maps, 1.0f, k, 0.9f};
    auto* ix = new clann_index();
    try {
        ix->puffinn_mode = true;
        ix->init(h->rows.data(), h->count, (uint32_t)h->dim, cfg);
        ix->seed = (uint64_t)std::random_device{}() << 32 | std::random_device{}();
        // a second rebuild keeps the functions drawn by the first (collection.hpp:257-263)
        if (h->ix && h->ix->g.L == ix->g.L && !h->ix->fsets.empty()) ix->fsets = h->ix->fsets;
        ix->build();
    } catch (...) {
        delete ix;
        throw;
    }
    delete h->ix;
    h->ix = ix;
    h->built_k = k;
}

uint64_t CPUFFINN_index_rebuild(CPUFFINN* index, unsigned int num_maps) {
    if (!index) return 0;
    try {
        if (num_maps == 0) return 0;  // collection.hpp:242-244 throws -> c_binder.cpp:57-59 returns 0
        if (index->count == 0 || index->loaded) return 0;
        index->num_maps = num_maps;
        legacy_build(index, 10);
        const HashGeom& g = index->ix->g;
        // same quantity the reference returns (collection.hpp:249-254), for the device-resident layout
        uint64_t per_point = (uint64_t)g.sl * 2 + kNumSketches * 8 + (uint64_t)g.L * 8;
        return per_point * index->count + (uint64_t)kNumPlanes * g.sl * 2 + (uint64_t)g.L * g.fph * kRotations * g.npts;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return 0;
    } catch (...) {
        return 0;
    }
}

uint32_t* CPUFFINN_search_cosine(CPUFFINN* index, float* query, unsigned int k, float recall, float max_sim, int dimension) {
    if (!query || dimension <= 0) {
        fprintf(stderr, "Error: Query is null or empty.\n");
        return nullptr;
    }
    if (!index || !index->ix || dimension != index->dim) return nullptr;
    try {
        const unsigned words = k > 1 ? k : 1;
        uint32_t* result = static_cast<uint32_t*>(malloc(sizeof(uint32_t) * words));
        if (!result) return nullptr;
        for (unsigned i = 0; i < words; i++) result[i] = 0xFFFFFFFFu;  // EMPTY_RESULT_SENTINEL, c_binder.h:8
        if (k == 0) return result;
        legacy_search(index, query, k, recall, max_sim, 0, result, nullptr);
        return result;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return nullptr;
    } catch (...) {
        return nullptr;
    }
}

// Index::search with its FilterType argument (collection.hpp:22-34,324-334), which c_binder.cpp never passes: 0 = Default (what
// CPUFFINN_search_cosine runs), 1 = None, 2 = Simple (:671-765). Caller-owned result buffer of k words.
int clann_puffinn_search(CPUFFINN* index, const float* query, uint32_t k, float recall, float max_sim, int filter_type,
                         uint32_t* out_ids, uint32_t* out_count, uint32_t* out_stop_depth) {
    return guarded([&] {
        if (!index || !query || !out_ids || !out_count) throw StatusError(CLANN_ERR_ARG, "null argument");
        if (!index->ix) throw StatusError(CLANN_ERR_NOT_BUILT, "CPUFFINN_index_rebuild has not been called");
        if (filter_type < 0 || filter_type > 2) throw StatusError(CLANN_ERR_ARG, "filter_type: 0 = Default, 1 = None, 2 = Simple");
        if (k == 0) throw StatusError(CLANN_ERR_CONFIG, "k must be at least 1");
        for (uint32_t i = 0; i < k; i++) out_ids[i] = 0xFFFFFFFFu;
        uint32_t stop = 0;
        *out_count = legacy_search(index, query, k, recall, max_sim, filter_type, out_ids, &stop);
        if (out_stop_depth) *out_stop_depth = stop >> 16;
    });
}

unsigned int CPUFFINN_get_distance_computations(void) { return g_legacy_distcomp.load(); }
void CPUFFINN_clear_distance_computations(void) { g_legacy_distcomp.store(0); }

// Persistence (c_binder.cpp:4-36,106-146). The reference appends the bytes of Index::serialize as dataset "index_{id}" of an
// HDF5 file; there is no HDF5 here, so the same bytes go into a flat record file instead: a sequence of
//   { char magic[8] = "CLB2REC"; u32 name_len; u32 reserved; u64 payload_len; name; payload }
// records. The payload is byte-compatible with the reference (tests/test_gpu_parity.py), only the container differs.
void CPUFFINN_save_index(CPUFFINN* index, const char* file_name, int index_number) {
    if (!index || !index->ix || !index->ix->built || !file_name) {
        fprintf(stderr, "Error: CPUFFINN_save_index needs a rebuilt index and a file name\n");
        return;
    }
    try {
        clann_index* ix = index->ix;
        const uint32_t nc = (uint32_t)ix->n, L = ix->g.L;
        std::vector<int16_t> rows = ix->d_q15.download((size_t)nc * ix->g.sl);
        std::vector<uint64_t> sks = ix->d_sketches.download((size_t)nc * kNumSketches);
        std::vector<uint32_t> th = ix->d_tbl_hash.download((size_t)L * nc), ti = ix->d_tbl_idx.download((size_t)L * nc);
        std::vector<uint8_t> blob = write_reference_stream(ix->g, ix->fsets[0], nc, rows.data(), sks.data(), th.data(), ti.data());
        const std::string name = "index_" + std::to_string(index_number);
        {
            // The reference appends an HDF5 dataset to the file ClusteredIndex::serialize has just created (index.rs:526-552,
            // c_binder.cpp:106-146). There is no HDF5 here: appending records to such a file would produce a container neither
            // side can read, so anything that is not empty and does not start with a CLB2REC record is refused, loudly.
            FILE* probe = fopen(file_name, "rb");
            if (probe) {
                char head[8];
                const size_t got = fread(head, 1, 8, probe);
                fclose(probe);
                if (got > 0 && (got < 8 || memcmp(head, "CLB2REC", 8) != 0)) {
                    g_last_error = std::string("CPUFFINN_save_index: ") + file_name + " is not a libclann_b200 record file (an HDF5 "
                                   "container written by the unmodified crate?); persistence through HDF5 is not supported, nothing written";
                    fprintf(stderr, "Error: %s\n", g_last_error.c_str());
                    return;
                }
            }
        }
        FILE* f = fopen(file_name, "ab");
        if (!f) {
            fprintf(stderr, "Error opening file: %s\n", file_name);
            return;
        }
        const char magic[8] = {'C', 'L', 'B', '2', 'R', 'E', 'C', 0};
        const uint32_t name_len = (uint32_t)name.size(), reserved = 0;
        const uint64_t payload_len = blob.size();
        bool ok = fwrite(magic, 1, 8, f) == 8 && fwrite(&name_len, 4, 1, f) == 1 && fwrite(&reserved, 4, 1, f) == 1 &&
                  fwrite(&payload_len, 8, 1, f) == 1 && fwrite(name.data(), 1, name_len, f) == name_len &&
                  fwrite(blob.data(), 1, blob.size(), f) == blob.size();
        if (fclose(f) != 0 || !ok) fprintf(stderr, "Error writing %s to file: %s\n", name.c_str(), file_name);
    } catch (const std::exception& e) {
        g_last_error = e.what();
        fprintf(stderr, "Error saving index_%d: %s\n", index_number, e.what());
    } catch (...) {
        fprintf(stderr, "Error saving index_%d\n", index_number);
    }
}

// Returns NULL on failure (the reference throws through extern "C", c_binder.cpp:8,15).
CPUFFINN* CPUFFINN_load_from_file(const char* file_name, const char* dataset_name) {
    if (!file_name || !dataset_name) return nullptr;
    FILE* f = fopen(file_name, "rb");
    if (!f) {
        fprintf(stderr, "Failed to open file %s\n", file_name);
        return nullptr;
    }
    CPUFFINN* h = nullptr;
    try {
        std::vector<uint8_t> blob;
        bool found = false;
        uint64_t file_size = 0;
        if (fseek(f, 0, SEEK_END) == 0) {
            const long e = ftell(f);
            file_size = e > 0 ? (uint64_t)e : 0;
        }
        rewind(f);
        for (;;) {
            char magic[8];
            uint32_t name_len, reserved;
            uint64_t payload_len;
            if (fread(magic, 1, 8, f) != 8) break;
            if (memcmp(magic, "CLB2REC", 8) != 0 || fread(&name_len, 4, 1, f) != 1 || fread(&reserved, 4, 1, f) != 1 ||
                fread(&payload_len, 8, 1, f) != 1 || name_len > 4096)
                throw StatusError(CLANN_ERR_SERIALIZE, "not a libclann_b200 record file");
            std::string name(name_len, '\0');
            if (fread(&name[0], 1, name_len, f) != name_len) throw StatusError(CLANN_ERR_SERIALIZE, "truncated record file");
            if (name == dataset_name) {  // the last record of that name wins (the file is append-only)
                if (payload_len > file_size) throw StatusError(CLANN_ERR_SERIALIZE, "truncated record file");
                blob.resize(payload_len);
                if (fread(blob.data(), 1, payload_len, f) != payload_len) throw StatusError(CLANN_ERR_SERIALIZE, "truncated record file");
                found = true;
            } else if (payload_len > (uint64_t)0x7fffffffffffll || fseek(f, (long)payload_len, SEEK_CUR) != 0) {
                throw StatusError(CLANN_ERR_SERIALIZE, "truncated record file");
            }
        }
        if (!found) throw StatusError(CLANN_ERR_SERIALIZE, std::string("no record named ") + dataset_name);
        h = new CPUFFINN();
        h->ix = new clann_index();
        h->ix->load_stream(blob.data(), blob.size());
        h->dim = (int)h->ix->g.d;
        h->count = (uint32_t)h->ix->n;
        h->num_maps = h->ix->g.L;
        h->built_k = 10;
        h->loaded = true;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        fprintf(stderr, "Failed to load %s from %s: %s\n", dataset_name, file_name, e.what());
        delete h;
        h = nullptr;
    } catch (...) {
        delete h;
        h = nullptr;
    }
    fclose(f);
    return h;
}

}  // extern "C"
