/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C restatement of the CLANN build + search hot path, used ONLY as the checker for the
 * CUDA path: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link it; nothing under clann_b200/
 * may. Every function cites the reference file:line (paths into /root/reference) it follows.
 *
 * Pinning status
 *   L0/L1 (PUFFINN: Q15 store, Q15 dot, SimHash sketches, FHT cross-polytope codes, table sort, anchors, ranges,
 *   search_maps incl. its quirks, MaxBuffer, stop rule): PINNED — checked bit-for-bit against the real reference
 *   compiled from /root/reference (oracle/_ref, g++ 13.3 -O3 -march=x86-64-v3) by tests/test_oracle_vs_ref.py, against
 *   the reference's own known-answer tests (tests/test_oracle_known_answers.py) and the committed golden fixtures.
 *   L3 (CLANN Rust layer: gmm.rs, index.rs search loop, heap.rs, angulardata.rs): PARITY UNPINNED — the crate cannot be
 *   built here (no cargo) and its fp32 dot is ndarray 0.16.1 `unrolled_dot` (Cargo.lock; not vendored). The published
 *   algorithm is restated; the only reference known-answer test at that level (index.rs:695-748) is checked.
 *
 * Compiler-dependent detail: UnitVectorFormat::store's sum of squares (unit_vector.hpp:71-74) is compiled by g++ -O3
 * into packed multiplies followed by in-order scalar adds for the first d - d%4 elements and scalar FMAs for the last
 * d%4 elements. orc_store_q15 reproduces exactly that (verified over d = 1..130 against oracle/_ref).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile). -ffp-contract=off matters: every FMA below is explicit.
 */
#include "clann_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ L0 */

/* format/unit_vector.hpp:40-45 — min(v * 2^15, 32767) then truncation toward zero. */
int16_t orc_to_q15(float v) {
    float s = v * 32768.0f;
    if (s > 32767.0f) s = 32767.0f;
    return (int16_t)s;
}

/* format/unit_vector.hpp:49-51 */
float orc_from_q15(int16_t v) { return (float)v / 32768.0f; }

/* format/generic.hpp:28-40 with ALIGNMENT 32 bytes / 2-byte type = multiples of 16. */
uint32_t orc_storage_len(uint32_t d) { return (d + 15) / 16 * 16; }

/* math.hpp:105-113 */
uint32_t orc_ceil_log(uint32_t v) {
    uint32_t lg = 0, p = 1;
    while (p < v) {
        lg++;
        p *= 2;
    }
    return lg;
}

/* format/unit_vector.hpp:61-89 (see header note on the summation shape). */
void orc_store_q15(const float* v, uint32_t d, uint32_t sl, int16_t* out) {
    float acc = 0.0f;
    uint32_t body = d & ~3u;
    for (uint32_t i = 0; i < body; i++) acc = acc + v[i] * v[i];
    for (uint32_t i = body; i < d; i++) acc = fmaf(v[i], v[i], acc);
    float len = sqrtf(acc);
    for (uint32_t i = 0; i < d; i++) {
        float x = v[i];
        if (len != 0.0f) x = x / len;
        out[i] = orc_to_q15(x);
    }
    for (uint32_t i = d; i < sl; i++) out[i] = 0;
}

/* math.hpp:37-44 (== the AVX2 vpmulhrsw path :11-35; integer, order independent, wrapping int16 accumulator). */
int16_t orc_dot_i16(const int16_t* a, const int16_t* b, uint32_t n) {
    uint16_t res = 0;
    for (uint32_t i = 0; i < n; i++) {
        int32_t precise = (int32_t)a[i] * (int32_t)b[i];
        res = (uint16_t)(res + (uint16_t)(int16_t)(((precise >> 14) + 1) >> 1));
    }
    return (int16_t)res;
}

/* similarity_measure/cosine.hpp:19-23. The reference passes desc.args (d) as the length; the AVX2 loop rounds it up to
 * the next multiple of 16 = storage_len, and the padding is zero. */
float orc_similarity(const int16_t* a, const int16_t* b, uint32_t sl) {
    float dot = orc_from_q15(orc_dot_i16(a, b, sl));
    return (dot + 1.0f) / 2.0f;
}

/* external/ffht/fht_avx.c:107-195 (log_n=5), :375-567 (log_n=7): in-place unnormalised WHT, butterflies of stride
 * 1,2,4,...,2^(m-1) in that order, each fl(u+v), fl(u-v). */
void orc_fht(float* buf, uint32_t m) {
    uint32_t n = 1u << m;
    for (uint32_t h = 1; h < n; h <<= 1) {
        for (uint32_t i = 0; i < n; i += 2 * h) {
            for (uint32_t j = i; j < i + h; j++) {
                float u = buf[j], v = buf[j + h];
                buf[j] = u + v;
                buf[j + h] = u - v;
            }
        }
    }
}

/* hash/crosspolytope.hpp:187-209 + encode_closest_axis :131-144 */
uint32_t orc_fht_hash(const orc_functions* fn, uint32_t fidx, const int16_t* q15) {
    float buf[1024];
    uint32_t n = 1u << fn->m;
    for (uint32_t i = 0; i < fn->d; i++) buf[i] = orc_from_q15(q15[i]);
    for (uint32_t i = fn->d; i < n; i++) buf[i] = 0.0f;
    const int8_t* s = fn->signs + (size_t)fidx * fn->rotations * n;
    for (uint32_t r = 0; r < fn->rotations; r++) {
        for (uint32_t i = 0; i < n; i++) buf[i] *= (float)s[r * n + i];
        orc_fht(buf, fn->m);
    }
    uint32_t res = 0;
    float best = 0.0f;
    for (uint32_t i = 0; i < n; i++) {
        if (buf[i] > best) {
            res = i;
            best = buf[i];
        } else if (-buf[i] > best) {
            res = i + n;
            best = -buf[i];
        }
    }
    return res;
}

/* hash_source/independent.hpp:70-86 */
void orc_codes(const orc_functions* fn, const int16_t* q15, uint32_t* out) {
    for (uint32_t rep = 0; rep < fn->L; rep++) {
        uint64_t res = 0;
        for (uint32_t i = 0; i < fn->fph; i++) {
            res <<= fn->bpf;
            res |= orc_fht_hash(fn, rep * fn->fph + i, q15);
        }
        res >>= fn->cut;
        out[rep] = (uint32_t)res;
    }
}

/* filterer.hpp:76-102 via independent.hpp:70-86 with 32 hashers x 64 one-bit SimHash functions (simhash.hpp:41-44). */
void orc_sketch(const orc_functions* fn, const int16_t* q15, uint64_t* out) {
    for (uint32_t s = 0; s < ORC_NUM_SKETCHES; s++) {
        uint64_t res = 0;
        for (uint32_t b = 0; b < ORC_SKETCH_BITS; b++) {
            const int16_t* plane = fn->planes + (size_t)(s * ORC_SKETCH_BITS + b) * fn->sl;
            res <<= 1;
            res |= (uint64_t)(orc_dot_i16(plane, q15, fn->sl) >= 0);
        }
        out[s] = res;
    }
}

/* sorthash.hpp:133-194 — stable LSD radix sort, three byte passes, key + payload. */
void orc_sort_pairs_24(const uint32_t* hashes_in, const uint32_t* idx_in, uint32_t n, uint32_t* hashes_out, uint32_t* idx_out) {
    uint32_t* ha = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* ia = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    memcpy(ha, hashes_in, sizeof(uint32_t) * n);
    memcpy(ia, idx_in, sizeof(uint32_t) * n);
    uint32_t *hsrc = ha, *isrc = ia, *hdst = hashes_out, *idst = idx_out;
    for (int pass = 0; pass < 3; pass++) {
        uint32_t bins[256];
        memset(bins, 0, sizeof(bins));
        int sh = 8 * pass;
        for (uint32_t i = 0; i < n; i++) bins[(hsrc[i] >> sh) & 0xff]++;
        uint32_t sum = 0;
        for (int b = 0; b < 256; b++) {
            uint32_t c = bins[b];
            bins[b] = sum;
            sum += c;
        }
        for (uint32_t i = 0; i < n; i++) {
            uint32_t t = bins[(hsrc[i] >> sh) & 0xff]++;
            hdst[t] = hsrc[i];
            idst[t] = isrc[i];
        }
        uint32_t* th = hsrc; hsrc = hdst; hdst = th;
        uint32_t* ti = isrc; isrc = idst; idst = ti;
    }
    /* after 3 passes the result sits in hsrc/isrc == hashes_out/idx_out (out, in, out) */
    if (hsrc != hashes_out) {
        memcpy(hashes_out, hsrc, sizeof(uint32_t) * n);
        memcpy(idx_out, isrc, sizeof(uint32_t) * n);
    }
    free(ha);
    free(ia);
}

/* hash/crosspolytope.hpp:116-118 */
static float est_lookup(const orc_functions* fn, float sim, uint32_t bits) {
    size_t bin = (size_t)(sim / fn->eps);
    return fn->est[(size_t)bits * ORC_EST_BINS + bin];
}

/* hash_source/hash_source.hpp:49-57: std::pow(float,int) promotes to double; double*float -> double -> float. */
static float concat_prob(const orc_functions* fn, uint32_t num_bits, float sim) {
    uint32_t whole = num_bits / fn->bpf;
    uint32_t rem = num_bits % fn->bpf;
    float wp = est_lookup(fn, sim, fn->bpf);
    float rp = est_lookup(fn, sim, rem);
    return (float)(pow((double)wp, (double)(int)whole) * (double)rp);
}

/* hash_source/independent.hpp:108-119: `1.0-col_prob` is double, `1-last_prob` is float. */
float orc_failure_probability(const orc_functions* fn, uint32_t hash_length, uint64_t tables, uint64_t max_tables, float kth) {
    float col = concat_prob(fn, hash_length, kth);
    float last = concat_prob(fn, hash_length + 1, kth);
    double a = pow(1.0 - (double)col, (double)tables);
    float one_minus_last = 1.0f - last;
    double b = pow((double)one_minus_last, (double)(max_tables - tables));
    return (float)(a * b);
}

/* filterer.hpp:108-111 with simhash.hpp:96-102: acos on float, division by M_PI in double, result rounded to float. */
uint32_t orc_max_sketch_diff(float kth) {
    float arg = 2.0f * kth - 1.0f;
    float cp = (float)(1.0 - (double)acosf(arg) / M_PI);
    float r = roundf((float)(64.0 * (1.0 - (double)cp)));
    return (uint32_t)(uint8_t)r;
}

/* ------------------------------------------------------------------------------------------------ MaxBuffer */

typedef struct {
    uint32_t idx;
    float val;
} mb_pair;

typedef struct {
    uint32_t size, inserted;
    float minval;
    mb_pair* data;
} maxbuffer;

static void mb_init(maxbuffer* mb, uint32_t k) { /* maxbuffer.hpp:50-60 */
    mb->size = k;
    mb->inserted = 0;
    mb->minval = (k == 0) ? 1.0f : 0.0f;
    mb->data = (mb_pair*)malloc(sizeof(mb_pair) * (2 * k + 1));
}

static int mb_cmp(const void* pa, const void* pb) { /* maxbuffer.hpp:27-31: value desc, then idx desc */
    const mb_pair* a = (const mb_pair*)pa;
    const mb_pair* b = (const mb_pair*)pb;
    if (a->val > b->val) return -1;
    if (a->val < b->val) return 1;
    if (a->idx > b->idx) return -1;
    if (a->idx < b->idx) return 1;
    return 0;
}

static void mb_filter(maxbuffer* mb) { /* maxbuffer.hpp:25-46 */
    qsort(mb->data, mb->inserted, sizeof(mb_pair), mb_cmp);
    uint32_t dedup = mb->inserted < 1 ? mb->inserted : 1;
    for (uint32_t i = 1; i < mb->inserted; i++) {
        if (mb->data[i].idx != mb->data[dedup - 1].idx) {
            mb->data[dedup] = mb->data[i];
            dedup++;
        }
    }
    mb->inserted = dedup < mb->size ? dedup : mb->size;
    if (mb->inserted == mb->size && mb->size != 0) mb->minval = mb->data[mb->inserted - 1].val;
}

static int mb_insert(maxbuffer* mb, uint32_t idx, float v) { /* maxbuffer.hpp:64-76 */
    v = fminf(1.0f, fmaxf(0.0f, v));
    if (v <= mb->minval) return 0;
    if (mb->inserted == 2 * mb->size) mb_filter(mb);
    mb->data[mb->inserted].idx = idx;
    mb->data[mb->inserted].val = v;
    mb->inserted++;
    return 1;
}

int orc_maxbuffer_run(uint32_t k, const uint32_t* ids, const float* vals, int n, uint32_t* out_ids, float* out_vals, float* minval) {
    maxbuffer mb;
    mb_init(&mb, k);
    for (int i = 0; i < n; i++) mb_insert(&mb, ids[i], vals[i]);
    if (minval) *minval = mb.minval;
    mb_filter(&mb); /* best_entries, maxbuffer.hpp:79-86 */
    for (uint32_t i = 0; i < mb.inserted; i++) {
        out_ids[i] = mb.data[i].idx;
        out_vals[i] = mb.data[i].val;
    }
    int r = (int)mb.inserted;
    free(mb.data);
    return r;
}

/* ------------------------------------------------------------------------------------------------ L1 index */

static void fn_copy(orc_functions* dst, const orc_functions* src) {
    *dst = *src;
    size_t np = (size_t)2048 * src->sl;
    size_t ns = (size_t)src->L * src->fph * src->rotations * ((size_t)1 << src->m);
    size_t ne = (size_t)(src->m + 2) * ORC_EST_BINS;
    dst->planes = (int16_t*)malloc(np * sizeof(int16_t));
    dst->signs = (int8_t*)malloc(ns);
    dst->est = (float*)malloc(ne * sizeof(float));
    memcpy(dst->planes, src->planes, np * sizeof(int16_t));
    memcpy(dst->signs, src->signs, ns);
    memcpy(dst->est, src->est, ne * sizeof(float));
}

/* prefixmap.hpp:169-247 for one table (padding + prefix_index). sorted_* hold n entries. */
static void table_finish(orc_index* ix, uint32_t t, const uint32_t* sorted_h, const uint32_t* sorted_i) {
    uint32_t n = ix->n;
    size_t len = (size_t)n + 2 * ORC_SEGMENT;
    uint32_t* H = ix->hashes + t * len;
    uint32_t* I = ix->indices + t * len;
    for (int i = 0; i < ORC_SEGMENT; i++) {
        H[i] = 0xffffffffu;
        I[i] = 0;
        H[ORC_SEGMENT + n + i] = 0xffffffffu;
        I[ORC_SEGMENT + n + i] = 0;
    }
    memcpy(H + ORC_SEGMENT, sorted_h, sizeof(uint32_t) * n);
    memcpy(I + ORC_SEGMENT, sorted_i, sizeof(uint32_t) * n);
    uint32_t* P = ix->prefix_index + (size_t)t * ((1u << ORC_PREFIX_BITS) + 1);
    uint32_t idx = 0;
    for (uint32_t prefix = 0; prefix < (1u << ORC_PREFIX_BITS); prefix++) {
        while (idx < n && (H[ORC_SEGMENT + idx] >> (ORC_MAX_HASHBITS - ORC_PREFIX_BITS)) < prefix) idx++;
        P[prefix] = ORC_SEGMENT + idx;
    }
    P[1u << ORC_PREFIX_BITS] = ORC_SEGMENT + n;
}

static orc_index* index_alloc(const orc_functions* fn, uint32_t n, int copy_fn) {
    orc_index* ix = (orc_index*)calloc(1, sizeof(orc_index));
    if (copy_fn) fn_copy(&ix->fn, fn);
    else ix->fn = *fn;
    ix->n = n;
    size_t len = (size_t)n + 2 * ORC_SEGMENT;
    ix->q15 = (int16_t*)calloc((size_t)(n ? n : 1) * fn->sl, sizeof(int16_t));
    ix->sketches = (uint64_t*)calloc((size_t)(n ? n : 1) * ORC_NUM_SKETCHES, sizeof(uint64_t));
    ix->hashes = (uint32_t*)malloc(sizeof(uint32_t) * len * fn->L);
    ix->indices = (uint32_t*)malloc(sizeof(uint32_t) * len * fn->L);
    ix->prefix_index = (uint32_t*)malloc(sizeof(uint32_t) * ((1u << ORC_PREFIX_BITS) + 1) * fn->L);
    return ix;
}

/* collection.hpp:219-222 (insert) + :241-306 (rebuild): Q15 store, sketches, table codes, per-table sort. */
orc_index* orc_index_build(const orc_functions* fn, const float* data, uint32_t n) {
    orc_index* ix = index_alloc(fn, n, 1);
    uint32_t L = fn->L;
    uint32_t* codes = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1) * L);
    for (uint32_t i = 0; i < n; i++) {
        int16_t* row = ix->q15 + (size_t)i * fn->sl;
        orc_store_q15(data + (size_t)i * fn->d, fn->d, fn->sl, row);
        orc_sketch(&ix->fn, row, ix->sketches + (size_t)i * ORC_NUM_SKETCHES);
        orc_codes(&ix->fn, row, codes + (size_t)i * L);
    }
    uint32_t* hin = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* iin = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* hout = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint32_t* iout = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (uint32_t t = 0; t < L; t++) {
        /* insertion order = ascending point id (per-thread staging concatenated in thread order under libgomp's static
         * schedule, prefixmap.hpp:157-159,192-197) */
        for (uint32_t i = 0; i < n; i++) {
            hin[i] = codes[(size_t)i * L + t];
            iin[i] = i;
        }
        orc_sort_pairs_24(hin, iin, n, hout, iout);
        table_finish(ix, t, hout, iout);
    }
    free(hin); free(iin); free(hout); free(iout); free(codes);
    return ix;
}

void orc_index_free(orc_index* ix) {
    if (!ix) return;
    free(ix->fn.planes); free(ix->fn.signs); free(ix->fn.est);
    free(ix->q15); free(ix->sketches); free(ix->hashes); free(ix->indices); free(ix->prefix_index);
    free(ix);
}

/* --- Index::serialize reader (collection.hpp:185-203 and the per-member serialize()s; layout in SURVEY.md 8c) */
typedef struct {
    const uint8_t* p;
    uint64_t len, off;
    int bad;
} rd;
static void rd_bytes(rd* r, void* dst, uint64_t n) {
    if (r->bad || r->off + n > r->len) {
        r->bad = 1;
        memset(dst, 0, n);
        return;
    }
    memcpy(dst, r->p + r->off, n);
    r->off += n;
}
static uint32_t rd_u32(rd* r) { uint32_t v; rd_bytes(r, &v, 4); return v; }
static uint64_t rd_u64(rd* r) { uint64_t v; rd_bytes(r, &v, 8); return v; }
static uint8_t rd_u8(rd* r) { uint8_t v; rd_bytes(r, &v, 1); return v; }
static float rd_f32(rd* r) { float v; rd_bytes(r, &v, 4); return v; }

orc_index* orc_index_import(const uint8_t* stream, uint64_t len) {
    rd r = {stream, len, 0, 0};
    orc_functions fn;
    memset(&fn, 0, sizeof(fn));
    /* Dataset (dataset.hpp:79-86) */
    fn.d = rd_u32(&r);
    fn.sl = rd_u32(&r);
    uint32_t n = rd_u32(&r);
    if (r.bad || fn.sl != orc_storage_len(fn.d)) return NULL;
    int16_t* q15 = (int16_t*)malloc(sizeof(int16_t) * (size_t)(n ? n : 1) * fn.sl);
    rd_bytes(&r, q15, sizeof(int16_t) * (size_t)n * fn.sl);
    /* Filterer (filterer.hpp:62-68): sketch args, SimHash source, sketches */
    uint32_t src_type = rd_u32(&r); /* HashSourceType::Independent == 0; SimHashArgs are empty */
    (void)rd_u32(&r); (void)rd_u32(&r); /* SimHash dataset description {d, sl} */
    uint64_t n_planes = rd_u64(&r);
    if (r.bad || src_type != 0 || n_planes != 2048) { free(q15); return NULL; }
    fn.planes = (int16_t*)malloc(sizeof(int16_t) * 2048 * (size_t)fn.sl);
    for (uint32_t f = 0; f < 2048; f++) {
        uint32_t dims = rd_u32(&r);
        if (dims != fn.sl) r.bad = 1;
        rd_bytes(&r, fn.planes + (size_t)f * fn.sl, sizeof(int16_t) * fn.sl);
    }
    uint32_t nh = rd_u32(&r), fph_s = rd_u32(&r);
    uint8_t bpf_s = rd_u8(&r);
    (void)rd_u32(&r); (void)rd_u32(&r); /* next_function, bits_to_cut */
    if (nh != 32 || fph_s != 64 || bpf_s != 1) r.bad = 1;
    uint64_t n_sk = rd_u64(&r);
    uint64_t* sketches = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n_sk ? n_sk : 1));
    rd_bytes(&r, sketches, sizeof(uint64_t) * n_sk);
    /* hash args (independent.hpp:135-139, crosspolytope.hpp:234-238) */
    uint32_t ht = rd_u32(&r);
    fn.rotations = rd_u32(&r);
    (void)rd_u32(&r); /* estimation_repetitions */
    (void)rd_f32(&r); /* estimation_eps */
    uint8_t has_source = rd_u8(&r);
    if (r.bad || ht != 0 || !has_source || n_sk != (uint64_t)n * 32) { free(q15); free(sketches); free(fn.planes); return NULL; }
    /* IndependentHashSource<FHTCrossPolytopeHash> (independent.hpp:56-68): family {desc, args, estimates}, functions */
    (void)rd_u32(&r); (void)rd_u32(&r);                 /* dataset description */
    (void)rd_u32(&r); (void)rd_u32(&r); (void)rd_f32(&r); /* args again */
    uint64_t rows = rd_u64(&r);
    fn.m = orc_ceil_log(fn.d);
    if (rows != fn.m + 2) r.bad = 1;
    fn.est = (float*)calloc((size_t)(fn.m + 2) * ORC_EST_BINS, sizeof(float));
    for (uint64_t b = 0; b < rows && !r.bad; b++) {
        uint64_t cols = rd_u64(&r);
        if (cols != ORC_EST_BINS) { r.bad = 1; break; }
        rd_bytes(&r, fn.est + b * ORC_EST_BINS, sizeof(float) * ORC_EST_BINS);
    }
    fn.eps = rd_f32(&r);
    uint64_t n_fn = rd_u64(&r);
    uint32_t npts = 1u << fn.m;
    fn.signs = (int8_t*)malloc((size_t)(n_fn ? n_fn : 1) * fn.rotations * npts);
    for (uint64_t f = 0; f < n_fn && !r.bad; f++) {
        uint32_t dd = rd_u32(&r), mm = rd_u32(&r), rr = rd_u32(&r);
        if (dd != fn.d || mm != fn.m || rr != fn.rotations) { r.bad = 1; break; }
        rd_bytes(&r, fn.signs + f * fn.rotations * npts, (uint64_t)fn.rotations * npts);
    }
    fn.L = rd_u32(&r);
    fn.fph = rd_u32(&r);
    fn.bpf = rd_u8(&r);
    (void)rd_u32(&r);
    fn.cut = rd_u32(&r);
    uint64_t num_maps = rd_u64(&r);
    uint8_t chunks = rd_u8(&r);
    if (r.bad || chunks || num_maps != fn.L || n_fn != (uint64_t)fn.L * fn.fph) {
        free(q15); free(sketches); free(fn.planes); free(fn.est); free(fn.signs);
        return NULL;
    }
    orc_index* ix = index_alloc(&fn, n, 0);
    memcpy(ix->q15, q15, sizeof(int16_t) * (size_t)n * fn.sl);
    memcpy(ix->sketches, sketches, sizeof(uint64_t) * (size_t)n * 32);
    free(q15); free(sketches);
    size_t tl = (size_t)n + 2 * ORC_SEGMENT;
    for (uint32_t t = 0; t < fn.L; t++) { /* prefixmap.hpp:128-154 */
        uint64_t l = rd_u64(&r);
        if (l != tl) { r.bad = 1; break; }
        rd_bytes(&r, ix->indices + t * tl, 4 * tl);
        rd_bytes(&r, ix->hashes + t * tl, 4 * tl);
        uint64_t rebuilding = rd_u64(&r);
        uint32_t hl = rd_u32(&r);
        if (rebuilding != 0 || hl != ORC_MAX_HASHBITS) { r.bad = 1; break; }
        rd_bytes(&r, ix->prefix_index + (size_t)t * 8193, 4 * 8193);
    }
    (void)rd_u32(&r); /* last_rebuild */
    if (r.bad || r.off != r.len) {
        orc_index_free(ix);
        return NULL;
    }
    return ix;
}

/* prefixmap.hpp:36-57 + :250-260 — halving search between the two prefix_index hints. */
uint32_t orc_anchor(const orc_index* ix, uint32_t t, uint32_t hash) {
    size_t tl = (size_t)ix->n + 2 * ORC_SEGMENT;
    const uint32_t* H = ix->hashes + t * tl;
    const uint32_t* P = ix->prefix_index + (size_t)t * 8193;
    uint32_t prefix = hash >> (ORC_MAX_HASHBITS - ORC_PREFIX_BITS);
    uint32_t start = P[prefix], end = P[prefix + 1];
    uint32_t half = end - start;
    while (half != 0) {
        half /= 2;
        uint32_t mid = start + half;
        start = (H[mid] < hash) ? (mid + 1) : start;
    }
    return start;
}

typedef struct {
    uint32_t hash, mask, start, end;
} pm_query;

/* prefixmap.hpp:267-304. Note prefix_start/prefix_end are never advanced by the reference. */
static void next_range(const orc_index* ix, uint32_t t, pm_query* q, uint32_t* rs, uint32_t* re) {
    size_t tl = (size_t)ix->n + 2 * ORC_SEGMENT;
    const uint32_t* H = ix->hashes + t * tl;
    uint32_t prev_mask = q->mask >> 1;
    uint32_t removed_bit = prev_mask & (0u - prev_mask);
    uint32_t bit_value = q->hash & removed_bit;
    uint32_t hash_prefix = q->hash & q->mask;
    if (bit_value == 0) {
        uint32_t next_idx = q->end;
        uint32_t start_idx = next_idx;
        while ((H[next_idx] & q->mask) == hash_prefix) next_idx += ORC_SEGMENT;
        uint32_t end_idx = next_idx;
        if (end_idx >= tl - ORC_SEGMENT) {
            uint32_t e2 = end_idx - ORC_SEGMENT;
            end_idx = start_idx > e2 ? start_idx : e2;
        }
        q->mask <<= 1;
        *rs = start_idx;
        *re = end_idx;
    } else {
        uint32_t next_idx = q->start - 1;
        uint32_t end_idx = next_idx + 1;
        while ((H[next_idx] & q->mask) == hash_prefix) next_idx -= ORC_SEGMENT;
        uint32_t start_idx = next_idx + 1;
        if (start_idx < ORC_SEGMENT) {
            uint32_t s2 = start_idx + ORC_SEGMENT;
            start_idx = end_idx < s2 ? end_idx : s2;
        }
        q->mask <<= 1;
        *rs = start_idx;
        *re = end_idx;
    }
}

void orc_query_ranges(const orc_index* ix, const uint32_t* codes, uint32_t* anchors, uint32_t* ranges) {
    uint32_t L = ix->fn.L;
    pm_query* qo = (pm_query*)malloc(sizeof(pm_query) * L);
    for (uint32_t t = 0; t < L; t++) {
        uint32_t a = orc_anchor(ix, t, codes[t]);
        anchors[t] = a;
        qo[t].hash = codes[t];
        qo[t].mask = 0xffffffffu;
        qo[t].start = qo[t].end = a;
    }
    for (uint32_t it = 0; it < ORC_MAX_HASHBITS; it++)
        for (uint32_t t = 0; t < L; t++) next_range(ix, t, &qo[t], &ranges[((size_t)it * L + t) * 2], &ranges[((size_t)it * L + t) * 2 + 1]);
    free(qo);
}

/* collection.hpp:524-541 */
static int search_bf(const orc_index* ix, const int16_t* q, uint32_t k, uint32_t* out) {
    maxbuffer mb;
    mb_init(&mb, k);
    for (uint32_t i = 0; i < ix->n; i++) mb_insert(&mb, i, orc_similarity(q, ix->q15 + (size_t)i * ix->fn.sl, ix->fn.sl));
    mb_filter(&mb);
    for (uint32_t i = 0; i < mb.inserted; i++) out[i] = mb.data[i].idx;
    int r = (int)mb.inserted;
    free(mb.data);
    return r;
}

static int passes(const uint64_t* qs, uint32_t max_diff, uint64_t sketch, uint32_t slot) { /* filterer.hpp:28-31 */
    return (uint32_t)__builtin_popcountll(sketch ^ qs[slot]) <= max_diff;
}

typedef struct {
    const uint32_t* cur;
    const uint32_t* end;
    uint32_t table;
} range_t;

/* collection.hpp:768-948 (search_maps) driven by :543-601 (search_formatted_query). */
int orc_index_search(const orc_index* ix, const float* qf, uint32_t k, float recall, float max_sim, uint32_t* out, orc_trace* tr) {
    const orc_functions* fn = &ix->fn;
    uint32_t L = fn->L;
    int16_t* q = (int16_t*)malloc(sizeof(int16_t) * fn->sl);
    orc_store_q15(qf, fn->d, fn->sl, q); /* collection.hpp:331-333 */
    orc_trace local;
    if (!tr) {
        memset(&local, 0, sizeof(local));
        tr = &local;
    }
    tr->distance_computations = tr->candidates = tr->stop_depth = tr->stop_table = tr->n_batches = 0;
    tr->passing_len = 0;
    tr->kth = 0;
    tr->max_sketch_diff = ORC_SKETCH_BITS;
    if (ix->n < 100) { /* collection.hpp:550-555 */
        int r = search_bf(ix, q, k, out);
        free(q);
        return r;
    }
    uint32_t* codes = (uint32_t*)malloc(sizeof(uint32_t) * L);
    uint64_t qs[ORC_NUM_SKETCHES];
    orc_codes(fn, q, codes);
    orc_sketch(fn, q, qs);
    uint32_t max_diff = ORC_SKETCH_BITS; /* filterer.hpp:101 */
    maxbuffer mb;
    mb_init(&mb, k);

    size_t tl = (size_t)ix->n + 2 * ORC_SEGMENT;
    pm_query* qo = (pm_query*)malloc(sizeof(pm_query) * L);
    for (uint32_t t = 0; t < L; t++) { /* collection.hpp:642-645 */
        uint32_t a = orc_anchor(ix, t, codes[t]);
        qo[t].hash = codes[t];
        qo[t].mask = 0xffffffffu;
        qo[t].start = qo[t].end = a;
    }
    static const uint32_t filler[2 * 32 * 4] = {0}; /* collection.hpp:609 */
    range_t* ranges = (range_t*)malloc(sizeof(range_t) * (L + 1));
    uint32_t passing[128 + 8 * 32];
    int stopped = 0;

    for (uint32_t depth = ORC_MAX_HASHBITS; depth > 0 && !stopped; depth--) {
        /* fill_ranges, collection.hpp:650-667 */
        uint32_t num_ranges = 0;
        for (uint32_t t = 0; t < L; t++) {
            uint32_t rs, re;
            next_range(ix, t, &qo[t], &rs, &re);
            ranges[num_ranges].cur = ix->indices + t * tl + rs;
            ranges[num_ranges].end = ix->indices + t * tl + re;
            ranges[num_ranges].table = t;
            num_ranges += (rs != re);
        }
        ranges[num_ranges].cur = filler;
        ranges[num_ranges].end = filler + 2 * 32 * 4;
        ranges[num_ranges].table = L;

        const uint32_t* ring[32];
        uint32_t range_idx = 0;
        int32_t missing = 32;
        for (int i = 0; i < 32; i++) { /* collection.hpp:802-808 */
            range_t* rg = &ranges[range_idx];
            ring[i] = rg->cur;
            rg->cur += 4;
            missing -= (range_idx < num_ranges);
            range_idx += (rg->cur == rg->end);
        }
        while (range_idx < num_ranges) { /* collection.hpp:810 */
            uint32_t np = 0;
            while (np < 128 && missing == 0) { /* :813-866 */
                for (int ri = 0; ri < 32; ri++) {
                    const uint32_t* seg = ring[ri];
                    for (int j = 0; j < 4; j++) {
                        uint32_t v = seg[j];
                        uint64_t s = ix->sketches[((size_t)v << 5) | (uint32_t)ri];
                        passing[np] = v;
                        np += passes(qs, max_diff, s, ri);
                    }
                    missing += (range_idx >= num_ranges);
                    range_t* rg = &ranges[range_idx];
                    ring[ri] = rg->cur;
                    rg->cur += 4;
                    range_idx += (rg->cur == rg->end);
                }
                tr->candidates += 32 * 4;
            }
            for (int ri = 32 - 1 - missing; ri >= 0; ri--) { /* :869-903 — the index itself is tested as a sketch */
                const uint32_t* seg = ring[ri];
                for (int j = 0; j < 4; j++) {
                    uint32_t v = seg[j];
                    passing[np] = v;
                    np += passes(qs, max_diff, (uint64_t)v, ri);
                }
            }
            tr->candidates += 4 * (32 - missing);
            /* empty the buffer, :909-925 */
            for (uint32_t p = 0; p < np; p++) {
                uint32_t idx = passing[p];
                float sim = orc_similarity(q, ix->q15 + (size_t)idx * fn->sl, fn->sl);
                mb_insert(&mb, idx, sim);
            }
            if (tr->passing) {
                for (uint32_t p = 0; p < np && tr->passing_len < tr->passing_cap; p++) tr->passing[tr->passing_len++] = passing[p];
            }
            if (tr->batch_sizes && tr->n_batches < tr->batch_cap) tr->batch_sizes[tr->n_batches] = np;
            tr->n_batches++;
            tr->distance_computations += np;
            float kth = mb.minval;
            max_diff = orc_max_sketch_diff(kth);
            /* stop rule, :927-943 */
            uint32_t table_idx = ranges[range_idx].table;
            uint32_t last_tables = (depth == ORC_MAX_HASHBITS) ? table_idx : L;
            float sim = kth > max_sim ? kth : max_sim; /* std::max(kth, max_sim) */
            float fp = orc_failure_probability(fn, depth, table_idx, last_tables, sim);
            if (fp <= 1 - recall) {
                tr->stop_depth = depth;
                tr->stop_table = table_idx;
                stopped = 1;
                break;
            }
        }
    }
    tr->kth = mb.minval;
    tr->max_sketch_diff = max_diff;
    mb_filter(&mb); /* best_indices, collection.hpp:598 */
    for (uint32_t i = 0; i < mb.inserted; i++) out[i] = mb.data[i].idx;
    int r = (int)mb.inserted;
    free(mb.data); free(q); free(codes); free(qo); free(ranges);
    return r;
}

/* collection.hpp:671-765 — search_maps_no_filter (filter_type 1) and search_maps_simple_filter (filter_type 2), driven by
 * search_formatted_query (:543-601). No ring, no passing buffer, no max_sim: every entry of every non-empty range, in table
 * order, goes to MaxBuffer::insert (after the sketch test of slot `range_idx % 32` in the Simple variant, whose threshold is
 * refreshed after each range, :739-740); one stop test per depth with table_idx = last_tables = L. Neither variant touches the
 * distance_computations / candidates counters (only :865,904,921 do). tr->stop_depth = hash_length (:704,758). */
int orc_index_search_filter(const orc_index* ix, const float* qf, uint32_t k, float recall, float max_sim, int filter_type,
                            uint32_t* out, orc_trace* tr) {
    if (filter_type != 1 && filter_type != 2) return orc_index_search(ix, qf, k, recall, max_sim, out, tr);
    const orc_functions* fn = &ix->fn;
    uint32_t L = fn->L;
    int16_t* q = (int16_t*)malloc(sizeof(int16_t) * fn->sl);
    orc_store_q15(qf, fn->d, fn->sl, q);
    orc_trace local;
    if (!tr) {
        memset(&local, 0, sizeof(local));
        tr = &local;
    }
    tr->distance_computations = tr->candidates = tr->stop_depth = tr->stop_table = tr->n_batches = 0;
    tr->passing_len = 0;
    tr->kth = 0;
    tr->max_sketch_diff = ORC_SKETCH_BITS;
    if (ix->n < 100) { /* collection.hpp:550-555, before the switch on the filter type */
        int r = search_bf(ix, q, k, out);
        free(q);
        return r;
    }
    uint32_t* codes = (uint32_t*)malloc(sizeof(uint32_t) * L);
    uint64_t qs[ORC_NUM_SKETCHES];
    orc_codes(fn, q, codes);
    orc_sketch(fn, q, qs);
    uint32_t max_diff = ORC_SKETCH_BITS;
    maxbuffer mb;
    mb_init(&mb, k);
    size_t tl = (size_t)ix->n + 2 * ORC_SEGMENT;
    pm_query* qo = (pm_query*)malloc(sizeof(pm_query) * L);
    for (uint32_t t = 0; t < L; t++) {
        uint32_t a = orc_anchor(ix, t, codes[t]);
        qo[t].hash = codes[t];
        qo[t].mask = 0xffffffffu;
        qo[t].start = qo[t].end = a;
    }
    for (uint32_t depth = ORC_MAX_HASHBITS; depth > 0; depth--) {
        uint32_t range_idx = 0; /* position among the non-empty ranges (fill_ranges skips the empty ones, :660) */
        for (uint32_t t = 0; t < L; t++) {
            uint32_t rs, re;
            next_range(ix, t, &qo[t], &rs, &re);
            if (rs == re) continue;
            uint32_t slot = range_idx % ORC_NUM_SKETCHES;
            for (uint32_t pos = rs; pos < re; pos++) {
                uint32_t idx = ix->indices[t * tl + pos];
                if (filter_type == 2 && !passes(qs, max_diff, ix->sketches[((size_t)idx << 5) | slot], slot)) continue;
                mb_insert(&mb, idx, orc_similarity(q, ix->q15 + (size_t)idx * fn->sl, fn->sl));
            }
            if (filter_type == 2) max_diff = orc_max_sketch_diff(mb.minval);
            range_idx++;
        }
        float fp = orc_failure_probability(fn, depth, L, L, mb.minval);
        if (fp <= 1 - recall) {
            tr->stop_depth = depth;
            tr->stop_table = L;
            break;
        }
    }
    tr->kth = mb.minval;
    tr->max_sketch_diff = max_diff;
    mb_filter(&mb);
    for (uint32_t i = 0; i < mb.inserted; i++) out[i] = mb.data[i].idx;
    int r = (int)mb.inserted;
    free(mb.data); free(q); free(codes); free(qo);
    return r;
}

/* ------------------------------------------------------------------------------------------------ L3 (CLANN) */

/* index.rs:78-80 */
uint64_t orc_num_clusters(float factor, uint64_t n) {
    double v = floor((double)factor * sqrt((double)n));
    uint64_t k = (uint64_t)v;
    return k < 1 ? 1 : k;
}

/* ndarray 0.16.1 numeric_util::unrolled_dot (dependency pinned in Cargo.lock, not vendored): eight partial sums over
 * chunks of 8, combined as (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7) into sum, then the <8 tail added in order. No FMA (Rust). */
float orc_ndarray_dot(const float* x, const float* y, size_t len) {
    float p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
    size_t i = 0;
    for (; i + 8 <= len; i += 8) {
        p0 = p0 + x[i + 0] * y[i + 0];
        p1 = p1 + x[i + 1] * y[i + 1];
        p2 = p2 + x[i + 2] * y[i + 2];
        p3 = p3 + x[i + 3] * y[i + 3];
        p4 = p4 + x[i + 4] * y[i + 4];
        p5 = p5 + x[i + 5] * y[i + 5];
        p6 = p6 + x[i + 6] * y[i + 6];
        p7 = p7 + x[i + 7] * y[i + 7];
    }
    float sum = 0;
    sum = sum + (p0 + p4);
    sum = sum + (p1 + p5);
    sum = sum + (p2 + p6);
    sum = sum + (p3 + p7);
    for (; i < len; i++) sum = sum + x[i] * y[i];
    return sum;
}

/* angulardata.rs:29-35 */
float orc_distance_point(const float* row, float row_norm, const float* q, uint32_t d) {
    float dot = orc_ndarray_dot(row, q, d);
    float s = 0;
    for (uint32_t i = 0; i < d; i++) s = s + q[i] * q[i];
    float nq = sqrtf(s);
    return 1.0f - dot / (row_norm * nq);
}

/* gmm.rs:21-62 over angulardata.rs:12-43 */
uint64_t orc_gmm(const float* data, uint64_t n, uint32_t d, uint64_t K, uint64_t* centers, uint64_t* assignment, float* radii) {
    if (n <= K) {
        for (uint64_t i = 0; i < n; i++) {
            centers[i] = i;
            assignment[i] = i;
            radii[i] = 0;
        }
        return n;
    }
    float* norms = (float*)malloc(sizeof(float) * n);
    float* dist = (float*)malloc(sizeof(float) * n);
    float* nd = (float*)malloc(sizeof(float) * n);
    for (uint64_t i = 0; i < n; i++) norms[i] = sqrtf(orc_ndarray_dot(data + i * d, data + i * d, d));
    for (uint64_t i = 0; i < n; i++) assignment[i] = 0;
    centers[0] = 0;
    for (uint64_t i = 0; i < n; i++) dist[i] = 1.0f - (orc_ndarray_dot(data + i * d, data, d) / (norms[i] * norms[0]));
    for (uint64_t c = 1; c < K; c++) {
        uint64_t far = 0;
        float m = dist[0];
        for (uint64_t i = 1; i < n; i++)
            if (dist[i] > m) {
                far = i;
                m = dist[i];
            }
        centers[c] = far;
        for (uint64_t i = 0; i < n; i++) nd[i] = 1.0f - (orc_ndarray_dot(data + i * d, data + far * d, d) / (norms[i] * norms[far]));
        for (uint64_t i = 0; i < n; i++)
            if (nd[i] < dist[i]) {
                assignment[i] = c;
                dist[i] = nd[i];
            }
    }
    for (uint64_t c = 0; c < K; c++) radii[c] = 0;
    for (uint64_t i = 0; i < n; i++) {
        float r = radii[assignment[i]];
        radii[assignment[i]] = dist[i] > r ? dist[i] : r; /* f32::max: a NaN distance (zero vector) is ignored, r is never NaN */
    }
    free(norms); free(dist); free(nd);
    return K;
}

/* heap.rs:5-49 as a bounded array: "max" = greatest (distance, index) pair (derived Ord). */
typedef struct {
    float dist;
    uint64_t idx;
} hp_elem;
typedef struct {
    hp_elem* e;
    uint32_t len, cap;
} topk;
/* ordered-float 4.6.0 total order (Cargo.lock): NaN is greater than every number and equal to itself */
static int of_cmp(float a, float b) {
    int an = a != a, bn = b != b;
    if (an || bn) return an - bn;
    return (a > b) - (a < b);
}
static int hp_less(hp_elem a, hp_elem b) {
    int c = of_cmp(a.dist, b.dist);
    return c < 0 || (c == 0 && a.idx < b.idx);
}
static uint32_t hp_max(const topk* h) {
    uint32_t m = 0;
    for (uint32_t i = 1; i < h->len; i++)
        if (hp_less(h->e[m], h->e[i])) m = i;
    return m;
}
static int hp_add(topk* h, hp_elem x) { /* heap.rs:23-36 */
    if (h->len < h->cap) {
        h->e[h->len++] = x;
        return 1;
    }
    if (h->len == 0) return 1; /* k == 0: `else if let Some(max)` is skipped and true is returned */
    uint32_t m = hp_max(h);
    if (of_cmp(x.dist, h->e[m].dist) < 0) {
        h->e[m] = x;
        return 1;
    }
    return 0;
}
static int hp_cmp(const void* a, const void* b) {
    const hp_elem* x = (const hp_elem*)a;
    const hp_elem* y = (const hp_elem*)b;
    int c = of_cmp(x->dist, y->dist); /* to_list's partial_cmp().unwrap() would panic on a NaN; here NaN sorts last */
    if (c) return c;
    if (x->idx < y->idx) return -1; /* tie order inside std's BinaryHeap::iter() is unspecified; ids ascending here */
    if (x->idx > y->idx) return 1;
    return 0;
}

typedef struct {
    uint64_t ci;
    float dist;
} cd_t;

/* Drives the TopKClosestHeap restatement for the reference's own unit tests (heap.rs:52-161): n adds, then to_list.
 * added[i] = return value of add; top_* = get_top() after all adds (id, distance), valid if the count is > 0. */
int orc_topk_run(uint32_t k, const float* dists, const uint64_t* ids, int n, float* out_d, uint64_t* out_ids, uint8_t* added,
                 uint64_t* top_id, float* top_dist) {
    topk h = {(hp_elem*)malloc(sizeof(hp_elem) * (k + 1)), 0, k};
    for (int i = 0; i < n; i++) {
        hp_elem e = {dists[i], ids[i]};
        added[i] = (uint8_t)hp_add(&h, e);
    }
    if (h.len > 0) {
        hp_elem t = h.e[hp_max(&h)];
        *top_id = t.idx;
        *top_dist = t.dist;
    }
    qsort(h.e, h.len, sizeof(hp_elem), hp_cmp);
    for (uint32_t i = 0; i < h.len; i++) {
        out_d[i] = h.e[i].dist;
        out_ids[i] = h.e[i].idx;
    }
    int r = (int)h.len;
    free(h.e);
    return r;
}

/* index.rs:592-616 on its own: stable ascending order of the centres by distance_point to the query. */
void orc_sort_clusters(const float* data, uint32_t d, const uint64_t* centers, uint64_t K, const float* q, uint64_t* order) {
    cd_t* cd = (cd_t*)malloc(sizeof(cd_t) * K);
    for (uint64_t ci = 0; ci < K; ci++) {
        const float* row = data + centers[ci] * d;
        cd[ci].ci = ci;
        cd[ci].dist = orc_distance_point(row, sqrtf(orc_ndarray_dot(row, row, d)), q, d);
    }
    for (uint64_t i = 1; i < K; i++) {
        cd_t x = cd[i];
        uint64_t j = i;
        while (j > 0 && x.dist < cd[j - 1].dist) {
            cd[j] = cd[j - 1];
            j--;
        }
        cd[j] = x;
    }
    for (uint64_t i = 0; i < K; i++) order[i] = cd[i].ci;
    free(cd);
}

struct orc_clann {
    const float* data;
    uint64_t n, K;
    uint32_t d, k;
    float delta;
    float* norms;
    uint64_t* centers;
    float* radii;
    uint64_t** members;
    uint64_t* sizes;
    uint8_t* brute;
    orc_index** indices;
};

orc_clann* orc_clann_create(const float* data, uint64_t n, uint32_t d, uint32_t k, float delta, uint64_t K,
                            const uint64_t* centers, const uint64_t* assignment, const float* radii) {
    orc_clann* c = (orc_clann*)calloc(1, sizeof(orc_clann));
    c->data = data; c->n = n; c->K = K; c->d = d; c->k = k; c->delta = delta;
    c->norms = (float*)malloc(sizeof(float) * n);
    for (uint64_t i = 0; i < n; i++) c->norms[i] = sqrtf(orc_ndarray_dot(data + i * d, data + i * d, d));
    c->centers = (uint64_t*)malloc(sizeof(uint64_t) * K);
    c->radii = (float*)malloc(sizeof(float) * K);
    memcpy(c->centers, centers, sizeof(uint64_t) * K);
    memcpy(c->radii, radii, sizeof(float) * K);
    c->sizes = (uint64_t*)calloc(K, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) c->sizes[assignment[i]]++;
    c->members = (uint64_t**)calloc(K, sizeof(uint64_t*));
    for (uint64_t ci = 0; ci < K; ci++) c->members[ci] = (uint64_t*)malloc(sizeof(uint64_t) * (c->sizes[ci] ? c->sizes[ci] : 1));
    uint64_t* fill = (uint64_t*)calloc(K, sizeof(uint64_t));
    for (uint64_t i = 0; i < n; i++) c->members[assignment[i]][fill[assignment[i]]++] = i; /* index.rs:188-192 */
    free(fill);
    c->brute = (uint8_t*)malloc(K);
    for (uint64_t ci = 0; ci < K; ci++) c->brute[ci] = c->sizes[ci] < 100 || c->sizes[ci] < k; /* index.rs:204-205 */
    c->indices = (orc_index**)calloc(K, sizeof(orc_index*));
    return c;
}

int orc_clann_set_cluster_stream(orc_clann* c, uint64_t ci, const uint8_t* stream, uint64_t len) {
    if (ci >= c->K) return -1;
    orc_index* ix = orc_index_import(stream, len);
    if (!ix || ix->n != c->sizes[ci]) {
        orc_index_free(ix);
        return -2;
    }
    orc_index_free(c->indices[ci]);
    c->indices[ci] = ix;
    return 0;
}

/* index.rs:259-262 -> puffinn.rs:15-59: subset rows in assignment order, insert, rebuild */
int orc_clann_build_cluster(orc_clann* c, uint64_t ci, const orc_functions* fn) {
    if (ci >= c->K || c->brute[ci]) return -1;
    uint64_t nc = c->sizes[ci];
    float* sub = (float*)malloc(sizeof(float) * nc * c->d);
    for (uint64_t j = 0; j < nc; j++) memcpy(sub + j * c->d, c->data + c->members[ci][j] * c->d, sizeof(float) * c->d);
    orc_index_free(c->indices[ci]);
    c->indices[ci] = orc_index_build(fn, sub, (uint32_t)nc);
    free(sub);
    return 0;
}

/* index.rs:311-439. visit_log (optional, visit_cap rows of 3): what the reference's RunMetrics records per visited cluster
 * (index.rs:426-431 -> metrics/mod.rs:84-112 -> sqlite.rs:248-283): {cluster, points_added = heap adds that returned true,
 * cluster_distance_computations = the prune-test evaluation (index.rs:348) + PUFFINN's counter or the brute-force list length}.
 * The visit that ends the walk at the prune test pushes no n_candidates, so the zip of sqlite.rs:248-253 drops it: not logged. */
int orc_clann_search_visits(orc_clann* c, const float* q, uint64_t* out_ids, float* out_dists, uint64_t* order_out, uint64_t* counters,
                            uint64_t* visit_log, uint64_t visit_cap) {
    uint64_t K = c->K;
    cd_t* cd = (cd_t*)malloc(sizeof(cd_t) * K);
    for (uint64_t ci = 0; ci < K; ci++) { /* index.rs:592-600 */
        cd[ci].ci = ci;
        cd[ci].dist = orc_distance_point(c->data + c->centers[ci] * c->d, c->norms[c->centers[ci]], q, c->d);
    }
    /* stable ascending sort (slice::sort_by), index.rs:609-613: insertion sort keeps it obviously stable */
    for (uint64_t i = 1; i < K; i++) {
        cd_t x = cd[i];
        uint64_t j = i;
        while (j > 0 && x.dist < cd[j - 1].dist) {
            cd[j] = cd[j - 1];
            j--;
        }
        cd[j] = x;
    }
    if (order_out)
        for (uint64_t i = 0; i < K; i++) order_out[i] = cd[i].ci;
    uint64_t visited = 0, distcomp = 0, cands = 0;
    topk pq = {(hp_elem*)malloc(sizeof(hp_elem) * (c->k + 1)), 0, c->k};
    topk loc = {(hp_elem*)malloc(sizeof(hp_elem) * (c->k + 1)), 0, c->k};
    uint32_t* res = (uint32_t*)malloc(sizeof(uint32_t) * (c->k + 1));
    float max_dist = INFINITY;
    int err = 0;
    for (uint64_t oi = 0; oi < K; oi++) {
        uint64_t ci = cd[oi].ci;
        uint64_t visit_dc = 0, points_added = 0;
        if (pq.len > 0) { /* index.rs:342-361 */
            hp_elem top = pq.e[hp_max(&pq)];
            max_dist = top.dist;
            visit_dc = 1; /* index.rs:348 */
            float cmin = orc_distance_point(c->data + c->centers[ci] * c->d, c->norms[c->centers[ci]], q, c->d) - c->radii[ci];
            if (cmin > top.dist) break;
        }
        visited++;
        if (c->brute[ci]) { /* index.rs:364-378, 666-685 */
            loc.len = 0;
            for (uint64_t j = 0; j < c->sizes[ci]; j++) {
                uint64_t p = c->members[ci][j];
                hp_elem e = {orc_distance_point(c->data + p * c->d, c->norms[p], q, c->d), p};
                hp_add(&loc, e);
            }
            qsort(loc.e, loc.len, sizeof(hp_elem), hp_cmp);
            for (uint32_t j = 0; j < loc.len; j++) points_added += (uint64_t)hp_add(&pq, loc.e[j]);
            visit_dc += loc.len; /* index.rs:378 */
            /* counters follow PUFFINN's own (performance.hpp:72-86): brute-force clusters add nothing */
        } else {
            if (!c->indices[ci]) { /* index.rs:386-388 IndexNotFound */
                err = -1;
                break;
            }
            float max_sim = 1.0f - max_dist / 2.0f; /* puffinn_types.rs:77-79 */
            orc_trace tr;
            memset(&tr, 0, sizeof(tr));
            int nr = orc_index_search(c->indices[ci], q, c->k, c->delta, max_sim, res, &tr);
            distcomp += tr.distance_computations;
            cands += tr.candidates;
            visit_dc += tr.distance_computations; /* index.rs:421 */
            for (int j = 0; j < nr; j++) { /* index.rs:392-416 */
                uint64_t p = c->members[ci][res[j]];
                hp_elem e = {orc_distance_point(c->data + p * c->d, c->norms[p], q, c->d), p};
                points_added += (uint64_t)hp_add(&pq, e);
            }
        }
        if (visit_log && visited <= visit_cap) {
            visit_log[(visited - 1) * 3 + 0] = ci;
            visit_log[(visited - 1) * 3 + 1] = points_added;
            visit_log[(visited - 1) * 3 + 2] = visit_dc;
        }
    }
    qsort(pq.e, pq.len, sizeof(hp_elem), hp_cmp); /* heap.rs:42-48 */
    for (uint32_t i = 0; i < pq.len; i++) {
        out_ids[i] = pq.e[i].idx;
        out_dists[i] = pq.e[i].dist;
    }
    int r = err ? err : (int)pq.len;
    if (counters) {
        counters[0] = visited;
        counters[1] = distcomp;
        counters[2] = cands;
    }
    free(cd); free(pq.e); free(loc.e); free(res);
    return r;
}

int orc_clann_search(orc_clann* c, const float* q, uint64_t* out_ids, float* out_dists, uint64_t* order_out, uint64_t* counters) {
    return orc_clann_search_visits(c, q, out_ids, out_dists, order_out, counters, NULL, 0);
}

void orc_clann_free(orc_clann* c) {
    if (!c) return;
    for (uint64_t ci = 0; ci < c->K; ci++) {
        free(c->members[ci]);
        orc_index_free(c->indices[ci]);
    }
    free(c->members); free(c->sizes); free(c->brute); free(c->indices);
    free(c->norms); free(c->centers); free(c->radii);
    free(c);
}
