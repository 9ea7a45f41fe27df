/* CPU oracle, C restatement of the reference's Java CPU algorithms for the hot path (BN254a G1 / Fr).
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (octopuszk_b200/, liboctozk.so) never links or calls it.
 *
 * It exists because (a) the Python oracle (oracle/dizk_oracle.py) is too slow beyond ~2^12 elements and (b) bench.py
 * needs the reference's CPU path timed on the box's host cores.  The Java itself cannot run here (no JVM), so this is
 * a "port": the same algorithms, in the same order of group operations, on 4x64-bit Montgomery limbs instead of
 * java.math.BigInteger (a JVM would be several times slower).  Parity status: validated against the Python oracle
 * (tests/test_oracle_c.py), which is pinned by the reference's own known-answer tests (tests/test_oracle.py).
 *
 * Restated (paths relative to the reference root, Java under src/main/java/):
 *   fq/fr arithmetic       algebra/fields/Fp.java:38-49 (add/sub/mul reduce after every operation)
 *   g1_add / g1_dbl        algebra/curves/barreto_naehrig/BNG1.java:38-97 (add-2007-bl + O and P==Q checks), :133-161 (dbl-2009-l)
 *   pippenger_msm          algebra/msm/VariableBaseMSM.java:134-188 (c = L - L/3, unsigned windows from the top, bucket 0
 *                          skipped, running-sum reduce, c doublings between windows); numBits = 254 as the native path
 *                          hard-codes (algebra_msm_VariableBaseMSM.cu:1267)
 *   threads                the reference's partition-then-reduce(add) structure: VariableBaseMSM.distributedMSM,
 *                          VariableBaseMSM.java:772-786 (mapPartitions + reduce)
 *   serial_radix2_fft      algebra/fft/FFTAuxiliary.java:100-123 (bit-reverse swap, log n DIT stages, w *= w_m)
 *   fixed-base             algebra/msm/FixedBaseMSM.java:71-99 (getWindowTable), :141-167 (serialMSM)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t v[4]; } fe;   /* field element, Montgomery form unless noted */

typedef struct {
    uint64_t m[4];
    uint64_t np;      /* -m^-1 mod 2^64 */
    fe one, rr;       /* 2^256 mod m, 2^512 mod m */
} field_t;

static field_t FQ, FR;

static int fe_is_zero(const fe* a) { return (a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0; }
static int fe_eq(const fe* a, const fe* b) { return a->v[0] == b->v[0] && a->v[1] == b->v[1] && a->v[2] == b->v[2] && a->v[3] == b->v[3]; }

static int geq(const uint64_t* a, const uint64_t* m) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static void sub_n(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - (uint64_t)br;
        r[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}
static void fe_add(const field_t* F, fe* r, const fe* a, const fe* b) {
    u128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a->v[i] + b->v[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq(t, F->m)) sub_n(r->v, t, F->m);
    else memcpy(r->v, t, 32);
}
static void fe_sub(const field_t* F, fe* r, const fe* a, const fe* b) {
    u128 br = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->v[i] - b->v[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + F->m[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->v, t, 32);
}
/* CIOS Montgomery multiplication */
static void fe_mul(const field_t* F, fe* r, const fe* a, const fe* b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->v[j] * b->v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->np;
        c = (u128)m * F->m[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->m[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq(t, F->m)) sub_n(r->v, t, F->m);
    else memcpy(r->v, t, 32);
}
static void fe_sqr(const field_t* F, fe* r, const fe* a) { fe_mul(F, r, a, a); }
static void fe_to_mont(const field_t* F, fe* r, const fe* a) { fe_mul(F, r, a, &F->rr); }
static void fe_from_mont(const field_t* F, fe* r, const fe* a) {
    fe o = {{1, 0, 0, 0}};
    fe_mul(F, r, a, &o);
}
static void fe_pow(const field_t* F, fe* r, const fe* a, const uint64_t e[4]) {
    fe acc = F->one, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) fe_mul(F, &acc, &acc, &base);
        fe_sqr(F, &base, &base);
    }
    *r = acc;
}
static void fe_inv(const field_t* F, fe* r, const fe* a) {      /* Fp.inverse (Fp.java:88-90) via Fermat */
    uint64_t e[4];
    memcpy(e, F->m, 32);
    e[0] -= 2;
    fe_pow(F, r, a, e);
}

static void field_init(field_t* F, const uint64_t m[4]) {
    memcpy(F->m, m, 32);
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - m[0] * inv;      /* Newton: m^-1 mod 2^64 */
    F->np = (uint64_t)0 - inv;
    /* 2^256 mod m by 256 modular doublings of 1, then 2^512 by 256 more */
    uint64_t x[4] = {1, 0, 0, 0};
    for (int k = 0; k < 512; k++) {
        uint64_t c = x[3] >> 63;
        x[3] = (x[3] << 1) | (x[2] >> 63);
        x[2] = (x[2] << 1) | (x[1] >> 63);
        x[1] = (x[1] << 1) | (x[0] >> 63);
        x[0] <<= 1;
        if (c || geq(x, m)) sub_n(x, x, m);
        if (k == 255) memcpy(F->one.v, x, 32);
    }
    memcpy(F->rr.v, x, 32);
}

static int g_init = 0;
static void init_once(void) {
    if (g_init) return;
    /* BN254aFqParameters.java:33, BN254aFrParameters.java:33 */
    static const uint64_t P[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const uint64_t R[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    field_init(&FQ, P);
    field_init(&FR, R);
    g_init = 1;
}

/* ---- G1, Jacobian ---------------------------------------------------------------------------------------------- */
typedef struct { fe x, y, z; } g1;

static void g1_zero(g1* r) {      /* (0, 1, 0): BN254aG1Parameters.java:52-55 */
    memset(r, 0, sizeof *r);
    r->y = FQ.one;
}
static int g1_is_zero(const g1* p) { return fe_is_zero(&p->z); }

static void g1_dbl(g1* r, const g1* p) {      /* BNG1.java:133-161 */
    if (g1_is_zero(p)) { *r = *p; return; }
    const field_t* F = &FQ;
    fe A, B, C, D, E, Fv, X3, Y3, Z3, t, eightC;
    fe_sqr(F, &A, &p->x);
    fe_sqr(F, &B, &p->y);
    fe_sqr(F, &C, &B);
    fe_add(F, &t, &p->x, &B);
    fe_sqr(F, &t, &t);
    fe_sub(F, &t, &t, &A);
    fe_sub(F, &D, &t, &C);
    fe_add(F, &D, &D, &D);
    fe_add(F, &E, &A, &A);
    fe_add(F, &E, &E, &A);
    fe_sqr(F, &Fv, &E);
    fe_add(F, &t, &D, &D);
    fe_sub(F, &X3, &Fv, &t);
    fe_add(F, &eightC, &C, &C);
    fe_add(F, &eightC, &eightC, &eightC);
    fe_add(F, &eightC, &eightC, &eightC);
    fe_sub(F, &t, &D, &X3);
    fe_mul(F, &t, &E, &t);
    fe_sub(F, &Y3, &t, &eightC);
    fe_mul(F, &t, &p->y, &p->z);
    fe_add(F, &Z3, &t, &t);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void g1_add(g1* r, const g1* p, const g1* q) {      /* BNG1.java:38-97 */
    if (g1_is_zero(p)) { *r = *q; return; }
    if (g1_is_zero(q)) { *r = *p; return; }
    const field_t* F = &FQ;
    fe Z1Z1, Z2Z2, U1, U2, Z1c, Z2c, S1, S2;
    fe_sqr(F, &Z1Z1, &p->z);
    fe_sqr(F, &Z2Z2, &q->z);
    fe_mul(F, &U1, &p->x, &Z2Z2);
    fe_mul(F, &U2, &q->x, &Z1Z1);
    fe_mul(F, &Z1c, &p->z, &Z1Z1);
    fe_mul(F, &Z2c, &q->z, &Z2Z2);
    fe_mul(F, &S1, &p->y, &Z2c);
    fe_mul(F, &S2, &q->y, &Z1c);
    if (fe_eq(&U1, &U2) && fe_eq(&S1, &S2)) { g1_dbl(r, p); return; }
    fe H, S2mS1, I, J, rr, V, X3, Y3, Z3, t, u;
    fe_sub(F, &H, &U2, &U1);
    fe_sub(F, &S2mS1, &S2, &S1);
    fe_add(F, &t, &H, &H);
    fe_sqr(F, &I, &t);
    fe_mul(F, &J, &H, &I);
    fe_add(F, &rr, &S2mS1, &S2mS1);
    fe_mul(F, &V, &U1, &I);
    fe_sqr(F, &t, &rr);
    fe_sub(F, &t, &t, &J);
    fe_add(F, &u, &V, &V);
    fe_sub(F, &X3, &t, &u);
    fe_mul(F, &t, &S1, &J);
    fe_add(F, &t, &t, &t);
    fe_sub(F, &u, &V, &X3);
    fe_mul(F, &u, &rr, &u);
    fe_sub(F, &Y3, &u, &t);
    fe_add(F, &t, &p->z, &q->z);
    fe_sqr(F, &t, &t);
    fe_sub(F, &t, &t, &Z1Z1);
    fe_sub(F, &t, &t, &Z2Z2);
    fe_mul(F, &Z3, &t, &H);
    r->x = X3; r->y = Y3; r->z = Z3;
}

static void load_fe(const field_t* F, fe* r, const uint8_t* b) {
    fe c;
    memcpy(c.v, b, 32);
    fe_to_mont(F, r, &c);
}
static void store_fe(const field_t* F, uint8_t* b, const fe* a) {
    fe c;
    fe_from_mont(F, &c, a);
    memcpy(b, c.v, 32);
}
static void load_g1(g1* r, const uint8_t* b) {
    load_fe(&FQ, &r->x, b);
    load_fe(&FQ, &r->y, b + 32);
    load_fe(&FQ, &r->z, b + 64);
}
static void store_g1(uint8_t* b, const g1* p) {
    store_fe(&FQ, b, &p->x);
    store_fe(&FQ, b + 32, &p->y);
    store_fe(&FQ, b + 64, &p->z);
}

static int java_log2(size_t x) { return (int)(log((double)x) / log(2.0)); }      /* MathUtils.java:12-14 */

static unsigned scalar_window(const uint64_t s[4], unsigned pos, unsigned c) {
    unsigned id = 0;
    for (unsigned j = 0; j < c; j++) {
        unsigned bit = pos + j;
        if (bit < 256 && ((s[bit >> 6] >> (bit & 63)) & 1)) id |= 1u << j;
    }
    return id;
}

/* VariableBaseMSM.pippengerMSM (VariableBaseMSM.java:134-188) on one slice */
static void pippenger_slice(g1* result, const uint8_t* scalars, const g1* bases, size_t length, int num_bits) {
    int L = java_log2(length);
    if (L < 1) L = 1;
    const unsigned c = (unsigned)(L - L / 3);
    const size_t num_buckets = (size_t)1 << c;
    const int num_groups = (num_bits + (int)c - 1) / (int)c;
    g1* buckets = (g1*)malloc(num_buckets * sizeof(g1));
    g1 res, zero;
    g1_zero(&zero);
    res = zero;
    for (int k = num_groups - 1; k >= 0; k--) {
        for (size_t i = 0; i < num_buckets; i++) buckets[i] = zero;
        for (size_t i = 0; i < length; i++) {
            uint64_t s[4];
            memcpy(s, scalars + 32 * i, 32);
            unsigned id = scalar_window(s, (unsigned)k * c, c);
            if (id == 0) continue;
            g1_add(&buckets[id], &buckets[id], &bases[i]);
        }
        g1 running = zero;
        for (size_t i = num_buckets - 1; i > 0; i--) {
            g1_add(&running, &running, &buckets[i]);
            g1_add(&res, &res, &running);
        }
        if (k > 0)
            for (unsigned i = 0; i < c; i++) g1_dbl(&res, &res);
    }
    free(buckets);
    *result = res;
}

/* out = sum scalars[i] * bases[i]; wire formats of include/octozk.h.  threads > 1: contiguous partitions, each
 * reduced by the serial algorithm, partial results added (VariableBaseMSM.java:772-786). */
int oracle_msm_g1(const uint8_t* scalars, const uint8_t* bases, size_t n, int threads, uint8_t out[96]) {
    init_once();
    g1 total;
    g1_zero(&total);
    if (n == 0) { store_g1(out, &total); return 0; }
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)n;
    g1* pts = (g1*)malloc(n * sizeof(g1));
    g1* partial = (g1*)malloc((size_t)threads * sizeof(g1));
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long i = 0; i < (long)n; i++) load_g1(&pts[i], bases + 96 * (size_t)i);
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; t++) {
        size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
        pippenger_slice(&partial[t], scalars + 32 * lo, pts + lo, hi - lo, 254);
    }
    for (int t = 0; t < threads; t++) g1_add(&total, &total, &partial[t]);
    store_g1(out, &total);
    free(pts);
    free(partial);
    return 0;
}

/* projective equality, BNG1.equals (BNG1.java:191-224) */
int oracle_g1_equal(const uint8_t a[96], const uint8_t b[96]) {
    init_once();
    g1 p, q;
    load_g1(&p, a);
    load_g1(&q, b);
    if (g1_is_zero(&p)) return g1_is_zero(&q);
    if (g1_is_zero(&q)) return 0;
    fe z1s, z2s, l, r, z1c, z2c;
    fe_sqr(&FQ, &z1s, &p.z);
    fe_sqr(&FQ, &z2s, &q.z);
    fe_mul(&FQ, &l, &p.x, &z2s);
    fe_mul(&FQ, &r, &q.x, &z1s);
    if (!fe_eq(&l, &r)) return 0;
    fe_mul(&FQ, &z1c, &p.z, &z1s);
    fe_mul(&FQ, &z2c, &q.z, &z2s);
    fe_mul(&FQ, &l, &p.y, &z2c);
    fe_mul(&FQ, &r, &q.y, &z1c);
    return fe_eq(&l, &r);
}

/* BNG1.toAffineCoordinates (BNG1.java:163-172): (x, y, 1) or (0, 1, 0) */
int oracle_g1_to_affine(const uint8_t in[96], uint8_t out[96]) {
    init_once();
    g1 p;
    load_g1(&p, in);
    if (g1_is_zero(&p)) { g1_zero(&p); store_g1(out, &p); return 0; }
    fe zi, z2, z3;
    fe_inv(&FQ, &zi, &p.z);
    fe_sqr(&FQ, &z2, &zi);
    fe_mul(&FQ, &z3, &z2, &zi);
    fe_mul(&FQ, &p.x, &p.x, &z2);
    fe_mul(&FQ, &p.y, &p.y, &z3);
    p.z = FQ.one;
    store_g1(out, &p);
    return 0;
}

/* ---- FFT: FFTAuxiliary.serialRadix2FFT (FFTAuxiliary.java:100-123), in place, canonical 32-byte elements -------- */
static unsigned bitreverse(unsigned n, int bits) {      /* MathUtils.java:47-56 */
    int count = bits - 1;
    unsigned reverse = n;
    n >>= 1;
    while (n > 0) {
        reverse = (reverse << 1) | (n & 1);
        n >>= 1;
        count--;
    }
    return (reverse << count) & ((1u << bits) - 1);
}

int oracle_fft_fr(uint8_t* data, size_t n, const uint8_t omega_b[32]) {
    init_once();
    if (n <= 1) return 0;
    int logn = 0;
    while (((size_t)1 << logn) < n) logn++;
    if (((size_t)1 << logn) != n) return -1;
    fe* a = (fe*)malloc(n * sizeof(fe));
    for (size_t i = 0; i < n; i++) load_fe(&FR, &a[i], data + 32 * i);
    fe omega;
    load_fe(&FR, &omega, omega_b);
    for (size_t k = 0; k < n; k++) {
        size_t rk = bitreverse((unsigned)k, logn);
        if (k < rk) { fe t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    size_t m = 1;
    for (int s = 1; s <= logn; s++) {
        uint64_t e[4] = {n / (2 * m), 0, 0, 0};
        fe w_m;
        fe_pow(&FR, &w_m, &omega, e);
        for (size_t k = 0; k < n; k += 2 * m) {
            fe w = FR.one;
            for (size_t j = 0; j < m; j++) {
                fe t;
                fe_mul(&FR, &t, &w, &a[k + j + m]);
                fe_sub(&FR, &a[k + j + m], &a[k + j], &t);
                fe_add(&FR, &a[k + j], &a[k + j], &t);
                fe_mul(&FR, &w, &w, &w_m);
            }
        }
        m *= 2;
    }
    for (size_t i = 0; i < n; i++) store_fe(&FR, data + 32 * i, &a[i]);
    free(a);
    return 0;
}

/* `batch` independent transforms of length n (data = batch * n elements), one per thread: how the reference's
 * distributed FFT uses its cores (FFTAuxiliary.java:151-207 runs one SerialFFT per group per executor thread). */
int oracle_fft_fr_batch(uint8_t* data, size_t n, size_t batch, const uint8_t omega_b[32], int threads) {
    if (threads < 1) threads = 1;
    int rc = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long b = 0; b < (long)batch; b++) {
        int r = oracle_fft_fr(data + 32 * n * (size_t)b, n, omega_b);
        if (r) rc = r;
    }
    return rc;
}

/* ---- fixed base: FixedBaseMSM.getWindowTable (:71-99) + serialMSM (:141-167) -------------------------------------- */
int oracle_fixed_g1(const uint8_t base_b[96], const uint8_t* scalars, size_t n, int scalar_size, int window, int threads,
                    uint8_t* out) {
    init_once();
    if (window < 1 || window > 24 || scalar_size < 1) return -1;
    const int num_windows = (scalar_size % window == 0) ? scalar_size / window : scalar_size / window + 1;
    const size_t inner = (size_t)1 << window;
    g1 base;
    load_g1(&base, base_b);
    g1* table = (g1*)malloc((size_t)num_windows * inner * sizeof(g1));
    g1 base_outer = base;
    for (int o = 0; o < num_windows; o++) {
        g1 base_inner;
        g1_zero(&base_inner);
        for (size_t i = 0; i < inner; i++) {
            table[(size_t)o * inner + i] = base_inner;
            g1_add(&base_inner, &base_inner, &base_outer);
        }
        for (int w = 0; w < window; w++) g1_dbl(&base_outer, &base_outer);
    }
    const int outerc = (scalar_size + window - 1) / window;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long i = 0; i < (long)n; i++) {
        uint64_t s[4];
        memcpy(s, scalars + 32 * (size_t)i, 32);
        g1 res = table[0];
        for (int o = 0; o < outerc; o++) {
            unsigned in = scalar_window(s, (unsigned)(o * window), (unsigned)window);
            g1_add(&res, &res, &table[(size_t)o * inner + in]);
        }
        store_g1(out + 96 * (size_t)i, &res);
    }
    free(table);
    return 0;
}

/* out[i] = a[i] * b mod r : field_MSM (algebra_msm_FixedBaseMSM.cu:1241-1266) */
int oracle_fr_scale(const uint8_t* a, size_t n, const uint8_t b[32], uint8_t* out) {
    init_once();
    fe bb;
    load_fe(&FR, &bb, b);
    for (size_t i = 0; i < n; i++) {
        fe x;
        load_fe(&FR, &x, a + 32 * i);
        fe_mul(&FR, &x, &x, &bb);
        store_fe(&FR, out + 32 * i, &x);
    }
    return 0;
}

/* ---- exact size-independent checks (SURVEY.md section 8c "scale-parity technique") -------------------------------- */
/* out = sum_i a[i] * b[i] mod r.  With bases P_i = k_i * G the definition of the MSM (NaiveMSM.variableBaseMSM,
 * src/main/java/algebra/msm/NaiveMSM.java:21-46: sum of scalar.mul(base)) gives sum_i s_i P_i = (sum_i s_i k_i mod r) * G,
 * so this dot product is the whole expected answer at any size. */
int oracle_fr_dot(const uint8_t* a, const uint8_t* b, size_t n, int threads, uint8_t out[32]) {
    init_once();
    if (threads < 1) threads = 1;
    fe* part = (fe*)calloc((size_t)threads, sizeof(fe));
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; t++) {
        size_t lo = n * (size_t)t / (size_t)threads, hi = n * (size_t)(t + 1) / (size_t)threads;
        fe acc;
        memset(&acc, 0, sizeof acc);
        for (size_t i = lo; i < hi; i++) {
            fe x, y;
            load_fe(&FR, &x, a + 32 * i);
            load_fe(&FR, &y, b + 32 * i);
            fe_mul(&FR, &x, &x, &y);
            fe_add(&FR, &acc, &acc, &x);
        }
        part[t] = acc;
    }
    fe total;
    memset(&total, 0, sizeof total);
    for (int t = 0; t < threads; t++) fe_add(&FR, &total, &total, &part[t]);
    store_fe(&FR, out, &total);
    free(part);
    return 0;
}

/* out[q] = sum_j coeffs[j] * xs[q]^j mod r (Horner): naive evaluation, the definition the reference's own FFT test compares
 * with (src/test/java/algebra/fft/SerialFFTTest.java:168-190: FFT output k == polynomial evaluated at omega^k).
 * One point per thread at a time; O(n) each. */
int oracle_fr_horner(const uint8_t* coeffs, size_t n, const uint8_t* xs, size_t npts, int threads, uint8_t* out) {
    init_once();
    if (threads < 1) threads = 1;
    fe* c = (fe*)malloc((n ? n : 1) * sizeof(fe));
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long j = 0; j < (long)n; j++) load_fe(&FR, &c[j], coeffs + 32 * (size_t)j);
    /* every point's polynomial is cut into chunks so that all threads have work: P(x) = sum_k x^(k L) P_k(x) */
    size_t nchunks = npts ? ((size_t)threads * 2 + npts - 1) / npts : 1;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > n / 1024 + 1) nchunks = n / 1024 + 1;
    const size_t L = (n + nchunks - 1) / nchunks;
    fe* part = (fe*)malloc((npts * nchunks + 1) * sizeof(fe));
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (long item = 0; item < (long)(npts * nchunks); item++) {
        const size_t q = (size_t)item / nchunks, k = (size_t)item % nchunks;
        const size_t lo = k * L, hi = (lo + L < n) ? lo + L : n;
        fe x, acc;
        load_fe(&FR, &x, xs + 32 * q);
        memset(&acc, 0, sizeof acc);
        for (size_t j = hi; j-- > lo;) {
            fe_mul(&FR, &acc, &acc, &x);
            fe_add(&FR, &acc, &acc, &c[j]);
        }
        part[item] = acc;
    }
    for (size_t q = 0; q < npts; q++) {
        fe x, xl, acc;
        load_fe(&FR, &x, xs + 32 * q);
        uint64_t e[4] = {(uint64_t)L, 0, 0, 0};
        fe_pow(&FR, &xl, &x, e);
        memset(&acc, 0, sizeof acc);
        for (size_t k = nchunks; k-- > 0;) {                  /* Horner over the chunks in x^L */
            fe_mul(&FR, &acc, &acc, &xl);
            fe_add(&FR, &acc, &acc, &part[q * nchunks + k]);
        }
        store_fe(&FR, out + 32 * q, &acc);
    }
    free(part);
    free(c);
    return 0;
}

/* out = sum_j vals[j] * L_j(t) over the domain {omega^j, j < n}: the polynomial with evaluations vals[] evaluated at t through
 * the Lagrange coefficients L_j(t) = (t^n - 1)/n * omega^j / (t - omega^j) of FFTAuxiliary.serialRadix2LagrangeCoefficients
 * (src/main/java/algebra/fft/FFTAuxiliary.java:249-302; t outside the domain).  Used to derive the expected Groth16 proof
 * "in the exponent": A(t) = sum_i z_i A_i(t) = sum_j L_j(t) <a_j, z> (R1CStoQAP.java:54-91 combined with :143-160).
 * Returns -2 when t is a point of the domain. */
int oracle_fr_lagrange_eval(const uint8_t* vals, size_t n, const uint8_t omega_b[32], const uint8_t t_b[32], int threads,
                            uint8_t out[32]) {
    init_once();
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    fe omega, t;
    load_fe(&FR, &omega, omega_b);
    load_fe(&FR, &t, t_b);
    fe* part = (fe*)calloc((size_t)threads, sizeof(fe));
    int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int th = 0; th < threads; th++) {
        size_t lo = n * (size_t)th / (size_t)threads, hi = n * (size_t)(th + 1) / (size_t)threads;
        size_t len = hi - lo;
        if (len == 0) continue;
        fe* den = (fe*)malloc(len * sizeof(fe));
        fe* pre = (fe*)malloc(len * sizeof(fe));
        fe* num = (fe*)malloc(len * sizeof(fe));
        uint64_t e[4] = {lo, 0, 0, 0};
        fe w;
        fe_pow(&FR, &w, &omega, e);
        for (size_t k = 0; k < len; k++) {
            num[k] = w;
            fe_sub(&FR, &den[k], &t, &w);
            if (fe_is_zero(&den[k])) { bad = 1; den[k] = FR.one; }
            if (k == 0) pre[k] = den[k];
            else fe_mul(&FR, &pre[k], &pre[k - 1], &den[k]);
            fe_mul(&FR, &w, &w, &omega);
        }
        fe inv, acc;
        fe_inv(&FR, &inv, &pre[len - 1]);
        memset(&acc, 0, sizeof acc);
        for (size_t k = len; k-- > 0;) {
            fe di, v, term;
            if (k) fe_mul(&FR, &di, &inv, &pre[k - 1]);
            else di = inv;
            fe_mul(&FR, &inv, &inv, &den[k]);
            load_fe(&FR, &v, vals + 32 * (lo + k));
            fe_mul(&FR, &term, &num[k], &di);
            fe_mul(&FR, &term, &term, &v);
            fe_add(&FR, &acc, &acc, &term);
        }
        part[th] = acc;
        free(den); free(pre); free(num);
    }
    fe total;
    memset(&total, 0, sizeof total);
    for (int th = 0; th < threads; th++) fe_add(&FR, &total, &total, &part[th]);
    free(part);
    if (bad) return -2;
    /* (t^n - 1) / n */
    fe tn = t, nn, ninv, f, c;
    for (size_t m = 1; m < n; m <<= 1) fe_sqr(&FR, &tn, &tn);
    fe_sub(&FR, &f, &tn, &FR.one);
    memset(&c, 0, sizeof c);
    c.v[0] = (uint64_t)n;
    fe_to_mont(&FR, &nn, &c);
    fe_inv(&FR, &ninv, &nn);
    fe_mul(&FR, &f, &f, &ninv);
    fe_mul(&FR, &total, &total, &f);
    store_fe(&FR, out, &total);
    return 0;
}

/* The assignment of the reference's synthetic circuit, R1CSConstruction.serialConstruct (src/main/java/profiler/generation/
 * R1CSConstruction.java:31-110): full[0] = 1, full[1] = a, full[2] = b, then for constraint i < numConstraints - 1:
 * odd i -> a*b, even i -> a+b, (a, b) <- (b, tmp); the last entry is (sum of full[1 .. numVariables-2])^2.  a0/b0 are the two
 * Fp.random values (equal under the reference's fixed seed).  out: numVariables = numConstraints + 3 canonical elements. */
int oracle_r1cs_chain(size_t num_constraints, const uint8_t a0[32], const uint8_t b0[32], uint8_t* out) {
    init_once();
    if (num_constraints < 1) return -1;
    const size_t nv = num_constraints + 3;
    fe a, b, sum;
    load_fe(&FR, &a, a0);
    load_fe(&FR, &b, b0);
    store_fe(&FR, out, &FR.one);
    store_fe(&FR, out + 32, &a);
    store_fe(&FR, out + 64, &b);
    fe_add(&FR, &sum, &a, &b);
    for (size_t i = 0; i + 1 < num_constraints; i++) {
        fe tmp;
        if (i % 2 != 0) fe_mul(&FR, &tmp, &a, &b);
        else fe_add(&FR, &tmp, &a, &b);
        a = b;
        b = tmp;
        store_fe(&FR, out + 32 * (3 + i), &tmp);
        fe_add(&FR, &sum, &sum, &tmp);
    }
    /* full has numVariables - 1 entries so far (indices 0 .. nv-2); res = sum of full[1 .. nv-2] */
    fe sq;
    fe_sqr(&FR, &sq, &sum);
    store_fe(&FR, out + 32 * (nv - 1), &sq);
    return 0;
}

/* out[i] = a[i] * g^i mod r: FFTAuxiliary.multiplyByCoset (src/main/java/algebra/fft/FFTAuxiliary.java:224-232) */
int oracle_fr_coset_scale(const uint8_t* a, size_t n, const uint8_t g_b[32], uint8_t* out) {
    init_once();
    fe g, u = FR.one;
    load_fe(&FR, &g, g_b);
    for (size_t i = 0; i < n; i++) {
        fe x;
        load_fe(&FR, &x, a + 32 * i);
        fe_mul(&FR, &x, &x, &u);
        store_fe(&FR, out + 32 * i, &x);
        fe_mul(&FR, &u, &u, &g);
    }
    return 0;
}

/* Evaluation vectors of R1CStoQAP.R1CStoQAPWitness (src/main/java/reductions/r1cs_to_qap/R1CStoQAP.java:143-160) for the
 * reference's synthetic circuit (R1CSConstruction.serialConstruct, R1CSConstruction.java:48-104): a_j = <A_j, z>, b_j, c_j for
 * every constraint j, the primary inputs appended to a at num_constraints + i (:151-153), zero up to the domain size n_dom.
 * Only variables with index < col_limit contribute (col_limit = numVariables: the full vectors; col_limit = numInputs: the
 * part of every row that belongs to the primary input, used to split sum_i z_i A_i(t) into its primary and auxiliary halves).
 * z: numConstraints + 3 canonical elements; outputs n_dom canonical elements each. */
int oracle_r1cs_synth_eval(size_t nc, size_t ni, const uint8_t* z, size_t col_limit, size_t n_dom, uint8_t* out_a, uint8_t* out_b,
                           uint8_t* out_c) {
    init_once();
    const size_t nv = nc + 3;
    if (n_dom < nc + ni || col_limit > nv) return -1;
    memset(out_a, 0, 32 * n_dom);
    memset(out_b, 0, 32 * n_dom);
    memset(out_c, 0, 32 * n_dom);
#define ZV(idx, dst) do { if ((size_t)(idx) < col_limit) load_fe(&FR, (dst), z + 32 * (size_t)(idx)); else memset((dst), 0, sizeof(fe)); } while (0)
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)nc - 1; i++) {
        fe a, b, c, t;
        if (i % 2 != 0) {                       /* A = x_{i+1}, B = x_{i+2}, C = x_{i+3} */
            ZV(i + 1, &a);
            ZV(i + 2, &b);
        } else {                                /* A = x_{i+1} + x_{i+2}, B = x_0, C = x_{i+3} */
            ZV(i + 1, &a);
            ZV(i + 2, &t);
            fe_add(&FR, &a, &a, &t);
            ZV(0, &b);
        }
        ZV(i + 3, &c);
        store_fe(&FR, out_a + 32 * (size_t)i, &a);
        store_fe(&FR, out_b + 32 * (size_t)i, &b);
        store_fe(&FR, out_c + 32 * (size_t)i, &c);
    }
    {                                           /* last constraint: A = B = sum_{1 <= j <= nv-2} x_j, C = x_{nv-1} */
        fe sum, v;
        memset(&sum, 0, sizeof sum);
        for (size_t j = 1; j + 1 < nv; j++) {
            ZV(j, &v);
            fe_add(&FR, &sum, &sum, &v);
        }
        store_fe(&FR, out_a + 32 * (nc - 1), &sum);
        store_fe(&FR, out_b + 32 * (nc - 1), &sum);
        ZV(nv - 1, &v);
        store_fe(&FR, out_c + 32 * (nc - 1), &v);
    }
    for (size_t i = 0; i < ni; i++) {
        fe v;
        ZV(i, &v);
        store_fe(&FR, out_a + 32 * (nc + i), &v);
    }
#undef ZV
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
