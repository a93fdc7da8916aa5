/*
 * oracle/lazy_gaussian.c -- CPU restatement of the reference's hash-seeded lazy Gaussian
 * DRM generator.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs as the checker.  The product path
 * (tt-sketch_b200/) never links or calls this file.
 *
 * What it restates (reference file:line, all under /root/reference):
 *   ora_hash64            tt_sketch/drm/fast_lazy_gaussian.pyx:13-37   (hash_int_c)
 *   ora_flat_index        tt_sketch/drm/fast_lazy_gaussian.pyx:56-71   (int32-wrapping stride)
 *   ora_uniform           tt_sketch/drm/fast_lazy_gaussian.pyx:73-102 + :48 (frexp*2-1)
 *   ora_ndtri             scipy.special.cython_special.ndtri, called at .pyx:49.  SciPy is a
 *                         third-party dependency absent from /root/reference and UNPINNED
 *                         there (setup.py:40-43, requirements.txt:2); the image has
 *                         SciPy 1.18.1 whose ndtri is cephes ndtri.c (via xsf).  The cephes
 *                         algorithm is restated below from its published form.
 *   ora_inds_to_normal    tt_sketch/drm/fast_lazy_gaussian.pyx:183-201 (inds_to_normal)
 *
 * Parity pin: tests/test_oracle_pin.py checks every function here bit-for-bit against the
 * compiled reference (oracle/_ref, built by oracle/build_ref.sh from the .pyx where it lies)
 * and against the committed golden vectors in tests/golden/ that the reference generated.
 *
 * Two logs are provided: ora_ndtri() calls the C library's log()/sqrt() exactly as the
 * reference's dependency does; ora_ndtri_restated() replaces log() by ora_log_glibc(), an
 * operation-by-operation restatement of glibc 2.39's FMA build of log() (the table comes
 * from the image's libm, see tools/gen_logtab.py).  The CUDA device code follows the second
 * form; the tests require both to agree bit-for-bit.
 *
 * Compile with -ffp-contract=off: the cephes polynomials are evaluated WITHOUT fused
 * multiply-add in the SciPy binary (baseline x86-64 build), so contraction would change bits.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "logtab.inc"

static inline double bits2d(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline uint64_t d2bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

/* ---- integer hash (splitmix64 finaliser with an additive constant) ---- .pyx:13-37 */
uint64_t ora_hash64(uint64_t r) {
    r += 0x4BE98134A5976FD3ULL;
    r ^= r >> 30;
    r *= 0xBF58476D1CE4E5B9ULL;
    r ^= r >> 27;
    r *= 0x94D049BB133111EBULL;
    r ^= r >> 31;
    return r;
}

/* ---- cephes ndtri ---- */
static const double P0[5] = {
    -5.99633501014107895267E1, 9.80010754185999661536E1, -5.66762857469070293439E1,
    1.39312609387279679503E1, -1.23916583867381258016E0};
static const double Q0[8] = {
    1.95448858338141759834E0, 4.67627912898881538453E0, 8.63602421390890590575E1,
    -2.25462687854119370527E2, 2.00260212380060660359E2, -8.20372256168333339912E1,
    1.59056225126211695515E1, -1.18331621121330003142E0};
static const double P1[9] = {
    4.05544892305962419923E0, 3.15251094599893866154E1, 5.71628192246421288162E1,
    4.40805073893200834700E1, 1.46849561928858024014E1, 2.18663306850790267539E0,
    -1.40256079171354495875E-1, -3.50424626827848203418E-2, -8.57456785154685413611E-4};
static const double Q1[8] = {
    1.57799883256466749731E1, 4.53907635128879210584E1, 4.13172038254672030440E1,
    1.50425385692907503408E1, 2.50464946208309415979E0, -1.42182922854787788574E-1,
    -3.80806407691578277194E-2, -9.33259480895457427372E-4};
static const double P2[9] = {
    3.23774891776946035970E0, 6.91522889068984211695E0, 3.93881025292474443415E0,
    1.33303460815807542389E0, 2.01485389549179081538E-1, 1.23716634817820021358E-2,
    3.01581553508235416007E-4, 2.65806974686737550832E-6, 6.23974539184983293730E-9};
static const double Q2[8] = {
    6.02427039364742014255E0, 3.67983563856160859403E0, 1.37702099489081330271E0,
    2.16236993594496635890E-1, 1.34204006088543189037E-2, 3.28014464682127739104E-4,
    2.89247864745380683936E-6, 6.79019408009981274425E-9};

static inline double polevl(double x, const double *c, int n) {
    double a = c[0];
    for (int i = 1; i <= n; i++) a = a * x + c[i];
    return a;
}
static inline double p1evl(double x, const double *c, int n) {
    double a = x + c[0];
    for (int i = 1; i < n; i++) a = a * x + c[i];
    return a;
}

/* glibc 2.39 log(), FMA variant, main path only (x not near 1, normal, positive).
 * Valid for the arguments ndtri produces here: y in [2^-52, 0.1354], x in (2, 8.5). */
double ora_log_glibc(double x) {
    static const uint64_t tabbits[256] = TTSK_LOG_TAB_INIT;
    const double ln2hi = bits2d(TTSK_LOG_LN2HI_BITS), ln2lo = bits2d(TTSK_LOG_LN2LO_BITS);
    const double A0 = bits2d(TTSK_LOG_A0_BITS), A1 = bits2d(TTSK_LOG_A1_BITS),
                 A2 = bits2d(TTSK_LOG_A2_BITS), A3 = bits2d(TTSK_LOG_A3_BITS),
                 A4 = bits2d(TTSK_LOG_A4_BITS);
    uint64_t ix = d2bits(x);
    uint64_t tmp = ix - 0x3fe6000000000000ULL;
    int i = (int)((tmp >> 45) & 127);
    int64_t k = (int64_t)tmp >> 52;
    uint64_t iz = ix - (tmp & 0xfff0000000000000ULL);
    double invc = bits2d(tabbits[2 * i]), logc = bits2d(tabbits[2 * i + 1]);
    double z = bits2d(iz);
    double kd = (double)k;
    double r = __builtin_fma(z, invc, -1.0);
    double w = __builtin_fma(kd, ln2hi, logc);
    double hi = w + r;
    double lo = __builtin_fma(kd, ln2lo, (w - hi) + r);
    double r2 = r * r;
    double p = __builtin_fma(r2, __builtin_fma(r, A4, A3), __builtin_fma(r, A2, A1));
    return __builtin_fma(r * r2, p, __builtin_fma(r2, A0, lo)) + hi;
}

static inline double ndtri_impl(double y0, int restated) {
    const double s2pi = 2.50662827463100050242E0;
    const double expm2 = 0.13533528323661269189; /* exp(-2) */
    if (y0 == 0.0) return -INFINITY;
    if (y0 == 1.0) return INFINITY;
    if (y0 < 0.0 || y0 > 1.0) return NAN;
    int code = 1;
    double y = y0;
    if (y > 1.0 - expm2) { y = 1.0 - y; code = 0; }
    if (y > expm2) {
        y = y - 0.5;
        double y2 = y * y;
        double x = y + y * (y2 * polevl(y2, P0, 4) / p1evl(y2, Q0, 8));
        return x * s2pi;
    }
    double ly = restated ? ora_log_glibc(y) : log(y);
    double x = sqrt(-2.0 * ly);
    double lx = restated ? ora_log_glibc(x) : log(x);
    double x0 = x - lx / x;
    double z = 1.0 / x;
    double x1;
    if (x < 8.0) x1 = z * polevl(z, P1, 8) / p1evl(z, Q1, 8);
    else         x1 = z * polevl(z, P2, 8) / p1evl(z, Q2, 8);
    x = x0 - x1;
    if (code) x = -x;
    return x;
}
double ora_ndtri(double y0) { return ndtri_impl(y0, 0); }
double ora_ndtri_restated(double y0) { return ndtri_impl(y0, 1); }

void ora_ndtri_array(const double *u, double *out, int64_t n, int restated) {
    for (int64_t i = 0; i < n; i++) out[i] = ndtri_impl(u[i], restated);
}
void ora_log_array(const double *x, double *out_libm, double *out_restated, int64_t n) {
    for (int64_t i = 0; i < n; i++) { out_libm[i] = log(x[i]); out_restated[i] = ora_log_glibc(x[i]); }
}

/* flat index with the reference's C `int prod` stride: truncated to 32 bits and
 * sign-extended before the 64-bit multiply (.pyx:60-71).  idx is (k, nnz) row-major. */
static inline uint64_t flat_index(const uint64_t *idx, int k, int64_t nnz, const uint64_t *shape, int64_t p) {
    uint64_t flat = idx[p];
    int prod = (int)shape[0];
    for (int i = 1; i < k; i++) {
        flat += idx[(int64_t)i * nnz + p] * (uint64_t)(int64_t)prod;
        prod = (int)((uint32_t)prod * (uint32_t)shape[i]);
    }
    return flat;
}
void ora_flat_index(const uint64_t *idx, int k, int64_t nnz, const uint64_t *shape, uint64_t *out) {
    for (int64_t p = 0; p < nnz; p++) out[p] = flat_index(idx, k, nnz, shape, p);
}

/* uniform in [0,1): the reference forces the top three bits of the hash to 001, reinterprets
 * as double and takes frexp()*2-1 (.pyx:91-102, :48) == low 52 bits scaled by 2^-52. */
static inline double uniform_from_hash(uint64_t h) {
    h = (h | 0x2000000000000000ULL) & 0x3FFFFFFFFFFFFFFFULL;
    int e;
    return frexp(bits2d(h), &e) * 2 - 1;
}

void ora_inds_to_uniform(const uint64_t *idx, int k, int64_t nnz, const uint64_t *shape,
                         int rank_min, int rank_max, uint64_t seed, double *out) {
    int rank = rank_max - rank_min;
    for (int64_t p = 0; p < nnz; p++) {
        uint64_t flat = flat_index(idx, k, nnz, shape, p);
        for (int a = 0; a < rank; a++) {
            uint64_t salt = ora_hash64((uint64_t)(rank_min + a)) + seed;
            out[p * rank + a] = uniform_from_hash(ora_hash64(flat + salt));
        }
    }
}

/* out is (nnz, rank) row-major, like the reference's return value (.pyx:201). */
void ora_inds_to_normal(const uint64_t *idx, int k, int64_t nnz, const uint64_t *shape,
                        int rank_min, int rank_max, uint64_t seed, double *out, int restated) {
    int rank = rank_max - rank_min;
    uint64_t salts[4096];
    for (int a = 0; a < rank && a < 4096; a++) salts[a] = ora_hash64((uint64_t)(rank_min + a)) + seed;
    for (int64_t p = 0; p < nnz; p++) {
        uint64_t flat = flat_index(idx, k, nnz, shape, p);
        for (int a = 0; a < rank; a++) {
            uint64_t salt = a < 4096 ? salts[a] : ora_hash64((uint64_t)(rank_min + a)) + seed;
            out[p * rank + a] = ndtri_impl(uniform_from_hash(ora_hash64(flat + salt)), restated);
        }
    }
}

/* ---- sparse sign DRM ---- tt_sketch/drm/fast_lazy_gaussian.pyx:121-180 (_inds_to_sparse_sign, inds_to_sparse_sign)
 * Per nonzero: nnz_row hashed doubles d_j (the Gaussian DRM's hash with the top bits forced to 001, columns
 * 0..nnz_row-1, .pyx:52-105).  frexp(d_j) splits each into the exponent e, whose parity gives the sign
 * (e % 2) * 2 - 1 -- Cython gives `%` on C ints PYTHON semantics (cdivision is off), so the entries are -1 / +1; the
 * exponent field is hash bits 52..60 under the forced 01, so the parity is hash bit 52 -- and the mantissa
 * m * 2 - 1 (the low 52 hash bits as a uniform), which drives a partial Fisher-Yates shuffle of the row.
 * out is (nnz, rank_max - rank_min) row-major int16 (0 / -1 / +1) like the reference's return value. */
void ora_inds_to_sparse_sign(const uint64_t *idx, int k, int64_t nnz, const uint64_t *shape, int rank, int rank_min,
                             int rank_max, int nnz_row, uint64_t seed, int16_t *out) {
    int16_t row[4096];
    double mant[4096];
    int width = rank_max - rank_min;
    for (int64_t p = 0; p < nnz; p++) {
        uint64_t flat = flat_index(idx, k, nnz, shape, p);
        for (int j = 0; j < rank; j++) row[j] = 0;
        for (int j = 0; j < nnz_row; j++) {
            uint64_t salt = ora_hash64((uint64_t)j) + seed;
            uint64_t h = (ora_hash64(flat + salt) | 0x2000000000000000ULL) & 0x3FFFFFFFFFFFFFFFULL;
            int e;
            mant[j] = frexp(bits2d(h), &e) * 2 - 1;
            row[j] = (int16_t)((((e % 2) + 2) % 2) * 2 - 1);  /* Python-style remainder */
        }
        for (int j = 0; j < nnz_row; j++) {
            int rn = (int)(mant[j] * (rank - j) + j);
            int16_t t = row[j];
            row[j] = row[rn];
            row[rn] = t;
        }
        for (int a = 0; a < width; a++) out[p * width + a] = row[rank_min + a];
    }
}
