// pow() with the HOST C library's results, bit for bit, on the device.
//
// Why: the strain energy of the triplet costs (reg_tools.cpp:596-597, DiscreteCostFunction.cpp:187, DiscreteGroupCostFunction.cpp:51)
// goes through three std::pow calls per cost, the costs feed a discrete optimiser, and glibc's pow is accurate to ~0.52 ulp but NOT
// correctly rounded: no independent implementation (CUDA's included) returns the same double for every argument. Round 1 therefore
// finished the costs on the host. This header evaluates glibc's own algorithm instead.
//
// What: glibc >= 2.28 computes pow(x, y) = exp(y * log(x)) with a table-driven log in double-double and a table-driven exp
// (the "optimized routines" algorithm: log: z = x / 2^k in [0x1.69555p-1, 0x1.69555p0), r = z * invc_i - 1 with one FMA, polynomial of
// degree 7; exp: k = round(x * 128 / ln2), 2^(k/128) from a table, polynomial of degree 5). The operation sequence below is a
// restatement of that algorithm as the x86-64 FMA build of glibc 2.39 executes it (which sums are fused and in which order: read off
// `objdump -d libm.so.6`, the ifunc target chosen on FMA + AVX2 hosts); every step is one IEEE-754 operation, which CUDA's
// add / mul / fma reproduce exactly (--fmad=false: nothing is contracted behind our back).
// The CONSTANTS (two polynomials and two tables, 6.3 KB) are not copied into this repository: hostpow.cu reads them out of the
// libm that is mapped into the running process, so the device reproduces whatever that library computes.
//
// Safety net: before the device path is enabled, hostpow.cu evaluates this same function on the HOST for ~2 M arguments (random,
// near-1, huge / tiny exponents, subnormal and special operands) and compares with std::pow bit for bit. Any mismatch (another libm,
// a non-FMA host, a future glibc) disables the device path and the costs are finished on the host as before (triplet.cu).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace msm {

struct PowTables {
    double ln2hi, ln2lo, A[7];                       // log: pow_log_data header
    double invln2N, shift, negln2hiN, negln2loN, C[4];   // exp: exp_data header (C2..C5)
    const double* logtab;                            // [128][4]: invc, (pad), logc, logctail
    const unsigned long long* exptab;                // [2 * 128]: tail bits, 2^(i/128) bits minus (i << 45)
};

__host__ __device__ __forceinline__ unsigned long long hp_bits(double x) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(x);
#else
    unsigned long long u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
__host__ __device__ __forceinline__ double hp_double(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}
__host__ __device__ __forceinline__ double hp_ld(const double* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
__host__ __device__ __forceinline__ unsigned long long hp_ldu(const unsigned long long* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// 0: y is not an integer, 1: odd integer, 2: even integer
__host__ __device__ __forceinline__ int hp_checkint(unsigned long long iy) {
    const int e = (int)(iy >> 52 & 0x7ff);
    if (e < 0x3ff) return 0;
    if (e > 0x3ff + 52) return 2;
    if (iy & ((1ULL << (0x3ff + 52 - e)) - 1)) return 0;
    if (iy & (1ULL << (0x3ff + 52 - e))) return 1;
    return 2;
}
__host__ __device__ __forceinline__ bool hp_zeroinfnan(unsigned long long i) { return 2 * i - 1 >= 2 * 0x7ff0000000000000ULL - 1; }

__host__ __device__ inline double host_pow(double x, double y, const PowTables& T) {
    const unsigned long long kOne = 0x3ff0000000000000ULL, kInf = 0x7ff0000000000000ULL;
    unsigned long long ix = hp_bits(x);
    const unsigned long long iy = hp_bits(y);
    unsigned topx = (unsigned)(ix >> 52);
    const unsigned topy = (unsigned)(iy >> 52);
    unsigned long long sign_bias = 0;
    if (topx - 0x001u >= 0x7ffu - 0x001u || (topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
        // x < 0x1p-1022, inf or nan; or |y| < 0x1p-65, |y| >= 0x1p63 or nan
        if (hp_zeroinfnan(iy)) {
            if (2 * iy == 0) return 1.0;                                   // (a signalling NaN x would give x + y: not produced here)
            if (ix == kOne) return 1.0;
            if (2 * ix > 2 * kInf || 2 * iy > 2 * kInf) return x + y;
            if (2 * ix == 2 * kOne) return 1.0;
            if ((2 * ix < 2 * kOne) == !(iy >> 63)) return 0.0;            // |x| < 1 and y = inf, or |x| > 1 and y = -inf
            return y * y;
        }
        if (hp_zeroinfnan(ix)) {
            double x2 = x * x;
            if (ix >> 63 && hp_checkint(iy) == 1) x2 = -x2;
            return iy >> 63 ? 1 / x2 : x2;                                  // x = 0, y < 0: +-inf (division by zero)
        }
        if (ix >> 63) {   // finite x < 0
            const int yint = hp_checkint(iy);
            if (yint == 0) return (x - x) / (x - x);                        // invalid: NaN
            if (yint == 1) sign_bias = 0x800ULL << 7;
            ix &= 0x7fffffffffffffffULL;
            topx &= 0x7ff;
        }
        if ((topy & 0x7ff) - 0x3beu >= 0x43eu - 0x3beu) {
            if (ix == kOne) return 1.0;
            if ((topy & 0x7ff) < 0x3beu) return ix > kOne ? 1.0 + y : 1.0 - y;   // |y| < 2^-65: x^y ~ 1 + y log x
            const bool over = (ix > kOne) == (topy < 0x800u);
            return over ? hp_double(kInf) : 0.0;                            // sign_bias is 0 here: y is not an odd integer
        }
        if (topx == 0) {   // subnormal x: normalise so that the exponent becomes negative
            ix = hp_bits(x * 0x1p52);
            ix &= 0x7fffffffffffffffULL;
            ix -= 52ULL << 52;
        }
    }
    // ---- log(x) = hi + lo: x = 2^k z, z in [0x1.69555p-1, 0x1.69555p0), log z = log c + log1p(z / c - 1)
    const unsigned long long tmp = ix - 0x3fe6955500000000ULL;
    const int i = (int)((tmp >> (52 - 7)) & 127);
    const int k = (int)((long long)tmp >> 52);
    const double z = hp_double(ix - (tmp & 0xfffULL << 52));
    const double kd = (double)k;
    const double* e = T.logtab + 4 * i;
    const double invc = hp_ld(e), logc = hp_ld(e + 2), logctail = hp_ld(e + 3);
    const double r = fma(z, invc, -1.0);
    const double t1 = fma(kd, T.ln2hi, logc);
    const double t2 = t1 + r;
    const double lo1 = fma(kd, T.ln2lo, logctail);
    const double lo2 = (t1 - t2) + r;
    const double ar = T.A[0] * r;
    const double ar2 = r * ar;
    const double ar3 = r * ar2;
    const double hi = t2 + ar2;
    const double lo3 = fma(ar, r, -ar2);
    const double lo4 = (t2 - hi) + ar2;
    const double q = fma(ar2, fma(ar2, fma(r, T.A[6], T.A[5]), fma(r, T.A[4], T.A[3])), fma(r, T.A[2], T.A[1]));
    const double lo = fma(ar3, q, ((lo1 + lo2) + lo3) + lo4);
    const double loghi = hi + lo;
    const double loglo = (hi - loghi) + lo;
    // ---- y * log(x) = ehi + elo
    const double ehi = y * loghi;
    const double elo = fma(y, loglo, fma(loghi, y, -ehi));
    // ---- exp(ehi + elo) with the sign of the result in sign_bias
    unsigned abstop = (unsigned)(hp_bits(ehi) >> 52) & 0x7ff;
    if (abstop - 0x3c9u >= 0x408u - 0x3c9u) {
        if (abstop - 0x3c9u >= 0x80000000u) {      // |ehi| < 2^-54 (0 is a common input): no spurious underflow
            const double one = 1.0 + ehi;
            return sign_bias ? -one : one;
        }
        if (abstop >= 0x409u) {                    // |ehi| >= 1024: overflow / underflow (inf and nan were handled above)
            const double v = hp_bits(ehi) >> 63 ? 0.0 : hp_double(kInf);
            return sign_bias ? -v : v;
        }
        abstop = 0;                                // 512 <= |ehi| < 1024: the result may over- or underflow, handled after the polynomial
    }
    const double zz = fma(ehi, T.invln2N, T.shift);
    const unsigned long long ki = hp_bits(zz);
    const double kd2 = zz - T.shift;
    double rr = fma(kd2, T.negln2hiN, ehi);
    rr = fma(kd2, T.negln2loN, rr);
    rr = elo + rr;
    const unsigned idx = 2 * (unsigned)(ki & 127);
    const unsigned long long top = (ki + sign_bias) << (52 - 7);
    const double tail = hp_double(hp_ldu(T.exptab + idx));
    unsigned long long sbits = hp_ldu(T.exptab + idx + 1) + top;
    const double r2 = rr * rr;
    const double p23 = fma(rr, T.C[1], T.C[0]);
    const double p45 = fma(rr, T.C[3], T.C[2]);
    const double tr = rr + tail;
    const double tmpv = fma(p45, r2 * r2, fma(p23, r2, tr));
    if (abstop == 0) {   // specialcase(): scale = 2^k may not be representable
        if ((ki & 0x80000000ULL) == 0) {   // k > 0: the exponent of scale might have overflowed by <= 460
            sbits -= 1009ULL << 52;
            const double scale = hp_double(sbits);
            return 0x1p1009 * fma(scale, tmpv, scale);
        }
        sbits += 1022ULL << 52;            // k < 0: careful in the subnormal range
        const double scale = hp_double(sbits);
        const double st = tmpv * scale;
        double yv = scale + st;
        if (fabs(yv) < 1.0) {              // round to the subnormal grid once (avoid double rounding)
            const double one = yv < 0.0 ? -1.0 : 1.0;
            double l = (scale - yv) + st;
            const double h = yv + one;
            l = (((one - h) + yv) + l);
            yv = (l + h) - one;
            if (yv == 0.0) yv = hp_double(sbits & 0x8000000000000000ULL);
        }
        return 0x1p-1022 * yv;
    }
    const double scale = hp_double(sbits);
    return fma(tmpv, scale, scale);
}

// the tables of the host C library on the device (hostpow.cu); enabled == false: finish on the host
struct DevicePow {
    bool enabled = false;
    PowTables t{};
};
const DevicePow& device_pow(int device);     // per device, built on first use (host self-test first)
const PowTables* host_pow_tables();          // NULL when the host library's tables were not found or the self-test failed

} // namespace msm
