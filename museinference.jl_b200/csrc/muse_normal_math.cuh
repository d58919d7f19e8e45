// muse_normal_math.cuh — the Box–Muller transform of muse_draws.cu, table-driven.
//
// Same generator as before, element for element (oracle/philox.py):
//     u₁ = (v₁ + ½)·2⁻⁵³,  u₂ = (v₂ + ½)·2⁻⁵³  (v: 53-bit integers from Philox4x32-10)
//     element 2p = r cos(2πu₂),  element 2p+1 = r sin(2πu₂),  r = sqrt(−2 ln u₁)
// but −2 ln u₁ and (cos, sin)(2πu₂) are evaluated without libm: the libm versions cost ≈ 110 issue slots per pair,
// 60 of them materialising polynomial constants; the draws kernel is issue-bound, so the transform was most of it.
//
//   −2 ln u₁:  y = v₁ + ½ = m·2^e, m ∈ [√½, √2) after moving m ≥ √2 down one octave; c = round(128·m)/128 from a 91-entry
//              table (1/c, −2 ln c); r = m·(1/c) − 1 (one FMA, |r| ≤ 0.0055; exact at c = 1, which keeps the result
//              relatively accurate as u₁ → 1); −2 ln(1+r) by a degree-7 polynomial; + (e − 53)·(−2 ln 2) in two parts.
//   cos, sin:  v₂ = j·2⁴⁵ + w: the angle is a_j + b with a_j = 2π(j + ½)/256 from a 256-entry table and
//              b = 2π(w + ½ − 2⁴⁴)/2⁵³ exactly the remainder, |b| ≤ π/256; sin b and cos b − 1 by short Taylor
//              polynomials (remainders < 2e-20), then the angle-addition formulas arranged as a_j-term + small correction.
// Absolute error of an output ≤ ≈ 6e-16·(1 + r) (checked against long-double references in tests/test_oracle.py
// through the host build of this header, tests/csrc/normal_math_host.cpp).
//
// Everything is written with explicit fma() so that the host build (checker) and the device build agree bit for bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

// nvcc: device-only functions reading a __constant__ table; g++ (the checker's host build): plain inline functions
#if defined(__CUDACC__)
#define MUSE_HD __device__ __forceinline__
#else
#define MUSE_HD inline
#endif

namespace muse {

// Polynomial and scaling constants.  In device code they are read as constant-bank operands of the FP64 instructions
// (c[3][…]): spelled as literals, the compiler materialises each through two UMOV/IMAD.MOV per use, ≈ 30 issue slots
// per pair in an issue-bound kernel.  The host build (checker) uses the same values from a plain array.
#if defined(__CUDACC__)
#define MUSE_NM_CONST static __constant__
#else
#define MUSE_NM_CONST static const
#endif
MUSE_NM_CONST double kNM[16] = {
    -2.0 / 7.0, 1.0 / 3.0, -0.4, -2.0 / 3.0,                 // 0-3: −2 ln(1+r) polynomial
    -0x1.62e42fefa38p+0, -0x1.ef35793c7673p-44,              // 4-5: −2 ln 2 = HI (42 significant bits) + LO
    0x1.921fb54442d18p-51,                                   // 6:   2π·2⁻⁵³
    -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0,                  // 7-9: sin b
    -1.0 / 720.0, 1.0 / 24.0,                                // 10-11: cos b − 1
    0.0, 0.0, 0.0, 0.0,
};

MUSE_HD double bits_to_double(uint64_t b) {
#if defined(__CUDACC__)
    return __longlong_as_double((long long)b);
#else
    double d; std::memcpy(&d, &b, 8); return d;
#endif
}
MUSE_HD uint64_t double_to_bits(double d) {
#if defined(__CUDACC__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t b; std::memcpy(&b, &d, 8); return b;
#endif
}

// 53-bit integer of a uniform from two Philox words (the same bits u53() of the libm path uses)
MUSE_HD uint64_t v53(uint32_t lo, uint32_t hi) { return ((uint64_t)(hi >> 5) << 26) + (uint64_t)(lo >> 6); }

// t = −2 ln((v + ½)·2⁻⁵³);  logtab: 91 × (1/c, −2 ln c)
MUSE_HD double neg2_log_u(uint64_t v, const double (*logtab)[2]) {
    const double y = (double)v + 0.5;                      // the rounding the oracle's u53 makes, kept bit for bit
    const uint64_t yb = double_to_bits(y);
    const uint32_t hi = (uint32_t)(yb >> 32);
    const uint32_t mant20 = hi & 0xFFFFFu;
    int e = (int)(hi >> 20) - 1023 - 53;                   // u₁ = y·2⁻⁵³
    int idx;
    uint64_t mb = (yb & 0x000FFFFFFFFFFFFFull);
    if (mant20 >= 0x6A09Fu) {                              // m ≥ √2 (to 20 bits): use m/2 ∈ [√½, 1)
        e += 1;
        mb |= 0x3FE0000000000000ull;
        idx = (int)((mant20 + 8192u) >> 14) - 27;          // c = (64 + round(64 f))/128 → table index 128c − 91
    } else {
        mb |= 0x3FF0000000000000ull;
        idx = (int)((mant20 + 4096u) >> 13) + 37;          // c = (128 + round(128 f))/128
    }
    const double m = bits_to_double(mb);
    const double rc = logtab[idx][0], tlogc = logtab[idx][1];
    const double r = fma(m, rc, -1.0);
    // −2 ln(1+r) = −2r + r²(1 − ⅔r + ½r² − ⅖r³ + ⅓r⁴ − 2⁄7 r⁵)   (|r| ≤ 0.0055: next term 2r⁸/8 < 3e-19)
    double p = fma(r, kNM[0], kNM[1]);
    p = fma(r, p, kNM[2]);
    p = fma(r, p, 0.5);
    p = fma(r, p, kNM[3]);
    p = fma(r, p, 1.0);
    const double P = fma(r * r, p, -2.0 * r);
    const double ed = (double)e;
    // −2 ln 2 = HI + LO with HI carrying 42 significant bits: e·HI is exact for |e| ≤ 2¹¹
    const double s1 = fma(ed, kNM[4], tlogc);
    const double s2 = fma(ed, kNM[5], P);
    return s1 + s2;
}

// (cos, sin)(2π(v + ½)·2⁻⁵³);  trigtab: 256 × (cos a_j, sin a_j), a_j = 2π(j + ½)/256
MUSE_HD void sincos_2pi_u(uint64_t v, const double (*trigtab)[2], double* cs, double* sn) {
    const int j = (int)(v >> 45);
    const long long w = (long long)(v & 0x1FFFFFFFFFFFull) - (1ll << 44);
    const double b = ((double)w + 0.5) * kNM[6];                     // 2π·2⁻⁵³; (w + ½) is exact (46 bits)
    const double ca = trigtab[j][0], sa = trigtab[j][1];
    const double b2 = b * b;
    double ps = fma(b2, kNM[7], kNM[8]);
    ps = fma(b2, ps, kNM[9]);
    const double sb = fma(b2 * b, ps, b);                            // sin b
    double pc = fma(b2, kNM[10], kNM[11]);
    pc = fma(b2, pc, -0.5);
    const double cm1 = b2 * pc;                                      // cos b − 1
    *cs = ca + fma(ca, cm1, -(sa * sb));
    *sn = sa + fma(sa, cm1, ca * sb);
}

// the pair of normals of one Philox block (r0..r3)
MUSE_HD void box_muller_tab(uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3, const double (*logtab)[2],
                            const double (*trigtab)[2], double* n0, double* n1) {
    const double rad = sqrt(neg2_log_u(v53(r0, r1), logtab));
    double cs, sn;
    sincos_2pi_u(v53(r2, r3), trigtab, &cs, &sn);
    *n0 = rad * cs;
    *n1 = rad * sn;
}

}  // namespace muse
