// parse_num.h -- one decimal token -> the correctly rounded double, host and device.
//
// The reference readers hand text tokens to numpy, i.e. to a correctly rounded
// string -> double conversion (io/vasp.py:94-103, io/cube.py:99-113), so the
// parser here has to be correctly rounded too (bit parity of every density
// value).  Three exact routes, everything else is reported for the host:
//   * zero;
//   * Clinger's fast path: significand w <= 2^53 and |q| <= 22, where
//     double(w) and 10^|q| are exact and one IEEE multiply / divide rounds once;
//   * q < 0 beyond that (small densities): w / 10^m = (w / 5^m) * 2^-m by exact
//     multi-word long division (5^m up to 2^256, m <= 110): 56 quotient bits, a
//     sticky remainder, round-to-nearest-even.
// Grammar: [+-] digits [. digits] [(e|E) [+-] digits] (what Python's float accepts of it), at most 19 significant
// digits.  Anything else ("****", nan, 20+ digits, huge exponents, subnormal
// results) returns 1 and the caller passes the token to the host's strtod.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define BDR_HD __host__ __device__ __forceinline__
#else
#define BDR_HD inline
#endif

namespace bdr {

#ifdef __CUDACC__
#define BDR_POW5_QUAL __device__ const
#define BDR_POW5_NAME d_pow5
#include "pow5_table.h"
#undef BDR_POW5_QUAL
#undef BDR_POW5_NAME
#endif
#define BDR_POW5_QUAL static const
#define BDR_POW5_NAME h_pow5
#include "pow5_table.h"
#undef BDR_POW5_QUAL
#undef BDR_POW5_NAME
#ifdef __CUDA_ARCH__
#define BDR_POW5(m) d_pow5[m]
#else
#define BDR_POW5(m) h_pow5[m]
#endif

BDR_HD int bitlen64(uint64_t x) {
#ifdef __CUDA_ARCH__
    return 64 - __clzll((long long)x);
#else
    return x ? 64 - __builtin_clzll(x) : 0;
#endif
}

BDR_HD double bits_to_double(uint64_t b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    union { uint64_t u; double d; } c;
    c.u = b;
    return c.d;
#endif
}

BDR_HD double pow10_exact(int k) {  // 10^k, 0 <= k <= 22: exactly representable
    const double t[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                          1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    return t[k];
}

// w / 10^m, 1 <= m <= BDR_POW5_MAX, 0 < w < 5^m, correctly rounded; 1 = not representable here
BDR_HD int decimal_long_path(uint64_t w, int m, uint64_t sign, double *out) {
    uint64_t D[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) D[i] = BDR_POW5(m)[i];
    int LD = 0;
#pragma unroll
    for (int i = 3; i >= 0; --i)
        if (LD == 0 && D[i]) LD = 64 * i + bitlen64(D[i]);
    const int LW = bitlen64(w);
    if (LD <= 64 && w >= D[0]) return 1;  // needs w < 5^m
    const int a = LD - LW - 1 > 0 ? LD - LW - 1 : 0;
    // R = w << a  (bit length LD - 1, so R < D), five words
    uint64_t R[5] = {0, 0, 0, 0, 0};
    {
        const int ws = a >> 6, bs = a & 63;
        R[ws] = w << bs;
        if (bs && ws + 1 < 5) R[ws + 1] = w >> (64 - bs);
    }
    uint64_t Q = 0;
    for (int it = 0; it < 56; ++it) {
        // R <<= 1
#pragma unroll
        for (int i = 4; i > 0; --i) R[i] = (R[i] << 1) | (R[i - 1] >> 63);
        R[0] <<= 1;
        // R >= D ?
        bool ge = R[4] != 0;
        if (!ge) {
            ge = true;
#pragma unroll
            for (int i = 3; i >= 0; --i) {
                if (R[i] != D[i]) {
                    ge = R[i] > D[i];
                    break;
                }
            }
        }
        Q <<= 1;
        if (ge) {
            uint64_t borrow = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint64_t d = D[i] + borrow;
                const uint64_t nb = (d < borrow) || (R[i] < d) ? 1u : 0u;  // d wrapped or R[i] < d
                R[i] = R[i] - d;
                borrow = nb;
            }
            R[4] -= borrow;
            Q |= 1;
        }
    }
    const bool sticky = (R[0] | R[1] | R[2] | R[3] | R[4]) != 0;
    const int nb = bitlen64(Q);  // 55 or 56
    const int sh = nb - 53;
    uint64_t M = Q >> sh;
    const uint64_t dropped = Q & ((1ULL << sh) - 1), half = 1ULL << (sh - 1);
    if (dropped > half || (dropped == half && (sticky || (M & 1)))) ++M;
    int E = sh - 56 - a - m;  // value = M * 2^E
    if (M == (1ULL << 53)) {
        M >>= 1;
        ++E;
    }
    const int biased = E + 52 + 1023;
    if (biased <= 0 || biased >= 2047) return 1;
    *out = bits_to_double(sign | ((uint64_t)biased << 52) | (M & ((1ULL << 52) - 1)));
    return 0;
}

BDR_HD bool is_space(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// parses the token starting at p (at most `avail` bytes readable); returns 0 and
// the value, or 1 when the token has to go to the host.  *len = token length.
BDR_HD int parse_token(const char *p, int64_t avail, double *out, int *len) {
    int i = 0;
    const int cap = avail < 64 ? (int)avail : 64;
    // token length first
    int n = 0;
    while (n < cap && !is_space((unsigned char)p[n])) ++n;
    *len = n;
    if (n == 0 || (n == cap && avail > 64)) return 1;
    uint64_t sign = 0;
    if (p[0] == '-') {
        sign = 1ULL << 63;
        ++i;
    } else if (p[0] == '+') {
        ++i;
    }
    uint64_t w = 0;
    int nd = 0, ndigits = 0, q = 0;
    bool inexact = false;
    for (; i < n && p[i] >= '0' && p[i] <= '9'; ++i) {
        const int d = p[i] - '0';
        ++ndigits;
        if (w == 0 && d == 0) continue;  // leading zeros
        if (nd < 19) {
            w = w * 10 + (uint64_t)d;
            ++nd;
        } else {
            ++q;  // an integer digit that did not fit: scales by 10
            inexact |= d != 0;
        }
    }
    if (i < n && p[i] == '.') {
        ++i;
        for (; i < n && p[i] >= '0' && p[i] <= '9'; ++i) {
            const int d = p[i] - '0';
            ++ndigits;
            if (nd < 19) {
                w = w * 10 + (uint64_t)d;
                if (w) ++nd;
                --q;
            } else {
                inexact |= d != 0;
            }
        }
    }
    if (ndigits == 0) return 1;
    if (i < n && (p[i] == 'e' || p[i] == 'E')) {
        ++i;
        bool eneg = false;
        if (i < n && (p[i] == '-' || p[i] == '+')) {
            eneg = p[i] == '-';
            ++i;
        }
        int e = 0, ed = 0;
        for (; i < n && p[i] >= '0' && p[i] <= '9'; ++i, ++ed)
            if (e < 100000) e = e * 10 + (p[i] - '0');
        if (ed == 0) return 1;
        q += eneg ? -e : e;
    }
    if (i != n) return 1;      // trailing junk ("****", "nan", "1.0x")
    if (inexact) return 1;     // more than 19 significant digits
    if (w == 0) {
        *out = bits_to_double(sign);
        return 0;
    }
    if (w <= (1ULL << 53)) {
        if (q >= 0 && q <= 22) {
            const double v = (double)w * pow10_exact(q);
            *out = sign ? -v : v;
            return 0;
        }
        if (q < 0 && q >= -22) {
            const double v = (double)w / pow10_exact(-q);
            *out = sign ? -v : v;
            return 0;
        }
    }
    if (q < 0 && -q <= BDR_POW5_MAX) return decimal_long_path(w, -q, sign, out);
    return 1;
}

}  // namespace bdr
