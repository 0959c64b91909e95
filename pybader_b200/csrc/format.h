// format.h -- fp64 grid -> text, the numeric blocks of the reference's writers
// (io/vasp.py:167-258, io/cube.py:159-222 through utils.python_format, utils.py:85-94):
// every value as " %.{prec}E" (or " % .{prec}E"), `per_line` values per line, a line break
// at the end of every row.  Host code: the conversion is glibc's correctly rounded printf
// (the same digits Python's format() produces), run on all host threads; a writer is
// bound by formatting, not by the device.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include <cmath>
#include <cstring>

#include "parse_num.h"   // the 5^m table and bit helpers (host side)

namespace bdr {

// " %.{prec}E" / " % .{prec}E" of one double without printf: the exact value m * 2^e is
// scaled by 10^s (s = prec - floor(log10|x|) >= 0) as the multi-word integer m * 5^s shifted by
// e + s bits, rounded to nearest-even on the exact remainder -- the digits glibc's printf and
// Python's format() print.  Returns the length written, or 0 when the value is outside this
// path (zero / inf / nan, |x| >= 10^(prec+1), s > BDR_POW5_MAX): the caller uses snprintf.
inline int format_e_exact(char *out, double x, int prec, int sign_space) {
    uint64_t bits;
    memcpy(&bits, &x, 8);
    const int be = (int)((bits >> 52) & 0x7ff);
    uint64_t m = bits & ((1ULL << 52) - 1);
    if (be == 0x7ff || (be == 0 && m == 0) || prec > 17) return 0;
    int e;
    if (be == 0) {
        e = -1074;
    } else {
        m |= 1ULL << 52;
        e = be - 1075;
    }
    // floor(log10|x|) estimate from the binary exponent (off by at most one; fixed below)
    const int b2 = e + bitlen64(m) - 1;                       // floor(log2|x|)
    int k = (int)std::floor(b2 * 0.30102999566398120);
    uint64_t N = 0;
    const uint64_t lo_lim = (uint64_t)std::llround(std::pow(10.0, prec));   // 10^prec (exact for prec <= 17)
    for (int attempt = 0; attempt < 3; ++attempt) {
        const int s = prec - k;
        if (s < 0 || s > BDR_POW5_MAX) return 0;
        // A = m * 5^s, five 64-bit words
        uint64_t A[6] = {0, 0, 0, 0, 0, 0};
        unsigned __int128 carry = 0;
        for (int i = 0; i < 4; ++i) {
            const unsigned __int128 p = (unsigned __int128)m * h_pow5[s][i] + carry;
            A[i] = (uint64_t)p;
            carry = p >> 64;
        }
        A[4] = (uint64_t)carry;
        const int sh = e + s;  // N = A * 2^sh
        if (sh >= 0) {
            if (sh > 63 || A[1] | A[2] | A[3] | A[4] || (sh && (A[0] >> (64 - sh)))) return 0;
            N = A[0] << sh;
        } else {
            const int r = -sh;               // shift right by r bits with round-to-nearest-even
            if (r > 320) return 0;
            const int ws = r >> 6, bs = r & 63;
            // the kept part must fit 64 bits
            uint64_t kept = 0;
            bool too_big = false;
            for (int i = 4; i >= 0; --i) {
                const int src = i;
                if (src < ws) continue;
                const int dst = src - ws;
                uint64_t lo = A[src] >> bs;
                uint64_t hi = (bs && src + 1 <= 5) ? (A[src + 1] << (64 - bs)) : 0;
                const uint64_t word = lo | hi;
                if (dst == 0) kept = word;
                else if (word) too_big = true;
            }
            if (too_big) return 0;
            // round bit and sticky from the r bits shifted out
            const int rb = r - 1;
            const uint64_t round_bit = (A[rb >> 6] >> (rb & 63)) & 1ULL;
            bool sticky = false;
            for (int i = 0; i < (rb >> 6); ++i) sticky |= A[i] != 0;
            if (rb & 63) sticky |= (A[rb >> 6] & ((1ULL << (rb & 63)) - 1)) != 0;
            N = kept + ((round_bit && (sticky || (kept & 1))) ? 1 : 0);
        }
        if (N >= lo_lim * 10) {
            ++k;
            continue;
        }
        if (N < lo_lim) {
            --k;
            continue;
        }
        break;
    }
    if (N < lo_lim || N >= lo_lim * 10) return 0;
    char *p = out;
    *p++ = ' ';
    if (bits >> 63) *p++ = '-';
    else if (sign_space) *p++ = ' ';
    char dig[24];
    for (int i = prec; i >= 0; --i) {
        dig[i] = (char)('0' + N % 10);
        N /= 10;
    }
    *p++ = dig[0];
    if (prec > 0) {
        *p++ = '.';
        memcpy(p, dig + 1, (size_t)prec);
        p += prec;
    }
    *p++ = 'E';
    int ak = k;
    if (k < 0) {
        *p++ = '-';
        ak = -k;
    } else {
        *p++ = '+';
    }
    if (ak >= 100) {
        *p++ = (char)('0' + ak / 100);
        ak %= 100;
        *p++ = (char)('0' + ak / 10);
        *p++ = (char)('0' + ak % 10);
    } else {
        *p++ = (char)('0' + ak / 10);
        *p++ = (char)('0' + ak % 10);
    }
    return (int)(p - out);
}

// rows x row_len values; x_fastest: value t of the single row is data[(x*ny + y)*nz + z] with
// t = (z*ny + y)*nx + x (CHGCAR order), otherwise data is read in C order.
inline int format_grid_append(const char *path, const double *data, int64_t nx, int64_t ny, int64_t nz,
                              int x_fastest, int64_t row_len, int per_line, int prec, int sign_space,
                              std::string *err) {
    const int64_t n = nx * ny * nz;
    if (n <= 0 || row_len <= 0 || n % row_len != 0 || per_line <= 0 || prec < 0 || prec > 17) {
        *err = "bdr_format_grid: bad argument";
        return 1;
    }
    FILE *f = fopen(path, "ab");
    if (!f) {
        *err = std::string("bdr_format_grid: cannot open ") + path;
        return 1;
    }
    const int64_t rows = n / row_len;
    const int64_t lines_per_row = (row_len + per_line - 1) / per_line;
    const int64_t total_lines = rows * lines_per_row;
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    const int64_t batch_lines = 1 << 18;  // lines formatted between two writes
    const char *fmt = sign_space ? " % .*E" : " %.*E";
    std::vector<std::string> bufs(hw);
    int rc = 0;
    for (int64_t l0 = 0; l0 < total_lines && rc == 0; l0 += batch_lines) {
        const int64_t l1 = std::min(total_lines, l0 + batch_lines);
        const int64_t per_thread = (l1 - l0 + hw - 1) / hw;
        std::vector<std::thread> ths;
        for (unsigned t = 0; t < hw; ++t) {
            const int64_t a = l0 + t * per_thread, b = std::min(l1, a + per_thread);
            bufs[t].clear();
            if (a >= b) continue;
            ths.emplace_back([&, t, a, b]() {
                std::string &s = bufs[t];
                s.reserve((size_t)(b - a) * per_line * (prec + 9));
                char tmp[64];
                for (int64_t line = a; line < b; ++line) {
                    const int64_t row = line / lines_per_row, k0 = (line % lines_per_row) * per_line;
                    const int64_t k1 = std::min<int64_t>(row_len, k0 + per_line);
                    for (int64_t k = k0; k < k1; ++k) {
                        const int64_t tkn = row * row_len + k;
                        double v;
                        if (x_fastest) {
                            const int64_t x = tkn % nx, y = (tkn / nx) % ny, z = tkn / (nx * ny);
                            v = data[(x * ny + y) * nz + z];
                        } else {
                            v = data[tkn];
                        }
                        int len = format_e_exact(tmp, v, prec, sign_space);
                        if (len == 0) len = snprintf(tmp, sizeof tmp, fmt, prec, v);
                        s.append(tmp, (size_t)len);
                    }
                    s.push_back('\n');
                }
            });
        }
        for (auto &th : ths) th.join();
        for (unsigned t = 0; t < hw; ++t)
            if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) {
                *err = "bdr_format_grid: write failed";
                rc = 1;
                break;
            }
    }
    fclose(f);
    return rc;
}

}  // namespace bdr
