// parse.cuh -- the numeric block of a CHGCAR / cube file, text -> fp64 grid, on the GPU
// (reference: io/vasp.py:90-137, io/cube.py:99-113, where numpy
// converts whitespace-separated tokens one Python string at a time).
//
//   k_tok_count   token starts per 1024-byte block (a start: non-space after space)
//   k_scan_*      exclusive scan of the block counts
//   k_tok_parse   every start parses its token (parse_num.h, correctly rounded) into
//                 vals[token index]; tokens it cannot do exactly are listed for the host
//   k_grid_finish token order -> C order [x][y][z] (CHGCAR writes x fastest) and the
//                 reader's one arithmetic operation (/ cell volume, * bohr^-3)
// Algorithmic traffic: R ~18 B (text) + W 8 per value, then R 8 + W 8.
#pragma once
#include "common.cuh"
#include "parse_num.h"

namespace bdr {

constexpr int TOK_BLOCK = 1024;  // bytes per CTA (256 threads x 4)

__device__ __forceinline__ unsigned tok_starts4(const char *__restrict__ text, int64_t n, int64_t i0,
                                                unsigned char prev) {
    // bit k set: byte i0 + k starts a token
    unsigned m = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = i0 + k;
        const unsigned char c = i < n ? (unsigned char)text[i] : (unsigned char)' ';
        if (!is_space(c) && is_space(prev)) m |= 1u << k;
        prev = c;
    }
    return m;
}

__global__ void __launch_bounds__(256)
k_tok_count(const char *__restrict__ text, int64_t n, unsigned *block_counts) {
    const int64_t i0 = (int64_t)blockIdx.x * TOK_BLOCK + 4 * threadIdx.x;
    const unsigned char prev = (i0 == 0 || i0 > n) ? (unsigned char)' ' : (unsigned char)text[i0 - 1];
    const unsigned c = i0 < n ? __popc(tok_starts4(text, n, i0, prev)) : 0u;
    __shared__ unsigned s_w[8];
    const unsigned wsum = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = wsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < 8; ++i) t += s_w[i];
        block_counts[blockIdx.x] = t;
    }
}

// exclusive scan of up to 1024 x 1024 counts: CTA-local scans, scan of the CTA totals, fix-up
__global__ void __launch_bounds__(1024)
k_scan_local(unsigned *v, int64_t n, unsigned *totals) {
    __shared__ unsigned s[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const unsigned x = i < n ? v[i] : 0u;
    s[threadIdx.x] = x;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned u = threadIdx.x >= o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += u;
        __syncthreads();
    }
    if (i < n) v[i] = s[threadIdx.x] - x;
    if (threadIdx.x == 1023) totals[blockIdx.x] = s[1023];
}
__global__ void __launch_bounds__(1024)
k_scan_add(unsigned *v, int64_t n, const unsigned *__restrict__ totals_scanned) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) v[i] += totals_scanned[blockIdx.x];
}

struct ParseOut {
    unsigned long long n_fallback;   // tokens handed to the host
    unsigned long long end_of_last;  // byte offset (in the whole text) just past token n_values - 1
};

__global__ void __launch_bounds__(256)
k_tok_parse(const char *__restrict__ text, int64_t n, const unsigned *__restrict__ block_off,
            int64_t token_base, int64_t byte_base, double *vals, int64_t n_values, ParseOut *po,
            int64_t *fallback, int64_t fallback_cap) {
    const int64_t i0 = (int64_t)blockIdx.x * TOK_BLOCK + 4 * threadIdx.x;
    const unsigned char prev = (i0 == 0 || i0 > n) ? (unsigned char)' ' : (unsigned char)text[i0 - 1];
    unsigned m = i0 < n ? tok_starts4(text, n, i0, prev) : 0u;
    const unsigned c = __popc(m);
    // rank of this thread's first token inside the CTA
    __shared__ unsigned s_w[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int i = 0; i < w; ++i) base += s_w[i];
    int64_t idx = token_base + block_off[blockIdx.x] + base + inc - c;
    while (m) {
        const int k = __ffs(m) - 1;
        m &= m - 1;
        if (idx < n_values) {
            double v;
            int len;
            const int st = parse_token(text + i0 + k, n - (i0 + k), &v, &len);
            if (st) {
                v = __longlong_as_double(0x7ff8000000000000LL);
                const unsigned long long o = atomicAdd(&po->n_fallback, 1ULL);
                if ((int64_t)o < fallback_cap) {
                    fallback[3 * o + 0] = idx;
                    fallback[3 * o + 1] = byte_base + i0 + k;
                    fallback[3 * o + 2] = len;
                }
            }
            vals[idx] = v;
            if (idx == n_values - 1) po->end_of_last = (unsigned long long)(byte_base + i0 + k + len);
        }
        ++idx;
    }
}

// token order -> C order [x][y][z] plus the reader's arithmetic.  x_fastest: token
// t = (z * ny + y) * nx + x (CHGCAR); otherwise token order is already C order (cube).
// op: 0 none, 1 divide by operand, 2 multiply by operand.
__global__ void __launch_bounds__(256)
k_grid_finish(const double *__restrict__ vals, double *out, int nx, int ny, int nz, int x_fastest,
              int op, double operand) {
    __shared__ double tile[32][33];
    if (!x_fastest) {
        const int64_t n = (int64_t)nx * ny * nz;
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
            const double v = vals[i];
            out[i] = op == 1 ? __ddiv_rn(v, operand) : (op == 2 ? __dmul_rn(v, operand) : v);
        }
        return;
    }
    // for one y: transpose the (z, x) matrix in 32 x 32 tiles; blockIdx.x -> (tile_x, tile_z), blockIdx.y -> y
    const int tiles_x = (nx + 31) / 32;
    const int tx0 = (blockIdx.x % tiles_x) * 32, tz0 = (blockIdx.x / tiles_x) * 32, y = blockIdx.y;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;  // 32 x 8 threads
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int z = tz0 + ly + 8 * r, x = tx0 + lx;
        if (z < nz && x < nx) tile[ly + 8 * r][lx] = vals[((int64_t)z * ny + y) * nx + x];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int x = tx0 + ly + 8 * r, z = tz0 + lx;
        if (x < nx && z < nz) {
            const double v = tile[lx][ly + 8 * r];
            out[((int64_t)x * ny + y) * nz + z] =
                op == 1 ? __ddiv_rn(v, operand) : (op == 2 ? __dmul_rn(v, operand) : v);
        }
    }
}

}  // namespace bdr
