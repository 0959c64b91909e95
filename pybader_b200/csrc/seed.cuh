// seed.cuh -- the pointer-jumpable seed of bader_calc('neargrid') and the
// tile-ordered pointer resolution shared by both methods.
//
// bader_calc('neargrid') only needs *a* strictly-ascending pointer field whose
// terminal voxels are exactly the ongrid maxima (methods.py:169-177): the
// labels it seeds are then driven to the fixed point of the reference's own
// refinement iteration, which does not depend on which
// ascending neighbour a seed pointer chose.  k_ongrid_pointers (kernels.cuh)
// pays ~340 issue slots per voxel for the bit-exact fp64 argmax the 'ongrid'
// method needs; the seed kernel here ranks the neighbours in fp32 instead:
//
//   * the density tile is converted to fp32 once while it is staged in shared
//     memory (half the shared memory, 64-bit loads feed two voxels);
//   * neighbours k and 26-k share a step weight, so only the larger of each
//     pair is scored: 13 x (FMNMX, FADD, FMUL, LOP3, FMNMX) per voxel, the
//     pair index riding in the four low mantissa bits of the score;
//   * the winner is accepted only when it is *clearly* uphill in fp32
//     (difference >= 2 fp32 ulps of the centre), which implies the exact fp64
//     ongrid criterion (rho_n - rho_c) * w + rho_c > rho_c for that neighbour;
//     anything else -- maxima, plateaus, denormal densities -- takes the exact
//     fp64 27-point evaluation of methods.py:87-117 from global memory.
//
// So every pointer goes strictly uphill in the exact density (the field is
// acyclic), and a voxel is a maximum here iff the reference's ongrid step
// says so.  All decisions are functions of the voxel's own neighbourhood, so
// slab windows of a sharded run produce the same pointers as one GPU.
#pragma once
#include "kernels.cuh"

namespace bdr {

constexpr int FX = 12, FY = 8, FZ = 64;  // seed tile: 256 threads x (2 z voxels) x 12 planes
constexpr int F_HX = FX + 2, F_HY = FY + 2, F_RS = 68, F_TILE = FX * FY * FZ;
constexpr size_t seed_smem() {
    return (size_t)F_HX * F_HY * F_RS * sizeof(float) + (size_t)F_TILE * sizeof(int32_t);
}

struct SeedWeights {
    float w[13];   // step weights of offsets k = 0..12 (== those of 26-k)
    float c1;      // "clearly uphill" threshold on the score, relative to |rho_c|
    float floor_;  // and its absolute floor (denormal densities go the exact way)
    unsigned tag_mask;  // 0xfffffff0, passed in so the tagging stays one LOP3 per score
};

// float bits -> unsigned with the same order (negative densities included)
__device__ __forceinline__ unsigned sortable_key(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}

// the exact fp64 step (methods.py:87-117), kept out of line: it runs for a
// handful of voxels per tile; the 27 step weights come from global memory
__device__ __noinline__ int seed_exact_step(const double *__restrict__ rho, Grid g,
                                            const double *__restrict__ w, int x, int y, int z) {
    const int self = lin3(g, x, y, z);
    const double rc = rho[self];
    double best = rc;
    int bi = self;
#pragma unroll 1
    for (int ix = -1; ix <= 1; ++ix) {
        const int tx = wrap1(x + ix, g.nx);
#pragma unroll 1
        for (int iy = -1; iy <= 1; ++iy) {
            const int ty = wrap1(y + iy, g.ny);
#pragma unroll
            for (int iz = -1; iz <= 1; ++iz) {
                const int q = lin3(g, tx, ty, wrap1(z + iz, g.nz));
                const double v = __dadd_rn(
                    __dmul_rn(__dsub_rn(rho[q], rc), w[(ix + 1) * 9 + (iy + 1) * 3 + (iz + 1)]), rc);
                if (v > best) {
                    best = v;
                    bi = q;
                }
            }
        }
    }
    return bi;
}

template <int PH, int ZO>
__device__ __forceinline__ float seed_best(const float (&P)[3][12], const SeedWeights &Wf, float rc,
                                           unsigned mask) {
    constexpr int A = PH % 3, B = (PH + 1) % 3, C = (PH + 2) % 3;
    float best = 0.f;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const int a = k / 9, b = (k / 3) % 3, c = k % 3;
        const int pa = a == 0 ? A : (a == 1 ? B : C);
        const int pb = a == 0 ? C : (a == 1 ? B : A);
        const float m = fmaxf(P[pa][b * 4 + c + ZO], P[pb][(2 - b) * 4 + (2 - c) + ZO]);
        float s = __fmul_rn(__fsub_rn(m, rc), Wf.w[k]);
        unsigned tagged;  // (bits & mask) | k in one LOP3
        asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(tagged) : "r"(__float_as_uint(s)), "r"(mask), "r"(k));
        best = fmaxf(best, __uint_as_float(tagged));  // NaN (both of a pair are vacuum) never wins
    }
    return best;
}

// Codes in the shared-memory tile are byte offsets while they point inside the
// tile (4 * tile index), so one hop of the in-tile chase is LDS + compare:
//     0 <= c < 4*F_TILE   pointer to the tile voxel at byte offset c
//     c >= 4*F_TILE       4*F_TILE + global index of a voxel outside the tile
//     c < 0               terminal (vacuum / maximum slot / exit slot)
constexpr int F_T4 = 4 * F_TILE;

template <int VAC, bool SLAB>
__global__ void __launch_bounds__(256, 3)
k_seed_pointers(const double *__restrict__ rho, int32_t *code, Grid g, SeedWeights Wf,
                const double *__restrict__ W,
                double vac_tol, unsigned long long *root_counter, int32_t *roots,
                int64_t roots_cap, int exit_base, int x_begin, uint32_t *tile_keys) {
    extern __shared__ float s_f[];  // [F_HX][F_HY][F_RS]: z0-1 at column 0, body 1..64, z0+64 at 65
    int32_t *s_code = reinterpret_cast<int32_t *>(s_f + F_HX * F_HY * F_RS);
    __shared__ TileIdx<1, FX, FY, FZ> idx;
    __shared__ int s_off[13];   // fp32-tile offset of neighbour k (k < 13; 26-k is the negative)
    __shared__ int s_t27[27];   // per move: code-tile byte delta << 8 | a | b << 2 | c << 4
    __shared__ unsigned s_key;
    const int x0 = x_begin + blockIdx.z * FX, y0 = blockIdx.y * FY, z0 = blockIdx.x * FZ;
    tile_index_tables(idx, g, x0, y0, z0);
    if (threadIdx.x < 27) {
        const int k = threadIdx.x, a = k / 9, b = (k / 3) % 3, c = k % 3;
        if (k < 13) s_off[k] = ((a - 1) * F_HY + (b - 1)) * F_RS + (c - 1);
        const int dl = 4 * (((a - 1) * FY + (b - 1)) * FZ + (c - 1));
        s_t27[k] = dl * 256 + (a | (b << 2) | (c << 4));
    }
    if (threadIdx.x == 0) s_key = 0u;
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane = g.ny * g.nz;
    {   // ---- stage the tile as fp32: a warp per (x,y) row, a lane per z pair; warp w
        // takes row y = w of every x plane, the two halo rows y = 8, 9 are dealt round ----
        const bool vec = ((g.nz & 1) == 0) && (z0 + FZ <= g.nz);
        const int zb0 = idx.zi[2 * lane + 1], zb1 = idx.zi[2 * lane + 2];
        const int zh = idx.zi[lane == 0 ? 0 : FZ + 1];
        const int ch = lane == 0 ? 0 : FZ + 1;
        auto rows = [&](auto lx_of, auto ly_of, int count) {
            constexpr int U = 7;
            for (int i0 = 0; i0 < count; i0 += U) {
                double va[U], vb[U], vh[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u < count) {
                        const double *row = rho + (idx.xi[lx_of(i0 + u)] * plane + idx.yi[ly_of(i0 + u)] * g.nz);
                        if (vec) {
                            const double2 q = *reinterpret_cast<const double2 *>(row + zb0);
                            va[u] = q.x;
                            vb[u] = q.y;
                        } else {
                            va[u] = row[zb0];
                            vb[u] = row[zb1];
                        }
                        if (lane < 2) vh[u] = row[zh];
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (i0 + u < count) {
                        float *dst = s_f + (lx_of(i0 + u) * F_HY + ly_of(i0 + u)) * F_RS;
                        float fa = __double2float_rn(va[u]), fb = __double2float_rn(vb[u]);
                        if (VAC == VAC_TOL) {
                            if (va[u] <= vac_tol) fa = -INFINITY;
                            if (vb[u] <= vac_tol) fb = -INFINITY;
                        }
                        dst[2 * lane + 1] = fa;
                        dst[2 * lane + 2] = fb;
                        if (lane < 2) {
                            float fh = __double2float_rn(vh[u]);
                            if (VAC == VAC_TOL && vh[u] <= vac_tol) fh = -INFINITY;
                            dst[ch] = fh;
                        }
                    }
            }
        };
        rows([](int i) { return i; }, [&](int) { return warp; }, F_HX);
        // 28 halo rows (x plane e >> 1, y row 8 + (e & 1)), e = warp, warp + 8, ...
        rows([&](int i) { return (warp + 8 * i) >> 1; }, [&](int i) { return 8 + ((warp + 8 * i) & 1); },
             (2 * F_HX - warp + 7) / 8);
    }
    __syncthreads();

    const int ty = warp, tz0 = 2 * lane;
    const int gy = y0 + ty, gz0 = z0 + tz0;
    const bool ok0 = gy < g.ny && gz0 < g.nz, ok1 = gy < g.ny && gz0 + 1 < g.nz;
    // extent of the tile inside the grid: a move stays in the tile iff it stays below these
    const unsigned ex = min(FX, g.nx - x0), ey = min(FY, g.ny - y0), ez = min(FZ, g.nz - z0);
    // vacuum flags of the two columns up front (bit tx: voxel A, bit 16+tx: voxel B)
    unsigned vac = 0;
    if (VAC == VAC_LABELS) {
#pragma unroll
        for (int tx = 0; tx < FX; ++tx)
            if (x0 + tx < g.nx) {
                const int v = lin3(g, x0 + tx, gy, gz0);
                if (ok0) vac |= (code[v] == -1 ? 1u : 0u) << tx;
                if (ok1) vac |= (code[v + 1] == -1 ? 1u : 0u) << (16 + tx);
            }
    }

    // P[p][r*4+j]: register plane p, row r (y-1..y+1), column j (z0-1 .. z0+2 of the pair)
    float P[3][12];
    const float *col = s_f + ty * F_RS + 2 * lane;
    auto load_plane = [&](float (&Q)[12], int p) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2 q0 = *reinterpret_cast<const float2 *>(col + (p * F_HY + r) * F_RS);
            const float2 q1 = *reinterpret_cast<const float2 *>(col + (p * F_HY + r) * F_RS + 2);
            Q[r * 4 + 0] = q0.x; Q[r * 4 + 1] = q0.y; Q[r * 4 + 2] = q1.x; Q[r * 4 + 3] = q1.y;
        }
    };
    load_plane(P[0], 0);
    load_plane(P[1], 1);
    float colmax = -INFINITY;
    unsigned mask;
    asm volatile("mov.b32 %0, %1;" : "=r"(mask) : "r"(Wf.tag_mask));
    const int eb0 = 4 * (ty * FZ + tz0);  // byte offset of voxel A of plane 0 in the code tile

    // one voxel: turn the fp32 winner (or the exact fallback) into a pointer code
    auto finish = [&](float best, float rc, int j, int tx, bool ok) {
        const int tz = tz0 + j, gx = x0 + tx;
        const int eb = eb0 + 4 * j + tx * (4 * FY * FZ);
        int32_t cde = -1;
        if (ok && gx < g.nx) {
            const bool is_vac = VAC == VAC_LABELS ? ((vac >> (16 * j + tx)) & 1u) != 0
                                                  : (VAC == VAC_TOL ? rc == -INFINITY : false);
            const bool is_exit = SLAB && (gx == 0 || gx == g.nx - 1);
            if (is_exit) {
                cde = -2 - ((gx == 0 ? 0 : plane) + gy * g.nz + gz0 + j);
            } else if (!is_vac) {
                colmax = fmaxf(colmax, rc);
                const float thr = fmaxf(fabsf(rc) * Wf.c1, Wf.floor_);
                if (best >= thr) {
                    const int k = __float_as_int(best) & 15;
                    const float *cp = col + ((tx + 1) * F_HY + 1) * F_RS + 1 + j;
                    const int off = s_off[k];
                    const int bk = cp[off] >= cp[-off] ? k : 26 - k;
                    const int t = s_t27[bk];
                    const unsigned a = t & 3, b = (t >> 2) & 3, c = (t >> 4) & 3;
                    // tile coordinates of the target, -1 .. extent (as unsigned: -1 is huge)
                    if (a + (unsigned)(tx - 1) < ex && b + (unsigned)(ty - 1) < ey &&
                        c + (unsigned)(tz - 1) < ez)
                        cde = eb + (t >> 8);
                    else
                        cde = F_T4 + lin3(g, idx.xi[tx + a], idx.yi[ty + b], idx.zi[tz + c]);
                } else {
                    // not clearly uphill in fp32: the reference's own step decides
                    const int self = lin3(g, gx, gy, gz0 + j);
                    const int tl = seed_exact_step(rho, g, W, gx, gy, gz0 + j);
                    if (tl == self) {
                        const unsigned long long s = atomicAdd(root_counter, 1ULL);
                        if ((int64_t)s < roots_cap) roots[s] = self;
                        cde = -2 - (exit_base + (int32_t)s);
                    } else {
                        cde = F_T4 + tl;
                    }
                }
            }
        }
        *reinterpret_cast<int32_t *>(reinterpret_cast<char *>(s_code) + eb) = cde;
    };

#pragma unroll 1
    for (int tx = 0; tx < FX; tx += 3) {
        load_plane(P[2], tx + 2);
        finish(seed_best<0, 0>(P, Wf, P[1][5], mask), P[1][5], 0, tx, ok0);
        finish(seed_best<0, 1>(P, Wf, P[1][6], mask), P[1][6], 1, tx, ok1);
        load_plane(P[0], tx + 3);
        finish(seed_best<1, 0>(P, Wf, P[2][5], mask), P[2][5], 0, tx + 1, ok0);
        finish(seed_best<1, 1>(P, Wf, P[2][6], mask), P[2][6], 1, tx + 1, ok1);
        load_plane(P[1], tx + 4 < F_HX ? tx + 4 : F_HX - 1);
        finish(seed_best<2, 0>(P, Wf, P[0][5], mask), P[0][5], 0, tx + 2, ok0);
        finish(seed_best<2, 1>(P, Wf, P[0][6], mask), P[0][6], 1, tx + 2, ok1);
    }
    // tile key for the resolve order: the largest density of the tile
    if (tile_keys) {
        const unsigned km = __reduce_max_sync(0xffffffffu, sortable_key(colmax));
        if (lane == 0) atomicMax(&s_key, km);
    }
    __syncthreads();
    if (tile_keys && threadIdx.x == 0)
        tile_keys[((x0 / FX) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s_key;
    if (!ok0) return;
    // chase every pointer inside the tile (the pair's two chains in lock step), store the pair
    const bool pair_store = ok1 && ((g.nz & 1) == 0);
    const char *sc = reinterpret_cast<const char *>(s_code);
    int v = lin3(g, x0, gy, gz0);
#pragma unroll 2
    for (int tx = 0; tx < FX; ++tx, v += plane) {
        if (x0 + tx >= g.nx) break;
        int2 *slot = reinterpret_cast<int2 *>(reinterpret_cast<char *>(s_code) + eb0 + tx * (4 * FY * FZ));
        int2 c = *slot;
        if ((unsigned)c.x < (unsigned)F_T4 || (unsigned)c.y < (unsigned)F_T4) {
            do {
                if ((unsigned)c.x < (unsigned)F_T4) c.x = *reinterpret_cast<const int32_t *>(sc + c.x);
                if ((unsigned)c.y < (unsigned)F_T4) c.y = *reinterpret_cast<const int32_t *>(sc + c.y);
            } while ((unsigned)c.x < (unsigned)F_T4 || (unsigned)c.y < (unsigned)F_T4);
            *slot = c;  // path compression for the voxels that point here
        }
        c.x = c.x >= F_T4 ? c.x - F_T4 : c.x;
        c.y = c.y >= F_T4 ? c.y - F_T4 : c.y;
        if (pair_store) {
            *reinterpret_cast<int2 *>(code + v) = c;
        } else {
            code[v] = c.x;
            if (ok1) code[v + 1] = c.y;
        }
    }
}

// -------------------------------------------------------------------------
// K2  tile-ordered pointer resolution.
//
// After the stencil pass every voxel holds a terminal code or the index of
// the voxel where its path enters another tile.  Tiles are visited in order of
// decreasing largest density (16-bit counting sort of the keys the stencil
// pass left behind): an ascent path only enters tiles that hold larger
// densities, and those were resolved thousands of CTAs earlier, so almost
// every chain is one hop long and the pass is one coalesced read, one gather
// and one coalesced write per voxel instead of a ~10-hop chase.
// Correctness does not depend on the order: a reader that meets an unresolved
// pointer simply keeps following the (acyclic) chain.
// Algorithmic traffic: R 4 + W 4 per voxel (+ the gathers, mostly L1/L2 hits).
// -------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tile_hist(const uint32_t *__restrict__ keys, int n, unsigned *hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + (65535u - (keys[i] >> 16)), 1u);
}
// exclusive scan of the 65536 bins by one CTA of 1024 threads
__global__ void __launch_bounds__(1024)
k_tile_scan(unsigned *hist) {
    __shared__ unsigned s_part[1024];
    const int t = threadIdx.x;
    unsigned loc[64], sum = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        loc[i] = hist[t * 64 + i];
        sum += loc[i];
    }
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned u = t >= o ? s_part[t - o] : 0u;
        __syncthreads();
        s_part[t] += u;
        __syncthreads();
    }
    unsigned run = s_part[t] - sum;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        hist[t * 64 + i] = run;
        run += loc[i];
    }
}
__global__ void __launch_bounds__(256)
k_tile_scatter(const uint32_t *__restrict__ keys, int n, unsigned *offsets, int32_t *order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(offsets + (65535u - (keys[i] >> 16)), 1u)] = i;
}

// one CTA per tile of TX x TY x TZ voxels (the stencil pass's tiling), in the
// given order.  VEC: nz % 4 == 0, so the four voxels of a thread move as one
// 16-byte access.
template <int TX, int TY, int TZ, bool VEC>
__global__ void __launch_bounds__(256)
k_resolve_tiles(int32_t *code, Grid g, const int32_t *__restrict__ order, int tiles_z, int tiles_y,
                int32_t *minidx) {
    constexpr int QPR = TZ / 4;          // 16-byte groups per row
    constexpr int ROWS = TX * TY;
    constexpr int RPP = 256 / QPR;       // rows per pass of the CTA
    constexpr int B = 3;                 // rows in flight per thread
    const int tile = order[blockIdx.x];
    const int bz = tile % tiles_z, by = (tile / tiles_z) % tiles_y, bx = tile / (tiles_z * tiles_y);
    const int q = threadIdx.x % QPR, r_in = threadIdx.x / QPR;
    const int z = bz * TZ + 4 * q;
    if (z >= g.nz) return;
    for (int rb = r_in; rb < ROWS; rb += RPP * B) {
        int32_t c[B][4];
        int v[B];
#pragma unroll
        for (int u = 0; u < B; ++u) {
            const int r = rb + u * RPP;
            const int lx = r / TY, ly = r - lx * TY;
            const int x = bx * TX + lx, y = by * TY + ly;
            v[u] = -1;
            if (r < ROWS && x < g.nx && y < g.ny) {
                v[u] = lin3(g, x, y, z);
                if (VEC) {
                    const int4 t = *reinterpret_cast<const int4 *>(code + v[u]);
                    c[u][0] = t.x; c[u][1] = t.y; c[u][2] = t.z; c[u][3] = t.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) c[u][k] = z + k < g.nz ? code[v[u] + k] : -1;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) c[u][k] = -1;
            }
        }
        // first hop of every chain together, then the (rare) rest
#pragma unroll
        for (int u = 0; u < B; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c[u][k] >= 0) c[u][k] = __ldca(code + c[u][k]);
#pragma unroll
        for (int u = 0; u < B; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                while (c[u][k] >= 0) c[u][k] = __ldcg(code + c[u][k]);
#pragma unroll
        for (int u = 0; u < B; ++u) {
            if (v[u] < 0) continue;
            if (VEC) {
                *reinterpret_cast<int4 *>(code + v[u]) = make_int4(c[u][0], c[u][1], c[u][2], c[u][3]);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (z + k < g.nz) code[v[u] + k] = c[u][k];
            }
            if (minidx) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (c[u][k] <= -2 && z + k < g.nz) {
                        const int s = -2 - c[u][k];
                        if (v[u] + k < minidx[s]) atomicMin(minidx + s, v[u] + k);
                    }
            }
        }
    }
}

}  // namespace bdr
