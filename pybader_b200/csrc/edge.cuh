// edge.cuh -- the label half of refinement.edge_find (refinement.py:339-376)
// through equality bits.
//
// A non-vacuum voxel is an edge candidate when some non-vacuum voxel of its
// 27-neighbourhood carries another label.  When the neighbourhood holds no
// vacuum voxel this is simply "the 27 labels are not all equal", and 27 labels
// are all equal iff 26 suitably chosen adjacent pairs are: the z-pairs of the
// nine rows, the y-pairs of the three planes at the centre column, and the two
// x-pairs at the centre.  So the streaming kernel only compares every voxel
// with its +z, +y and +x neighbour (three ISETP + three VOTE per voxel; the
// warp votes *are* the 32-voxel words of the bit volumes) and a per-word kernel
// ANDs 26 words of those bit volumes together.  Voxels that do have a vacuum
// voxel within Chebyshev distance 1 (the skin of the vacuum region; found by
// dilating the vacuum bits) cannot use the chain argument -- vacuum neighbours
// are ignored by the reference (refinement.py:370-371), not compared -- and
// are evaluated exactly from a small deferred list.
// Result: the same ebits / vbits as the min/max kernel k_edge_bits (kernels.cuh),
// at ~30 instead of ~67 issue slots per voxel (ncu).
// Algorithmic traffic: R 4 (labels) + W 4/8 (four bit volumes) per voxel, then
// R ~27/8 + W 1/8 per voxel for the word pass.
#pragma once
#include "kernels.cuh"

namespace bdr {

// bit z of word j of row (x,y):  eqz: L[z] == L[z+1],  eqy: L[y] == L[y+1],
// eqx: L[x] == L[x+1] (all periodic), vbits: L == -1.  Bits past nz are 0.
// A warp owns one y row and WPT consecutive words of it and marches along x.
// ALLOK: every voxel of the warp's segment exists (no predicates in the loop)
template <int WPT, int CX, bool ALLOK>
__device__ __forceinline__ void label_eq_bits_body(const int32_t *__restrict__ lab, const Grid &g, int nzw,
                                                   uint32_t *__restrict__ eqz, uint32_t *__restrict__ eqy,
                                                   uint32_t *__restrict__ eqx, uint32_t *__restrict__ vbits,
                                                   unsigned long long *vac_seen, int lane, int j0, int y,
                                                   int x0, int x_end) {
    constexpr unsigned FULL = 0xffffffffu;
    const int nplanes = min(CX, x_end - x0);
    const int plane = g.ny * g.nz;
    const int z0 = 32 * j0 + lane;  // word i of the segment holds voxel z0 + 32 i of this lane
    bool ok[WPT];
    uint32_t vm[WPT];
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
        ok[i] = ALLOK || (j0 + i < nzw && z0 + 32 * i < g.nz);
        vm[i] = ALLOK ? FULL : __ballot_sync(FULL, ok[i]);
    }
    // the +z neighbour is the next lane's voxel, except after the thread segment's
    // last voxel and after the row's last voxel (periodic wrap): those lanes load it
    int sp_i = -1, sp_z = 0;
#pragma unroll
    for (int i = 0; i < WPT; ++i) {
        const int z = z0 + 32 * i;
        if (ok[i] && (z + 1 == g.nz || (i == WPT - 1 && lane == 31))) {
            sp_i = i;
            sp_z = z + 1 == g.nz ? 0 : z + 1;
        }
    }
    // all addresses of a plane are p + a warp-uniform or per-thread constant
    const int32_t *p = lab + ((int64_t)x0 * plane + y * g.nz + (ok[0] ? z0 : 0));
    const int d_up = ((y + 1 == g.ny ? 0 : y + 1) - y) * g.nz;
    const int d_sp = sp_z - (ok[0] ? z0 : 0);
    int32_t cur[WPT], nxt[WPT];
#pragma unroll
    for (int i = 0; i < WPT; ++i) cur[i] = ok[i] ? p[32 * i] : -1;
    {
        const int d_next = (x0 + 1 == g.nx ? -x0 : 1) * plane;
#pragma unroll
        for (int i = 0; i < WPT; ++i) nxt[i] = ok[i] ? p[d_next + 32 * i] : -1;
    }
    const bool vec_store = WPT == 4 && (nzw & 3) == 0;
    unsigned anyv = 0;
    int wd = (x0 * g.ny + y) * nzw + j0;
    const int wd_step = g.ny * nzw;
#pragma unroll 2
    for (int pl = 0; pl < nplanes; ++pl) {
        const int x = x0 + pl;
        // plane x+2 is fetched while plane x is compared, so two planes of loads are in flight
        int xn2 = x + 2;
        if (xn2 >= g.nx) xn2 -= g.nx;
        if (xn2 >= g.nx) xn2 -= g.nx;
        const int32_t *q = p + (int64_t)(xn2 - x) * plane;
        int32_t nx2[WPT], up[WPT];
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
            up[i] = ok[i] ? p[d_up + 32 * i] : -1;
            nx2[i] = ok[i] ? q[32 * i] : -1;
        }
        const int32_t zv = sp_i >= 0 ? p[d_sp] : 0;
        uint32_t bz[WPT], by[WPT], bx[WPT], bv[WPT];
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
            int32_t nb = __shfl_down_sync(FULL, cur[i], 1);
            if (i + 1 < WPT) {
                const int32_t first = __shfl_sync(FULL, cur[i + 1], 0);
                if (lane == 31) nb = first;
            }
            if (sp_i == i) nb = zv;
            bz[i] = __ballot_sync(FULL, cur[i] == nb) & vm[i];
            by[i] = __ballot_sync(FULL, cur[i] == up[i]) & vm[i];
            bx[i] = __ballot_sync(FULL, cur[i] == nxt[i]) & vm[i];
            bv[i] = __ballot_sync(FULL, cur[i] == -1) & vm[i];
            anyv |= bv[i];
        }
        if (lane == 0) {
            if (vec_store) {
                *reinterpret_cast<uint4 *>(eqz + wd) = make_uint4(bz[0], bz[1], bz[2], bz[3]);
                *reinterpret_cast<uint4 *>(eqy + wd) = make_uint4(by[0], by[1], by[2], by[3]);
                *reinterpret_cast<uint4 *>(eqx + wd) = make_uint4(bx[0], bx[1], bx[2], bx[3]);
                *reinterpret_cast<uint4 *>(vbits + wd) = make_uint4(bv[0], bv[1], bv[2], bv[3]);
            } else {
#pragma unroll
                for (int i = 0; i < WPT; ++i)
                    if (j0 + i < nzw) {
                        eqz[wd + i] = bz[i];
                        eqy[wd + i] = by[i];
                        eqx[wd + i] = bx[i];
                        vbits[wd + i] = bv[i];
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < WPT; ++i) {
            cur[i] = nxt[i];
            nxt[i] = nx2[i];
        }
        p += plane;  // never past the grid: the loads of planes x+1, x+2 wrapped above
        wd += wd_step;
    }
    if (anyv && lane == 0) atomicOr(vac_seen, 1ULL);
}

template <int WPT, int CX>
__global__ void __launch_bounds__(256)
k_label_eq_bits(const int32_t *__restrict__ lab, Grid g, int nzw, uint32_t *__restrict__ eqz,
                uint32_t *__restrict__ eqy, uint32_t *__restrict__ eqx,
                uint32_t *__restrict__ vbits, unsigned long long *vac_seen, int x_begin, int x_end) {
    // planes [x_begin, x_end): the whole grid, or the few planes next to a slab's halos whose
    // labels an exchange has just replaced
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int j0 = blockIdx.x * WPT, y = blockIdx.y * 8 + w, x0 = x_begin + blockIdx.z * CX;
    if (y >= g.ny || x0 >= x_end) return;  // the whole warp; the kernel has no barriers
    if (32 * (j0 + WPT) <= g.nz)
        label_eq_bits_body<WPT, CX, true>(lab, g, nzw, eqz, eqy, eqx, vbits, vac_seen, lane, j0, y, x0, x_end);
    else
        label_eq_bits_body<WPT, CX, false>(lab, g, nzw, eqz, eqy, eqx, vbits, vac_seen, lane, j0, y, x0, x_end);
}

// Incremental maintenance of the four bit volumes: after a trace relabelled a few voxels, only
// the bits that compare one of them with a neighbour can have changed -- eqz at z-1 and z, eqy
// at y-1 and y, eqx at x-1 and x, and the voxel's own vacuum bit.  One thread per relabelled
// voxel recomputes those seven bits from the current labels (two voxels that are neighbours
// both write the shared bit, with the same value).  Replaces a full R 4 B/voxel pass when the
// labels have not changed otherwise since the bits were made (renumbering keeps equalities).
__device__ __forceinline__ void put_bit(uint32_t *vol, const Grid &g, int nzw, int x, int y, int z, bool on) {
    uint32_t *w = vol + ((int64_t)x * g.ny + y) * nzw + (z >> 5);
    const uint32_t m = 1u << (z & 31);
    if (on) atomicOr(w, m);
    else atomicAnd(w, ~m);
}
__global__ void __launch_bounds__(128)
k_eq_update(const int32_t *__restrict__ lab, Grid g, int nzw, uint32_t *eqz, uint32_t *eqy, uint32_t *eqx,
            uint32_t *vbits, const int32_t *__restrict__ list, int64_t n) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int v = list[t];
    int x, y, z;
    unlin3(g, v, x, y, z);
    const int xm = x == 0 ? g.nx - 1 : x - 1, xp = x + 1 == g.nx ? 0 : x + 1;
    const int ym = y == 0 ? g.ny - 1 : y - 1, yp = y + 1 == g.ny ? 0 : y + 1;
    const int zm = z == 0 ? g.nz - 1 : z - 1, zp = z + 1 == g.nz ? 0 : z + 1;
    const int32_t l = lab[v];
    put_bit(eqz, g, nzw, x, y, z, l == lab[lin3(g, x, y, zp)]);
    put_bit(eqz, g, nzw, x, y, zm, l == lab[lin3(g, x, y, zm)]);
    put_bit(eqy, g, nzw, x, y, z, l == lab[lin3(g, x, yp, z)]);
    put_bit(eqy, g, nzw, x, ym, z, l == lab[lin3(g, x, ym, z)]);
    put_bit(eqx, g, nzw, x, y, z, l == lab[lin3(g, xp, y, z)]);
    put_bit(eqx, g, nzw, xm, y, z, l == lab[lin3(g, xm, y, z)]);
    put_bit(vbits, g, nzw, x, y, z, l == -1);
}

// Edge-candidate bits from the equality bits.  A thread owns one word position
// (y, j) and marches along x: per plane it folds the 3 x 3 (y,z) patch into one
// word P(x) (9 loads), and the candidate word of plane x is
// ~(P(x-1) & P(x) & P(x+1) & eqx(x-1) & eqx(x)).  The vacuum dilation is folded
// the same way.  Candidates whose neighbourhood touches vacuum go to the
// deferred list instead.
template <int CXB>
__global__ void __launch_bounds__(256)
k_edge_from_eq(const uint32_t *__restrict__ eqz, const uint32_t *__restrict__ eqy,
               const uint32_t *__restrict__ eqx, const uint32_t *__restrict__ vbits, Grid g, int nzw,
               uint32_t *__restrict__ ebits, const unsigned long long *__restrict__ vac_seen,
               unsigned long long *defer_cnt, int32_t *defer, int64_t defer_cap) {
    const int wpp = g.ny * nzw;  // words per plane; word indices fit 32 bits
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= wpp) return;
    const int y = t / nzw, j = t - y * nzw;
    const int nvalid = min(32, g.nz - 32 * j);
    const unsigned valid = nvalid == 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    const int jp = j == 0 ? nzw - 1 : j - 1, jn = j + 1 == nzw ? 0 : j + 1;
    const int sp = jp == nzw - 1 ? ((g.nz - 1) & 31) : 31;  // last valid bit of the previous word
    const int rows[3] = {(y == 0 ? g.ny - 1 : y - 1) * nzw, y * nzw, (y + 1 == g.ny ? 0 : y + 1) * nzw};
    const int x0 = blockIdx.y * CXB;
    const int nplanes = min(CXB, g.nx - x0);
    const bool vac_any = *vac_seen != 0ULL;
    auto fold = [&](int x, unsigned &P, unsigned &V) {
        const uint32_t *ez = eqz + x * wpp, *ey = eqy + x * wpp;
        unsigned a = ey[rows[0] + j] & ey[rows[1] + j];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const unsigned wz = ez[rows[b] + j], carry = (ez[rows[b] + jp] >> sp) & 1u;
            a &= wz & ((wz << 1) | carry);  // bit z: L[z-1] == L[z] == L[z+1] in this row
        }
        P = a;
        V = 0;
        if (vac_any) {
            const uint32_t *vb = vbits + x * wpp;
            unsigned m = 0, lc = 0, rc = 0;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                m |= vb[rows[b] + j];
                lc |= (vb[rows[b] + jp] >> sp) & 1u;
                rc |= vb[rows[b] + jn] & 1u;
            }
            V = (m | (m << 1) | (m >> 1) | lc | (rc << (nvalid - 1))) & valid;
        }
    };
    const int xm0 = x0 == 0 ? g.nx - 1 : x0 - 1;
    unsigned Pm, Pc, Pn, Vm, Vc, Vn;
    fold(xm0, Pm, Vm);
    fold(x0, Pc, Vc);
    unsigned Xm = eqx[xm0 * wpp + rows[1] + j];
    for (int pl = 0; pl < nplanes; ++pl) {
        const int x = x0 + pl, xn = x + 1 == g.nx ? 0 : x + 1;
        fold(xn, Pn, Vn);
        const int wid = x * wpp + rows[1] + j;
        const unsigned Xc = eqx[wid];
        const unsigned all = Pm & Pc & Pn & Xm & Xc;
        unsigned cand = ~all & ~vbits[wid] & valid;
        if (vac_any) {
            unsigned dirty = cand & (Vm | Vc | Vn);
            cand &= ~dirty;
            if (dirty) {
                const int n = __popc(dirty);
                const unsigned long long base = atomicAdd(defer_cnt, (unsigned long long)n);
                int64_t pos = (int64_t)base;
                const int v0 = (x * g.ny + y) * g.nz + 32 * j;
                while (dirty) {
                    const int bit = __ffs(dirty) - 1;
                    dirty &= dirty - 1;
                    if (pos < defer_cap) defer[pos] = v0 + bit;
                    ++pos;
                }
            }
        }
        ebits[wid] = cand;
        Pm = Pc; Pc = Pn;
        Vm = Vc; Vc = Vn;
        Xm = Xc;
    }
}

// exact evaluation of the deferred voxels (non-vacuum, vacuum next to them):
// an edge candidate iff a non-vacuum neighbour carries another label
__global__ void __launch_bounds__(128)
k_edge_deferred(const int32_t *__restrict__ lab, Grid g, int nzw, uint32_t *ebits,
                const int32_t *__restrict__ defer, const unsigned long long *__restrict__ defer_cnt,
                int64_t defer_cap) {
    const int64_t n = min((int64_t)*defer_cnt, defer_cap);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int v = defer[t];
        int x, y, z;
        unlin3(g, v, x, y, z);
        const int32_t mine = lab[v];
        bool edge = false;
        for (int ix = -1; ix <= 1 && !edge; ++ix) {
            const int tx = wrap1(x + ix, g.nx);
            for (int iy = -1; iy <= 1 && !edge; ++iy) {
                const int ty = wrap1(y + iy, g.ny);
                for (int iz = -1; iz <= 1; ++iz) {
                    const int32_t l = lab[lin3(g, tx, ty, wrap1(z + iz, g.nz))];
                    edge |= l != -1 && l != mine;
                }
            }
        }
        if (edge) atomicOr(ebits + ((int64_t)x * g.ny + y) * nzw + (z >> 5), 1u << (z & 31));
    }
}

}  // namespace bdr
