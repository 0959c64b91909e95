/*
 * oracle/bader_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the pybader (v0.3.12) hot path:
 * the algorithm of pybader/methods.py, pybader/refinement.py and the jitted
 * helpers of pybader/utils.py, restricted to the "one brick == whole volume"
 * case (threads=1, idx = 0) so none of the brick-growth plumbing
 * (methods.py:118-164) exists here.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's .so;
 * the shipped CUDA path never does.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function below against
 * fixtures in tests/golden/ that were produced by running the real reference
 * (numba) in the build container -- see tests/golden/make_golden.py.
 *
 * Arithmetic contract: numba emits neither FMA nor fast-math,
 * so this file must be compiled with -ffp-contract=off and without -ffast-math.
 * Array layout: C-contiguous [x][y][z], z fastest (io/vasp.py:102-103).
 * dist_mat is the reference's 3x3x3 table indexed with Python negative-index
 * semantics: offset -1 reads element 2 (interface.py:242-259, methods.py:110).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int32_t lab_t;

typedef struct {
    int64_t nx, ny, nz;
} dims_t;

static inline int64_t wrap1(int64_t i, int64_t n) {
    /* single periodic wrap, as the reference does (methods.py:89-93) */
    if (i < 0) return i + n;
    if (i >= n) return i - n;
    return i;
}

static inline int64_t lin(const dims_t *d, int64_t x, int64_t y, int64_t z) {
    return (x * d->ny + y) * d->nz + z;
}

/* weight of the step (ix,iy,iz) in {-1,0,1}^3 out of the reference's table */
static inline double wgt(const double *dist, int ix, int iy, int iz) {
    return dist[((ix + 3) % 3) * 9 + ((iy + 3) % 3) * 3 + ((iz + 3) % 3)];
}

/*
 * One ongrid step (methods.py:87-117, and its copies at methods.py:418-447,
 * refinement.py:206-235): steepest ascent over the 26 neighbours weighted by
 * 1/distance; strict '>' against a running maximum that starts at the centre
 * value; neighbours visited in (ix,iy,iz) lexicographic order so ties go to
 * the first.  Returns the linear index of the winner (== centre if none).
 */
static int64_t ongrid_step(const double *rho, const dims_t *d, const double *dist,
                           int64_t px, int64_t py, int64_t pz,
                           int64_t *ox, int64_t *oy, int64_t *oz) {
    const double ctr = rho[lin(d, px, py, pz)];
    double best = ctr;
    int64_t bx = px, by = py, bz = pz;
    for (int ix = -1; ix <= 1; ++ix) {
        const int64_t tx = wrap1(px + ix, d->nx);
        for (int iy = -1; iy <= 1; ++iy) {
            const int64_t ty = wrap1(py + iy, d->ny);
            for (int iz = -1; iz <= 1; ++iz) {
                const int64_t tz = wrap1(pz + iz, d->nz);
                double v = (rho[lin(d, tx, ty, tz)] - ctr) * wgt(dist, ix, iy, iz);
                v += ctr;
                if (v > best) {
                    best = v;
                    bx = tx; by = ty; bz = tz;
                }
            }
        }
    }
    *ox = bx; *oy = by; *oz = bz;
    return lin(d, bx, by, bz);
}

/*
 * One neargrid gradient step (methods.py:302-363 with strict_axis=0,
 * refinement.py:89-154 with strict_axis=1).  p is the current voxel, dr the
 * running residual.  Writes the target into t[] and returns 1 if a step was
 * taken, 0 if the gradient vanished (target == p).
 */
static int neargrid_step(const double *rho, const dims_t *d, const double *T,
                         const int64_t p[3], double dr[3], int strict_axis,
                         int64_t t[3]) {
    const int64_t n[3] = {d->nx, d->ny, d->nz};
    const double here = rho[lin(d, p[0], p[1], p[2])];
    double g[3], gd[3];
    for (int j = 0; j < 3; ++j) {
        int64_t q[3] = {p[0], p[1], p[2]};
        q[j] = wrap1(p[j] + 1, n[j]);
        const double up = rho[lin(d, q[0], q[1], q[2])];
        q[j] = wrap1(p[j] - 1, n[j]);
        const double dn = rho[lin(d, q[0], q[1], q[2])];
        int flat;
        if (strict_axis) flat = (up < here) && (here > dn);   /* refinement.py:111 */
        else             flat = (up <= here) && (here >= dn); /* methods.py:324 */
        g[j] = flat ? 0.0 : (up - dn) / 2.0;
    }
    double gmax = 0.0;
    for (int j = 0; j < 3; ++j) {
        gd[j] = ((T[j * 3 + 0] * g[0]) + (T[j * 3 + 1] * g[1])) + (T[j * 3 + 2] * g[2]);
        if (gd[j] > gmax) gmax = gd[j];
        else if (-gd[j] > gmax) gmax = -gd[j];
    }
    if (gmax < 1E-14) {
        t[0] = p[0]; t[1] = p[1]; t[2] = p[2];
        return 0;
    }
    for (int j = 0; j < 3; ++j) {
        gd[j] /= gmax;
        int64_t ig = (gd[j] > 0) ? (int64_t)(gd[j] + .5) : (int64_t)(gd[j] - .5);
        int64_t q = p[j] + ig;
        dr[j] += gd[j] - (double)ig;
        int64_t ir = (dr[j] > 0) ? (int64_t)(dr[j] + .5) : (int64_t)(dr[j] - .5);
        q += ir;
        dr[j] -= (double)ir;
        if (q >= n[j]) q -= n[j];
        else if (q < 0) q += n[j];
        t[j] = q;
    }
    return 1;
}

/* ------------------------------------------------------------------------ */
/* utils.vacuum_assign (utils.py:383-401)                                     */
void orc_vacuum_assign(const double *reference, lab_t *vol, int64_t n, double tol,
                       const double *density, double voxel_volume,
                       double *out_charge, double *out_volume) {
    double charge = 0, volume = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (reference[i] <= tol) {
            vol[i] = -1;
            charge += density[i];
            volume += voxel_volume;
        }
    }
    *out_charge = charge * voxel_volume;
    *out_volume = volume;
}

/* ------------------------------------------------------------------------ */
/* methods.ongrid (methods.py:15-219), whole-volume case.  vol holds 0 / -1 on
 * entry, 1-based labels on exit.  maxima receives voxel indices [n][3].
 * Returns number of maxima, or -1 if max_cap was too small.                  */
int64_t orc_ongrid(const double *rho, lab_t *vol, int64_t nx, int64_t ny, int64_t nz,
                   const double *dist, int64_t *maxima, int64_t max_cap) {
    const dims_t d = {nx, ny, nz};
    const int64_t N = nx * ny * nz;
    int64_t path_cap = 1024, n_max = 0;
    int64_t *path = (int64_t *)malloc(path_cap * sizeof(int64_t));
    for (int64_t i = 0; i < N; ++i) {
        if (vol[i] != 0) continue;
        int64_t px = i / (ny * nz), py = (i / nz) % ny, pz = i % nz;
        int64_t plen = 0;
        path[plen++] = i;
        lab_t label;
        for (;;) {
            int64_t tx, ty, tz;
            const int64_t cur = lin(&d, px, py, pz);
            const int64_t nxt = ongrid_step(rho, &d, dist, px, py, pz, &tx, &ty, &tz);
            if (vol[nxt] != 0) { label = vol[nxt]; break; }       /* methods.py:166 */
            if (nxt == cur) {                                     /* methods.py:169 */
                if (n_max >= max_cap) { free(path); return -1; }
                maxima[n_max * 3 + 0] = tx;
                maxima[n_max * 3 + 1] = ty;
                maxima[n_max * 3 + 2] = tz;
                label = (lab_t)(++n_max);
                break;
            }
            if (plen == path_cap) {
                path_cap *= 2;
                path = (int64_t *)realloc(path, path_cap * sizeof(int64_t));
            }
            path[plen++] = nxt;
            px = tx; py = ty; pz = tz;
        }
        for (int64_t k = 0; k < plen; ++k) vol[path[k]] = label;  /* methods.py:211 */
    }
    free(path);
    return n_max;
}

/* ------------------------------------------------------------------------ */
/* methods.neargrid (methods.py:222-611), whole-volume case.  Scan-order
 * dependent path painting with the `known` interior cache.                   */
static int inb(int64_t v, int64_t n) { return v >= 0 && v < n; }

static void promote_if_interior(const dims_t *d, const lab_t *vol, int8_t *known,
                                int64_t x, int64_t y, int64_t z) {
    /* methods.py:556-577: a face neighbour whose own six face neighbours (no
     * periodic wrap) all carry its label becomes known == 2 */
    const lab_t v = vol[lin(d, x, y, z)];
    if (v == -1 || v == 0) return;
    const int64_t c[3] = {x, y, z};
    const int64_t n[3] = {d->nx, d->ny, d->nz};
    for (int h = 0; h < 3; ++h) {
        for (int s = 1; s >= -1; s -= 2) {
            int64_t q[3] = {c[0], c[1], c[2]};
            q[h] += s;
            if (!inb(q[h], n[h])) return;
            if (vol[lin(d, q[0], q[1], q[2])] != v) return;
        }
    }
    known[lin(d, x, y, z)] = 2;
}

int64_t orc_neargrid(const double *rho, lab_t *vol, int64_t nx, int64_t ny, int64_t nz,
                     const double *dist, const double *T, int64_t *maxima,
                     int64_t max_cap) {
    const dims_t d = {nx, ny, nz};
    const int64_t n3[3] = {nx, ny, nz};
    const int64_t N = nx * ny * nz;
    int8_t *known = (int8_t *)calloc((size_t)N, 1);
    int64_t path_cap = 1024, n_max = 0;
    int64_t *path = (int64_t *)malloc(path_cap * sizeof(int64_t));
    for (int64_t i = 0; i < N; ++i) {
        if (vol[i] == -1) continue;
        if (known[i] == 2) continue;
        known[i] = 1;
        int64_t p[3] = {i / (ny * nz), (i / nz) % ny, i % nz};
        int64_t t[3];
        double dr[3] = {0., 0., 0.};
        int64_t plen = 0;
        path[plen++] = i;
        lab_t label;
        int64_t endp = i;
        for (;;) {
            neargrid_step(rho, &d, T, p, dr, 0, t);
            int64_t tl = lin(&d, t[0], t[1], t[2]);
            if (known[tl] == 1) {                                 /* methods.py:411 */
                dr[0] = dr[1] = dr[2] = 0.;
                const int64_t cur = lin(&d, p[0], p[1], p[2]);
                tl = ongrid_step(rho, &d, dist, p[0], p[1], p[2], &t[0], &t[1], &t[2]);
                if (tl == cur) {                                  /* methods.py:496 */
                    label = vol[cur];          /* 0 => brand-new maximum */
                    endp = cur;
                    break;
                }
            }
            if (known[tl] == 2) {                                 /* methods.py:509 */
                label = vol[tl];
                endp = tl;
                break;
            }
            if (plen == path_cap) {
                path_cap *= 2;
                path = (int64_t *)realloc(path, path_cap * sizeof(int64_t));
            }
            p[0] = t[0]; p[1] = t[1]; p[2] = t[2];
            path[plen++] = tl;
            known[tl] = 1;
        }
        if (label == 0) {                                         /* methods.py:533 */
            if (n_max >= max_cap) { free(path); free(known); return -1; }
            maxima[n_max * 3 + 0] = endp / (ny * nz);
            maxima[n_max * 3 + 1] = (endp / nz) % ny;
            maxima[n_max * 3 + 2] = endp % nz;
            label = (lab_t)(++n_max);
        }
        for (int64_t k = 0; k < plen; ++k) {                      /* methods.py:543 */
            const int64_t q = path[k];
            const int64_t c[3] = {q / (ny * nz), (q / nz) % ny, q % nz};
            vol[q] = label;
            if (known[q] != 2) known[q] = 0;
            for (int a = 0; a < 3; ++a) {
                for (int s = 1; s >= -1; s -= 2) {
                    int64_t m[3] = {c[0], c[1], c[2]};
                    m[a] += s;
                    if (!inb(m[a], n3[a])) continue;
                    promote_if_interior(&d, vol, known, m[0], m[1], m[2]);
                }
            }
        }
    }
    free(path);
    free(known);
    return n_max;
}

/* ------------------------------------------------------------------------ */
/* classification shared by edge_find / edge_check (refinement.py:345-375,
 * 446-476): vacuum neighbours are ignored for both flags.                     */
static void classify(const double *rho, const lab_t *vol, const dims_t *d,
                     int64_t x, int64_t y, int64_t z, int *is_edge, int *is_max) {
    const int64_t c = lin(d, x, y, z);
    const lab_t mine = vol[c];
    const double here = rho[c];
    int e = 0, m = 1;
    for (int ix = -1; ix <= 1; ++ix) {
        const int64_t tx = wrap1(x + ix, d->nx);
        for (int iy = -1; iy <= 1; ++iy) {
            const int64_t ty = wrap1(y + iy, d->ny);
            for (int iz = -1; iz <= 1; ++iz) {
                const int64_t tz = wrap1(z + iz, d->nz);
                const int64_t q = lin(d, tx, ty, tz);
                if (vol[q] == -1) continue;
                if (vol[q] != mine) e = 1;
                if (rho[q] > here) m = 0;
            }
        }
    }
    *is_edge = e;
    *is_max = m;
}

static void dilate_near(int8_t *known, const dims_t *d, int64_t x, int64_t y, int64_t z) {
    /* refinement.py:385-404 / 484-503 */
    for (int ix = -1; ix <= 1; ++ix) {
        const int64_t tx = wrap1(x + ix, d->nx);
        for (int iy = -1; iy <= 1; ++iy) {
            const int64_t ty = wrap1(y + iy, d->ny);
            for (int iz = -1; iz <= 1; ++iz) {
                const int64_t tz = wrap1(z + iz, d->nz);
                const int64_t q = lin(d, tx, ty, tz);
                if (known[q] >= 0) known[q] = -1;
            }
        }
    }
}

/* refinement.edge_find (refinement.py:326-405) */
int64_t orc_edge_find(int8_t *known, const double *rho, const lab_t *vol,
                      int64_t nx, int64_t ny, int64_t nz) {
    const dims_t d = {nx, ny, nz};
    int64_t edges = 0;
    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                const int64_t c = lin(&d, x, y, z);
                if (known[c] == 2) continue;
                if (vol[c] == -1) continue;
                int e, m;
                classify(rho, vol, &d, x, y, z, &e, &m);
                if (!e || m) {
                    if (known[c] >= 0) known[c] = 2;
                } else {
                    known[c] = -2;
                    ++edges;
                    dilate_near(known, &d, x, y, z);
                }
            }
    return edges;
}

/* refinement.edge_check (refinement.py:409-508) */
void orc_edge_check(int8_t *known, const double *rho, const lab_t *vol,
                    int64_t nx, int64_t ny, int64_t nz,
                    int64_t *out_checked, int64_t *out_edges) {
    const dims_t d = {nx, ny, nz};
    int64_t checked = 0, edges = 0;
    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                if (known[lin(&d, x, y, z)] != -2) continue;
                for (int ex = -1; ex <= 1; ++ex) {
                    const int64_t qx = wrap1(x + ex, nx);
                    for (int ey = -1; ey <= 1; ++ey) {
                        const int64_t qy = wrap1(y + ey, ny);
                        for (int ez = -1; ez <= 1; ++ez) {
                            const int64_t qz = wrap1(z + ez, nz);
                            const int64_t q = lin(&d, qx, qy, qz);
                            int e, m;
                            classify(rho, vol, &d, qx, qy, qz, &e, &m);
                            if (!e) {
                                known[q] = -1;
                                ++checked;
                            } else if (!m) {
                                if (known[q] != -3) {
                                    known[q] = -3;
                                    ++edges;
                                    dilate_near(known, &d, qx, qy, qz);
                                    ++checked;
                                }
                            }
                        }
                    }
                }
            }
    const int64_t N = nx * ny * nz;
    for (int64_t i = 0; i < N; ++i)
        if (known[i] == -3) known[i] = -2;
    *out_checked = checked;
    *out_edges = edges;
}

/* ------------------------------------------------------------------------ */
/* refinement.neargrid (refinement.py:17-322), whole-volume case.  For every
 * voxel flagged -2: follow its own neargrid trajectory (strict axis rule)
 * until it lands on a voxel that rknown says is interior, or on a maximum,
 * and take that voxel's label.  Returns the number of relabelled voxels, or
 * -1 if a trajectory exceeded step_cap steps.                                */
int64_t orc_refine_neargrid(int8_t *known, const int8_t *rknown, const double *rho,
                            lab_t *vol, int64_t nx, int64_t ny, int64_t nz,
                            const double *dist, const double *T, int64_t step_cap) {
    const dims_t d = {nx, ny, nz};
    const int64_t N = nx * ny * nz;
    int64_t path_cap = 1024, changed = 0;
    int64_t *path = (int64_t *)malloc(path_cap * sizeof(int64_t));
    for (int64_t i = 0; i < N; ++i) {
        if (known[i] != -2) continue;
        int64_t p[3] = {i / (ny * nz), (i / nz) % ny, i % nz};
        int64_t t[3];
        double dr[3] = {0., 0., 0.};
        const lab_t mine = vol[i];
        int64_t plen = 0, steps = 0;
        path[plen++] = i;
        known[i] += 5;                                            /* refinement.py:84 */
        for (;;) {
            if (++steps > step_cap) { free(path); return -1; }
            neargrid_step(rho, &d, T, p, dr, 1, t);
            int64_t tl = lin(&d, t[0], t[1], t[2]);
            int done = 0;
            if (known[tl] >= 3 && known[tl] <= 5) {               /* refinement.py:200 */
                dr[0] = dr[1] = dr[2] = 0.;
                const int64_t cur = lin(&d, p[0], p[1], p[2]);
                tl = ongrid_step(rho, &d, dist, p[0], p[1], p[2], &t[0], &t[1], &t[2]);
                if (tl == cur) done = 1;                          /* refinement.py:283 */
            }
            if (done || rknown[tl] == 2) {                        /* refinement.py:294 */
                const lab_t other = vol[tl];
                if (other != mine) {
                    vol[i] += other - mine;
                    ++changed;
                } else {
                    known[i] += 1;
                }
                break;
            }
            if (plen == path_cap) {
                path_cap *= 2;
                path = (int64_t *)realloc(path, path_cap * sizeof(int64_t));
            }
            p[0] = t[0]; p[1] = t[1]; p[2] = t[2];
            path[plen++] = tl;
            if (known[tl] < 2) known[tl] += 5;                    /* refinement.py:314 */
        }
        for (int64_t k = 0; k < plen; ++k)                        /* refinement.py:317 */
            if (known[path[k]] > 2) known[path[k]] -= 5;
    }
    free(path);
    return changed;
}

/* ------------------------------------------------------------------------ */
/* utils.charge_sum (utils.py:236-252) */
void orc_charge_sum(double *charge, double *volume, int64_t n_lab, double voxel_volume,
                    const double *density, const lab_t *vol, int64_t N) {
    for (int64_t i = 0; i < N; ++i) {
        const lab_t a = vol[i];
        if (a >= 0) {
            charge[a] += density[i];
            volume[a] += voxel_volume;
        }
    }
    for (int64_t j = 0; j < n_lab; ++j) charge[j] *= voxel_volume;
}

/* utils.atom_assign (utils.py:186-232) */
void orc_atom_assign(const double *bmax, int64_t n_max, const double *atoms, int64_t n_atoms,
                     const double *lat, int64_t *out_atom, double *out_dist) {
    for (int64_t i = 0; i < n_max; ++i) {
        const double *b = bmax + 3 * i;
        double best = (b[0] - (atoms[0] + 0.)) * (b[0] - (atoms[0] + 0.))
                    + (b[1] - (atoms[1] + 0.)) * (b[1] - (atoms[1] + 0.))
                    + (b[2] - (atoms[2] + 0.)) * (b[2] - (atoms[2] + 0.));
        int64_t who = 0;
        for (int64_t j = 0; j < n_atoms; ++j) {
            const double *a = atoms + 3 * j;
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y)
                    for (int z = -1; z <= 1; ++z) {
                        double s[3];
                        for (int k = 0; k < 3; ++k)
                            s[k] = (lat[0 * 3 + k] * x + lat[1 * 3 + k] * y) + lat[2 * 3 + k] * z;
                        const double e0 = b[0] - (a[0] + s[0]);
                        const double e1 = b[1] - (a[1] + s[1]);
                        const double e2 = b[2] - (a[2] + s[2]);
                        const double dd = (e0 * e0 + e1 * e1) + e2 * e2;
                        if (dd < best) { best = dd; who = j; }
                    }
        }
        out_atom[i] = who;
        out_dist[i] = sqrt(best);
    }
}

/* utils.volume_assign (utils.py:405-421) */
void orc_volume_assign(lab_t *vol, int64_t N, const int64_t *swap) {
    for (int64_t i = 0; i < N; ++i)
        if (vol[i] >= 0) vol[i] = (lab_t)swap[vol[i]];
}

/* utils.surface_dist (utils.py:321-379) over the whole volume (one brick).
 * distance[a] ends as the minimum, over edge voxels (known == -2) labelled a,
 * of the 27-image distance to atom a, clamped from above by the reference's
 * starting value sqrt(nx^2+ny^2+nz^2); 0 when atom a owns no edge voxel.      */
void orc_surface_dist(const int8_t *known, const lab_t *vol, int64_t nx, int64_t ny,
                      int64_t nz, const double *lat, const double *atoms,
                      int64_t n_atoms, double *distance) {
    const dims_t d = {nx, ny, nz};
    double *best = (double *)malloc(sizeof(double) * (size_t)n_atoms);
    for (int64_t a = 0; a < n_atoms; ++a) {
        best[a] = (double)(nx * nx + ny * ny + nz * nz);
        distance[a] = 0.;
    }
    for (int64_t x = 0; x < nx; ++x)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t z = 0; z < nz; ++z) {
                const int64_t c = lin(&d, x, y, z);
                if (known[c] != -2) continue;
                const lab_t a = vol[c];
                double pc[3], mind = best[a];
                for (int j = 0; j < 3; ++j) {
                    pc[j] = lat[0 * 3 + j] * (double)x / (double)nx;
                    pc[j] += lat[1 * 3 + j] * (double)y / (double)ny;
                    pc[j] += lat[2 * 3 + j] * (double)z / (double)nz;
                }
                for (int ix = -1; ix <= 1; ++ix)
                    for (int iy = -1; iy <= 1; ++iy)
                        for (int iz = -1; iz <= 1; ++iz) {
                            double s[3];
                            for (int k = 0; k < 3; ++k)
                                s[k] = (lat[0 * 3 + k] * ix + lat[1 * 3 + k] * iy) + lat[2 * 3 + k] * iz;
                            const double e0 = pc[0] - (atoms[3 * a + 0] + s[0]);
                            const double e1 = pc[1] - (atoms[3 * a + 1] + s[1]);
                            const double e2 = pc[2] - (atoms[3 * a + 2] + s[2]);
                            const double dd = (e0 * e0 + e1 * e1) + e2 * e2;
                            if (dd < mind) mind = dd;
                        }
                distance[a] = sqrt(mind);
                best[a] = mind;
            }
    free(best);
}

/* utils.volume_mask (utils.py:462-476) */
void orc_volume_mask(const lab_t *vol, const double *density, int64_t N, lab_t which,
                     double *out) {
    for (int64_t i = 0; i < N; ++i) out[i] = (vol[i] == which) ? density[i] : 0.0;
}
