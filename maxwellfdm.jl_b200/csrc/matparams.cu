// N4 - material-parameter pipeline on the GPU: object assignment + Kottke subpixel smoothing in ONE kernel.
//
// Replaces calc_matparams!(mdl::ModelFull) (reference src/model/full.jl:16-70), i.e. the assign_param! /
// smooth_param! pair of the un-vendored MaxwellBase: the reference rasterises object indices into four global Int
// arrays (full.jl:47-61) and then smooths every voxel from them (full.jl:64-67), all on one CPU thread.  Here a CTA
// owns an 8x8x4 tile of cells:
//   1. it culls the shape list against the tile's bounding box (shared-memory list; Bloch axes at the domain edge are
//      not culled because their ghost corners wrap to the far side);
//   2. it evaluates the object index at every voxel corner the tile needs - the four parity classes of the half-step
//      lattice, (T+1)^3 points each - ONCE, into shared memory (the reference's four oind3d arrays, tile-sized);
//   3. one thread per cell smooths its four voxels (E_x, E_y, E_z locations and the corner location that holds the
//      off-diagonal entries) and writes the nine entries straight into the Julia-layout (Nx,Ny,nzl,3,3) array that
//      fdfd_set_eps / fdfd_set_mu consume: each of the nine planes is a coalesced 16-byte stream.
// Bound: instruction throughput of the interface voxels; HBM traffic is the 144 B/cell output, nothing is read but
// the shape list.  Decision rules and formulas: oracle/matparams.py (same restatement, same tie-breaking).
#include <cmath>
#include <cstring>
#include <vector>

#include "cplx.cuh"
#include "fdfd_internal.h"

namespace fdfd {

namespace {

constexpr int MTX = 8, MTY = 8, MTZ = 4, MNT = MTX * MTY * MTZ;
constexpr int LPX = MTX + 1, LPY = MTY + 1, LPZ = MTZ + 1, LPN = LPX * LPY * LPZ;   // corner lattice of one class
constexpr int MAXLIST = 512;                                                           // culled shapes per tile
constexpr double VOLFRAC_TOL = 1e-6;

struct MatParams {
    int32_t N[3];
    int32_t isbloch[3];
    int32_t g[3];                 // grid type of the field planes per axis (0 PRIM, 1 DUAL): ft2gt(ft, boundft[w])
    int32_t ortho;                // field orthogonal to the shape dimensions: arithmetic averages (model.jl:65-69)
    int32_t symmetric;            // every material tensor is symmetric: write P_ab (a < b) into both (a,b) and (b,a)
    int32_t k0, k1;               // planes of this slab
    int32_t nshape, nparam;
    const double *H[3];           // half-step lattice per axis: H[2i] = ghosted dual i, H[2i+1] = primal i (2N+2 entries)
    double lo[3], hi[3], L[3];    // domain bounds
    const fdfd_shape *shapes;
    const double2 *params;        // nparam x 9 (row-major 3x3)
    const double2 *params_inv;    // inverses (harmonic mean of >= 3 materials)
    double2 *out;                 // Julia layout (Nx,Ny,nzl,3,3)
    int32_t *err;                 // set to 1 when a voxel corner is covered by no shape
};

// membership tests compare sums of products: keep them un-contracted so that a point decides the same way here, in
// the CPU logic-check build and in the numpy oracle
__device__ __forceinline__ double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__device__ __forceinline__ double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

__device__ __forceinline__ bool shape_contains(const fdfd_shape &s, const double x[3]) {
    const double d0 = x[0] - s.c[0], d1 = x[1] - s.c[1], d2 = x[2] - s.c[2];
    if (s.kind == FDFD_SHAPE_BOX) return fabs(d0) <= s.r[0] && fabs(d1) <= s.r[1] && fabs(d2) <= s.r[2];
    if (s.kind == FDFD_SHAPE_BALL)
        return add_rn(add_rn(mul_rn(d0, d0), mul_rn(d1, d1)), mul_rn(d2, d2)) <= mul_rn(s.r[0], s.r[0]);
    const double d[3] = {d0, d1, d2};
    const int a = s.axis, b = (a + 1) % 3, c = (a + 2) % 3;
    return fabs(d[a]) <= s.r[1] && add_rn(mul_rn(d[b], d[b]), mul_rn(d[c], d[c])) <= mul_rn(s.r[0], s.r[0]);
}

// nearest surface point r0 and outward normal n of shape s seen from x0 (oracle: surfpt_nearby)
__device__ void surfpt_nearby(const fdfd_shape &s, const double x0[3], double r0[3], double n[3]) {
    const double d[3] = {x0[0] - s.c[0], x0[1] - s.c[1], x0[2] - s.c[2]};
    n[0] = n[1] = n[2] = 0.0;
    if (s.kind == FDFD_SHAPE_BALL) {
        const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (nd > 0) { n[0] = d[0] / nd; n[1] = d[1] / nd; n[2] = d[2] / nd; }
        else n[0] = 1.0;
        for (int w = 0; w < 3; ++w) r0[w] = s.c[w] + s.r[0] * n[w];
        return;
    }
    if (s.kind == FDFD_SHAPE_BOX) {
        const double q[3] = {fabs(d[0]) - s.r[0], fabs(d[1]) - s.r[1], fabs(d[2]) - s.r[2]};
        if (q[0] <= 0 && q[1] <= 0 && q[2] <= 0) {
            int a = 0;
            if (q[1] > q[a]) a = 1;
            if (q[2] > q[a]) a = 2;
            n[a] = d[a] >= 0 ? 1.0 : -1.0;
            for (int w = 0; w < 3; ++w) r0[w] = x0[w];
            r0[a] = s.c[a] + n[a] * s.r[a];
            return;
        }
        double v[3], vv = 0.0;
        for (int w = 0; w < 3; ++w) {
            r0[w] = s.c[w] + fmin(fmax(d[w], -s.r[w]), s.r[w]);
            v[w] = x0[w] - r0[w];
            vv += v[w] * v[w];
        }
        vv = sqrt(vv);
        for (int w = 0; w < 3; ++w) n[w] = v[w] / vv;
        return;
    }
    const int a = s.axis, b = (a + 1) % 3, c = (a + 2) % 3;
    const double R = s.r[0], h = s.r[1];
    const double rho = sqrt(d[b] * d[b] + d[c] * d[c]);
    const double qa = fabs(d[a]) - h, qr = rho - R;
    for (int w = 0; w < 3; ++w) r0[w] = x0[w];
    if (qa <= 0 && qr <= 0) {
        if (qa > qr) {
            n[a] = d[a] >= 0 ? 1.0 : -1.0;
            r0[a] = s.c[a] + n[a] * h;
        } else {
            if (rho > 0) { n[b] = d[b] / rho; n[c] = d[c] / rho; }
            else n[b] = 1.0;
            r0[b] = s.c[b] + R * n[b];
            r0[c] = s.c[c] + R * n[c];
        }
        return;
    }
    r0[a] = s.c[a] + fmin(fmax(d[a], -h), h);
    const double f = rho <= R ? 1.0 : R / rho;
    r0[b] = s.c[b] + f * d[b];
    r0[c] = s.c[c] + f * d[c];
    double v[3], vv = 0.0;
    for (int w = 0; w < 3; ++w) { v[w] = x0[w] - r0[w]; vv += v[w] * v[w]; }
    vv = sqrt(vv);
    for (int w = 0; w < 3; ++w) n[w] = v[w] / vv;
}

// fraction of the box [lo,hi] with n.(x - r0) <= 0 (oracle: volfrac)
__device__ double volfrac(const double lo[3], const double hi[3], const double n[3], const double r0[3]) {
    double a[3], dmin = 0.0, amax = 0.0;
    for (int w = 0; w < 3; ++w) {
        a[w] = fabs(n[w]) * (hi[w] - lo[w]);
        dmin += fmin(n[w] * lo[w], n[w] * hi[w]);
        amax = fmax(amax, a[w]);
    }
    const double d = (n[0] * r0[0] + n[1] * r0[1] + n[2] * r0[2]) - dmin;
    int act[3], k = 0;
    for (int w = 0; w < 3; ++w)
        if (a[w] > VOLFRAC_TOL * amax) act[k++] = w;
    double tot = 0.0, prod = 1.0;
    for (int q = 0; q < k; ++q) prod *= a[act[q]];
    // subsets in the oracle's order: by size, then lexicographic
    for (int m = 0; m <= k; ++m) {
        for (int mask = 0; mask < (1 << k); ++mask) {
            if (__popc(mask) != m) continue;
            double t = d;
            double sub = 0.0;
            for (int q = 0; q < k; ++q)
                if (mask & (1 << q)) sub += a[act[q]];
            t -= sub;
            if (t > 0) {
                double p = 1.0;
                for (int q = 0; q < k; ++q) p *= t;
                tot += (m & 1) ? -p : p;
            }
        }
    }
    const double fact = k == 3 ? 6.0 : k == 2 ? 2.0 : 1.0;
    const double f = tot / (prod * fact);
    return fmin(fmax(f, 0.0), 1.0);
}

struct M3 { double2 m[9]; };

__device__ __forceinline__ M3 m3_load(const double2 *p) {
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.m[i] = p[i];
    return r;
}
// S^T P S (fwd) or S P S^T (back) for a real orthonormal S (row-major 3x3)
__device__ M3 m3_rotate(const double S[9], const M3 &P, bool back) {
    M3 t, r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double2 acc = c_zero();
            for (int k = 0; k < 3; ++k) {
                const double s = back ? S[i * 3 + k] : S[k * 3 + i];
                acc = c_add(acc, c_scale(s, P.m[k * 3 + j]));
            }
            t.m[i * 3 + j] = acc;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double2 acc = c_zero();
            for (int k = 0; k < 3; ++k) {
                const double s = back ? S[j * 3 + k] : S[k * 3 + j];
                acc = c_add(acc, c_scale(s, t.m[i * 3 + k]));
            }
            r.m[i * 3 + j] = acc;
        }
    return r;
}
__device__ M3 m3_tau(const M3 &P, bool inverse) {
    // tau and its inverse have the same form up to the sign of the first row / column
    M3 T;
    const double2 one = c_make(1.0, 0.0);
    const double2 p00 = P.m[0];
    T.m[0] = c_neg(c_div(one, p00));
    for (int j = 1; j < 3; ++j) {
        const double2 r = c_div(P.m[j], p00), c = c_div(P.m[j * 3], p00);
        T.m[j] = inverse ? c_neg(r) : r;
        T.m[j * 3] = inverse ? c_neg(c) : c;
    }
    for (int i = 1; i < 3; ++i)
        for (int j = 1; j < 3; ++j) T.m[i * 3 + j] = c_sub(P.m[i * 3 + j], c_div(c_mul(P.m[i * 3], P.m[j]), p00));
    return T;
}
__device__ M3 kottke_avg(const M3 &P1, const M3 &P2, const double n[3], double rvol) {
    // frame S = [n t1 t2] (columns), t1 from the coordinate axis least aligned with n (first of equals)
    int e = 0;
    if (fabs(n[1]) < fabs(n[e])) e = 1;
    if (fabs(n[2]) < fabs(n[e])) e = 2;
    double t1[3] = {-n[e] * n[0], -n[e] * n[1], -n[e] * n[2]};
    t1[e] += 1.0;
    const double nt = sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]);
    for (int w = 0; w < 3; ++w) t1[w] /= nt;
    const double t2[3] = {n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]};
    const double S[9] = {n[0], t1[0], t2[0], n[1], t1[1], t2[1], n[2], t1[2], t2[2]};
    const M3 T1 = m3_tau(m3_rotate(S, P1, false), false), T2 = m3_tau(m3_rotate(S, P2, false), false);
    M3 T;
    for (int i = 0; i < 9; ++i) T.m[i] = c_add(c_scale(rvol, T1.m[i]), c_scale(1.0 - rvol, T2.m[i]));
    return m3_rotate(S, m3_tau(T, true), true);
}
__device__ M3 m3_inverse(const M3 &A) {
    const double2 *a = A.m;
    M3 C;   // cofactors (transposed -> adjugate)
    C.m[0] = c_sub(c_mul(a[4], a[8]), c_mul(a[5], a[7]));
    C.m[1] = c_sub(c_mul(a[2], a[7]), c_mul(a[1], a[8]));
    C.m[2] = c_sub(c_mul(a[1], a[5]), c_mul(a[2], a[4]));
    C.m[3] = c_sub(c_mul(a[5], a[6]), c_mul(a[3], a[8]));
    C.m[4] = c_sub(c_mul(a[0], a[8]), c_mul(a[2], a[6]));
    C.m[5] = c_sub(c_mul(a[2], a[3]), c_mul(a[0], a[5]));
    C.m[6] = c_sub(c_mul(a[3], a[7]), c_mul(a[4], a[6]));
    C.m[7] = c_sub(c_mul(a[1], a[6]), c_mul(a[0], a[7]));
    C.m[8] = c_sub(c_mul(a[0], a[4]), c_mul(a[1], a[3]));
    const double2 det = c_add(c_add(c_mul(a[0], C.m[0]), c_mul(a[1], C.m[3])), c_mul(a[2], C.m[6]));
    for (int i = 0; i < 9; ++i) C.m[i] = c_div(C.m[i], det);
    return C;
}

// tau-transform of a (ghost) coordinate back into the domain: wrap (Bloch) or mirror (symmetry boundary)
__device__ __forceinline__ double to_domain(const MatParams &p, int w, double x) {
    if (x < p.lo[w]) return p.isbloch[w] ? x + p.L[w] : 2.0 * p.lo[w] - x;
    if (x > p.hi[w]) return p.isbloch[w] ? x - p.L[w] : 2.0 * p.hi[w] - x;
    return x;
}

__global__ void __launch_bounds__(MNT) matparams_kernel(const __grid_constant__ MatParams p) {
    __shared__ int s_list[MAXLIST];
    __shared__ int s_nlist;
    __shared__ int s_oind[4][LPN];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * MTX, j0 = blockIdx.y * MTY, k0 = p.k0 + blockIdx.z * MTZ;
    const int base[3] = {i0, j0, k0};
    const int ext[3] = {MTX, MTY, MTZ};

    // ---- 1. cull the shape list against the tile (conservative) ---------------------------------------------------
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    double tlo[3], thi[3];
    bool open_axis[3];
    for (int w = 0; w < 3; ++w) {
        const int iend = min(base[w] + ext[w], p.N[w]);
        tlo[w] = p.H[w][2 * base[w]];
        thi[w] = p.H[w][2 * iend + 1];
        // ghost corners below the domain wrap to the far side on Bloch axes: no culling along that axis; at a
        // symmetry boundary they mirror into the first cell, which the interval already covers once it starts at lo
        open_axis[w] = base[w] == 0 && p.isbloch[w];
        if (base[w] == 0) tlo[w] = p.lo[w];
    }
    for (int o = tid; o < p.nshape; o += MNT) {
        const fdfd_shape s = p.shapes[o];
        double hw[3];
        if (s.kind == FDFD_SHAPE_BOX) { hw[0] = s.r[0]; hw[1] = s.r[1]; hw[2] = s.r[2]; }
        else if (s.kind == FDFD_SHAPE_BALL) { hw[0] = hw[1] = hw[2] = s.r[0]; }
        else { hw[0] = hw[1] = hw[2] = s.r[0]; hw[s.axis] = s.r[1]; }
        bool hit = true;
        for (int w = 0; w < 3; ++w) {
            const double margin = 1e-12 * (fabs(tlo[w]) + fabs(thi[w]) + hw[w]);   // rounding slack: never cull a toucher
            if (!open_axis[w] && (s.c[w] + hw[w] < tlo[w] - margin || s.c[w] - hw[w] > thi[w] + margin)) hit = false;
        }
        if (hit) {
            const int slot = atomicAdd(&s_nlist, 1);
            if (slot < MAXLIST) s_list[slot] = o;
        }
    }
    __syncthreads();
    const int nlist = s_nlist;
    const bool use_list = nlist <= MAXLIST;   // overflow: scan every shape (slow, still correct)

    // ---- 2. object index at the voxel corners: parity class v has corners H[2i + g_w(v)] ------------------------------
    for (int t = tid; t < 4 * LPN; t += MNT) {
        const int v = t / LPN;
        int q = t % LPN;
        const int lx = q % LPX;
        q /= LPX;
        const int ly = q % LPY, lz = q / LPY;
        const int lidx[3] = {lx, ly, lz};
        double x[3];
        bool inside_grid = true;
        for (int w = 0; w < 3; ++w) {
            const int gw = (v < 3 && v == w) ? 1 - p.g[w] : p.g[w];
            const int cell = base[w] + lidx[w];
            if (cell > p.N[w]) { inside_grid = false; x[w] = 0.0; continue; }
            x[w] = to_domain(p, w, p.H[w][2 * cell + gw]);
        }
        int best = -1;
        if (inside_grid) {
            if (use_list) {
                for (int q2 = 0; q2 < nlist; ++q2) {
                    const int o = s_list[q2];
                    if (o > best && shape_contains(p.shapes[o], x)) best = o;
                }
            } else {
                for (int o = p.nshape - 1; o >= 0; --o)
                    if (shape_contains(p.shapes[o], x)) { best = o; break; }
            }
        }
        s_oind[v][t % LPN] = best;
    }
    __syncthreads();

    // ---- 3. one thread per cell: four voxels ---------------------------------------------------------------------
    const int cx = tid % MTX, cy = (tid / MTX) % MTY, cz = tid / (MTX * MTY);
    const int ci[3] = {i0 + cx, j0 + cy, k0 + cz};
    if (ci[0] >= p.N[0] || ci[1] >= p.N[1] || ci[2] >= p.k1) return;
    const int64_t Nxy = (int64_t)p.N[0] * p.N[1], nzl = p.k1 - p.k0;
    const int64_t cell = ((int64_t)(ci[2] - p.k0) * p.N[1] + ci[1]) * p.N[0] + ci[0];
    for (int v = 0; v < 4; ++v) {
        int oc[8], omax = -1, nobj = 0, objs[8];
        bool uncovered = false;
        for (int c = 0; c < 8; ++c) {
            const int o = s_oind[v][((cz + ((c >> 2) & 1)) * LPY + cy + ((c >> 1) & 1)) * LPX + cx + (c & 1)];
            oc[c] = o;
            uncovered |= o < 0;
            omax = max(omax, o);
            bool seen = false;
            for (int q = 0; q < nobj; ++q) seen |= objs[q] == o;
            if (!seen) objs[nobj++] = o;
        }
        if (uncovered) { *p.err = 1; return; }
        int pc[8], np = 0, ps[8];
        for (int c = 0; c < 8; ++c) {
            pc[c] = p.shapes[oc[c]].pind;
            bool seen = false;
            for (int q = 0; q < np; ++q) seen |= ps[q] == pc[c];
            if (!seen) ps[np++] = pc[c];
        }
        M3 P;
        if (np == 1) {
            P = m3_load(p.params + 9 * pc[0]);
        } else {
            bool mean_only = np >= 3;
            const int p_fg = p.shapes[omax].pind;
            int p_bg = ps[0] == p_fg ? ps[1] : ps[0];
            double nout[3] = {0.0, 0.0, 0.0}, rvol = 0.0;
            if (!mean_only) {
                double lo[3], hi[3], x0[3];
                for (int w = 0; w < 3; ++w) {
                    const int gw = (v < 3 && v == w) ? 1 - p.g[w] : p.g[w];
                    lo[w] = p.H[w][2 * ci[w] + gw];
                    hi[w] = p.H[w][2 * ci[w] + gw + 2];
                    x0[w] = 0.5 * (lo[w] + hi[w]);
                }
                if (nobj == 2) {
                    // the shape is seen from the in-domain image of the voxel centre (a centre can sit a hair outside
                    // a Bloch boundary when the first and last cells differ in size); r0 is shifted back
                    double r0[3], xt[3];
                    for (int w = 0; w < 3; ++w) xt[w] = to_domain(p, w, x0[w]);
                    surfpt_nearby(p.shapes[omax], xt, r0, nout);
                    for (int w = 0; w < 3; ++w) r0[w] += x0[w] - xt[w];
                    rvol = volfrac(lo, hi, nout, r0);
                } else {
                    // two materials, more than two objects: normal and fraction from the corner occupancy
                    int nfg = 0;
                    for (int c = 0; c < 8; ++c) {
                        const bool fg = pc[c] == p_fg;
                        nfg += fg;
                        const double sgn = fg ? -1.0 : 1.0;
                        for (int w = 0; w < 3; ++w) nout[w] += sgn * (((c >> w) & 1) ? 1.0 : -1.0);
                    }
                    rvol = nfg / 8.0;
                    const double nn = sqrt(nout[0] * nout[0] + nout[1] * nout[1] + nout[2] * nout[2]);
                    if (nn == 0.0) mean_only = true;
                    else for (int w = 0; w < 3; ++w) nout[w] /= nn;
                }
            }
            if (mean_only) {
                // harmonic mean over the corners (arithmetic when the field is orthogonal to the shape dimensions)
                for (int i = 0; i < 9; ++i) P.m[i] = c_zero();
                const double2 *src = p.ortho ? p.params : p.params_inv;
                for (int c = 0; c < 8; ++c)
                    for (int i = 0; i < 9; ++i) P.m[i] = c_add(P.m[i], src[9 * pc[c] + i]);
                for (int i = 0; i < 9; ++i) P.m[i] = c_scale(0.125, P.m[i]);
                if (!p.ortho) P = m3_inverse(P);
            } else if (p.ortho) {
                const M3 A = m3_load(p.params + 9 * p_fg), B = m3_load(p.params + 9 * p_bg);
                for (int i = 0; i < 9; ++i) P.m[i] = c_add(c_scale(rvol, A.m[i]), c_scale(1.0 - rvol, B.m[i]));
            } else {
                P = kottke_avg(m3_load(p.params + 9 * p_fg), m3_load(p.params + 9 * p_bg), nout, rvol);
            }
        }
        // Julia layout (Nx,Ny,nzl,3,3): entry [i,j,k,a,b] at ((b*3 + a)*nzl*Nxy + cell)
        if (v < 3) {
            p.out[(int64_t)(v * 3 + v) * nzl * Nxy + cell] = P.m[v * 3 + v];
        } else {
            // symmetric materials give a symmetric average; rounding in the frame rotation breaks that by an ulp, so
            // the upper triangle is written to both places - the operator then stores the off-diagonals once
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    if (a != b)
                        p.out[(int64_t)(b * 3 + a) * nzl * Nxy + cell] = (p.symmetric && a > b) ? P.m[b * 3 + a] : P.m[a * 3 + b];
        }
    }
}

void host_inverse3(const cplx *A, cplx *Ai) {
    const cplx c0 = A[4] * A[8] - A[5] * A[7], c1 = A[2] * A[7] - A[1] * A[8], c2 = A[1] * A[5] - A[2] * A[4];
    const cplx det = A[0] * c0 + A[3] * c1 + A[6] * c2;
    Ai[0] = c0 / det; Ai[1] = c1 / det; Ai[2] = c2 / det;
    Ai[3] = (A[5] * A[6] - A[3] * A[8]) / det; Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det; Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    Ai[6] = (A[3] * A[7] - A[4] * A[6]) / det; Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det; Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
}

}  // namespace

int calc_matparams(const fdfd_matparams_desc *d, fdfd_c128 *out, int where, std::string &err) {
    auto fail = [&](int code, const std::string &msg) { err = msg; return code; };
#define MP_CUDA(call)                                                                                     \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess) { cleanup(); return fail(FDFD_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } \
    } while (0)
    if (!d || !out) return fail(FDFD_EINVAL, "null argument");
    for (int w = 0; w < 3; ++w)
        if (d->N[w] < 1 || d->N[w] > (1 << 28) || !d->lprim[w]) return fail(FDFD_EINVAL, "bad grid");
    if (d->k0 < 0 || d->k1 > d->N[2] || d->k0 >= d->k1) return fail(FDFD_EINVAL, "bad plane range");
    if (d->nshape < 1 || d->nparam < 1 || !d->shapes || !d->params) return fail(FDFD_EINVAL, "no shapes / parameters");
    for (int o = 0; o < d->nshape; ++o) {
        const fdfd_shape &s = d->shapes[o];
        if (s.kind < FDFD_SHAPE_BOX || s.kind > FDFD_SHAPE_CYLINDER || s.pind < 0 || s.pind >= d->nparam || s.axis < 0 || s.axis > 2)
            return fail(FDFD_EINVAL, "bad shape " + std::to_string(o));
    }
    if (d->device >= 0) {
        cudaError_t e = cudaSetDevice(d->device);
        if (e != cudaSuccess) return fail(FDFD_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    }
    MatParams p{};
    std::vector<double> H[3];
    for (int w = 0; w < 3; ++w) {
        const int64_t N = d->N[w];
        const double *lp = d->lprim[w];
        p.N[w] = (int32_t)N;
        p.isbloch[w] = d->isbloch[w] != 0;
        // PRIM iff the field type equals boundft on this axis (ft2gt, call sites model.jl:130-131)
        const bool boundft_e = d->boundft_is_E[w] != 0, ft_e = d->field_type == FDFD_FT_EE;
        p.g[w] = (boundft_e == ft_e) ? 0 : 1;
        p.lo[w] = lp[0];
        p.hi[w] = lp[N];
        p.L[w] = lp[N] - lp[0];
        // half-step lattice: H[2i] = ghosted dual i, H[2i+1] = primal i (i = 0..N)
        H[w].resize(2 * N + 2);
        for (int64_t i = 0; i <= N; ++i) H[w][2 * i + 1] = lp[i];
        for (int64_t i = 1; i <= N; ++i) H[w][2 * i] = 0.5 * (lp[i - 1] + lp[i]);
        H[w][0] = p.isbloch[w] ? H[w][2 * N] - p.L[w] : 2.0 * lp[0] - H[w][2];
    }
    p.ortho = d->field_ortho_shape != 0;
    p.k0 = (int32_t)d->k0;
    p.k1 = (int32_t)d->k1;
    p.nshape = d->nshape;
    p.nparam = d->nparam;
    std::vector<cplx> prm(9 * (size_t)d->nparam), prm_inv(9 * (size_t)d->nparam);
    std::memcpy(prm.data(), d->params, prm.size() * sizeof(cplx));
    for (int q = 0; q < d->nparam; ++q) host_inverse3(&prm[9 * q], &prm_inv[9 * q]);
    p.symmetric = 1;
    for (int q = 0; q < d->nparam; ++q)
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (prm[9 * q + 3 * a + b] != prm[9 * q + 3 * b + a]) p.symmetric = 0;

    const int64_t nzl = d->k1 - d->k0;
    const size_t out_bytes = (size_t)9 * nzl * d->N[0] * d->N[1] * sizeof(double2);
    double *dH[3] = {nullptr, nullptr, nullptr};
    fdfd_shape *dshapes = nullptr;
    double2 *dprm = nullptr, *dinv = nullptr, *dout = nullptr;
    int32_t *derr = nullptr;
    auto cleanup = [&]() {
        for (int w = 0; w < 3; ++w) if (dH[w]) cudaFree(dH[w]);
        if (dshapes) cudaFree(dshapes);
        if (dprm) cudaFree(dprm);
        if (dinv) cudaFree(dinv);
        if (derr) cudaFree(derr);
        if (dout && where == FDFD_HOST) cudaFree(dout);
    };
    for (int w = 0; w < 3; ++w) {
        MP_CUDA(cudaMalloc((void **)&dH[w], H[w].size() * sizeof(double)));
        MP_CUDA(cudaMemcpy(dH[w], H[w].data(), H[w].size() * sizeof(double), cudaMemcpyHostToDevice));
        p.H[w] = dH[w];
    }
    MP_CUDA(cudaMalloc((void **)&dshapes, (size_t)d->nshape * sizeof(fdfd_shape)));
    MP_CUDA(cudaMemcpy(dshapes, d->shapes, (size_t)d->nshape * sizeof(fdfd_shape), cudaMemcpyHostToDevice));
    MP_CUDA(cudaMalloc((void **)&dprm, prm.size() * sizeof(cplx)));
    MP_CUDA(cudaMemcpy(dprm, prm.data(), prm.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    MP_CUDA(cudaMalloc((void **)&dinv, prm.size() * sizeof(cplx)));
    MP_CUDA(cudaMemcpy(dinv, prm_inv.data(), prm.size() * sizeof(cplx), cudaMemcpyHostToDevice));
    MP_CUDA(cudaMalloc((void **)&derr, sizeof(int32_t)));
    MP_CUDA(cudaMemset(derr, 0, sizeof(int32_t)));
    if (where == FDFD_HOST) MP_CUDA(cudaMalloc((void **)&dout, out_bytes));
    else dout = reinterpret_cast<double2 *>(out);
    MP_CUDA(cudaMemset(dout, 0, out_bytes));
    p.shapes = dshapes;
    p.params = dprm;
    p.params_inv = dinv;
    p.out = dout;
    p.err = derr;
    const dim3 grid((unsigned)((d->N[0] + MTX - 1) / MTX), (unsigned)((d->N[1] + MTY - 1) / MTY), (unsigned)((nzl + MTZ - 1) / MTZ));
    matparams_kernel<<<grid, MNT, 0, 0>>>(p);
    MP_CUDA(cudaGetLastError());
    int32_t herr = 0;
    MP_CUDA(cudaMemcpy(&herr, derr, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (herr == 0 && where == FDFD_HOST) MP_CUDA(cudaMemcpy(out, dout, out_bytes, cudaMemcpyDeviceToHost));
    MP_CUDA(cudaDeviceSynchronize());
    cleanup();
#undef MP_CUDA
    if (herr) return fail(FDFD_EINVAL, "a voxel corner is covered by no shape: add a background shape first");
    return FDFD_OK;
}

}  // namespace fdfd
