// General (any boundft / layout / field type) matrix-free apply: one thread per cell, gathers through
// L1/L2.  It is the on-device cross-check of the tiled kernel and the path for configurations the
// tiled kernel does not cover.  Not tuned; see apply_tiled.cu for the roofline kernel.
//
// Computes, per SURVEY.md App. A.4-A.5 (reference composition model.jl:236-237):
//   y = C2 ( q .* (C1 x) ) + massd .* x + Mout[ masso .* (Min x) ]
// where C1/C2/Min/Mout are given by 1-D coefficient arrays (coeffs.cpp) and massd/masso are the
// material entries pre-multiplied by -w^2.
#include <cstdlib>

#include "cplx.cuh"
#include "fdfd_internal.h"

namespace fdfd {

namespace {

__device__ __forceinline__ int g_wrap(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }

struct Gather {
    const ApplyParams &p;
    __device__ __forceinline__ int wrapx(int i) const { return i < 0 ? i + p.Nx : (i >= p.Nx ? i - p.Nx : i); }
    __device__ __forceinline__ int wrapy(int j) const { return j < 0 ? j + p.Ny : (j >= p.Ny ? j - p.Ny : j); }
    __device__ __forceinline__ int kglob(int kl) const {
        int k = p.kz0 + kl;
        return k < 0 ? k + p.Nz : (k >= p.Nz ? k - p.Nz : k);
    }
    // E_c at (i,j,kl): i,j in range, kl in [-1, nzl]
    __device__ __forceinline__ double2 E(int c, int i, int j, int kl) const {
        const double2 *base;
        int64_t cs;
        if (kl < 0) { base = p.x.lo; cs = p.x.cs_lo; }
        else if (kl >= p.nzl) { base = p.x.hi; cs = p.x.cs_hi; }
        else { base = p.x.base + (int64_t)kl * p.x.pstride; cs = p.x.cs; }
        return base[(int64_t)c * cs + ((int64_t)j * p.Nx + i) * p.x.es];
    }
    __device__ __forceinline__ int64_t gidx(int i, int j, int kl) const {
        return ((int64_t)(kl + 1) * p.Ny + j) * p.Nx + i;
    }
    // E_c at cell shifted by s along axis w
    __device__ __forceinline__ double2 Esh(int c, int i, int j, int kl, int w, int s) const {
        if (w == 0) return E(c, wrapx(i + s), j, kl);
        if (w == 1) return E(c, i, wrapy(j + s), kl);
        return E(c, i, j, kl + s);
    }
    __device__ __forceinline__ int cidx(int w, int i, int j, int kl) const {
        return w == 0 ? i : (w == 1 ? j : kglob(kl));
    }
    // first curl, component u, at (i,j,kl)
    __device__ double2 H(int u, int i, int j, int kl) const {
        const int wa = (u + 1) % 3, ca = (u + 2) % 3;
        const int wb = (u + 2) % 3, cb = (u + 1) % 3;
        const int ia = cidx(wa, i, j, kl), ib = cidx(wb, i, j, kl);
        double2 t = c_mul(p.c.a0[wa][ia], E(ca, i, j, kl));
        t = c_fma(p.c.a1[wa][ia], Esh(ca, i, j, kl, wa, p.s1[wa]), t);
        t = c_fms(p.c.a0[wb][ib], E(cb, i, j, kl), t);
        t = c_fms(p.c.a1[wb][ib], Esh(cb, i, j, kl, wb, p.s1[wb]), t);
        if (p.has_q) t = c_mul(p.q[u][gidx(i, j, kl)], t);
        return t;
    }
    __device__ __forceinline__ double2 Hsh(int u, int i, int j, int kl, int w, int s) const {
        if (w == 0) return H(u, wrapx(i + s), j, kl);
        if (w == 1) return H(u, i, wrapy(j + s), kl);
        return H(u, i, j, kl + s);
    }
    // input average of component u along its own axis, at corner (i,j,kl)
    __device__ double2 Ain(int u, int i, int j, int kl) const {
        const int iu = cidx(u, i, j, kl);
        double2 t = c_mul(p.c.mi0[u][iu], E(u, i, j, kl));
        return c_fma(p.c.mi1[u][iu], Esh(u, i, j, kl, u, -p.s1[u]), t);
    }
    // corner quantity G_v = sum_{u != v} masso_vu * Ain_u
    __device__ double2 G(int v, int i, int j, int kl) const {
        const int u1 = (v + 1) % 3, u2 = (v + 2) % 3;
        const int64_t g = gidx(i, j, kl);
        // index of (v,u) in the off-diagonal list (0,1),(0,2),(1,0),(1,2),(2,0),(2,1)
        const int e1 = 2 * v + (u1 > v ? u1 - 1 : u1);
        const int e2 = 2 * v + (u2 > v ? u2 - 1 : u2);
        double2 t = c_mul(p.mo[e1][g], Ain(u1, i, j, kl));
        return c_fma(p.mo[e2][g], Ain(u2, i, j, kl), t);
    }
    __device__ __forceinline__ double2 Gsh(int v, int i, int j, int kl, int s) const {
        if (v == 0) return G(v, wrapx(i + s), j, kl);
        if (v == 1) return G(v, i, wrapy(j + s), kl);
        return G(v, i, j, kl + s);
    }
};

__global__ void __launch_bounds__(128) apply_naive_kernel(const __grid_constant__ ApplyParams p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    if (i >= p.Nx) return;
    Gather g{p};
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const int w1 = (v + 1) % 3, u1 = (v + 2) % 3;
        const int w2 = (v + 2) % 3, u2 = (v + 1) % 3;
        const int i1 = g.cidx(w1, i, j, kl), i2 = g.cidx(w2, i, j, kl);
        double2 y = c_mul(p.c.b0[w1][i1], g.H(u1, i, j, kl));
        y = c_fma(p.c.b1[w1][i1], g.Hsh(u1, i, j, kl, w1, -p.s1[w1]), y);
        y = c_fms(p.c.b0[w2][i2], g.H(u2, i, j, kl), y);
        y = c_fms(p.c.b1[w2][i2], g.Hsh(u2, i, j, kl, w2, -p.s1[w2]), y);
        if (p.has_mass) {
            y = c_fma(p.md[v] ? p.md[v][g.gidx(i, j, kl)] : p.md_uniform, g.E(v, i, j, kl), y);
            if (p.has_off) {
                const int iv = g.cidx(v, i, j, kl);
                y = c_fma(p.c.mo0[v][iv], g.G(v, i, j, kl), y);
                y = c_fma(p.c.mo1[v][iv], g.Gsh(v, i, j, kl, p.s1[v]), y);
            }
        }
        p.y[(int64_t)kl * p.y_pstride + (int64_t)v * p.y_cs + ((int64_t)j * p.Nx + i) * p.y_es] = y;
    }
}

// h = alpha * q .* (C1 e + sj * jm)  (h_from_e on an FT_EE handle, e_from_h on an FT_HH handle; model.jl:276-284)
__global__ void __launch_bounds__(128) curl1_kernel(const __grid_constant__ ApplyParams p, const double2 *jm,
                                                     double2 alpha, double sj) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    if (i >= p.Nx) return;
    Gather g{p};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int wa = (u + 1) % 3, ca = (u + 2) % 3;
        const int wb = (u + 2) % 3, cb = (u + 1) % 3;
        const int ia = g.cidx(wa, i, j, kl), ib = g.cidx(wb, i, j, kl);
        double2 t = c_mul(p.c.a0[wa][ia], g.E(ca, i, j, kl));
        t = c_fma(p.c.a1[wa][ia], g.Esh(ca, i, j, kl, wa, p.s1[wa]), t);
        t = c_fms(p.c.a0[wb][ib], g.E(cb, i, j, kl), t);
        t = c_fms(p.c.a1[wb][ib], g.Esh(cb, i, j, kl, wb, p.s1[wb]), t);
        const int64_t o = (int64_t)kl * p.y_pstride + (int64_t)u * p.y_cs + ((int64_t)j * p.Nx + i) * p.y_es;
        if (jm) t = c_add(t, c_scale(sj, jm[o]));
        if (p.has_q) t = c_mul(p.q[u][g.gidx(i, j, kl)], t);
        p.y[o] = c_mul(alpha, t);
    }
}


// y = beta * C2 (q .* h) + gamma * je      (create_b, reference model.jl:262-265: b = -Cm (Pmu \ jm) - i w je)
// p.x holds h (may be null planes when has_h == 0).
__global__ void __launch_bounds__(128) curl2_kernel(const __grid_constant__ ApplyParams p, const double2 *je,
                                                     double2 beta, double2 gamma, int has_h, int divide_by_md) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    if (i >= p.Nx) return;
    Gather g{p};
    auto Hq = [&](int u, int ii, int jj, int kk) -> double2 {
        double2 t = g.E(u, ii, jj, kk);
        if (p.has_q) t = c_mul(p.q[u][g.gidx(ii, jj, kk)], t);
        return t;
    };
    auto Hqsh = [&](int u, int w, int s) -> double2 {
        if (w == 0) return Hq(u, g.wrapx(i + s), j, kl);
        if (w == 1) return Hq(u, i, g.wrapy(j + s), kl);
        return Hq(u, i, j, kl + s);
    };
#pragma unroll
    for (int v = 0; v < 3; ++v) {
        const int64_t o = (int64_t)kl * p.y_pstride + (int64_t)v * p.y_cs + ((int64_t)j * p.Nx + i) * p.y_es;
        double2 y = c_zero();
        if (has_h) {
            const int w1 = (v + 1) % 3, u1 = (v + 2) % 3;
            const int w2 = (v + 2) % 3, u2 = (v + 1) % 3;
            const int i1 = g.cidx(w1, i, j, kl), i2 = g.cidx(w2, i, j, kl);
            double2 t = c_mul(p.c.b0[w1][i1], Hq(u1, i, j, kl));
            t = c_fma(p.c.b1[w1][i1], Hqsh(u1, w1, -p.s1[w1]), t);
            t = c_fms(p.c.b0[w2][i2], Hq(u2, i, j, kl), t);
            t = c_fms(p.c.b1[w2][i2], Hqsh(u2, w2, -p.s1[w2]), t);
            y = c_mul(beta, t);
        }
        if (je) y = c_fma(gamma, je[o], y);
        if (divide_by_md) y = c_div(y, p.md[v] ? p.md[v][g.gidx(i, j, kl)] : p.md_uniform);   // e_from_h: divide by -w^2 eps_vv
        p.y[o] = y;
    }
}


// out_w = (M f)_w : two-point weighted mean of component w along its own axis w (create_Mcs, model.jl:287-306)
__global__ void __launch_bounds__(128) interp_kernel(const __grid_constant__ ApplyParams p, int which_other) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    const int kl = blockIdx.z;
    if (i >= p.Nx) return;
    Gather g{p};
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        const int iw = g.cidx(w, i, j, kl);
        const double2 *t0 = which_other ? p.c.mh0[w] : p.c.mi0[w], *t1 = which_other ? p.c.mh1[w] : p.c.mi1[w];
        const int sh = which_other ? p.s1[w] : -p.s1[w];
        double2 t = c_mul(t0[iw], g.E(w, i, j, kl));
        t = c_fma(t1[iw], g.Esh(w, i, j, kl, w, sh), t);
        p.y[(int64_t)kl * p.y_pstride + (int64_t)w * p.y_cs + ((int64_t)j * p.Nx + i) * p.y_es] = t;
    }
}

// y += Mout[ masso .* (Min x) ]: the off-diagonal part of the mass operator, added after the diagonal-material
// kernel when off-diagonal entries are sparse (material interfaces only).  Work item = (tile, [ks,ke)): a run of
// consecutive z-planes of one 32x8 thread tile (outputs: inner 30x6) whose corner terms can be non-zero.  The CTA
// marches the run: the corner quantity G(k+1) is computed once per plane (11 loads per thread) and reused as
// G(k) in the next step; x/y neighbours of G travel through a double-buffered shared tile.  Per axis the in-average
// looks towards -s1 and the out-average towards +s1 (s1 = direction of the first curl); the march follows s1_z.
// SKIPZ (opt-in, FDFD_CORR_SKIP_ZERO): inside a flagged block most cells still carry no off-diagonal entry (an
// interface is a surface).  A thread whose six entries are exactly zero loads no field values (its G is zero), and an
// output cell whose three corner terms are exactly zero is not read-modified-written: the pass then moves the material
// streams plus x / y only around the interface cells instead of 64-80 B per DOF of the whole block.  Results are
// identical (only additions of exact zeros are dropped).
template <bool SKIPZ>
__global__ void __launch_bounds__(256, SKIPZ ? 2 : 0) offdiag_march_kernel(const __grid_constant__ ApplyParams p,
                                                             const int4 *__restrict__ items, int ntx, int kl_begin,
                                                             int kl_end) {
    const int4 item = items[blockIdx.x];
    const int ks = max(item.y, kl_begin), ke = min(item.z, kl_end);
    __shared__ double2 gs[2][2][256];   // [plane parity][G_x, G_y][thread]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    // Fused Krylov dots: this pass changes (y,x) and (y,y) by exact deltas.  They are added to the sums the main kernel has
    // published in a FIXED order: every CTA stores its deltas in its own slot at the far end of the partial buffer, the last
    // CTA (atomic ticket) adds the slots in index order - bit-reproducible from run to run, unlike per-CTA atomic adds.
    __shared__ double red[8][3];
    __shared__ bool last_cta;
    auto publish_deltas = [&](double d_re, double d_im, double d_tt) {
        double v3[3] = {d_re, d_im, d_tt};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v3[q] += __shfl_xor_sync(0xffffffffu, v3[q], o);
            if ((tid & 31) == 0) red[tid >> 5][q] = v3[q];
        }
        __syncthreads();
        if (tid == 0) {
            double *pp = p.dot_partial + ((size_t)p.dot_cap - 1 - blockIdx.x) * 4;
            for (int q = 0; q < 3; ++q) {
                double a = 0.0;
                for (int w = 0; w < 8; ++w) a += red[w][q];
                pp[q] = a;
            }
            __threadfence();
            last_cta = atomicAdd(p.dot_ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (!last_cta) return;
        __threadfence();
        double s3[3] = {0.0, 0.0, 0.0};
        for (int b = tid; b < (int)gridDim.x; b += 256) {
            const double *pp = p.dot_partial + ((size_t)p.dot_cap - 1 - b) * 4;
            s3[0] += __ldcg(pp); s3[1] += __ldcg(pp + 1); s3[2] += __ldcg(pp + 2);
        }
        __syncthreads();   // red[] is reused
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s3[q] += __shfl_xor_sync(0xffffffffu, s3[q], o);
            if ((tid & 31) == 0) red[tid >> 5][q] = s3[q];
        }
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < 3; ++q) {
                double a = 0.0;
                for (int w = 0; w < 8; ++w) a += red[w][q];
                p.dot_out[q] += a;       // single writer: the main kernel's sums were published before this launch began
            }
            *p.dot_ticket = 0;
        }
    };
    if (ks >= ke) {
        if (p.dot_mode == 2) publish_deltas(0.0, 0.0, 0.0);
        return;
    }
    const int gi = (item.x % ntx) * 30 - 1 + tx, gj = (item.x / ntx) * 6 - 1 + ty;
    const int ci = ((gi % p.Nx) + p.Nx) % p.Nx, cj = ((gj % p.Ny) + p.Ny) % p.Ny;
    const int SGX = p.s1[0], SGY = p.s1[1], SG = p.s1[2];
    const int cim = g_wrap(ci - SGX, p.Nx), cjm = g_wrap(cj - SGY, p.Ny);   // in-average neighbour (shift -s1)
    const bool out_ok = tx >= 1 && tx <= 30 && ty >= 1 && ty <= 6 && gi < p.Nx && gj < p.Ny;
    Gather g{p};
    const double2 mi0x = p.c.mi0[0][ci], mi1x = p.c.mi1[0][ci], mi0y = p.c.mi0[1][cj], mi1y = p.c.mi1[1][cj];
    const int txp = min(max(tx + SGX, 0), 31), typ = min(max(ty + SGY, 0), 7);   // out-average neighbour (shift +s1)
    auto nonzero = [](double2 a) { return (a.x != 0.0) | (a.y != 0.0); };

    // corner quantity G(kk) at this thread's cell; ezm = E_z of plane kk-SG at this cell (in), E_z(kk) (out)
    // own-cell field of the plane handled last by corner() (needed as `s` by the fused-dot deltas); have_e tells
    // whether corner() loaded it (SKIPZ: only where the cell carries off-diagonal material)
    double2 ex = c_zero(), ey = c_zero(), ez = c_zero();
    bool have_e = false;
    auto corner = [&](int kk, double2 &ezm, double2 &Gx, double2 &Gy, double2 &Gz) {
        const int64_t m = g.gidx(ci, cj, kk);
        if (SKIPZ) {
            // a tensor stored once (aliased slots, api.cu) needs three loads, not six
            const bool sym = p.mo[2] == p.mo[0] && p.mo[4] == p.mo[1] && p.mo[5] == p.mo[3];
            const double2 o01 = p.mo[0][m], o02 = p.mo[1][m], o12 = p.mo[3][m];
            const double2 o10 = sym ? o01 : p.mo[2][m], o20 = sym ? o02 : p.mo[4][m], o21 = sym ? o12 : p.mo[5][m];
            have_e = nonzero(o01) | nonzero(o02) | nonzero(o10) | nonzero(o12) | nonzero(o20) | nonzero(o21);
            if (!have_e) {
                Gx = Gy = Gz = c_zero();
                return;
            }
            ex = g.E(0, ci, cj, kk); ey = g.E(1, ci, cj, kk); ez = g.E(2, ci, cj, kk);
            const int kg = g.kglob(kk);
            const double2 Ax = c_fma(mi1x, g.E(0, cim, cj, kk), c_mul(mi0x, ex));
            const double2 Ay = c_fma(mi1y, g.E(1, ci, cjm, kk), c_mul(mi0y, ey));
            const double2 Az = c_fma(p.c.mi1[2][kg], g.E(2, ci, cj, kk - SG), c_mul(p.c.mi0[2][kg], ez));
            Gx = c_fma(o02, Az, c_mul(o01, Ay));
            Gy = c_fma(o12, Az, c_mul(o10, Ax));
            Gz = c_fma(o21, Ay, c_mul(o20, Ax));
            return;
        }
        ex = g.E(0, ci, cj, kk); ey = g.E(1, ci, cj, kk); ez = g.E(2, ci, cj, kk);
        have_e = true;
        const int kg = g.kglob(kk);
        const double2 Ax = c_fma(mi1x, g.E(0, cim, cj, kk), c_mul(mi0x, ex));
        const double2 Ay = c_fma(mi1y, g.E(1, ci, cjm, kk), c_mul(mi0y, ey));
        const double2 Az = c_fma(p.c.mi1[2][kg], ezm, c_mul(p.c.mi0[2][kg], ez));
        Gx = c_fma(p.mo[1][m], Az, c_mul(p.mo[0][m], Ay));
        Gy = c_fma(p.mo[3][m], Az, c_mul(p.mo[2][m], Ax));
        Gz = c_fma(p.mo[5][m], Ay, c_mul(p.mo[4][m], Ax));
        ezm = ez;
    };

    const int kfirst = SG < 0 ? ke - 1 : ks;
    double2 ezm = SKIPZ ? c_zero() : g.E(2, ci, cj, kfirst - SG);
    double2 Gcx, Gcy, Gcz;
    corner(kfirst, ezm, Gcx, Gcy, Gcz);
    gs[0][0][tid] = Gcx;
    gs[0][1][tid] = Gcy;
    __syncthreads();
    double d_re = 0.0, d_im = 0.0, d_tt = 0.0;   // fused Krylov dots: exact change of (y,x), (y,y) caused by this pass
    for (int st = 0; st < ke - ks; ++st) {
        const int k = kfirst + SG * st;
        double2 sx = ex, sy = ey, sz = ez;            // x at this cell, plane k (valid if have_s)
        const bool have_s = have_e;
        // neighbours of G(k): written one step ago, visible since the last barrier; read BEFORE this step's barrier
        const double2 Gx_xp = gs[st & 1][0][ty * 32 + txp], Gy_yp = gs[st & 1][1][typ * 32 + tx];
        double2 Gnx, Gny, Gnz;
        corner(k + SG, ezm, Gnx, Gny, Gnz);
        gs[(st + 1) & 1][0][tid] = Gnx;
        gs[(st + 1) & 1][1][tid] = Gny;
        if (out_ok) {
            const int kg = g.kglob(k);
            const double2 tx_ = c_fma(p.c.mo1[0][ci], Gx_xp, c_mul(p.c.mo0[0][ci], Gcx));
            const double2 ty_ = c_fma(p.c.mo1[1][cj], Gy_yp, c_mul(p.c.mo0[1][cj], Gcy));
            const double2 tz_ = c_fma(p.c.mo1[2][kg], Gnz, c_mul(p.c.mo0[2][kg], Gcz));
            if (!SKIPZ || (nonzero(tx_) | nonzero(ty_) | nonzero(tz_))) {
                double2 *yo = &p.y[(int64_t)k * p.y_pstride + ((int64_t)gj * p.Nx + gi) * p.y_es];
                const double2 o0 = yo[0], o1 = yo[p.y_cs], o2 = yo[2 * p.y_cs];
                const double2 n0 = c_add(o0, tx_), n1 = c_add(o1, ty_), n2 = c_add(o2, tz_);
                yo[0] = n0;
                yo[p.y_cs] = n1;
                yo[2 * p.y_cs] = n2;
                if (p.dot_mode == 2) {
                    if (SKIPZ && !have_s) { sx = g.E(0, ci, cj, k); sy = g.E(1, ci, cj, k); sz = g.E(2, ci, cj, k); }
                    d_re += tx_.x * sx.x + tx_.y * sx.y + ty_.x * sy.x + ty_.y * sy.y + tz_.x * sz.x + tz_.y * sz.y;
                    d_im += tx_.x * sx.y - tx_.y * sx.x + ty_.x * sy.y - ty_.y * sy.x + tz_.x * sz.y - tz_.y * sz.x;
                    // |o + d|^2 - |o|^2 = 2 Re(conj(o) d) + |d|^2  (no cancellation when the correction is small)
                    d_tt += 2.0 * (o0.x * tx_.x + o0.y * tx_.y + o1.x * ty_.x + o1.y * ty_.y + o2.x * tz_.x + o2.y * tz_.y) +
                            (tx_.x * tx_.x + tx_.y * tx_.y + ty_.x * ty_.x + ty_.y * ty_.y + tz_.x * tz_.x + tz_.y * tz_.y);
                }
            }
        }
        Gcx = Gnx;
        Gcy = Gny;
        Gcz = Gnz;
        __syncthreads();
    }
    if (p.dot_mode == 2) publish_deltas(d_re, d_im, d_tt);
}

}  // namespace

cudaError_t launch_apply_naive(const ApplyParams &p, cudaStream_t s) {
    dim3 block(128, 1, 1);
    dim3 grid((p.Nx + 127) / 128, p.Ny, p.nzl);
    apply_naive_kernel<<<grid, block, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_offdiag_correction(const ApplyParams &p, const int4 *items, int count, int ntx, int kl_begin,
                                      int kl_end, cudaStream_t s) {
    if (count <= 0) return cudaSuccess;
    if (p.dot_mode == 2 && (int64_t)count + 8192 > p.dot_cap) return cudaErrorInvalidConfiguration;   // delta slots
    static const bool skipz = getenv("FDFD_CORR_SKIP_ZERO") != nullptr;   // opt-in until it has been timed on hardware
    if (skipz) offdiag_march_kernel<true><<<count, 256, 0, s>>>(p, items, ntx, kl_begin, kl_end);
    else       offdiag_march_kernel<false><<<count, 256, 0, s>>>(p, items, ntx, kl_begin, kl_end);
    return cudaGetLastError();
}

cudaError_t launch_interp(const ApplyParams &p, int which_other, cudaStream_t s) {
    dim3 block(128, 1, 1);
    dim3 grid((p.Nx + 127) / 128, p.Ny, p.nzl);
    interp_kernel<<<grid, block, 0, s>>>(p, which_other);
    return cudaGetLastError();
}

cudaError_t launch_curl1(const ApplyParams &p, const double2 *jm, double2 alpha, double sj, cudaStream_t s) {
    dim3 block(128, 1, 1);
    dim3 grid((p.Nx + 127) / 128, p.Ny, p.nzl);
    curl1_kernel<<<grid, block, 0, s>>>(p, jm, alpha, sj);
    return cudaGetLastError();
}

}  // namespace fdfd

namespace fdfd {
cudaError_t launch_curl2(const ApplyParams &p, const double2 *je, double2 beta, double2 gamma, int has_h,
                         cudaStream_t s, int divide_by_md) {
    dim3 block(128, 1, 1);
    dim3 grid((p.Nx + 127) / 128, p.Ny, p.nzl);
    curl2_kernel<<<grid, block, 0, s>>>(p, je, beta, gamma, has_h, divide_by_md);
    return cudaGetLastError();
}
}  // namespace fdfd
