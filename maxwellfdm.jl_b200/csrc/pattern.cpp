// K5 - debug export of the assembled operator as a Julia SparseMatrixCSC{ComplexF64,Int64}
// (1-based colptr / rowval, rows sorted within a column), bit-exact in the index arrays.
//
// Reference-defined integer work reproduced here: the result of create_A (src/model/model.jl:236-237)
//     A = Cm * (Pmu \ Ce);  iszero(w) || (A -= w^2 * Peps)
// under the sparse-pattern rule set of SURVEY.md App. A.6:
//   1. sparse(I,J,V) keeps explicit zeros (boundary rules zero VALUES, never positions);
//   2. `Pmu \ Ce` rescales, pattern unchanged;   3. SpGEMM keeps the structural product pattern;
//   4. w != 0: sparse subtraction drops results that are exactly zero; w == 0: subtraction skipped,
//      so the structural pattern (explicit zeros included) is the answer.
// Host code (integer/pointer work, one-off, debug only); columns are independent -> OpenMP.
#include <algorithm>
#include <cstring>

#include "fdfd_internal.h"

namespace fdfd {

namespace {

struct Ent {
    int64_t row;
    cplx val;
};

inline int levi(int v, int w, int u) { return (v - w) * (w - u) * (u - v) / 2; }

struct Exporter {
    const Ctx *c;
    CoefHost cf;
    int64_t N[3], M;
    int s1[3];
    bool cmpfirst, has_mass, has_off;
    const std::vector<cplx> *mass, *mid;

    int64_t cell_of(const int64_t ijk[3]) const { return ijk[0] + N[0] * (ijk[1] + N[1] * ijk[2]); }
    int64_t dof(int64_t cell, int cmp) const { return cmpfirst ? 3 * cell + cmp : M * cmp + cell; }
    static int64_t wrap(int64_t i, int64_t n) { return ((i % n) + n) % n; }

    // all stored entries of column (cell0, u), unsorted, duplicates not merged; returns count
    int column(int64_t cell0, int u, Ent *out) const {
        int n = 0;
        int64_t c0[3] = {cell0 % N[0], (cell0 / N[0]) % N[1], cell0 / (N[0] * N[1])};
        // ---- curl-curl: paths  E_u(cell0) --C1--> H_v'(cellh) --q--> --C2--> y_v(cell)
        for (int vp = 0; vp < 3; ++vp) {
            if (vp == u) continue;
            const int w = 3 - u - vp;
            const double sg1 = levi(vp, w, u);
            for (int e1 = 0; e1 < 2; ++e1) {
                int64_t ch[3] = {c0[0], c0[1], c0[2]};
                if (e1 == 1) ch[w] = wrap(c0[w] - s1[w], N[w]);
                const cplx a = (e1 == 0 ? cf.a[w].t0 : cf.a[w].t1)[ch[w]];
                const int64_t cellh = cell_of(ch);
                cplx x = sg1 * a;
                if (mid && !mid->empty()) x = x / (*mid)[cellh + M * (vp + 3 * vp)];
                for (int v = 0; v < 3; ++v) {
                    if (v == vp) continue;
                    const int w2 = 3 - v - vp;
                    const double sg2 = levi(v, w2, vp);
                    for (int e2 = 0; e2 < 2; ++e2) {
                        int64_t cr[3] = {ch[0], ch[1], ch[2]};
                        if (e2 == 1) cr[w2] = wrap(ch[w2] + s1[w2], N[w2]);
                        const cplx b = (e2 == 0 ? cf.b[w2].t0 : cf.b[w2].t1)[cr[w2]];
                        out[n++] = Ent{dof(cell_of(cr), v), (sg2 * b) * x};
                    }
                }
            }
        }
        if (!has_mass) return n;
        // ---- mass operator: -w^2 * P
        const cplx w2 = -(c->omega * c->omega);
        const cplx pd = (mass && !mass->empty()) ? (*mass)[cell0 + M * (u + 3 * u)] : cplx(1.0);
        out[n++] = Ent{dof(cell0, u), w2 * pd};
        if (has_off) {
            for (int v = 0; v < 3; ++v) {
                if (v == u) continue;
                for (int e1 = 0; e1 < 2; ++e1) {
                    int64_t cg[3] = {c0[0], c0[1], c0[2]};
                    if (e1 == 1) cg[u] = wrap(c0[u] + s1[u], N[u]);
                    const cplx mi = (e1 == 0 ? cf.mi[u].t0 : cf.mi[u].t1)[cg[u]];
                    const int64_t cellg = cell_of(cg);
                    const cplx pvu = (*mass)[cellg + M * (v + 3 * u)];
                    for (int e2 = 0; e2 < 2; ++e2) {
                        int64_t cr[3] = {cg[0], cg[1], cg[2]};
                        if (e2 == 1) cr[v] = wrap(cg[v] - s1[v], N[v]);
                        const cplx mo = (e2 == 0 ? cf.mo[v].t0 : cf.mo[v].t1)[cr[v]];
                        out[n++] = Ent{dof(cell_of(cr), v), w2 * (mo * (pvu * mi))};
                    }
                }
            }
        }
        return n;
    }

    // sort by row, merge duplicates, apply the zero-dropping rule; returns final count
    int finish(Ent *e, int n) const {
        std::sort(e, e + n, [](const Ent &a, const Ent &b) { return a.row < b.row; });
        int m = 0;
        for (int i = 0; i < n;) {
            Ent acc = e[i];
            int j = i + 1;
            for (; j < n && e[j].row == acc.row; ++j) acc.val += e[j].val;
            i = j;
            if (has_mass && acc.val == cplx(0.0)) continue;  // rule 4 (w != 0)
            e[m++] = acc;
        }
        return m;
    }
};

}  // namespace

int export_pattern(Ctx *c, int64_t *colptr, int64_t *rowval, fdfd_c128 *nzval, int64_t *nnz_inout) {
    if (!nnz_inout) return set_err(c, FDFD_EINVAL, "fdfd_export_pattern: nnz_inout is null");
    if (c->d.nranks != 1) return set_err(c, FDFD_EINVAL, "fdfd_export_pattern: single-slab handles only");
    if (!c->have_coeffs) return set_err(c, FDFD_ESTATE, "fdfd_set_coeffs has not been called");
    const bool ee = c->d.field_type == FDFD_FT_EE;
    Exporter ex;
    ex.c = c;
    build_coefs(c->d, c->sdl_e, c->sdl_m, c->phase, ex.cf);
    for (int w = 0; w < 3; ++w) { ex.N[w] = c->d.N[w]; ex.s1[w] = ex.cf.a[w].shift; }
    ex.M = ex.N[0] * ex.N[1] * ex.N[2];
    ex.cmpfirst = c->d.order_cmpfirst != 0;
    ex.has_mass = c->omega != cplx(0.0);
    ex.mass = ee ? &c->eps_host : &c->mu_host;
    ex.mid = ee ? &c->mu_host : &c->eps_host;
    if (ex.has_mass && ee && !c->have_eps) return set_err(c, FDFD_ESTATE, "fdfd_set_eps has not been called");
    if (!ee && !c->have_eps) return set_err(c, FDFD_ESTATE, "FT_HH needs fdfd_set_eps");
    ex.has_off = ex.has_mass && (ee ? c->eps_off : c->mu_off);
    const int64_t n = 3 * ex.M;
    const bool fill = colptr && rowval;
    std::vector<int64_t> counts;
    try {
        counts.assign((size_t)n + 1, 0);
    } catch (const std::bad_alloc &) {
        return set_err(c, FDFD_ENOMEM, "fdfd_export_pattern: out of host memory");
    }
#pragma omp parallel for schedule(static)
    for (int64_t col = 0; col < n; ++col) {
        Ent buf[32];
        const int64_t cell = ex.cmpfirst ? col / 3 : col % ex.M;
        const int u = ex.cmpfirst ? (int)(col % 3) : (int)(col / ex.M);
        counts[col + 1] = ex.finish(buf, ex.column(cell, u, buf));
    }
    for (int64_t col = 0; col < n; ++col) counts[col + 1] += counts[col];
    const int64_t nnz = counts[n];
    if (!fill) {
        *nnz_inout = nnz;
        return FDFD_OK;
    }
    if (*nnz_inout < nnz) {
        *nnz_inout = nnz;
        return set_err(c, FDFD_EINVAL, "fdfd_export_pattern: rowval/nzval capacity too small");
    }
    *nnz_inout = nnz;
#pragma omp parallel for schedule(static)
    for (int64_t col = 0; col < n; ++col) {
        Ent buf[32];
        const int64_t cell = ex.cmpfirst ? col / 3 : col % ex.M;
        const int u = ex.cmpfirst ? (int)(col % 3) : (int)(col / ex.M);
        const int m = ex.finish(buf, ex.column(cell, u, buf));
        int64_t o = counts[col];
        for (int k = 0; k < m; ++k, ++o) {
            rowval[o] = buf[k].row + 1;  // Julia is 1-based
            if (nzval) {
                nzval[o].re = buf[k].val.real();
                nzval[o].im = buf[k].val.imag();
            }
        }
        colptr[col] = counts[col] + 1;
    }
    colptr[n] = nnz + 1;
    return FDFD_OK;
}

}  // namespace fdfd
