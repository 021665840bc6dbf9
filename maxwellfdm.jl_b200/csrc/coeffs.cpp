// Host-side construction of the 1-D coefficient arrays that encode the reference's difference and
// averaging operators including every boundary rule.
//
// What this replaces: the sparse-matrix builders the reference calls at
//   src/model/model.jl:171-172  create_curl(isfwd, s∆l⁻¹, isbloch, e⁻ⁱᵏᴸ)      -> create_∂
//   src/model/model.jl:152-155  create_paramop(arr, isfwd_in, s∆l, s∆l′⁻¹, ...)  -> create_mean
// (StaggeredGridCalculus / MaxwellBase, not vendored; algorithm per SURVEY.md App. A.2-A.5).
// Instead of COO triplets we keep, per axis, the two coefficients of the two-term stencil
//   (T f)[i] = t0[i] f[i] + t1[i] f[(i+s) mod N].
#include "fdfd_internal.h"

namespace fdfd {

static inline int wrap(int i, int N) { return ((i % N) + N) % N; }

// create_∂
AxisOp make_diff(bool isfwd, const std::vector<cplx> &dinv, bool isbloch, cplx ph) {
    const int N = (int)dinv.size();
    AxisOp t;
    t.shift = isfwd ? +1 : -1;
    t.t0.resize(N);
    t.t1.resize(N);
    const double sgn = isfwd ? 1.0 : -1.0;
    for (int i = 0; i < N; ++i) {
        t.t0[i] = -sgn * dinv[i];
        t.t1[i] = +sgn * dinv[i];
    }
    if (isbloch) {
        if (isfwd) t.t1[N - 1] *= ph;   // f[N+1] := e^{-ikL} f[1]
        else       t.t1[0] /= ph;       // g[0]   := g[N] / e^{-ikL}
    } else {
        if (isfwd) { t.t0[0] = 0.0; t.t1[N - 1] = 0.0; }
        else       { t.t0[0] = 0.0; t.t1[0] = 0.0; }
    }
    return t;
}

// create_mean; dl / dlo_inv == nullptr -> weights 1
AxisOp make_mean(bool isfwd, const std::vector<cplx> *dl, const std::vector<cplx> *dlo_inv, bool isbloch,
                 cplx ph, int N) {
    AxisOp t;
    t.shift = isfwd ? +1 : -1;
    t.t0.resize(N);
    t.t1.resize(N);
    for (int i = 0; i < N; ++i) {
        const cplx wo = dlo_inv ? (*dlo_inv)[i] : cplx(1.0);
        const cplx w0 = dl ? (*dl)[i] : cplx(1.0);
        const cplx w1 = dl ? (*dl)[wrap(i + t.shift, N)] : cplx(1.0);
        t.t0[i] = 0.5 * wo * w0;
        t.t1[i] = 0.5 * wo * w1;
    }
    if (isbloch) {
        if (isfwd) t.t1[N - 1] *= ph;
        else       t.t1[0] /= ph;
    } else {
        if (isfwd) { t.t0[0] = 0.0; t.t1[N - 1] = 0.0; }
        else       { t.t0[0] *= 2.0; t.t1[0] = 0.0; }   // even image: "0's and 2's"
    }
    return t;
}

// (T^T g)[j] = t0[j] g[j] + t1[j-s] g[j-s]
AxisOp transpose_op(const AxisOp &t) {
    const int N = (int)t.t0.size();
    AxisOp r;
    r.shift = -t.shift;
    r.t0 = t.t0;
    r.t1.resize(N);
    for (int j = 0; j < N; ++j) r.t1[j] = t.t1[wrap(j - t.shift, N)];
    return r;
}

static std::vector<cplx> inv(const std::vector<cplx> &v) {
    std::vector<cplx> r(v.size());
    for (size_t i = 0; i < v.size(); ++i) r[i] = 1.0 / v[i];
    return r;
}

void build_coefs(const fdfd_desc &d, const std::vector<cplx> sdl_e[3], const std::vector<cplx> sdl_m[3],
                 const cplx phase[3], CoefHost &out) {
    for (int w = 0; w < 3; ++w) {
        const bool bE = d.boundft_is_E[w] != 0;
        const bool bl = d.isbloch[w] != 0;
        const int N = (int)d.N[w];
        const std::vector<cplx> sei = inv(sdl_e[w]), smi = inv(sdl_m[w]);
        if (d.field_type == FDFD_FT_EE) {
            // Ce: isfwd = boundft.==EE, 1/sdl_m (model.jl:168,171); Cm: isfwd = boundft.==HH, 1/sdl_e (:169,172)
            out.a[w] = make_diff(bE, smi, bl, phase[w]);
            out.b[w] = make_diff(!bE, sei, bl, phase[w]);
            // Peps: isfwd_in = boundft.!=EE, weights (sdl_m, 1/sdl_e) (model.jl:149,153)
            out.mi[w] = make_mean(!bE, &sdl_m[w], &sei, bl, phase[w], N);
            if (d.weighted_out_avg) out.mo[w] = make_mean(bE, &sdl_e[w], &smi, bl, phase[w], N);
            else                    out.mo[w] = make_mean(bE, nullptr, nullptr, bl, phase[w], N);
            // create_Mcs (model.jl:299-303): Mc_e == the in-average above; Mc_m: isfwd = boundft.!=HH, (sdl_e, 1/sdl_m)
            out.mh[w] = make_mean(bE, &sdl_e[w], &smi, bl, phase[w], N);
        } else {
            // A = Ce (Peps \ Cm) - w^2 Pmu (model.jl:238-240); Pmu: isfwd_in = boundft.!=HH, (sdl_e, 1/sdl_m) (:150,155)
            out.a[w] = make_diff(!bE, sei, bl, phase[w]);
            out.b[w] = make_diff(bE, smi, bl, phase[w]);
            out.mi[w] = make_mean(bE, &sdl_e[w], &smi, bl, phase[w], N);
            if (d.weighted_out_avg) out.mo[w] = make_mean(!bE, &sdl_m[w], &sei, bl, phase[w], N);
            else                    out.mo[w] = make_mean(!bE, nullptr, nullptr, bl, phase[w], N);
            out.mh[w] = make_mean(!bE, &sdl_m[w], &sei, bl, phase[w], N);   // Mc_e for an FT_HH handle
        }
    }
}

// A^T = C1^T q C2^T + mass^T.  The transpose of a curl built from D is the curl built from -D^T,
// so the transposed operator has the SAME stencil structure (first curl shift s1, second -s1):
//   first curl'  = -(second curl)^T,  second curl' = -(first curl)^T,
//   in-average'  = (out-average)^T,   out-average' = (in-average)^T,   P'_{uv} = P_{vu}.
void transpose_coefs(const CoefHost &in, CoefHost &out) {
    for (int w = 0; w < 3; ++w) {
        out.a[w] = transpose_op(in.b[w]);
        out.b[w] = transpose_op(in.a[w]);
        for (auto &z : out.a[w].t0) z = -z;
        for (auto &z : out.a[w].t1) z = -z;
        for (auto &z : out.b[w].t0) z = -z;
        for (auto &z : out.b[w].t1) z = -z;
        out.mi[w] = transpose_op(in.mo[w]);
        out.mo[w] = transpose_op(in.mi[w]);
        out.mh[w] = in.mh[w];
    }
}

}  // namespace fdfd
