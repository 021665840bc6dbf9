// QMR (Freund & Nachtigal, coupled two-term recurrences without look-ahead, unpreconditioned) for the
// complex non-Hermitian FDFD system.  Uses A and the plain transpose A^T (fdfd_apply_transpose: same
// stencil kernel with transposed 1-D coefficient tables) and the bilinear form w^T v.
// BASELINE.json north_star names QMR next to BiCGSTAB; the reference has no solver (README.md:27-33).
//
// With M = I the textbook vectors y == v~ and z == w~, so the state is
//   vt (v~), wt (w~), p, q, pt = A p, qt = A^T q, d, s, r      (9 work vectors) + x.
// The residual recurrence (s = eta pt + c s, r -= s: five vector passes per iteration) only serves the residual
// norm, so it runs only when a per-iteration history is requested; otherwise the TRUE residual b - A x is
// evaluated at the convergence checks (one extra apply per `check_every` iterations) and an iteration moves
// 2 applies + 19 vector passes instead of 24.
// Scalars are double-buffered by iteration parity so that a kernel never reads a slot another block of
// the same launch is writing.
#include <cmath>
#include <cstdlib>

#include "krylov_common.cuh"

namespace fdfd {

namespace {

using namespace kry;

// slot layout: parity block of 8 at QB(par) = 8*par: RHO2, XI2, DELTA, EPS, GAMMA, THETA, ETA
enum { Q_RHO2 = 0, Q_XI2 = 1, Q_DELTA = 2, Q_EPS = 3, Q_GAMMA = 4, Q_THETA = 5, Q_ETA = 6 };
enum { Q_RR = 16, Q_BNORM = 17 };
__host__ __device__ inline int QB(int par) { return 8 * par; }

// r = b - r ; vt = wt = r ; p = q = d = s = 0 ; RHO2 = XI2 = ||r||^2, DELTA = r^T r ; RR, BNORM
__global__ void __launch_bounds__(RB) q_init(int64_t n, const double2 *__restrict__ b, double2 *__restrict__ r,
                                             double2 *__restrict__ vt, double2 *__restrict__ wt,
                                             double2 *__restrict__ p, double2 *__restrict__ q,
                                             double2 *__restrict__ d, double2 *__restrict__ s, Red rd) {
    double2 acc[3] = {c_zero(), c_zero(), c_zero()};
    double2 acc2[2] = {c_zero(), c_zero()};
    GRID_STRIDE(i, n) {
        const double2 bb = b[i];
        const double2 rr = c_sub(bb, r[i]);
        r[i] = rr;
        vt[i] = rr;
        wt[i] = rr;
        p[i] = q[i] = d[i] = s[i] = c_zero();
        dot_acc(acc[0], rr, rr);
        dotu_acc(acc[2], rr, rr);
        dot_acc(acc2[1], bb, bb);
    }
    acc[1] = acc[0];
    acc2[0] = acc[0];
    reduce_publish<3>(acc, rd, QB(0) + Q_RHO2);
    reduce_publish<2>(acc2, rd, Q_RR);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rd.scal[QB(0) + Q_EPS] = c_make(1.0, 0.0);
        rd.scal[QB(0) + Q_GAMMA] = c_make(1.0, 0.0);
        rd.scal[QB(0) + Q_THETA] = c_make(0.0, 0.0);
        rd.scal[QB(0) + Q_ETA] = c_make(-1.0, 0.0);
    }
}

// p = vt/rho - (xi delta/eps_prev) p ; q = wt/xi - (rho delta/eps_prev) q
__global__ void __launch_bounds__(RB) q_pq(int64_t n, const double2 *__restrict__ vt, const double2 *__restrict__ wt,
                                           double2 *__restrict__ p, double2 *__restrict__ q, Red rd, int par) {
    const double rho = sqrt(rd.scal[QB(par) + Q_RHO2].x), xi = sqrt(rd.scal[QB(par) + Q_XI2].x);
    const double2 delta = c_scale(1.0 / (rho * xi), rd.scal[QB(par) + Q_DELTA]);
    const double2 de = c_div(delta, rd.scal[QB(par) + Q_EPS]);
    const double2 cp = c_scale(xi, de), cq = c_scale(rho, de);
    const double ir = 1.0 / rho, ix = 1.0 / xi;
    GRID_STRIDE(i, n) {
        p[i] = c_fms(cp, p[i], c_scale(ir, vt[i]));
        q[i] = c_fms(cq, q[i], c_scale(ix, wt[i]));
    }
}

// EPS' = q^T pt
__global__ void __launch_bounds__(RB) q_eps(int64_t n, const double2 *__restrict__ q, const double2 *__restrict__ pt,
                                            Red rd, int par) {
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) dotu_acc(acc[0], q[i], pt[i]);
    reduce_publish<1>(acc, rd, QB(par ^ 1) + Q_EPS);
}

__device__ __forceinline__ double2 q_beta(const Red &rd, int par, double &rho, double &xi) {
    rho = sqrt(rd.scal[QB(par) + Q_RHO2].x);
    xi = sqrt(rd.scal[QB(par) + Q_XI2].x);
    const double2 delta = c_scale(1.0 / (rho * xi), rd.scal[QB(par) + Q_DELTA]);
    return c_div(rd.scal[QB(par ^ 1) + Q_EPS], delta);
}

// beta = eps/delta ; vt = pt - (beta/rho) vt ; wt = qt - (beta/xi) wt ; RHO2', XI2', DELTA' of the new vt, wt
__global__ void __launch_bounds__(RB) q_vw(int64_t n, const double2 *__restrict__ pt, const double2 *__restrict__ qt,
                                           double2 *__restrict__ vt, double2 *__restrict__ wt, Red rd, int par) {
    double rho, xi;
    const double2 beta = q_beta(rd, par, rho, xi);
    const double2 bv = c_scale(1.0 / rho, beta), bw = c_scale(1.0 / xi, beta);
    double2 acc[3] = {c_zero(), c_zero(), c_zero()};
    GRID_STRIDE(i, n) {
        const double2 v = c_fms(bv, vt[i], pt[i]);
        const double2 w = c_fms(bw, wt[i], qt[i]);
        vt[i] = v;
        wt[i] = w;
        dot_acc(acc[0], v, v);
        dot_acc(acc[1], w, w);
        dotu_acc(acc[2], w, v);
    }
    reduce_publish<3>(acc, rd, QB(par ^ 1) + Q_RHO2);
}

// theta, gamma, eta ; d = eta p + (theta_prev gamma)^2 d ; s = eta pt + (theta_prev gamma)^2 s ; x += d ; r -= s
__global__ void __launch_bounds__(RB) q_xr(int64_t n, const double2 *__restrict__ p, const double2 *__restrict__ pt,
                                           double2 *__restrict__ d, double2 *__restrict__ s, double2 *__restrict__ x,
                                           double2 *__restrict__ r, Red rd, int par) {
    double rho, xi;
    const double2 beta = q_beta(rd, par, rho, xi);
    const double rho1 = sqrt(rd.scal[QB(par ^ 1) + Q_RHO2].x);
    const double g0 = rd.scal[QB(par) + Q_GAMMA].x, th0 = rd.scal[QB(par) + Q_THETA].x;
    const double2 eta0 = rd.scal[QB(par) + Q_ETA];
    const double absb = sqrt(beta.x * beta.x + beta.y * beta.y);
    const double th = rho1 / (g0 * absb);
    const double g = 1.0 / sqrt(1.0 + th * th);
    // eta = -eta0 * rho * g^2 / (beta * g0^2)
    const double2 eta = c_div(c_scale(-rho * g * g / (g0 * g0), eta0), beta);
    const double c2 = (th0 * g) * (th0 * g);
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) {
        const double2 dd = c_fma(eta, p[i], c_scale(c2, d[i]));
        const double2 ss = c_fma(eta, pt[i], c_scale(c2, s[i]));
        d[i] = dd;
        s[i] = ss;
        x[i] = c_add(x[i], dd);
        const double2 rr = c_sub(r[i], ss);
        r[i] = rr;
        dot_acc(acc[0], rr, rr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rd.scal[QB(par ^ 1) + Q_GAMMA] = c_make(g, 0.0);
        rd.scal[QB(par ^ 1) + Q_THETA] = c_make(th, 0.0);
        rd.scal[QB(par ^ 1) + Q_ETA] = eta;
    }
    reduce_publish<1>(acc, rd, Q_RR);
}

// the same update without the residual recurrence: d = eta p + (theta_prev gamma)^2 d ; x += d
__global__ void __launch_bounds__(RB) q_xd(int64_t n, const double2 *__restrict__ p, double2 *__restrict__ d,
                                           double2 *__restrict__ x, Red rd, int par) {
    double rho, xi;
    const double2 beta = q_beta(rd, par, rho, xi);
    const double rho1 = sqrt(rd.scal[QB(par ^ 1) + Q_RHO2].x);
    const double g0 = rd.scal[QB(par) + Q_GAMMA].x, th0 = rd.scal[QB(par) + Q_THETA].x;
    const double2 eta0 = rd.scal[QB(par) + Q_ETA];
    const double absb = sqrt(beta.x * beta.x + beta.y * beta.y);
    const double th = rho1 / (g0 * absb);
    const double g = 1.0 / sqrt(1.0 + th * th);
    const double2 eta = c_div(c_scale(-rho * g * g / (g0 * g0), eta0), beta);
    const double c2 = (th0 * g) * (th0 * g);
    GRID_STRIDE(i, n) {
        const double2 dd = c_fma(eta, p[i], c_scale(c2, d[i]));
        d[i] = dd;
        x[i] = c_add(x[i], dd);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rd.scal[QB(par ^ 1) + Q_GAMMA] = c_make(g, 0.0);
        rd.scal[QB(par ^ 1) + Q_THETA] = c_make(th, 0.0);
        rd.scal[QB(par ^ 1) + Q_ETA] = eta;
    }
}

// r = b - r (r holds A x on entry) ; RR = ||r||^2
__global__ void __launch_bounds__(RB) q_resid(int64_t n, const double2 *__restrict__ b, double2 *__restrict__ r, Red rd) {
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) {
        const double2 rr = c_sub(b[i], r[i]);
        r[i] = rr;
        dot_acc(acc[0], rr, rr);
    }
    reduce_publish<1>(acc, rd, Q_RR);
}

__global__ void q_store_hist(double *hist, int idx, const double2 *scal) {
    hist[idx] = sqrt(scal[Q_RR].x / scal[Q_BNORM].x);
}

}  // namespace

int qmr(Ctx *c, const double2 *b, double2 *x, double rtol, int maxit, int check_every, bool fixed_iters, int *iters,
        double *relres, double *hist) {
    int rc = ensure_ready(c);
    if (rc != FDFD_OK) return rc;
    if ((rc = kry::workspace(c, 9)) != FDFD_OK) return rc;
    const int64_t n = c->nloc;
    double2 *r = c->work, *vt = r + n, *wt = vt + n, *p = wt + n, *q = p + n, *pt = q + n, *qt = pt + n, *d = qt + n,
            *s = d + n;
    kry::Red rd = kry::make_red(c);
    double *sc = c->scal;
    const int g = kry::grid_for(n);
    cudaStream_t st = c->stream;
    double *hist_dev = nullptr;
    if (hist) FDFD_CUDA(c, cudaMalloc((void **)&hist_dev, sizeof(double) * (size_t)(maxit + 1)));
    auto cleanup = [&]() { if (hist_dev) cudaFree(hist_dev); };
#define KCHK(expr) do { int r__ = (expr); if (r__ != FDFD_OK) { cleanup(); return r__; } } while (0)
#define LCHK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { cleanup(); return set_err(c, FDFD_ECUDA, cudaGetErrorString(e__)); } } while (0)

    KCHK(apply_device(c, x, r, false));
    q_init<<<g, kry::RB, 0, st>>>(n, b, r, vt, wt, p, q, d, s, rd);
    LCHK();
    c->launches += 1;
    KCHK(allreduce_sum(c, sc + 2 * (QB(0) + Q_RHO2), 6, st));
    KCHK(allreduce_sum(c, sc + 2 * Q_RR, 4, st));
    if (hist_dev) { q_store_hist<<<1, 1, 0, st>>>(hist_dev, 0, rd.scal); c->launches += 1; }

    auto read_relres = [&](double &out) -> int {
        FDFD_CUDA(c, cudaMemcpyAsync(c->scal_host, sc, sizeof(double2) * kry::NSLOT, cudaMemcpyDeviceToHost, st));
        FDFD_CUDA(c, cudaStreamSynchronize(st));
        const double rr = c->scal_host[2 * Q_RR], bn = c->scal_host[2 * Q_BNORM];
        out = bn > 0 ? std::sqrt(rr / bn) : std::sqrt(rr);
        return FDFD_OK;
    };
    double rel = 1.0;
    int it = 0;
    bool converged = false;
    if (!fixed_iters) {
        KCHK(read_relres(rel));
        if (c->scal_host[2 * Q_BNORM] == 0.0) {
            FDFD_CUDA(c, cudaMemsetAsync(x, 0, sizeof(double2) * (size_t)n, st));
            FDFD_CUDA(c, cudaStreamSynchronize(st));
            if (iters) *iters = 0;
            if (relres) *relres = 0.0;
            if (hist) hist[0] = 0.0;
            cleanup();
            return FDFD_OK;
        }
        converged = rel <= rtol;
    }
    const bool track = hist_dev != nullptr;     // per-iteration residual history: keep the s / r recurrence
    // true residual at a convergence check: r = b - A x
    auto true_residual = [&]() -> int {
        int r1 = apply_device(c, x, r, false);
        if (r1 != FDFD_OK) return r1;
        q_resid<<<g, kry::RB, 0, st>>>(n, b, r, rd);
        cudaError_t e1 = cudaGetLastError();
        if (e1 != cudaSuccess) return set_err(c, FDFD_ECUDA, cudaGetErrorString(e1));
        c->launches += 1;
        return allreduce_sum(c, sc + 2 * Q_RR, 2, st);
    };
    auto enqueue_iter = [&](int i) -> int {
        const int par = i & 1;
        int r1;
        q_pq<<<g, kry::RB, 0, st>>>(n, vt, wt, p, q, rd, par);
        if ((r1 = apply_device(c, p, pt, false)) != FDFD_OK) return r1;
        if ((r1 = apply_device(c, q, qt, true)) != FDFD_OK) return r1;
        q_eps<<<g, kry::RB, 0, st>>>(n, q, pt, rd, par);
        if ((r1 = allreduce_sum(c, sc + 2 * (QB(par ^ 1) + Q_EPS), 2, st)) != FDFD_OK) return r1;
        q_vw<<<g, kry::RB, 0, st>>>(n, pt, qt, vt, wt, rd, par);
        if ((r1 = allreduce_sum(c, sc + 2 * (QB(par ^ 1) + Q_RHO2), 6, st)) != FDFD_OK) return r1;
        if (track) {
            q_xr<<<g, kry::RB, 0, st>>>(n, p, pt, d, s, x, r, rd, par);
            if ((r1 = allreduce_sum(c, sc + 2 * Q_RR, 2, st)) != FDFD_OK) return r1;
        } else {
            q_xd<<<g, kry::RB, 0, st>>>(n, p, d, x, rd, par);
        }
        cudaError_t e1 = cudaGetLastError();
        if (e1 != cudaSuccess) return set_err(c, FDFD_ECUDA, cudaGetErrorString(e1));
        c->launches += 4;
        return FDFD_OK;
    };
    // small grids: replay two iterations (both scalar parities) from a CUDA graph (see krylov.cu)
    cudaGraphExec_t gexec = nullptr;
    int64_t graph_launches = 0;
    if (c->d.nranks == 1 && n <= 6000000 && !hist_dev && maxit >= 4 && !getenv("FDFD_NO_GRAPH")) {
        cudaGraph_t graph = nullptr;
        const int64_t l0 = c->launches;
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            int rc1 = enqueue_iter(0);
            if (rc1 == FDFD_OK) rc1 = enqueue_iter(1);
            cudaError_t ec = cudaStreamEndCapture(st, &graph);
            if (rc1 == FDFD_OK && ec == cudaSuccess && graph) {
                if (cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
            }
            if (graph) cudaGraphDestroy(graph);
            (void)cudaGetLastError();
        }
        graph_launches = c->launches - l0;
        c->launches = l0;
    }
    auto cleanup2 = [&]() { if (gexec) cudaGraphExecDestroy(gexec); };
    while (!converged && it < maxit) {
        const int to_check = check_every - (it % check_every);
        if (gexec && (it & 1) == 0 && it + 2 <= maxit && (fixed_iters || to_check >= 2)) {
            cudaError_t eg = cudaGraphLaunch(gexec, st);
            if (eg != cudaSuccess) { cleanup2(); cleanup(); return set_err(c, FDFD_ECUDA, cudaGetErrorString(eg)); }
            c->launches += graph_launches;
            it += 2;
        } else {
            int ri;
            {
                BurstTurn turn(c);   // slabs of one process take turns issuing an iteration (fdfd_internal.h)
                ri = enqueue_iter(it);
            }
            if (ri != FDFD_OK) { cleanup2(); cleanup(); return ri; }
            ++it;
        }
        if (hist_dev) { q_store_hist<<<1, 1, 0, st>>>(hist_dev, it, rd.scal); c->launches += 1; }
        if (!fixed_iters && (it % check_every == 0 || it >= maxit)) {
            if (!track) {
                int rt = true_residual();
                if (rt != FDFD_OK) { cleanup2(); cleanup(); return rt; }
            }
            int rq = read_relres(rel);
            if (rq != FDFD_OK) { cleanup2(); cleanup(); return rq; }
            if (!(rel == rel)) break;
            converged = rel <= rtol;
        }
    }
    cleanup2();
    if (fixed_iters) {
        if (!track) KCHK(true_residual());
        KCHK(read_relres(rel));
    }
    if (hist) {
        FDFD_CUDA(c, cudaMemcpyAsync(hist, hist_dev, sizeof(double) * (size_t)(it + 1), cudaMemcpyDeviceToHost, st));
        FDFD_CUDA(c, cudaStreamSynchronize(st));
    }
    cleanup();
#undef KCHK
#undef LCHK
    if (iters) *iters = it;
    if (relres) *relres = rel;
    if (fixed_iters) return FDFD_OK;
    if (!converged) return set_err(c, FDFD_ENOCONV, "QMR: not converged within maxit");
    return FDFD_OK;
}

}  // namespace fdfd
