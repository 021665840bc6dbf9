// K2 - Krylov drivers on the device: BiCGSTAB (van der Vorst) and QMR for A x = b with the matrix-free
// operator.  The reference stops at create_linsys -> (A, b) (src/model/model.jl:209-220) and leaves the
// solve to the user (README.md:27-33); BASELINE.json's north_star defines this loop.
//
// All scalars (rho, alpha, omega, inner products) live in device memory; the vector kernels read them
// and derive what they need, so an iteration is a pure stream of launches with no host round trip.
// Inner products are reduced with warp shuffles -> one partial per block -> the last block (atomic
// ticket) adds the partials in a fixed order, so results are deterministic for a given grid.  With
// several z-slabs the local sums are combined by ncclAllReduce on the same stream.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "krylov_common.cuh"

namespace fdfd {

namespace {

using namespace kry;

// complex scalar slots (index into double2 array)
enum { S_RHO0 = 0, S_RR0 = 1, S_RHO1 = 2, S_RR1 = 3, S_SIGMA = 4, S_TS = 5, S_TT = 6, S_ALPHA = 7, S_OMEGA = 8,
       S_BNORM = 9, S_TMP0 = 10, S_TMP1 = 11, S_TMP2 = 12, S_SIG0 = 13, S_SIG1 = 14, S_SIG2 = 15 };

// r = b - r ; rhat = r ; p = r ;  RHO0 = RR0 = (r,r) ; BNORM = (b,b) ; cj = conj(rhat) when asked for
__global__ void __launch_bounds__(RB) k_init(int64_t n, const double2 *__restrict__ b, double2 *__restrict__ r,
                                             double2 *__restrict__ rhat, double2 *__restrict__ p,
                                             double2 *__restrict__ cj, Red rd) {
    double2 acc[2] = {c_zero(), c_zero()};
    double2 accb[1] = {c_zero()};
    GRID_STRIDE(i, n) {
        const double2 bb = b[i];
        const double2 rr = c_sub(bb, r[i]);
        r[i] = rr;
        rhat[i] = rr;
        p[i] = rr;
        if (cj) cj[i] = make_double2(rr.x, -rr.y);
        dot_acc(acc[0], rr, rr);
        dot_acc(accb[0], bb, bb);
    }
    acc[1] = acc[0];
    reduce_publish<2>(acc, rd, S_RHO0);
    reduce_publish<1>(accb, rd, S_BNORM);
}

// SIGMA = (rhat, v)
__global__ void __launch_bounds__(RB) k_dot1(int64_t n, const double2 *__restrict__ a, const double2 *__restrict__ b,
                                             Red rd, int slot) {
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) dot_acc(acc[0], a[i], b[i]);
    reduce_publish<1>(acc, rd, slot);
}

// sum_i u_i p_i (no conjugation) -> slot
__global__ void __launch_bounds__(RB) k_dotu1(int64_t n, const double2 *__restrict__ u, const double2 *__restrict__ p,
                                              Red rd, int slot) {
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) dotu_acc(acc[0], u[i], p[i]);
    reduce_publish<1>(acc, rd, slot);
}

// sigma = (rhat, v): slot SIGMA, or - when it was accumulated by the kernels that produced p - the sum of SIG0..2
__device__ __forceinline__ double2 load_sigma(const Red &rd, int sig3) {
    if (!sig3) return rd.scal[S_SIGMA];
    const double2 a = rd.scal[S_SIG0], b = rd.scal[S_SIG1], c = rd.scal[S_SIG2];
    return make_double2(a.x + b.x + c.x, a.y + b.y + c.y);
}

// Convergence far below any tolerance between two residual checks (tiny systems converge exactly within a few
// iterations; b in an invariant subspace) would drive rho, sigma, (t,t) to 0 / underflow and fill x with NaN before the
// host looks.  Once ||r|| <= 1e-30 ||b|| the iteration therefore idles: alpha = omega = beta = 0, x and r stay.  A
// genuine breakdown (sigma or rho = 0 with a residual above that) still shows as NaN / ENOCONV.
__device__ __forceinline__ bool idle(const Red &rd, int rr_slot) {
    return rd.scal[rr_slot].x <= 1e-60 * rd.scal[S_BNORM].x;
}

// alpha = rho/sigma ; s = r - alpha v
__global__ void __launch_bounds__(RB) k_s(int64_t n, const double2 *__restrict__ r, const double2 *__restrict__ v,
                                          double2 *__restrict__ s, Red rd, int rho_slot, int sig3) {
    const double2 alpha = idle(rd, rho_slot + 1) ? c_zero() : c_div(rd.scal[rho_slot], load_sigma(rd, sig3));
    GRID_STRIDE(i, n) s[i] = c_fms(alpha, v[i], r[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) rd.scal[S_ALPHA] = alpha;
}

// TS = (t,s), TT = (t,t)
__global__ void __launch_bounds__(RB) k_dot2(int64_t n, const double2 *__restrict__ t, const double2 *__restrict__ s,
                                             Red rd) {
    double2 acc[2] = {c_zero(), c_zero()};
    GRID_STRIDE(i, n) {
        const double2 tt = t[i];
        dot_acc(acc[0], tt, s[i]);
        dot_acc(acc[1], tt, tt);
    }
    reduce_publish<2>(acc, rd, S_TS);
}

// omega = TS/TT ; x += alpha p + omega s ; r = s - omega t ; RHO' = (rhat, r) ; RR' = (r,r)
__global__ void __launch_bounds__(RB) k_xr(int64_t n, double2 *__restrict__ x, const double2 *__restrict__ p,
                                           const double2 *__restrict__ s, const double2 *__restrict__ t,
                                           const double2 *__restrict__ rhat, double2 *__restrict__ r, Red rd,
                                           int rho_new_slot) {
    const double2 alpha = rd.scal[S_ALPHA];
    const int rho_old_slot = rho_new_slot == S_RHO0 ? S_RHO1 : S_RHO0;    // RR of the residual this iteration started from
    const double2 omega = idle(rd, rho_old_slot + 1) ? c_zero() : c_div(rd.scal[S_TS], rd.scal[S_TT]);
    double2 acc[2] = {c_zero(), c_zero()};
    GRID_STRIDE(i, n) {
        const double2 ss = s[i];
        double2 xx = c_fma(alpha, p[i], x[i]);
        xx = c_fma(omega, ss, xx);
        x[i] = xx;
        const double2 rr = c_fms(omega, t[i], ss);
        r[i] = rr;
        dot_acc(acc[0], rhat[i], rr);
        dot_acc(acc[1], rr, rr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) rd.scal[S_OMEGA] = omega;
    reduce_publish<2>(acc, rd, rho_new_slot);
}

// beta = (rho'/rho)(alpha/omega) ; p = r + beta (p - omega v)
// SIG: also the next iteration's sigma = (rhat, A p) = sum_i u_i p_i with u = A^T conj(rhat) fixed for the whole solve,
// accumulated while p is in registers -> sig_slot (the separate pass over rhat and v after the apply disappears)
template <bool SIG>
__global__ void __launch_bounds__(RB) k_p(int64_t n, const double2 *__restrict__ r, const double2 *__restrict__ v,
                                          double2 *__restrict__ p, const double2 *__restrict__ u, Red rd,
                                          int rho_old_slot, int rho_new_slot, int sig_slot) {
    const double2 omega = rd.scal[S_OMEGA];
    const double2 beta = idle(rd, rho_new_slot + 1) || idle(rd, rho_old_slot + 1)
                             ? c_zero()
                             : c_mul(c_div(rd.scal[rho_new_slot], rd.scal[rho_old_slot]), c_div(rd.scal[S_ALPHA], omega));
    double2 acc[1] = {c_zero()};
    GRID_STRIDE(i, n) {
        const double2 q = c_fms(omega, v[i], p[i]);
        const double2 pn = c_fma(beta, q, r[i]);
        p[i] = pn;
        if (SIG) dotu_acc(acc[0], u[i], pn);
    }
    if (SIG) reduce_publish<1>(acc, rd, sig_slot);
}

__global__ void k_store_hist(double *hist, int idx, const double2 *scal, int rr_slot) {
    hist[idx] = sqrt(scal[rr_slot].x / scal[S_BNORM].x);
}

}  // namespace

using kry::Red;
using kry::RB;
using kry::NSLOT;

static int bicgstab(Ctx *c, const double2 *b, double2 *x, double rtol, int maxit, int check_every, bool fixed_iters,
                    int *iters, double *relres, double *hist) {
    int rc = ensure_ready(c);
    if (rc != FDFD_OK) return rc;
    // sigma = (rhat, A p) is taken as (A^H rhat, p): u = A^T conj(rhat) is computed once (one transposed apply, one more
    // workspace vector) and the sum rides on the kernel that writes p, so an iteration has one vector pass (32 B/DOF
    // read, one launch, one reduction tail) less and the allreduce of sigma no longer sits behind the apply.
    // FDFD_BICGSTAB_CLASSIC restores the separate (rhat, v) pass.
    const bool sigf = getenv("FDFD_BICGSTAB_CLASSIC") == nullptr;
    if ((rc = kry::workspace(c, sigf ? 7 : 6)) != FDFD_OK) return rc;
    if ((rc = peer_direct_map(c)) != FDFD_OK) return rc;   // opt-in (FDFD_PEER_DIRECT): neighbours' workspaces, read in place
    if ((rc = ensure_dot_buffers(c)) != FDFD_OK) return rc;
    const int64_t n = c->nloc;
    double2 *r = c->work, *rhat = r + n, *p = rhat + n, *v = p + n, *s = v + n, *t = s + n, *u = sigf ? t + n : nullptr;
    Red rd = kry::make_red(c);
    double *sc = c->scal;
    const int g = kry::grid_for(n);
    cudaStream_t st = c->stream;
    double *hist_dev = nullptr;
    if (hist) FDFD_CUDA(c, cudaMalloc((void **)&hist_dev, sizeof(double) * (size_t)(maxit + 1)));
    auto cleanup = [&]() { if (hist_dev) cudaFree(hist_dev); };
#define KCHK(expr) do { int r__ = (expr); if (r__ != FDFD_OK) { cleanup(); return r__; } } while (0)
#define LCHK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) { cleanup(); return set_err(c, FDFD_ECUDA, cudaGetErrorString(e__)); } } while (0)

    // r = b - A x0
    KCHK(apply_device(c, x, r, false));
    k_init<<<g, RB, 0, st>>>(n, b, r, rhat, p, sigf ? v : nullptr, rd);
    LCHK();
    c->launches += 1;
    KCHK(allreduce_sum(c, sc + 2 * S_RHO0, 4, st));
    KCHK(allreduce_sum(c, sc + 2 * S_BNORM, 2, st));
    if (hist_dev) { k_store_hist<<<1, 1, 0, st>>>(hist_dev, 0, rd.scal, S_RR0); c->launches += 1; }

    auto read_relres = [&](int rr_slot, double &out) -> int {
        FDFD_CUDA(c, cudaMemcpyAsync(c->scal_host, sc, sizeof(double2) * NSLOT, cudaMemcpyDeviceToHost, st));
        FDFD_CUDA(c, cudaStreamSynchronize(st));
        const double rr = c->scal_host[2 * rr_slot], bn = c->scal_host[2 * S_BNORM];
        out = bn > 0 ? std::sqrt(rr / bn) : std::sqrt(rr);
        return FDFD_OK;
    };

    double rel = 1.0;
    int it = 0;
    bool converged = false;
    if (!fixed_iters) {
        KCHK(read_relres(S_RR0, rel));
        if (c->scal_host[2 * S_BNORM] == 0.0) {
            // b == 0: x = 0 is the solution (matches what a direct solve would return)
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
    if (sigf && !converged && maxit > 0) {
        // u = A^T conj(rhat) (conj(rhat) was left in v by k_init), sigma of the first iteration = sum u_i p_i
        FDFD_CUDA(c, cudaMemsetAsync(sc + 2 * S_SIG0, 0, 3 * sizeof(double2), st));
        KCHK(apply_device(c, v, u, true));
        k_dotu1<<<g, RB, 0, st>>>(n, u, p, rd, S_SIG0);
        LCHK();
        c->launches += 1;
        KCHK(allreduce_sum(c, sc + 2 * S_SIG0, 6, st));
    }
    // z-slabs: the kernels that produce s and p handle the two boundary planes first so that the NCCL halo exchange
    // of the next apply runs behind their interior part (halo_prefetch)
    const bool pre = halo_prefetch_usable(c);
    const int64_t pl = c->plane;
    const int gb = kry::grid_for(pl);
    // one iteration = a pure stream of launches (no host round trip), so it can be captured into a CUDA graph
    auto enqueue_iter = [&](int i) -> int {
        const int par = i & 1;
        const int rho_old = par ? S_RHO1 : S_RHO0, rho_new = par ? S_RHO0 : S_RHO1;
        int r1 = apply_device(c, p, v, false);
        if (r1 != FDFD_OK) return r1;
        if (!sigf) {
            k_dot1<<<g, RB, 0, st>>>(n, rhat, v, rd, S_SIGMA);
            if ((r1 = allreduce_sum(c, sc + 2 * S_SIGMA, 2, st)) != FDFD_OK) return r1;
            c->launches += 1;
        }
        const int sig3 = sigf ? 1 : 0;
        if (pre) {   // boundary planes first, start their halo exchange, then the interior behind which it hides
            k_s<<<gb, RB, 0, st>>>(pl, r, v, s, rd, rho_old, sig3);
            k_s<<<gb, RB, 0, st>>>(pl, r + n - pl, v + n - pl, s + n - pl, rd, rho_old, sig3);
            if ((r1 = halo_prefetch(c, s)) != FDFD_OK) return r1;
            k_s<<<g, RB, 0, st>>>(n - 2 * pl, r + pl, v + pl, s + pl, rd, rho_old, sig3);
            c->launches += 2;
        } else {
            k_s<<<g, RB, 0, st>>>(n, r, v, s, rd, rho_old, sig3);
        }
        // t = A s with (t,s) and (t,t) accumulated in the kernel epilogue when the tiled path is taken
        bool fused = false;
        if ((r1 = apply_device_dots(c, s, t, sc + 2 * S_TS, &fused)) != FDFD_OK) return r1;
        if (!fused) { k_dot2<<<g, RB, 0, st>>>(n, t, s, rd); c->launches += 1; }
        if ((r1 = allreduce_sum(c, sc + 2 * S_TS, 4, st)) != FDFD_OK) return r1;
        k_xr<<<g, RB, 0, st>>>(n, x, p, s, t, rhat, r, rd, rho_new);
        if ((r1 = allreduce_sum(c, sc + 2 * rho_new, 4, st)) != FDFD_OK) return r1;
        if (pre && sigf) {   // the three parts of p publish their share of the next sigma into SIG1, SIG2, SIG0
            k_p<true><<<gb, RB, 0, st>>>(pl, r, v, p, u, rd, rho_old, rho_new, S_SIG1);
            k_p<true><<<gb, RB, 0, st>>>(pl, r + n - pl, v + n - pl, p + n - pl, u + n - pl, rd, rho_old, rho_new, S_SIG2);
            if ((r1 = halo_prefetch(c, p)) != FDFD_OK) return r1;
            k_p<true><<<g, RB, 0, st>>>(n - 2 * pl, r + pl, v + pl, p + pl, u + pl, rd, rho_old, rho_new, S_SIG0);
            c->launches += 2;
        } else if (pre) {
            k_p<false><<<gb, RB, 0, st>>>(pl, r, v, p, nullptr, rd, rho_old, rho_new, 0);
            k_p<false><<<gb, RB, 0, st>>>(pl, r + n - pl, v + n - pl, p + n - pl, nullptr, rd, rho_old, rho_new, 0);
            if ((r1 = halo_prefetch(c, p)) != FDFD_OK) return r1;
            k_p<false><<<g, RB, 0, st>>>(n - 2 * pl, r + pl, v + pl, p + pl, nullptr, rd, rho_old, rho_new, 0);
            c->launches += 2;
        } else if (sigf) {
            k_p<true><<<g, RB, 0, st>>>(n, r, v, p, u, rd, rho_old, rho_new, S_SIG0);
        } else {
            k_p<false><<<g, RB, 0, st>>>(n, r, v, p, nullptr, rd, rho_old, rho_new, 0);
        }
        cudaError_t e1 = cudaGetLastError();
        if (e1 != cudaSuccess) return set_err(c, FDFD_ECUDA, cudaGetErrorString(e1));
        if (sigf && (r1 = allreduce_sum(c, sc + 2 * S_SIG0, 6, st)) != FDFD_OK) return r1;
        c->launches += 3;
        return FDFD_OK;
    };
    // Small grids are launch-bound (40^3: ~50 us of launches per iteration): replay two iterations (both scalar
    // parities) from a CUDA graph.  Single slab only (no NCCL in the graph), no per-iteration history.
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
        const int to_check = check_every - (it % check_every);   // iterations until the next residual read-back
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
        const int rr_new = ((it & 1) ? S_RHO1 : S_RHO0) + 1;   // RR slot written by the last finished iteration
        if (hist_dev) { k_store_hist<<<1, 1, 0, st>>>(hist_dev, it, rd.scal, rr_new); c->launches += 1; }
        if (!fixed_iters && (it % check_every == 0 || it >= maxit)) {
            int rq = read_relres(rr_new, rel);
            if (rq != FDFD_OK) { cleanup2(); cleanup(); return rq; }
            if (!(rel == rel)) break;  // NaN: breakdown
            converged = rel <= rtol;
        }
    }
    cleanup2();
    if (fixed_iters) {
        const int rr_last = (it & 1) ? S_RR1 : S_RR0;
        KCHK(read_relres(rr_last, rel));
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
    if (!converged) return set_err(c, FDFD_ENOCONV, "BiCGSTAB: not converged within maxit");
    return FDFD_OK;
}

int qmr(Ctx *c, const double2 *b, double2 *x, double rtol, int maxit, int check_every, bool fixed_iters, int *iters,
        double *relres, double *hist);

int krylov_solve(Ctx *c, int method, const double2 *b, double2 *x, double rtol, int maxit, int check_every,
                 bool fixed_iters, int *iters, double *relres, double *hist) {
    if (method == FDFD_BICGSTAB) return bicgstab(c, b, x, rtol, maxit, check_every, fixed_iters, iters, relres, hist);
    if (method == FDFD_QMR) return qmr(c, b, x, rtol, maxit, check_every, fixed_iters, iters, relres, hist);
    return set_err(c, FDFD_EINVAL, "unknown Krylov method");
}

}  // namespace fdfd
