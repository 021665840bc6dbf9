// fdfd_multi_*: ONE call, N GPUs ("the caller sees a single call").
//
// The reference's seam is one value from one process - `A = create_A(...)` followed by `A * x` / `A \ b`
// (src/model/model.jl:209-246).  The slab handles of this library are one-per-GPU, meant for one process per GPU; a host
// language that is a single process (a Julia session) would need an MPI launcher to use more than one.  A multi handle
// hides that: it owns one slab handle per device, one host thread per handle (the NCCL communicator is built across the
// threads), takes FULL-GRID host arrays and splits them into z-slabs itself.
//
//   fdfd_multi_apply   each thread runs the sub-slab pipeline of the host-buffer apply on its slab (H2D, kernel and D2H
//                      overlapped); the neighbour planes it needs are read from the caller's host vector together with
//                      the slab, so the apply involves no exchange between GPUs and every GPU uses its own PCIe link;
//   fdfd_multi_solve   each thread runs the slab Krylov loop (halos over NCCL send/recv, dots over NCCL allreduce),
//                      b copied in and x copied out slab by slab.
//
// Built only on the public C ABI of the slab handles (include/fdfd_b200.h) plus threads: nothing here touches CUDA.
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fdfd_b200.h"

namespace {

struct Multi {
    int n = 0;
    fdfd_desc d{};
    std::vector<int> dev;
    std::vector<fdfd_handle> h;
    std::vector<int64_t> k0, k1;
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::function<int(int)> job;
    uint64_t gen = 0;
    int pending = 0;
    bool stop = false;
    std::vector<int> rc;
    std::string err;
    int64_t plane = 0;   // DOFs per z-plane

    void worker(int r) {
        uint64_t seen = 0;
        for (;;) {
            std::function<int(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_job.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                f = job;
            }
            int v = FDFD_EINVAL;
            try { v = f(r); } catch (...) { v = FDFD_ENOMEM; }
            {
                std::lock_guard<std::mutex> lk(mu);
                rc[r] = v;
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }

    // run f(rank) on every worker thread; returns the first non-zero status (FDFD_ENOCONV only if no rank failed harder)
    int run(const std::function<int(int)> &f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            job = f;
            pending = n;
            ++gen;
        }
        cv_job.notify_all();
        {
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return pending == 0; });
        }
        int out = FDFD_OK;
        for (int r = 0; r < n; ++r) {
            if (rc[r] == FDFD_OK) continue;
            if (out == FDFD_OK || out == FDFD_ENOCONV) {
                out = rc[r];
                const char *m = h[r] ? fdfd_last_error(h[r]) : fdfd_last_error(nullptr);
                err = "slab " + std::to_string(r) + " (device " + std::to_string(dev[r]) + "): " + (m ? m : "");
            }
        }
        return out;
    }
};

thread_local std::string g_err;

int fail(Multi *m, int code, const std::string &msg) {
    if (m) m->err = msg;
    else g_err = msg;
    return code;
}

// slab <-> full-grid copies of a DOF vector (cmp-first layout: one contiguous piece; component-major: three)
void gather_slab(const Multi *m, int r, const fdfd_c128 *full, fdfd_c128 *slab) {
    const int64_t nxy = m->d.N[0] * m->d.N[1], nzl = m->k1[r] - m->k0[r], Nz = m->d.N[2];
    if (m->d.order_cmpfirst) {
        std::memcpy(slab, full + 3 * nxy * m->k0[r], (size_t)(3 * nxy * nzl) * sizeof(fdfd_c128));
    } else {
        for (int c = 0; c < 3; ++c)
            std::memcpy(slab + nxy * nzl * c, full + nxy * (m->k0[r] + Nz * c), (size_t)(nxy * nzl) * sizeof(fdfd_c128));
    }
}
void scatter_slab(const Multi *m, int r, const fdfd_c128 *slab, fdfd_c128 *full) {
    const int64_t nxy = m->d.N[0] * m->d.N[1], nzl = m->k1[r] - m->k0[r], Nz = m->d.N[2];
    if (m->d.order_cmpfirst) {
        std::memcpy(full + 3 * nxy * m->k0[r], slab, (size_t)(3 * nxy * nzl) * sizeof(fdfd_c128));
    } else {
        for (int c = 0; c < 3; ++c)
            std::memcpy(full + nxy * (m->k0[r] + Nz * c), slab + nxy * nzl * c, (size_t)(nxy * nzl) * sizeof(fdfd_c128));
    }
}
// one plane of the full-grid vector in the halo layout (three components after one another)
void gather_plane(const Multi *m, int64_t k, const fdfd_c128 *full, fdfd_c128 *out) {
    const int64_t nxy = m->d.N[0] * m->d.N[1], Nz = m->d.N[2];
    if (m->d.order_cmpfirst) std::memcpy(out, full + 3 * nxy * k, (size_t)(3 * nxy) * sizeof(fdfd_c128));
    else
        for (int c = 0; c < 3; ++c) std::memcpy(out + nxy * c, full + nxy * (k + Nz * c), (size_t)nxy * sizeof(fdfd_c128));
}

// this slab's part of a Julia-layout material array (Nx,Ny,Nz,3,3) -> (Nx,Ny,nzl,3,3)
void gather_material(const Multi *m, int r, const fdfd_c128 *full, std::vector<fdfd_c128> &slab) {
    const int64_t nxy = m->d.N[0] * m->d.N[1], nzl = m->k1[r] - m->k0[r], Nz = m->d.N[2];
    slab.resize((size_t)(9 * nxy * nzl));
    for (int e = 0; e < 9; ++e)
        std::memcpy(slab.data() + (size_t)(nxy * nzl) * e, full + nxy * (Nz * e + m->k0[r]), (size_t)(nxy * nzl) * sizeof(fdfd_c128));
}

// a vector operation slab by slab through the slab handles' host-buffer entry points
template <class F>
int per_slab_vectors(Multi *m, std::initializer_list<const fdfd_c128 *> ins, fdfd_c128 *out, F call) {
    const bool direct = m->d.order_cmpfirst != 0;   // slabs are contiguous pieces of the caller's arrays
    std::vector<const fdfd_c128 *> inv(ins);
    return m->run([&, direct](int r) -> int {
        const int64_t nloc = m->plane * (m->k1[r] - m->k0[r]);
        std::vector<std::vector<fdfd_c128>> tmp;
        std::vector<const fdfd_c128 *> a(inv.size(), nullptr);
        std::vector<fdfd_c128> o;
        for (size_t i = 0; i < inv.size(); ++i) {
            if (!inv[i]) continue;
            if (direct) a[i] = inv[i] + m->plane * m->k0[r];
            else {
                tmp.emplace_back((size_t)nloc);
                gather_slab(m, r, inv[i], tmp.back().data());
                a[i] = tmp.back().data();
            }
        }
        fdfd_c128 *op = direct ? out + m->plane * m->k0[r] : (o.resize((size_t)nloc), o.data());
        const int rc = call(r, m->h[r], a, op);
        if ((rc == FDFD_OK || rc == FDFD_ENOCONV) && !direct) scatter_slab(m, r, op, out);
        return rc;
    });
}

}  // namespace

extern "C" {

int fdfd_multi_create(fdfd_multi *out, const fdfd_desc *desc, int32_t ngpu, const int32_t *devices_or_null) {
    if (!out || !desc) return fail(nullptr, FDFD_EINVAL, "fdfd_multi_create: null argument");
    *out = nullptr;
    if (ngpu < 1 || ngpu > 64) return fail(nullptr, FDFD_EINVAL, "fdfd_multi_create: ngpu must be 1..64");
    if (desc->N[2] < ngpu) return fail(nullptr, FDFD_EINVAL, "fdfd_multi_create: fewer z-planes than GPUs");
    Multi *m = new (std::nothrow) Multi();
    if (!m) return fail(nullptr, FDFD_ENOMEM, "out of host memory");
    m->n = ngpu;
    m->d = *desc;
    m->plane = 3 * desc->N[0] * desc->N[1];
    m->dev.resize(ngpu);
    m->h.assign(ngpu, nullptr);
    m->k0.assign(ngpu, 0);
    m->k1.assign(ngpu, 0);
    m->rc.assign(ngpu, 0);
    for (int r = 0; r < ngpu; ++r) {
        m->dev[r] = devices_or_null ? devices_or_null[r] : r;
        fdfd_partition(desc->N[2], ngpu, r, &m->k0[r], &m->k1[r]);
    }
    char uid[128];
    std::memset(uid, 0, sizeof(uid));
    if (ngpu > 1) {
        const int ru = fdfd_comm_unique_id(uid);
        if (ru != FDFD_OK) {
            const char *e = fdfd_last_error(nullptr);
            std::string msg = std::string("fdfd_multi_create: ") + (e ? e : "NCCL unique id");
            delete m;
            return fail(nullptr, ru, msg);
        }
    }
    for (int r = 0; r < ngpu; ++r) m->workers.emplace_back([m, r] { m->worker(r); });
    std::vector<std::string> cerr_(ngpu);
    int rc = m->run([&](int r) -> int {
        fdfd_desc d = m->d;
        d.device = m->dev[r];
        d.rank = r;
        d.nranks = m->n;
        const int v = fdfd_create(&m->h[r], &d);
        if (v != FDFD_OK) { const char *e = fdfd_last_error(nullptr); cerr_[r] = e ? e : ""; }
        else if (m->n > 1) fdfd_set_shared_process(m->h[r], 1);   // the slab threads take turns in the Krylov loops
        return v;
    });
    if (rc != FDFD_OK) {
        for (int r = 0; r < ngpu; ++r)
            if (!m->h[r] && !cerr_[r].empty()) m->err = "slab " + std::to_string(r) + ": " + cerr_[r];
    }
    // the communicator is built by all ranks together (ncclCommInitRank blocks until every rank has called it)
    if (rc == FDFD_OK && ngpu > 1) rc = m->run([&](int r) -> int { return fdfd_comm_init(m->h[r], uid); });
    if (rc != FDFD_OK) {
        const std::string msg = m->err;
        fdfd_multi_destroy(reinterpret_cast<fdfd_multi>(m));
        return fail(nullptr, rc, "fdfd_multi_create: " + msg);
    }
    *out = reinterpret_cast<fdfd_multi>(m);
    return FDFD_OK;
}

int fdfd_multi_destroy(fdfd_multi mh) {
    Multi *m = reinterpret_cast<Multi *>(mh);
    if (!m) return FDFD_OK;
    if (!m->workers.empty()) {
        m->run([&](int r) -> int {
            if (m->h[r]) fdfd_destroy(m->h[r]);
            m->h[r] = nullptr;
            return FDFD_OK;
        });
        {
            std::lock_guard<std::mutex> lk(m->mu);
            m->stop = true;
        }
        m->cv_job.notify_all();
        for (auto &t : m->workers) t.join();
    }
    delete m;
    return FDFD_OK;
}

const char *fdfd_multi_last_error(fdfd_multi mh) {
    Multi *m = reinterpret_cast<Multi *>(mh);
    return m ? m->err.c_str() : g_err.c_str();
}

int fdfd_multi_ngpu(fdfd_multi mh) { return mh ? reinterpret_cast<Multi *>(mh)->n : 0; }

int fdfd_multi_slab(fdfd_multi mh, int32_t slab, fdfd_handle *h, int64_t *k0, int64_t *k1) {
    Multi *m = reinterpret_cast<Multi *>(mh);
    if (!m) return fail(nullptr, FDFD_EINVAL, "null handle");
    if (slab < 0 || slab >= m->n) return fail(m, FDFD_EINVAL, "fdfd_multi_slab: slab index out of range");
    if (h) *h = m->h[slab];
    if (k0) *k0 = m->k0[slab];
    if (k1) *k1 = m->k1[slab];
    return FDFD_OK;
}

#define MULTI(mh)                                                     \
    Multi *m = reinterpret_cast<Multi *>(mh);                         \
    if (!m) return fail(nullptr, FDFD_EINVAL, "null handle")

int fdfd_multi_set_coeffs(fdfd_multi mh, const fdfd_c128 *const sdl_e[3], const fdfd_c128 *const sdl_m[3]) {
    MULTI(mh);
    return m->run([&](int r) { return fdfd_set_coeffs(m->h[r], sdl_e, sdl_m); });
}
int fdfd_multi_set_bloch(fdfd_multi mh, const fdfd_c128 e_mikL[3]) {
    MULTI(mh);
    return m->run([&](int r) { return fdfd_set_bloch(m->h[r], e_mikL); });
}
int fdfd_multi_set_omega(fdfd_multi mh, fdfd_c128 omega) {
    MULTI(mh);
    return m->run([&](int r) { return fdfd_set_omega(m->h[r], omega); });
}
int fdfd_multi_set_eps(fdfd_multi mh, const fdfd_c128 *eps, int has_offdiag) {
    MULTI(mh);
    if (!eps) return fail(m, FDFD_EINVAL, "fdfd_multi_set_eps: null argument");
    return m->run([&](int r) {
        std::vector<fdfd_c128> slab;
        gather_material(m, r, eps, slab);
        return fdfd_set_eps(m->h[r], slab.data(), has_offdiag);
    });
}
int fdfd_multi_set_mu(fdfd_multi mh, const fdfd_c128 *mu_or_null) {
    MULTI(mh);
    return m->run([&](int r) {
        if (!mu_or_null) return fdfd_set_mu(m->h[r], nullptr);
        std::vector<fdfd_c128> slab;
        gather_material(m, r, mu_or_null, slab);
        return fdfd_set_mu(m->h[r], slab.data());
    });
}
int fdfd_multi_set_eps_objects(fdfd_multi mh, const fdfd_matparams_desc *desc) {
    MULTI(mh);
    return m->run([&](int r) { return fdfd_set_eps_objects(m->h[r], desc); });
}

static int multi_apply(Multi *m, const fdfd_c128 *x, fdfd_c128 *y, int transpose) {
    if (!x || !y) return fail(m, FDFD_EINVAL, "fdfd_multi_apply: null argument");
    if (x == y) return fail(m, FDFD_EINVAL, "fdfd_multi_apply: x and y must not alias");
    const int64_t Nz = m->d.N[2];
    const bool wrapz = m->d.isbloch[2] != 0, direct = m->d.order_cmpfirst != 0;
    return m->run([&, Nz, wrapz, direct](int r) -> int {
        const int64_t k0 = m->k0[r], k1 = m->k1[r];
        const int64_t kb = k0 > 0 ? k0 - 1 : (wrapz ? Nz - 1 : -1), ka = k1 < Nz ? k1 : (wrapz ? 0 : -1);
        if (direct) {
            return fdfd_apply_host_halos(m->h[r], x + m->plane * k0, kb >= 0 ? x + m->plane * kb : nullptr,
                                         ka >= 0 ? x + m->plane * ka : nullptr, y + m->plane * k0, transpose);
        }
        std::vector<fdfd_c128> xs((size_t)(m->plane * (k1 - k0))), ys(xs.size()), lo, hi;
        gather_slab(m, r, x, xs.data());
        if (kb >= 0) { lo.resize((size_t)m->plane); gather_plane(m, kb, x, lo.data()); }
        if (ka >= 0) { hi.resize((size_t)m->plane); gather_plane(m, ka, x, hi.data()); }
        const int rc = fdfd_apply_host_halos(m->h[r], xs.data(), kb >= 0 ? lo.data() : nullptr, ka >= 0 ? hi.data() : nullptr,
                                             ys.data(), transpose);
        if (rc == FDFD_OK) scatter_slab(m, r, ys.data(), y);
        return rc;
    });
}
int fdfd_multi_apply(fdfd_multi mh, const fdfd_c128 *x, fdfd_c128 *y) {
    MULTI(mh);
    return multi_apply(m, x, y, 0);
}
int fdfd_multi_apply_transpose(fdfd_multi mh, const fdfd_c128 *x, fdfd_c128 *y) {
    MULTI(mh);
    return multi_apply(m, x, y, 1);
}

int fdfd_multi_solve(fdfd_multi mh, int method, const fdfd_c128 *b, fdfd_c128 *x, double rtol, int maxit, int check_every,
                     int *iters, double *relres, double *hist_or_null) {
    MULTI(mh);
    if (!b || !x) return fail(m, FDFD_EINVAL, "fdfd_multi_solve: null argument");
    std::vector<int> it(m->n, 0);
    std::vector<double> rr(m->n, 0.0);
    const int rc = per_slab_vectors(m, {b, x}, x, [&](int r, fdfd_handle h, const std::vector<const fdfd_c128 *> &a, fdfd_c128 *xo) -> int {
        // x is both the initial guess and the result: the slab solver works in place on xo (cmp-first layout: xo IS the
        // slab of the caller's x; component-major: a packed copy that starts from the packed guess)
        if (xo != a[1]) std::memcpy(xo, a[1], (size_t)(m->plane * (m->k1[r] - m->k0[r])) * sizeof(fdfd_c128));
        // every rank returns the same residual history (allreduced norms); rank 0 writes it
        return fdfd_solve(h, method, a[0], xo, FDFD_HOST, rtol, maxit, check_every, &it[r], &rr[r], r == 0 ? hist_or_null : nullptr);
    });
    if (iters) *iters = it[0];
    if (relres) *relres = rr[0];
    return rc;
}

int fdfd_multi_create_b(fdfd_multi mh, const fdfd_c128 *je, const fdfd_c128 *jm_or_null, fdfd_c128 *b) {
    MULTI(mh);
    if (!je || !b) return fail(m, FDFD_EINVAL, "fdfd_multi_create_b: null argument");
    return per_slab_vectors(m, {je, jm_or_null}, b, [](int, fdfd_handle h, const std::vector<const fdfd_c128 *> &a, fdfd_c128 *o) {
        return fdfd_create_b(h, a[0], a[1], o, FDFD_HOST);
    });
}
int fdfd_multi_h_from_e(fdfd_multi mh, const fdfd_c128 *e, const fdfd_c128 *jm_or_null, fdfd_c128 *hout) {
    MULTI(mh);
    if (!e || !hout) return fail(m, FDFD_EINVAL, "fdfd_multi_h_from_e: null argument");
    return per_slab_vectors(m, {e, jm_or_null}, hout, [](int, fdfd_handle h, const std::vector<const fdfd_c128 *> &a, fdfd_c128 *o) {
        return fdfd_h_from_e(h, a[0], a[1], o, FDFD_HOST);
    });
}
int fdfd_multi_e_from_h(fdfd_multi mh, const fdfd_c128 *hfield, const fdfd_c128 *je_or_null, fdfd_c128 *eout) {
    MULTI(mh);
    if (!hfield || !eout) return fail(m, FDFD_EINVAL, "fdfd_multi_e_from_h: null argument");
    return per_slab_vectors(m, {hfield, je_or_null}, eout, [](int, fdfd_handle h, const std::vector<const fdfd_c128 *> &a, fdfd_c128 *o) {
        return fdfd_e_from_h(h, a[0], a[1], o, FDFD_HOST);
    });
}

}  // extern "C"
