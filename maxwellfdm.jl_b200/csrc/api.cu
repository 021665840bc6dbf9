// C ABI of libfdfd_b200.so (see include/fdfd_b200.h for the contract and the reference lines each entry
// point stands in for).  Host-side state handling, device array construction, launches.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <new>

#include "cplx.cuh"
#include "fdfd_internal.h"

namespace fdfd {

static thread_local std::string g_create_err;

namespace {
std::mutex g_turn_mu;
std::condition_variable g_turn_cv;
uint64_t g_turn_next = 0, g_turn_serving = 0;
}  // namespace
BurstTurn::BurstTurn(const Ctx *c) : held(c && c->shared_process) {
    static const bool off = getenv("FDFD_NO_TURNS") != nullptr;   // A/B timing: let the slab threads contend
    if (off) held = false;
    if (!held) return;
    std::unique_lock<std::mutex> lk(g_turn_mu);
    const uint64_t mine = g_turn_next++;
    g_turn_cv.wait(lk, [&] { return g_turn_serving == mine; });
}
BurstTurn::~BurstTurn() {
    if (!held) return;
    {
        std::lock_guard<std::mutex> lk(g_turn_mu);
        ++g_turn_serving;
    }
    g_turn_cv.notify_all();
}

int set_err(Ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    else g_create_err = msg;
    return code;
}

// ---------------------------------------------------------------------------------------------------
// small device kernels used while building the material arrays
// ---------------------------------------------------------------------------------------------------
__global__ void scale_copy_kernel(double2 *dst, const double2 *src, int64_t n, double2 f) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = c_mul(f, src[i]);
}
__global__ void recip_copy_kernel(double2 *dst, const double2 *src, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 a = src[i];
        const double d = a.x * a.x + a.y * a.y;
        dst[i] = make_double2(a.x / d, -a.y / d);
    }
}
__global__ void fill_kernel(double2 *dst, int64_t n, double2 v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = v;
}
// dst[i*3 + c] = src_c[i]: the diagonal mass entries interleaved like a cmp-first DOF vector, so that the row-pair
// kernel's TMA row copies fetch them with the geometry of the x rows
__global__ void interleave3_kernel(double2 *dst, const double2 *s0, const double2 *s1, const double2 *s2, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        dst[3 * i] = s0[i];
        dst[3 * i + 1] = s1[i];
        dst[3 * i + 2] = s2[i];
    }
}
// real-valued variant (dst doubles); *any_imag is raised when an entry has a non-zero imaginary part
__global__ void interleave3_real_kernel(double *dst, const double2 *s0, const double2 *s1, const double2 *s2, int64_t n,
                                        int Nx, int64_t pitch, int *any_imag) {
    bool bad = false;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 a = s0[i], b = s1[i], c = s2[i];
        double *d = dst + (i / Nx) * pitch + (i % Nx) * 3;     // rows padded to a whole number of 16-byte units
        d[0] = a.x;
        d[1] = b.x;
        d[2] = c.x;
        bad |= (a.y != 0.0) | (b.y != 0.0) | (c.y != 0.0);
    }
    if (bad) *any_imag = 1;
}
// the three off-diagonal arrays of a symmetric tensor, real parts, interleaved and padded by one cell / row on either side
// in x and y (wrapped copies on Bloch axes, zeros otherwise): the fused row-pair kernel's tensor-map boxes then deliver
// the forward neighbours of edge tiles without any patching.  n = padded elements; *any_imag as above.
__global__ void interleave3_real_pad_kernel(double *dst, const double2 *s0, const double2 *s1, const double2 *s2, int64_t n,
                                            int Nx, int Ny, int wrapx, int wrapy, int64_t pitch, int *any_imag) {
    bool bad = false;
    const int Px = Nx + 2, Py = Ny + 2;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int ip = (int)(t % Px);
        const int64_t row = t / Px;                  // g * Py + jp
        const int jp = (int)(row % Py);
        const int64_t g = row / Py;
        int i = ip - 1, j = jp - 1;
        bool in = true;
        if (i < 0 || i >= Nx) { in = in && wrapx; i = (i + Nx) % Nx; }
        if (j < 0 || j >= Ny) { in = in && wrapy; j = (j + Ny) % Ny; }
        double a = 0.0, b = 0.0, c = 0.0;
        if (in) {
            const int64_t src = (g * Ny + j) * Nx + i;
            const double2 va = s0[src], vb = s1[src], vc = s2[src];
            a = va.x; b = vb.x; c = vc.x;
            bad |= (va.y != 0.0) | (vb.y != 0.0) | (vc.y != 0.0);
        }
        double *d = dst + row * pitch + (int64_t)ip * 3;
        d[0] = a; d[1] = b; d[2] = c;
    }
    if (bad) *any_imag = 1;
}
__global__ void flush_kernel(float4 *p, int64_t n, float v) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}

static inline int nblocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    return (int)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b));
}

static void free_device(Ctx *c) {
    auto F = [](auto *&p) { if (p) cudaFree((void *)p); p = nullptr; };
    F(c->coef_dev); F(c->mat_dev); F(c->md_aos); F(c->md_aos_r); F(c->mo_aos_r); F(c->halo_lo); F(c->halo_hi); F(c->work); F(c->scal); F(c->partial);
    F(c->stage_x); F(c->stage_y); F(c->flush_buf); F(c->offmask); F(c->corr_list); F(c->dot_partial); F(c->dot_ticket);
    F(c->halo_flag);
    if (c->scal_host) cudaFreeHost(c->scal_host);
    c->scal_host = nullptr;
}

// ---------------------------------------------------------------------------------------------------
// device arrays
// ---------------------------------------------------------------------------------------------------
static int upload_coefs(Ctx *c) {
    CoefHost f, t;
    build_coefs(c->d, c->sdl_e, c->sdl_m, c->phase, f);
    transpose_coefs(f, t);
    for (int w = 0; w < 3; ++w) c->s1[w] = f.a[w].shift;
    const int64_t Ns = c->d.N[0] + c->d.N[1] + c->d.N[2];
    const size_t bytes = (size_t)(2 * 10 * Ns) * sizeof(double2);
    std::vector<cplx> host((size_t)2 * 10 * Ns);
    if (c->coef_bytes != bytes) {
        if (c->coef_dev) cudaFree(c->coef_dev);
        c->coef_dev = nullptr;
        FDFD_CUDA(c, cudaMalloc((void **)&c->coef_dev, bytes));
        c->coef_bytes = bytes;
    }
    size_t off = 0;
    auto put = [&](const std::vector<cplx> &v, const double2 *&dst) {
        std::memcpy(&host[off], v.data(), v.size() * sizeof(cplx));
        dst = c->coef_dev + off;
        off += v.size();
    };
    for (int set = 0; set < 2; ++set) {
        const CoefHost &h = set == 0 ? f : t;
        CoefDev &d = set == 0 ? c->cf : c->ct;
        for (int w = 0; w < 3; ++w) {
            put(h.a[w].t0, d.a0[w]); put(h.a[w].t1, d.a1[w]);
            put(h.b[w].t0, d.b0[w]); put(h.b[w].t1, d.b1[w]);
            put(h.mi[w].t0, d.mi0[w]); put(h.mi[w].t1, d.mi1[w]);
            put(h.mo[w].t0, d.mo0[w]); put(h.mo[w].t1, d.mo1[w]);
            put(h.mh[w].t0, d.mh0[w]); put(h.mh[w].t1, d.mh1[w]);
        }
    }
    FDFD_CUDA(c, cudaMemcpyAsync(c->coef_dev, host.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

// Build md (3), mo (6, optional), q (3, optional) as z-ghosted component-major arrays on the device.
static int upload_materials(Ctx *c) {
    const int64_t Nxy = c->d.N[0] * c->d.N[1];
    const int64_t nzl = c->k1 - c->k0;
    const int64_t M = Nxy * nzl, Mg = Nxy * (nzl + 2);
    const bool ee = c->d.field_type == FDFD_FT_EE;
    // mass parameter / middle parameter in reference terms
    const std::vector<cplx> &mass = ee ? c->eps_host : c->mu_host;
    const std::vector<cplx> &mid = ee ? c->mu_host : c->eps_host;
    const bool obj = ee && c->eps_obj.set;      // eps from objects: rasterised on the device below, no host array
    const bool mass_given = !mass.empty() || obj, mid_given = !mid.empty();
    const bool has_mass = c->omega != cplx(0.0);
    bool has_off = has_mass && mass_given && (ee ? c->eps_off : c->mu_off);
    // A pointwise symmetric mass tensor (P_vu == P_uv exactly, the usual outcome of subpixel smoothing of reciprocal
    // media) is stored as three off-diagonal arrays; the other three slots alias them, so every kernel reads the same
    // values through the same code while the off-diagonal streams cost 16 instead of 32 B/DOF of HBM traffic.
    bool off_sym = has_off;
    if (has_off && obj) {
        off_sym = c->eps_obj.symmetric;         // the kernel writes the upper triangle to both places
    } else if (has_off) {
        int nonsym = 0;
#pragma omp parallel for reduction(| : nonsym) schedule(static)
        for (int64_t i = 0; i < M; ++i)
            for (int v = 0; v < 3; ++v)
                for (int u = v + 1; u < 3; ++u)
                    nonsym |= mass[(size_t)M * (v + 3 * u) + i] != mass[(size_t)M * (u + 3 * v) + i];
        off_sym = nonsym == 0;
    }
    if (c->d.nranks > 1 && has_mass && mass_given) {
        // z-slabs: every rank must build (and exchange the ghost planes of) the same set of arrays - a slab without
        // any off-diagonal entry next to one that has them (the C4 sphere) builds zero-filled arrays, and the tensor
        // counts as symmetric only if it is on every slab
        if (!c->comm) return set_err(c, FDFD_ESTATE, "nranks > 1: call fdfd_comm_init before the first apply");
        double *flags = nullptr;
        const double h2[2] = {has_off ? 1.0 : 0.0, (has_off && !off_sym) ? 1.0 : 0.0};
        FDFD_CUDA(c, cudaMalloc((void **)&flags, 2 * sizeof(double)));
        FDFD_CUDA(c, cudaMemcpyAsync(flags, h2, sizeof(h2), cudaMemcpyHostToDevice, c->stream));
        int ra = allreduce_sum(c, flags, 2, c->stream);
        double g2[2] = {0.0, 0.0};
        if (ra == FDFD_OK) {
            FDFD_CUDA(c, cudaMemcpyAsync(g2, flags, sizeof(g2), cudaMemcpyDeviceToHost, c->stream));
            FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
        }
        cudaFree(flags);
        if (ra != FDFD_OK) return ra;
        has_off = g2[0] > 0.0;
        off_sym = has_off && g2[1] == 0.0;
    }
    c->off_sym = off_sym;
    // an identity mass parameter (mu == 1 of the HH formulation) is a scalar, not three arrays
    const int narr = (has_mass && mass_given ? 3 : 0) + (has_off ? (off_sym ? 3 : 6) : 0) + (mid_given ? 3 : 0);
    const size_t bytes = (size_t)narr * Mg * sizeof(double2);
    if (bytes != c->mat_bytes) {
        if (c->mat_dev) cudaFree(c->mat_dev);
        c->mat_dev = nullptr;
        if (bytes) FDFD_CUDA(c, cudaMalloc((void **)&c->mat_dev, bytes));
        c->mat_bytes = bytes;
    }
    for (int i = 0; i < 3; ++i) c->md[i] = c->q[i] = nullptr;
    for (int i = 0; i < 6; ++i) c->mo[i] = c->mo_t[i] = nullptr;
    c->has_mass = has_mass;
    {
        const cplx w2u = -(c->omega * c->omega);
        c->md_uniform = make_double2(w2u.real(), w2u.imag());
    }
    if (c->md_aos) { cudaFree(c->md_aos); c->md_aos = nullptr; }
    if (c->md_aos_r) { cudaFree(c->md_aos_r); c->md_aos_r = nullptr; }
    if (c->mo_aos_r) { cudaFree(c->mo_aos_r); c->mo_aos_r = nullptr; }
    if (!narr) return FDFD_OK;
    double2 *objbuf = nullptr;                  // objects: the smoothed slab in Julia layout, on the device
    if (obj && has_mass) {
        fdfd_matparams_desc md{};
        for (int w = 0; w < 3; ++w) {
            md.N[w] = c->d.N[w];
            md.isbloch[w] = c->d.isbloch[w];
            md.boundft_is_E[w] = c->d.boundft_is_E[w];
            md.lprim[w] = c->eps_obj.lprim[w].data();
        }
        md.field_type = FDFD_FT_EE;
        md.field_ortho_shape = c->eps_obj.ortho;
        md.k0 = c->k0;
        md.k1 = c->k1;
        md.nshape = (int32_t)c->eps_obj.shapes.size();
        md.nparam = (int32_t)(c->eps_obj.params.size() / 9);
        md.shapes = c->eps_obj.shapes.data();
        md.params = reinterpret_cast<const fdfd_c128 *>(c->eps_obj.params.data());
        md.device = -1;
        FDFD_CUDA(c, cudaMalloc((void **)&objbuf, (size_t)9 * M * sizeof(double2)));
        std::string merr;
        const int rm = calc_matparams(&md, reinterpret_cast<fdfd_c128 *>(objbuf), FDFD_DEVICE, merr);
        c->launches += 1;
        if (rm != FDFD_OK) { cudaFree(objbuf); return set_err(c, rm, "fdfd_set_eps_objects: " + merr); }
    }
    double2 *tmp = nullptr;
    FDFD_CUDA(c, cudaMalloc((void **)&tmp, (size_t)M * sizeof(double2)));
    double2 *cur = c->mat_dev;
    const cplx w2 = -(c->omega * c->omega);
    const double2 f = make_double2(w2.real(), w2.imag());
    std::vector<double2 *> ghosted;
    auto build = [&](const std::vector<cplx> *src, int v, int u, int mode, const double2 *&slot) -> int {
        // mode 0: f * src, mode 1: 1/src; src == nullptr: identity parameter (1 on the diagonal)
        if (src) {
            const double2 *from = tmp;
            if (objbuf && src == &mass) {
                from = objbuf + (size_t)M * (v + 3 * u);      // already on the device
            } else {
                FDFD_CUDA(c, cudaMemcpyAsync(tmp, src->data() + (size_t)M * (v + 3 * u), (size_t)M * sizeof(double2),
                                             cudaMemcpyHostToDevice, c->stream));
            }
            if (mode == 0) scale_copy_kernel<<<nblocks(M), 256, 0, c->stream>>>(cur + Nxy, from, M, f);
            else           recip_copy_kernel<<<nblocks(M), 256, 0, c->stream>>>(cur + Nxy, from, M);
        } else {
            fill_kernel<<<nblocks(M), 256, 0, c->stream>>>(cur + Nxy, M, mode == 0 ? f : make_double2(1.0, 0.0));
        }
        FDFD_CUDA(c, cudaGetLastError());
        FDFD_CUDA(c, cudaStreamSynchronize(c->stream));  // tmp is reused
        slot = cur;
        ghosted.push_back(cur);
        cur += Mg;
        return FDFD_OK;
    };
    int rc = FDFD_OK;
    if (has_mass && mass_given)
        for (int v = 0; v < 3 && rc == FDFD_OK; ++v) rc = build(&mass, v, v, 0, c->md[v]);
    if (has_off && rc == FDFD_OK) {
        int idx[3][3], e = 0;
        for (int v = 0; v < 3; ++v)
            for (int u = 0; u < 3; ++u)
                if (u != v) idx[v][u] = e++;
        for (int v = 0; v < 3; ++v)
            for (int u = 0; u < 3; ++u) {
                if (u == v || rc != FDFD_OK) continue;
                if (off_sym && v > u) c->mo[idx[v][u]] = c->mo[idx[u][v]];   // (u,v) with u < v was built earlier
                else rc = build(&mass, v, u, 0, c->mo[idx[v][u]]);
            }
        // transposed operator uses P'_{uv} = P_{vu}
        for (int v = 0; v < 3; ++v)
            for (int u = 0; u < 3; ++u)
                if (u != v) c->mo_t[idx[v][u]] = c->mo[idx[u][v]];
    }
    if (mid_given && rc == FDFD_OK)
        for (int v = 0; v < 3 && rc == FDFD_OK; ++v) rc = build(&mid, v, v, 1, c->q[v]);
    cudaFree(tmp);
    if (objbuf) cudaFree(objbuf);
    if (rc != FDFD_OK) return rc;
    // ghost planes: periodic wrap (single slab), neighbour exchange (multi slab), zero at symmetry ends
    for (double2 *g : ghosted) {
        FDFD_CUDA(c, cudaMemsetAsync(g, 0, (size_t)Nxy * sizeof(double2), c->stream));
        FDFD_CUDA(c, cudaMemsetAsync(g + (nzl + 1) * Nxy, 0, (size_t)Nxy * sizeof(double2), c->stream));
        if (c->d.nranks == 1) {
            if (c->d.isbloch[2]) {
                FDFD_CUDA(c, cudaMemcpyAsync(g, g + nzl * Nxy, (size_t)Nxy * sizeof(double2),
                                             cudaMemcpyDeviceToDevice, c->stream));
                FDFD_CUDA(c, cudaMemcpyAsync(g + (nzl + 1) * Nxy, g + Nxy, (size_t)Nxy * sizeof(double2),
                                             cudaMemcpyDeviceToDevice, c->stream));
            }
        } else {
            if (!c->comm) return set_err(c, FDFD_ESTATE, "nranks > 1: call fdfd_comm_init before the first apply");
            int r = halo_exchange_ghosted(c, g, c->stream);
            if (r != FDFD_OK) return r;
        }
    }
    // cmp-first DOF layout: a second, interleaved copy of the (ghosted) diagonal mass arrays for the row-pair kernel,
    // which stages material through the same TMA ring as x (apply_rowpair.cu); +16 B/DOF of HBM capacity
    if (c->md[0] && c->d.order_cmpfirst && c->d.kernel != FDFD_KERNEL_NAIVE) {
        // Real diagonal entries (real omega, real eps_vv - every lossless dielectric) are kept as doubles: the kernel's
        // tensor-map boxes then move 8 instead of 16 bytes per entry (40 instead of 48 B/DOF per apply).  Needs the
        // tensor-map path; FDFD_RP_MDR=0 / FDFD_RP_TMAP=0 keep the complex rows (A/B timing).
        static const bool allow_real = [] {
            const char *a = getenv("FDFD_RP_MDR"), *t = getenv("FDFD_RP_TMAP");
            return !(a && atoi(a) == 0) && !(t && atoi(t) == 0);
        }();
        bool real_rows = false;
        // z-slabs: every rank takes the same kernel variant (the fused dots and the graphs assume one plan)
        int r_any = FDFD_OK;
        auto any_rank = [&](int &flag_io) -> int {
            if (c->d.nranks <= 1) return FDFD_OK;
            double *f = nullptr;
            const double hv = flag_io ? 1.0 : 0.0;
            FDFD_CUDA(c, cudaMalloc((void **)&f, sizeof(double)));
            FDFD_CUDA(c, cudaMemcpyAsync(f, &hv, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            int ra = allreduce_sum(c, f, 1, c->stream);
            double gv = 1.0;
            if (ra == FDFD_OK) {
                FDFD_CUDA(c, cudaMemcpyAsync(&gv, f, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
            }
            cudaFree(f);
            if (ra != FDFD_OK) return ra;
            flag_io = gv > 0.0;
            return FDFD_OK;
        };
        if (allow_real && tmap_probe(c->md[0])) {
            int *flag = nullptr;
            FDFD_CUDA(c, cudaMalloc((void **)&flag, sizeof(int)));
            FDFD_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
            const int64_t pitch = mdr_row_pitch((int)c->d.N[0]), nrow = c->d.N[1] * (nzl + 2);
            FDFD_CUDA(c, cudaMalloc((void **)&c->md_aos_r, (size_t)(pitch * nrow) * sizeof(double)));
            FDFD_CUDA(c, cudaMemsetAsync(c->md_aos_r, 0, (size_t)(pitch * nrow) * sizeof(double), c->stream));
            interleave3_real_kernel<<<nblocks(Mg), 256, 0, c->stream>>>(c->md_aos_r, c->md[0], c->md[1], c->md[2], Mg,
                                                                        (int)c->d.N[0], pitch, flag);
            FDFD_CUDA(c, cudaGetLastError());
            int any_imag = 1;
            FDFD_CUDA(c, cudaMemcpyAsync(&any_imag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(flag);
            if ((r_any = any_rank(any_imag)) != FDFD_OK) return r_any;
            real_rows = !any_imag;
            if (!real_rows) { cudaFree(c->md_aos_r); c->md_aos_r = nullptr; }
        }
        if (!real_rows) {
            FDFD_CUDA(c, cudaMalloc((void **)&c->md_aos, (size_t)3 * Mg * sizeof(double2)));
            interleave3_kernel<<<nblocks(Mg), 256, 0, c->stream>>>(c->md_aos, c->md[0], c->md[1], c->md[2], Mg);
            FDFD_CUDA(c, cudaGetLastError());
        }
        // Symmetric off-diagonal entries that are real as well (the smoothed tensor of lossless media): a third,
        // interleaved and padded copy for the fused full-tensor shape of the row-pair kernel (8 B/DOF on the flagged blocks
        // instead of 16).  FDFD_RP_FUSED=0 keeps the two-pass / first-generation plans (A/B timing).
        static const bool allow_fused = [] { const char *a = getenv("FDFD_RP_FUSED"); return !(a && atoi(a) == 0); }();
        if (real_rows && allow_fused && c->mo[0] && c->off_sym) {
            int *flag = nullptr;
            FDFD_CUDA(c, cudaMalloc((void **)&flag, sizeof(int)));
            FDFD_CUDA(c, cudaMemsetAsync(flag, 0, sizeof(int), c->stream));
            const int Nx = (int)c->d.N[0], Ny = (int)c->d.N[1];
            const int64_t pitch = mdr_row_pitch(Nx + 2), nrow = (int64_t)(Ny + 2) * (nzl + 2);
            const int64_t npad = (int64_t)(Nx + 2) * nrow;
            FDFD_CUDA(c, cudaMalloc((void **)&c->mo_aos_r, (size_t)(pitch * nrow) * sizeof(double)));
            FDFD_CUDA(c, cudaMemsetAsync(c->mo_aos_r, 0, (size_t)(pitch * nrow) * sizeof(double), c->stream));
            // order of c->mo: (0,1), (0,2), (1,0), (1,2), (2,0), (2,1)
            interleave3_real_pad_kernel<<<nblocks(npad), 256, 0, c->stream>>>(c->mo_aos_r, c->mo[0], c->mo[1], c->mo[3], npad, Nx, Ny,
                                                                             c->d.isbloch[0] ? 1 : 0, c->d.isbloch[1] ? 1 : 0, pitch, flag);
            FDFD_CUDA(c, cudaGetLastError());
            int any_imag = 1;
            FDFD_CUDA(c, cudaMemcpyAsync(&any_imag, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(flag);
            if ((r_any = any_rank(any_imag)) != FDFD_OK) return r_any;
            if (any_imag) { cudaFree(c->mo_aos_r); c->mo_aos_r = nullptr; }
        }
    }
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

int ensure_ready(Ctx *c) {
    if (c->dev == -2) return set_err(c, FDFD_ESTATE, "host-only handle (device = -2): only fdfd_export_pattern is available");
    if (!c->dirty) return FDFD_OK;
    if (!c->have_coeffs) return set_err(c, FDFD_ESTATE, "fdfd_set_coeffs has not been called");
    const bool ee = c->d.field_type == FDFD_FT_EE;
    if (c->omega != cplx(0.0) && ee && !c->have_eps)
        return set_err(c, FDFD_ESTATE, "fdfd_set_eps has not been called (needed when omega != 0)");
    if (!ee && !c->have_eps) return set_err(c, FDFD_ESTATE, "FT_HH needs fdfd_set_eps (A = Ce (Peps \\ Cm) - w^2 Pmu)");
    FDFD_CUDA(c, cudaSetDevice(c->dev));
    int r = upload_coefs(c);
    if (r != FDFD_OK) return r;
    r = upload_materials(c);
    if (r != FDFD_OK) return r;
    // occupancy mask of the off-diagonal material for the tiled kernel
    if (c->offmask) { cudaFree(c->offmask); c->offmask = nullptr; }
    if (c->corr_list) { cudaFree(c->corr_list); c->corr_list = nullptr; }
    c->corr_count = 0;
    c->offmask_ty = 0;
    c->off_frac = c->mo[0] ? 1.0 : 0.0;
    if (c->mo[0] && c->d.kernel != FDFD_KERNEL_NAIVE) {
        ApplyParams p;
        fill_params(c, p, nullptr, nullptr, false);
        // the fused row-pair shape takes the operator from a few per cent of flagged blocks on (FDFD_RP_FUSE_MIN, measured
        // choice - DESIGN.md section 5); without it the first-generation fused kernel pays off above 25 %
        static const double fuse_min_rp = [] { const char *e = getenv("FDFD_RP_FUSE_MIN"); return e ? atof(e) : 0.03; }();
        FDFD_CUDA(c, tiled_build_offmask(p, &c->offmask, &c->offmask_ty, &c->off_frac, &c->corr_list, &c->corr_count,
                                         c->stream, c->mo_aos_r ? fuse_min_rp : 0.25));
    }
    c->dirty = false;
    return FDFD_OK;
}

void fill_params(Ctx *c, ApplyParams &p, const double2 *x, double2 *y, bool transpose) {
    std::memset(&p, 0, sizeof(p));
    const int64_t Nx = c->d.N[0], Ny = c->d.N[1], nzl = c->k1 - c->k0;
    p.Nx = (int)Nx; p.Ny = (int)Ny; p.nzl = (int)nzl; p.Nz = (int)c->d.N[2]; p.kz0 = (int)c->k0;
    for (int w = 0; w < 3; ++w) { p.s1[w] = c->s1[w]; p.wrap[w] = c->d.isbloch[w] ? 1 : 0; }
    p.cmpfirst = c->d.order_cmpfirst ? 1 : 0;
    p.has_mass = c->has_mass ? 1 : 0;
    p.md_uniform = c->md_uniform;
    p.has_off = c->mo[0] != nullptr;
    p.has_q = c->q[0] != nullptr;
    p.c = transpose ? c->ct : c->cf;
    for (int i = 0; i < 3; ++i) { p.md[i] = c->md[i]; p.q[i] = c->q[i]; }
    p.md_aos = c->md_aos;
    p.md_aos_r = c->md_aos_r;
    p.mo_aos_r = c->mo_aos_r;
    for (int i = 0; i < 6; ++i) p.mo[i] = transpose ? c->mo_t[i] : c->mo[i];
    const int64_t Nxy = Nx * Ny;
    p.x.base = x;
    if (p.cmpfirst) { p.x.pstride = 3 * Nxy; p.x.cs = 1; p.x.es = 3; }
    else            { p.x.pstride = Nxy; p.x.cs = Nxy * nzl; p.x.es = 1; }
    if (c->d.nranks == 1) {
        // periodic wrap inside the slab (for symmetry boundaries the planes are only multiplied by zeros)
        p.x.lo = x + (nzl - 1) * p.x.pstride; p.x.cs_lo = p.x.cs;
        p.x.hi = x;                           p.x.cs_hi = p.x.cs;
    } else {
        p.x.lo = c->halo_lo; p.x.hi = c->halo_hi;
        p.x.cs_lo = p.x.cs_hi = p.cmpfirst ? 1 : Nxy;
    }
    p.y = y; p.y_pstride = p.x.pstride; p.y_cs = p.x.cs; p.y_es = p.x.es;
    p.offmask = c->offmask; p.offmask_ty = c->offmask_ty;
    p.corr_list = c->corr_list; p.corr_count = c->corr_count;
    if (c->dot_req) {
        p.dot_mode = 2; p.dot_out = c->dot_req; p.dot_partial = c->dot_partial; p.dot_ticket = c->dot_ticket;
        p.dot_cap = c->dot_cap;
    }
}

int apply_device(Ctx *c, const double2 *x, double2 *y, bool transpose) {
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    if (x == y) return set_err(c, FDFD_EINVAL, "fdfd_apply: x and y must not alias");
    ApplyParams p;
    fill_params(c, p, x, y, transpose);
    const bool can_tile = tiled_supported(p);
    if (c->d.kernel == FDFD_KERNEL_TILED && !can_tile)
        return set_err(c, FDFD_EINVAL, "tiled kernel: unsupported arrangement");
    const bool use_tiled = c->d.kernel != FDFD_KERNEL_NAIVE && can_tile;
    // timing experiments only (results are wrong with FDFD_DEBUG_SKIP_HALO): where does the multi-slab overhead go?
    static const bool env_skip_halo = getenv("FDFD_DEBUG_SKIP_HALO") != nullptr;
    // halo planes already in c->halo_lo / halo_hi: the caller supplied them with the host vector (fdfd_apply_host_halos)
    const bool dbg_skip_halo = env_skip_halo || c->halo_preloaded;
    // The interior/boundary split (exchange overlapped with the interior planes) costs more than it hides on
    // NVLink (measured: +21 us for the split vs 25 us for the exchange), so it is opt-in; the default is exchange,
    // then one launch - and inside the Krylov loops the exchange is started early by the kernel that produces the
    // vector (halo_prefetch), so the apply only waits on an event.
    static const bool split_overlap = getenv("FDFD_SPLIT_OVERLAP") != nullptr;
    if (c->d.nranks > 1 && c->halo_for == x && x != nullptr && c->peer.direct_pending) {
        // peer-direct (opt-in): the neighbours announced their boundary planes of this workspace vector; the kernel
        // reads them in place over NVLink - no halo copy
        c->halo_for = nullptr;
        if (!peer_direct_planes(c, x, &p.x.lo, &p.x.hi)) return set_err(c, FDFD_ESTATE, "peer-direct: vector is not in the workspace");
        if ((r = peer_direct_wait(c, c->stream)) != FDFD_OK) return r;
        static const bool verbose = getenv("FDFD_VERBOSE") != nullptr;
        if (verbose && c->peer.depoch == 1) fprintf(stderr, "fdfd rank %d: peer-direct halo reads active\n", c->d.rank);
        if (use_tiled) {
            int nl = 0;
            FDFD_CUDA(c, launch_apply_tiled(p, 0, p.nzl, c->stream, &nl));
            c->launches += nl;
        } else {
            FDFD_CUDA(c, launch_apply_naive(p, c->stream));
            c->launches += 1;
        }
        return FDFD_OK;
    }
    if (c->d.nranks > 1 && c->halo_for == x && x != nullptr) {
        c->halo_for = nullptr;
        c->comm_pending = false;
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        if (use_tiled) {
            int nl = 0;
            FDFD_CUDA(c, launch_apply_tiled(p, 0, p.nzl, c->stream, &nl));
            c->launches += nl;
        } else {
            FDFD_CUDA(c, launch_apply_naive(p, c->stream));
            c->launches += 1;
        }
        return FDFD_OK;
    }
    c->halo_for = nullptr;
    if (c->comm_pending) {   // NCCL operations on one communicator must not overlap: order after the prefetch
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        c->comm_pending = false;
    }
    // EXPERIMENTAL, opt-in, not validated for back-to-back applies: ONE launch that does not wait for the exchange.
    // The kernel schedules the two z-chunks that touch a neighbour's plane last and gates them on a flag word that
    // a stream memory operation sets behind the NCCL exchange, so the transfer runs under the interior chunks.
    // Needs at least two interior chunks per tile column (several waves of CTAs ahead of the gated ones).  Status
    // (2x B200): results equal the single-slab operator when every apply is followed by a host sync; a stream of
    // back-to-back applies (bench.py) did not finish - the NCCL kernel of apply i+1 and the apply kernel i+1
    // become runnable together and, if the apply kernel is dispatched first, its gated CTAs can occupy every SM
    // before the NCCL kernel gets one.  The gated spin is bounded (traps after ~4 s) so this shows as an error,
    // not a hang.  See DESIGN.md section 10 for the fix that is planned (copy-engine / peer-memory transfer).
    static const bool inkernel_wait = getenv("FDFD_INKERNEL_HALO_WAIT") != nullptr;
    // Second generation of the same idea, on the persistent row-pair kernel: the boundary z-chunks of every tile column
    // are walked last and only the producer warp of a CTA waits - for the flag word, right before its first load of a
    // neighbour's plane.  ON by default when the planes travel by the copy-engine peer exchange (peer.cpp: the transfer
    // needs no SM, the apply keeps all 148 CTAs); with NCCL as the data plane it is opt-in (FDFD_HALO_OVERLAP=1) and
    // leaves FDFD_HALO_SM_RESERVE (8) SMs to the NCCL kernels.  Measured on 2x B200, C2 per GPU (gpurun_out/r02c16_*,
    // r02c18_*): NCCL, exchange then launch 219 GDOF/s; peer exchange then launch 217; NCCL overlapped, 8 SMs reserved
    // 189 (item quantisation: 735 items over 140 CTAs; with 4 or 2 free SMs the NCCL kernel made no progress and the
    // bounded spin trapped); peer exchange overlapped 245 = 2 x the single-GPU rate.  FDFD_HALO_OVERLAP=0 turns it off.
    static const int halo_overlap_env = [] { const char *e = getenv("FDFD_HALO_OVERLAP"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool halo_overlap = halo_overlap_env >= 0 ? halo_overlap_env == 1 : c->peer.ready;
    p.halo_sm_free = c->peer.ready ? 1 : 0;
    const bool overlap_rp = halo_overlap && c->d.nranks > 1 && use_tiled && !dbg_skip_halo && stream_write_u32_available() &&
                            rowpair_halo_overlap_ok(p);
    if (overlap_rp || (c->d.nranks > 1 && use_tiled && inkernel_wait && c->d.order_cmpfirst && !dbg_skip_halo &&
                       stream_write_u32_available() && tiled_plan_nchunk(p) >= 4)) {
        if (!c->halo_flag) {
            FDFD_CUDA(c, cudaMalloc((void **)&c->halo_flag, sizeof(uint32_t)));
            FDFD_CUDA(c, cudaMemset(c->halo_flag, 0, sizeof(uint32_t)));
        }
        const uint32_t epoch = ++c->halo_epoch;
        FDFD_CUDA(c, cudaEventRecord(c->ev_x, c->stream));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_comm, c->ev_x, 0));
        if ((r = halo_exchange(c, x, c->halo_lo, c->halo_hi, c->stream_comm)) != FDFD_OK) return r;
        if ((r = stream_write_u32(c, c->stream_comm, c->halo_flag, epoch)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaEventRecord(c->ev_halo, c->stream_comm));
        c->comm_pending = true;     // later NCCL operations order themselves after this exchange through ev_halo
        p.halo_flag = c->halo_flag;
        p.halo_expect = epoch;
        int nl = 0;
        FDFD_CUDA(c, launch_apply_tiled(p, 0, p.nzl, c->stream, &nl));
        c->launches += nl;
        return FDFD_OK;
    }
    if (c->d.nranks > 1 && use_tiled && p.nzl >= 4 && split_overlap) {
        // z-slabs: the halo exchange (NCCL, own stream) overlaps the interior planes, which only need this rank's
        // own planes; the two boundary planes run on a high-priority stream as soon as the halos have landed,
        // concurrently with the interior kernel (SURVEY.md 8e "Overlap").
        FDFD_CUDA(c, cudaEventRecord(c->ev_x, c->stream));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_comm, c->ev_x, 0));
        if (!dbg_skip_halo && (r = halo_exchange(c, x, c->halo_lo, c->halo_hi, c->stream_comm)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaEventRecord(c->ev_halo, c->stream_comm));
        int nl = 0;
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_bnd, c->ev_halo, 0));
        FDFD_CUDA(c, launch_apply_tiled(p, 0, 1, c->stream_bnd, &nl));
        FDFD_CUDA(c, launch_apply_tiled(p, p.nzl - 1, p.nzl, c->stream_bnd, &nl));
        FDFD_CUDA(c, cudaEventRecord(c->ev_bnd, c->stream_bnd));
        FDFD_CUDA(c, launch_apply_tiled(p, 1, p.nzl - 1, c->stream, &nl));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_bnd, 0));
        c->launches += nl;
        return FDFD_OK;
    }
    if (c->d.nranks > 1 && !dbg_skip_halo) {
        r = halo_exchange(c, x, c->halo_lo, c->halo_hi, c->stream);
        if (r != FDFD_OK) return r;
    }
    if (use_tiled) {
        int nl = 0;
        FDFD_CUDA(c, launch_apply_tiled(p, 0, p.nzl, c->stream, &nl));
        c->launches += nl;
    } else {
        FDFD_CUDA(c, launch_apply_naive(p, c->stream));
        c->launches += 1;
    }
    return FDFD_OK;
}

int ensure_dot_buffers(Ctx *c) {   // outside any stream capture
    if (c->dot_partial) return FDFD_OK;
    c->dot_cap = 1 << 20;
    FDFD_CUDA(c, cudaMalloc((void **)&c->dot_partial, (size_t)c->dot_cap * 4 * sizeof(double)));
    FDFD_CUDA(c, cudaMalloc((void **)&c->dot_ticket, sizeof(unsigned int)));
    FDFD_CUDA(c, cudaMemset(c->dot_ticket, 0, sizeof(unsigned int)));
    return FDFD_OK;
}

int apply_device_dots(Ctx *c, const double2 *x, double2 *y, double *dot_out, bool *fused) {
    *fused = false;
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    static const bool no_fuse = getenv("FDFD_NO_DOT_FUSION") != nullptr;
    static const bool split_overlap = getenv("FDFD_SPLIT_OVERLAP") != nullptr;
    ApplyParams p0;
    fill_params(c, p0, x, y, false);
    const bool tiled = c->d.kernel != FDFD_KERNEL_NAIVE && tiled_supported(p0);
    if (no_fuse || !tiled || (c->d.nranks > 1 && split_overlap)) return apply_device(c, x, y, false);
    if (!c->dot_partial) return apply_device(c, x, y, false);   // ensure_dot_buffers() was not called
    c->dot_req = dot_out;
    r = apply_device(c, x, y, false);
    c->dot_req = nullptr;
    *fused = (r == FDFD_OK);
    return r;
}

bool halo_prefetch_usable(const Ctx *c) {
    return c->d.nranks > 1 && c->d.order_cmpfirst && (c->k1 - c->k0) >= 3 && c->comm != nullptr &&
           getenv("FDFD_NO_HALO_PREFETCH") == nullptr;
}

int halo_prefetch(Ctx *c, const double2 *v) {
    if (!halo_prefetch_usable(c)) return FDFD_OK;
    FDFD_CUDA(c, cudaEventRecord(c->ev_x, c->stream));
    FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_comm, c->ev_x, 0));
    const double2 *plo = nullptr, *phi = nullptr;
    if (c->peer.direct && peer_direct_planes(c, v, &plo, &phi)) {
        // peer-direct (opt-in): nothing is copied - only tell the neighbours that the planes they will read are final
        int rs = peer_direct_signal(c, c->stream_comm);
        if (rs != FDFD_OK) return rs;
        c->halo_for = v;
        return FDFD_OK;
    }
    int r = halo_exchange(c, v, c->halo_lo, c->halo_hi, c->stream_comm);
    if (r != FDFD_OK) return r;
    FDFD_CUDA(c, cudaEventRecord(c->ev_halo, c->stream_comm));
    c->halo_for = v;
    c->comm_pending = true;
    return FDFD_OK;
}

static int stage_buffers(Ctx *c) {
    if (c->dev == -2) return set_err(c, FDFD_ESTATE, "host-only handle (device = -2): only fdfd_export_pattern is available");
    if (!c->stage_x) FDFD_CUDA(c, cudaMalloc((void **)&c->stage_x, (size_t)c->nloc * sizeof(double2)));
    if (!c->stage_y) FDFD_CUDA(c, cudaMalloc((void **)&c->stage_y, (size_t)c->nloc * sizeof(double2)));
    return FDFD_OK;
}

// fdfd_apply with HOST buffers, pipelined over z sub-slabs: H2D of sub-slab s+1, the kernel of sub-slab s and
// D2H of sub-slab s-1 run concurrently on three streams (PCIe is full duplex), so the call costs about one
// direction's transfer time instead of H2D + kernel + D2H.  Needs the cmp-first layout (a z sub-slab is contiguous)
// and the tiled kernel; other configurations use the plain staged path.  On z-slabs the halo planes come either from
// the caller's host vector (fdfd_apply_host_halos) or from a device exchange of the two boundary planes, which are
// copied up first.
static bool can_pipeline(Ctx *c, bool host_halos) {
    return (c->d.nranks == 1 || host_halos || c->comm != nullptr) && c->d.order_cmpfirst && c->d.kernel != FDFD_KERNEL_NAIVE &&
           (c->k1 - c->k0) >= 16;
}

// xlo_h / xhi_h (z-slabs only): host pointers to the plane below / above this slab (null: symmetry boundary, the halo
// buffer keeps its zeros) - the caller holds the whole vector in host memory, so no exchange between GPUs is needed
static int apply_host_pipelined(Ctx *c, const double2 *xh, double2 *yh, bool transpose, bool host_halos = false,
                                const double2 *xlo_h = nullptr, const double2 *xhi_h = nullptr) {
    const int64_t nzl = c->k1 - c->k0, pl = c->plane;
    static const int s_env = [] { const char *e = getenv("FDFD_PIPE_SLABS"); return e ? atoi(e) : 0; }();
    const int S = (int)std::max<int64_t>(1, std::min<int64_t>(s_env > 0 ? s_env : 16, nzl / 2));
    if (!c->stream_d2h) FDFD_CUDA(c, cudaStreamCreateWithFlags(&c->stream_d2h, cudaStreamNonBlocking));
    while ((int)c->ev_h2d.size() < S + 1) {
        cudaEvent_t e;
        FDFD_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_h2d.push_back(e);
    }
    while ((int)c->ev_k.size() < S) {
        cudaEvent_t e;
        FDFD_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_k.push_back(e);
    }
    ApplyParams p;
    fill_params(c, p, c->stage_x, c->stage_y, transpose);
    auto bound = [&](int s) { return (int)((nzl * s) / S); };
    if (c->d.nranks == 1) {
        // the wrap plane (x_lo = plane nzl-1) first, then the sub-slabs in order
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x + (nzl - 1) * pl, xh + (nzl - 1) * pl, (size_t)pl * sizeof(double2),
                                     cudaMemcpyHostToDevice, c->stream_copy));
    } else if (host_halos) {
        // z-slab: the neighbours' boundary planes come straight from the caller's host vector
        if (xlo_h) FDFD_CUDA(c, cudaMemcpyAsync(c->halo_lo, xlo_h, (size_t)pl * sizeof(double2), cudaMemcpyHostToDevice, c->stream_copy));
        if (xhi_h) FDFD_CUDA(c, cudaMemcpyAsync(c->halo_hi, xhi_h, (size_t)pl * sizeof(double2), cudaMemcpyHostToDevice, c->stream_copy));
    } else {
        // z-slab, one process per GPU: this rank's two boundary planes go up first and are exchanged on the device
        // (peer exchange or NCCL) while the sub-slabs follow; the kernels wait for the exchange below
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, xh, (size_t)pl * sizeof(double2), cudaMemcpyHostToDevice, c->stream_copy));
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x + (nzl - 1) * pl, xh + (nzl - 1) * pl, (size_t)pl * sizeof(double2),
                                     cudaMemcpyHostToDevice, c->stream_copy));
        FDFD_CUDA(c, cudaEventRecord(c->ev_x, c->stream_copy));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_comm, c->ev_x, 0));
        // the exchange releases the halo buffers of the previous epoch to the neighbours: it must also be ordered behind
        // whatever this handle's compute stream still holds (nothing, after a blocking API call - but cheap to state)
        FDFD_CUDA(c, cudaEventRecord(c->ev_bnd, c->stream));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_comm, c->ev_bnd, 0));
        int rh = halo_exchange(c, c->stage_x, c->halo_lo, c->halo_hi, c->stream_comm);
        if (rh != FDFD_OK) return rh;
        FDFD_CUDA(c, cudaEventRecord(c->ev_halo, c->stream_comm));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
    }
    FDFD_CUDA(c, cudaEventRecord(c->ev_h2d[S], c->stream_copy));
    for (int s = 0; s < S; ++s) {
        const int64_t o = (int64_t)bound(s) * pl, n = (int64_t)(bound(s + 1) - bound(s)) * pl;
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x + o, xh + o, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice,
                                     c->stream_copy));
        FDFD_CUDA(c, cudaEventRecord(c->ev_h2d[s], c->stream_copy));
    }
    FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_h2d[S], 0));
    for (int s = 0; s < S; ++s) {
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_h2d[std::min(s + 1, S - 1)], 0));
        int nl = 0;
        FDFD_CUDA(c, launch_apply_tiled(p, bound(s), bound(s + 1), c->stream, &nl));
        c->launches += nl;
        FDFD_CUDA(c, cudaEventRecord(c->ev_k[s], c->stream));
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream_d2h, c->ev_k[s], 0));
        const int64_t o = (int64_t)bound(s) * pl, n = (int64_t)(bound(s + 1) - bound(s)) * pl;
        FDFD_CUDA(c, cudaMemcpyAsync(yh + o, c->stage_y + o, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost,
                                     c->stream_d2h));
    }
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream_d2h));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

static int apply_host(Ctx *c, const fdfd_c128 *x, fdfd_c128 *y, bool transpose) {
    int r;
    if ((r = stage_buffers(c)) != FDFD_OK) return r;
    if ((r = ensure_ready(c)) != FDFD_OK) return r;
    if (can_pipeline(c, false)) {
        if (c->d.nranks > 1) {   // order after anything a Krylov prefetch left on the communication stream
            if (c->comm_pending) { FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0)); FDFD_CUDA(c, cudaStreamSynchronize(c->stream)); }
            c->comm_pending = false;
            c->halo_for = nullptr;
        }
        return apply_host_pipelined(c, reinterpret_cast<const double2 *>(x), reinterpret_cast<double2 *>(y), transpose);
    }
    const size_t bytes = (size_t)c->nloc * sizeof(double2);
    FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, x, bytes, cudaMemcpyHostToDevice, c->stream));
    if ((r = apply_device(c, c->stage_x, c->stage_y, transpose)) != FDFD_OK) return r;
    FDFD_CUDA(c, cudaMemcpyAsync(y, c->stage_y, bytes, cudaMemcpyDeviceToHost, c->stream));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

// host-buffer apply of one z-slab whose halo planes the caller supplies from host memory (no exchange between GPUs)
static int apply_host_halos(Ctx *c, const fdfd_c128 *x, const fdfd_c128 *xlo, const fdfd_c128 *xhi, fdfd_c128 *y, bool transpose) {
    int r;
    if (c->d.nranks == 1) return apply_host(c, x, y, transpose);
    if ((r = stage_buffers(c)) != FDFD_OK) return r;
    if ((r = ensure_ready(c)) != FDFD_OK) return r;
    if (c->comm_pending) {
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
        c->comm_pending = false;
    }
    c->halo_for = nullptr;
    if (can_pipeline(c, true))
        return apply_host_pipelined(c, reinterpret_cast<const double2 *>(x), reinterpret_cast<double2 *>(y), transpose, true,
                                    reinterpret_cast<const double2 *>(xlo), reinterpret_cast<const double2 *>(xhi));
    const size_t bytes = (size_t)c->nloc * sizeof(double2), pb = (size_t)c->plane * sizeof(double2);
    FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, x, bytes, cudaMemcpyHostToDevice, c->stream));
    if (xlo) FDFD_CUDA(c, cudaMemcpyAsync(c->halo_lo, xlo, pb, cudaMemcpyHostToDevice, c->stream));
    if (xhi) FDFD_CUDA(c, cudaMemcpyAsync(c->halo_hi, xhi, pb, cudaMemcpyHostToDevice, c->stream));
    c->halo_preloaded = true;
    r = apply_device(c, c->stage_x, c->stage_y, transpose);
    c->halo_preloaded = false;
    if (r != FDFD_OK) return r;
    FDFD_CUDA(c, cudaMemcpyAsync(y, c->stage_y, bytes, cudaMemcpyDeviceToHost, c->stream));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

}  // namespace fdfd

using namespace fdfd;

#define CHECK_H(h)                                                             \
    if (!(h)) return set_err(nullptr, FDFD_EINVAL, "null handle");            \
    Ctx *c = static_cast<Ctx *>(h);                                            \
    if (c->dev != -2) {                                                        \
        cudaError_t e_ = cudaSetDevice(c->dev);                                \
        if (e_ != cudaSuccess) return set_err(c, FDFD_ECUDA, cudaGetErrorString(e_)); \
    }

extern "C" {

int fdfd_offdiag_symmetric(fdfd_handle h, int *symmetric) {
    Ctx *c = static_cast<Ctx *>(h);
    if (!c || !symmetric) return FDFD_EINVAL;
    int r = fdfd::ensure_ready(c);
    if (r != FDFD_OK) return r;
    *symmetric = (c->mo[0] != nullptr && c->off_sym) ? 1 : 0;
    return FDFD_OK;
}

int fdfd_calc_matparams(const fdfd_matparams_desc *desc, fdfd_c128 *out, int where) {
    std::string err;
    int r;
    try {
        r = fdfd::calc_matparams(desc, out, where, err);
    } catch (const std::exception &e) {
        r = FDFD_ENOMEM;
        err = e.what();
    }
    if (r != FDFD_OK) fdfd::set_err(nullptr, r, err);
    return r;
}

const char *fdfd_version(void) { return "fdfd_b200 0.1.0 (sm_100a)"; }

const char *fdfd_last_error(fdfd_handle h) { return h ? static_cast<Ctx *>(h)->err.c_str() : g_create_err.c_str(); }

int fdfd_partition(int64_t Nz, int32_t nranks, int32_t rank, int64_t *k0, int64_t *k1) {
    if (Nz < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > Nz || !k0 || !k1) return FDFD_EINVAL;
    *k0 = (Nz * rank) / nranks;
    *k1 = (Nz * (rank + 1)) / nranks;
    return FDFD_OK;
}

int fdfd_halo_plan(int32_t nranks, int32_t rank, int32_t wrapz, int32_t *up, int32_t *dn) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !up || !dn) return FDFD_EINVAL;
    halo_neighbours(nranks, rank, wrapz != 0, up, dn);
    return FDFD_OK;
}

int fdfd_create(fdfd_handle *out, const fdfd_desc *d) {
    if (!out || !d) return set_err(nullptr, FDFD_EINVAL, "fdfd_create: null argument");
    *out = nullptr;
    for (int w = 0; w < 3; ++w)
        if (d->N[w] < 1 || d->N[w] > (1 << 30)) return set_err(nullptr, FDFD_EINVAL, "fdfd_create: bad grid size");
    if (d->field_type != FDFD_FT_EE && d->field_type != FDFD_FT_HH)
        return set_err(nullptr, FDFD_EINVAL, "ft is unsupported.");  // reference @error model.jl:242
    if (d->nranks < 1 || d->rank < 0 || d->rank >= d->nranks || d->nranks > d->N[2])
        return set_err(nullptr, FDFD_EINVAL, "fdfd_create: bad rank/nranks (need 1 <= nranks <= Nz)");
    if (d->kernel < FDFD_KERNEL_AUTO || d->kernel > FDFD_KERNEL_TILED)
        return set_err(nullptr, FDFD_EINVAL, "fdfd_create: bad kernel selector");
    for (int w = 0; w < 3; ++w)
        if (!d->isbloch[w] && d->N[w] < 2 && false) return FDFD_EINVAL;
    if (d->device == -2) {
        // host-only handle: debug pattern export (integer work on the host) and nothing else
        fdfd_ctx *hc = new (std::nothrow) fdfd_ctx();
        if (!hc) return set_err(nullptr, FDFD_ENOMEM, "out of host memory");
        hc->d = *d;
        hc->dev = -2;
        fdfd_partition(d->N[2], d->nranks, d->rank, &hc->k0, &hc->k1);
        hc->plane = 3 * d->N[0] * d->N[1];
        hc->nloc = hc->plane * (hc->k1 - hc->k0);
        *out = hc;
        return FDFD_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return set_err(nullptr, FDFD_ECUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                " (libfdfd_b200 has no CPU fallback)");
    fdfd_ctx *c = new (std::nothrow) fdfd_ctx();
    if (!c) return set_err(nullptr, FDFD_ENOMEM, "out of host memory");
    c->d = *d;
    if (d->device >= 0) c->dev = d->device;
    else cudaGetDevice(&c->dev);
    if (c->dev >= ndev) { delete c; return set_err(nullptr, FDFD_EINVAL, "fdfd_create: device ordinal out of range"); }
    fdfd_partition(d->N[2], d->nranks, d->rank, &c->k0, &c->k1);
    c->plane = 3 * d->N[0] * d->N[1];
    c->nloc = c->plane * (c->k1 - c->k0);
    auto fail = [&](cudaError_t ee, const char *what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(ee);
        free_device(c);
        delete c;
        return set_err(nullptr, FDFD_ECUDA, m);
    };
    if ((e = cudaSetDevice(c->dev)) != cudaSuccess) return fail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&c->stream_copy, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "cudaStreamCreate");
    if (d->nranks > 1) {
        const size_t pb = (size_t)c->plane * sizeof(double2);
        if ((e = cudaMalloc((void **)&c->halo_lo, pb)) != cudaSuccess) return fail(e, "cudaMalloc(halo)");
        if ((e = cudaMalloc((void **)&c->halo_hi, pb)) != cudaSuccess) return fail(e, "cudaMalloc(halo)");
        cudaMemset(c->halo_lo, 0, pb);
        cudaMemset(c->halo_hi, 0, pb);
        if (getenv("FDFD_INKERNEL_HALO_WAIT")) {
            // experimental path only: the NCCL kernels must win the work distributor against an apply kernel whose
            // last CTAs spin on them, so their stream gets the highest priority (the default paths are unchanged)
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            e = cudaStreamCreateWithPriority(&c->stream_comm, cudaStreamNonBlocking, hi_pri);
        } else {
            e = cudaStreamCreateWithFlags(&c->stream_comm, cudaStreamNonBlocking);
        }
        if (e != cudaSuccess) return fail(e, "cudaStreamCreate");
        if ((e = cudaEventCreateWithFlags(&c->ev_x, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        if ((e = cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming)) != cudaSuccess) return fail(e, "cudaEventCreate");
        {
            int lo_pri = 0, hi_pri = 0;
            cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
            if ((e = cudaStreamCreateWithPriority(&c->stream_bnd, cudaStreamNonBlocking, hi_pri)) != cudaSuccess)
                return fail(e, "cudaStreamCreate");
        }
    }
    *out = c;
    return FDFD_OK;
}

int fdfd_destroy(fdfd_handle h) {
    if (!h) return FDFD_OK;
    Ctx *c = static_cast<Ctx *>(h);
    if (c->dev == -2) { delete static_cast<fdfd_ctx *>(h); return FDFD_OK; }
    cudaSetDevice(c->dev);
    cudaDeviceSynchronize();
    comm_destroy(c);
    free_device(c);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream_copy) cudaStreamDestroy(c->stream_copy);
    if (c->stream_d2h) cudaStreamDestroy(c->stream_d2h);
    if (c->stream_comm) cudaStreamDestroy(c->stream_comm);
    if (c->ev_x) cudaEventDestroy(c->ev_x);
    if (c->ev_halo) cudaEventDestroy(c->ev_halo);
    if (c->ev_bnd) cudaEventDestroy(c->ev_bnd);
    if (c->stream_bnd) cudaStreamDestroy(c->stream_bnd);
    for (auto e : c->ev_h2d) cudaEventDestroy(e);
    for (auto e : c->ev_k) cudaEventDestroy(e);
    delete static_cast<fdfd_ctx *>(h);
    return FDFD_OK;
}

int fdfd_slab_range(fdfd_handle h, int64_t *k0, int64_t *k1) {
    if (!h || !k0 || !k1) return FDFD_EINVAL;
    Ctx *c = static_cast<Ctx *>(h);
    *k0 = c->k0;
    *k1 = c->k1;
    return FDFD_OK;
}

int fdfd_set_coeffs(fdfd_handle h, const fdfd_c128 *const sdl_e[3], const fdfd_c128 *const sdl_m[3]) {
    CHECK_H(h);
    if (!sdl_e || !sdl_m) return set_err(c, FDFD_EINVAL, "fdfd_set_coeffs: null argument");
    for (int w = 0; w < 3; ++w) {
        if (!sdl_e[w] || !sdl_m[w]) return set_err(c, FDFD_EINVAL, "fdfd_set_coeffs: null axis array");
        const cplx *pe = reinterpret_cast<const cplx *>(sdl_e[w]);
        const cplx *pm = reinterpret_cast<const cplx *>(sdl_m[w]);
        c->sdl_e[w].assign(pe, pe + c->d.N[w]);
        c->sdl_m[w].assign(pm, pm + c->d.N[w]);
        for (int64_t i = 0; i < c->d.N[w]; ++i)
            if (c->sdl_e[w][i] == cplx(0.0) || c->sdl_m[w][i] == cplx(0.0))
                return set_err(c, FDFD_EINVAL, "fdfd_set_coeffs: zero cell size");
    }
    c->have_coeffs = true;
    c->dirty = true;
    return FDFD_OK;
}

int fdfd_set_bloch(fdfd_handle h, const fdfd_c128 e_mikL[3]) {
    CHECK_H(h);
    if (!e_mikL) return set_err(c, FDFD_EINVAL, "fdfd_set_bloch: null argument");
    for (int w = 0; w < 3; ++w) {
        c->phase[w] = cplx(e_mikL[w].re, e_mikL[w].im);
        if (c->phase[w] == cplx(0.0)) return set_err(c, FDFD_EINVAL, "fdfd_set_bloch: zero phase factor");
    }
    c->dirty = true;
    return FDFD_OK;
}

int fdfd_set_omega(fdfd_handle h, fdfd_c128 omega) {
    CHECK_H(h);
    c->omega = cplx(omega.re, omega.im);
    c->have_omega = true;
    c->dirty = true;
    return FDFD_OK;
}

static bool offdiag_nonzero(const std::vector<cplx> &a, int64_t M) {
    for (int v = 0; v < 3; ++v)
        for (int u = 0; u < 3; ++u)
            if (u != v) {
                const cplx *p = a.data() + (size_t)M * (v + 3 * u);
                for (int64_t i = 0; i < M; ++i)
                    if (p[i] != cplx(0.0)) return true;
            }
    return false;
}

int fdfd_set_eps(fdfd_handle h, const fdfd_c128 *eps, int has_offdiag) {
    CHECK_H(h);
    if (!eps) return set_err(c, FDFD_EINVAL, "fdfd_set_eps: null argument");
    const int64_t M = c->d.N[0] * c->d.N[1] * (c->k1 - c->k0);
    const cplx *p = reinterpret_cast<const cplx *>(eps);
    c->eps_obj = ObjMaterial();
    try {
        c->eps_host.assign(p, p + 9 * M);
    } catch (const std::bad_alloc &) {
        return set_err(c, FDFD_ENOMEM, "fdfd_set_eps: out of host memory");
    }
    if (!has_offdiag) {
        c->eps_off = false;
        // the caller promised zeros and they are normally never read; a neighbouring z-slab WITH off-diagonal entries
        // makes this rank build (zero) arrays too, so make the promise true in the copy
        for (int v = 0; v < 3; ++v)
            for (int u = 0; u < 3; ++u)
                if (u != v) std::fill(c->eps_host.begin() + (size_t)M * (v + 3 * u), c->eps_host.begin() + (size_t)M * (v + 3 * u + 1), cplx(0.0));
    } else {
        c->eps_off = offdiag_nonzero(c->eps_host, M);
    }
    if (c->d.field_type == FDFD_FT_HH && c->eps_off)
        return set_err(c, FDFD_EINVAL, "FT_HH: Peps must be diagonal (reference model.jl:239)");
    c->have_eps = true;
    c->dirty = true;
    return FDFD_OK;
}

int fdfd_set_eps_objects(fdfd_handle h, const fdfd_matparams_desc *d) {
    CHECK_H(h);
    if (!d || !d->shapes || !d->params || d->nshape < 1 || d->nparam < 1)
        return set_err(c, FDFD_EINVAL, "fdfd_set_eps_objects: null argument / no shapes");
    if (c->d.field_type != FDFD_FT_EE)
        return set_err(c, FDFD_EINVAL, "fdfd_set_eps_objects: FT_HH needs a diagonal Peps (reference model.jl:239); smoothed objects are not");
    for (int w = 0; w < 3; ++w)
        if (d->N[w] != c->d.N[w] || !d->lprim[w] || (d->isbloch[w] != 0) != (c->d.isbloch[w] != 0) ||
            (d->boundft_is_E[w] != 0) != (c->d.boundft_is_E[w] != 0))
            return set_err(c, FDFD_EINVAL, "fdfd_set_eps_objects: grid / boundary description differs from the handle's");
    for (int o = 0; o < d->nshape; ++o)
        if (d->shapes[o].pind < 0 || d->shapes[o].pind >= d->nparam)
            return set_err(c, FDFD_EINVAL, "fdfd_set_eps_objects: bad material index");
    ObjMaterial &m = c->eps_obj;
    try {
        m.shapes.assign(d->shapes, d->shapes + d->nshape);
        const cplx *pp = reinterpret_cast<const cplx *>(d->params);
        m.params.assign(pp, pp + 9 * (size_t)d->nparam);
        for (int w = 0; w < 3; ++w) m.lprim[w].assign(d->lprim[w], d->lprim[w] + d->N[w] + 1);
    } catch (const std::bad_alloc &) {
        return set_err(c, FDFD_ENOMEM, "fdfd_set_eps_objects: out of host memory");
    }
    m.ortho = d->field_ortho_shape != 0;
    m.symmetric = true;
    for (int q = 0; q < d->nparam; ++q)
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b)
                if (m.params[9 * q + 3 * a + b] != m.params[9 * q + 3 * b + a]) m.symmetric = false;
    m.set = true;
    c->eps_host.clear();
    c->eps_host.shrink_to_fit();
    c->eps_off = d->nshape > 1;        // an interface may be cut obliquely; an all-zero result is skipped through the occupancy mask
    c->have_eps = true;
    c->dirty = true;
    return FDFD_OK;
}

int fdfd_set_mu(fdfd_handle h, const fdfd_c128 *mu) {
    CHECK_H(h);
    const int64_t M = c->d.N[0] * c->d.N[1] * (c->k1 - c->k0);
    if (!mu) {
        c->mu_host.clear();
        c->mu_host.shrink_to_fit();
        c->have_mu = false;
        c->mu_off = false;
    } else {
        const cplx *p = reinterpret_cast<const cplx *>(mu);
        try {
            c->mu_host.assign(p, p + 9 * M);
        } catch (const std::bad_alloc &) {
            return set_err(c, FDFD_ENOMEM, "fdfd_set_mu: out of host memory");
        }
        c->mu_off = offdiag_nonzero(c->mu_host, M);
        if (c->mu_off && c->d.field_type == FDFD_FT_EE) {
            c->mu_host.clear();
            c->mu_off = false;
            // reference: `Pmu \ Ce` (model.jl:236) is unsupported for non-diagonal Pmu.  (For FT_HH mu is the mass
            // parameter, A = Ce (Peps \ Cm) - w^2 Pmu, model.jl:238-240, and may be a full tensor.)
            return set_err(c, FDFD_EINVAL, "mu must be diagonal for FT_EE (reference model.jl:236)");
        }
        c->have_mu = true;
    }
    c->dirty = true;
    return FDFD_OK;
}

int fdfd_apply(fdfd_handle h, const fdfd_c128 *x, fdfd_c128 *y, int where) {
    CHECK_H(h);
    if (!x || !y) return set_err(c, FDFD_EINVAL, "fdfd_apply: null argument");
    if (where == FDFD_HOST) return apply_host(c, x, y, false);
    if (where != FDFD_DEVICE) return set_err(c, FDFD_EINVAL, "fdfd_apply: bad `where`");
    int r = apply_device(c, reinterpret_cast<const double2 *>(x), reinterpret_cast<double2 *>(y), false);
    if (r != FDFD_OK) return r;
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

int fdfd_apply_host_halos(fdfd_handle h, const fdfd_c128 *x, const fdfd_c128 *x_below_or_null, const fdfd_c128 *x_above_or_null,
                          fdfd_c128 *y, int transpose) {
    CHECK_H(h);
    if (!x || !y) return set_err(c, FDFD_EINVAL, "fdfd_apply_host_halos: null argument");
    return apply_host_halos(c, x, x_below_or_null, x_above_or_null, y, transpose != 0);
}

int fdfd_apply_transpose(fdfd_handle h, const fdfd_c128 *x, fdfd_c128 *y, int where) {
    CHECK_H(h);
    if (!x || !y) return set_err(c, FDFD_EINVAL, "fdfd_apply_transpose: null argument");
    if (where == FDFD_HOST) return apply_host(c, x, y, true);
    if (where != FDFD_DEVICE) return set_err(c, FDFD_EINVAL, "fdfd_apply_transpose: bad `where`");
    int r = apply_device(c, reinterpret_cast<const double2 *>(x), reinterpret_cast<double2 *>(y), true);
    if (r != FDFD_OK) return r;
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

int fdfd_solve(fdfd_handle h, int method, const fdfd_c128 *b, fdfd_c128 *x, int where, double rtol, int maxit,
               int check_every, int *iters, double *relres, double *hist) {
    CHECK_H(h);
    if (!b || !x || maxit < 0 || !(rtol >= 0)) return set_err(c, FDFD_EINVAL, "fdfd_solve: bad argument");
    if (method != FDFD_BICGSTAB && method != FDFD_QMR) return set_err(c, FDFD_EINVAL, "fdfd_solve: unknown method");
    if (check_every < 1) check_every = 1;
    int r, it = 0;
    double rr = 0;
    if (where == FDFD_DEVICE) {
        r = krylov_solve(c, method, reinterpret_cast<const double2 *>(b), reinterpret_cast<double2 *>(x), rtol, maxit,
                         check_every, false, &it, &rr, hist);
    } else if (where == FDFD_HOST) {
        if ((r = stage_buffers(c)) != FDFD_OK) return r;
        const size_t bytes = (size_t)c->nloc * sizeof(double2);
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, b, bytes, cudaMemcpyHostToDevice, c->stream));
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_y, x, bytes, cudaMemcpyHostToDevice, c->stream));
        r = krylov_solve(c, method, c->stage_x, c->stage_y, rtol, maxit, check_every, false, &it, &rr, hist);
        if (r == FDFD_OK || r == FDFD_ENOCONV) {
            FDFD_CUDA(c, cudaMemcpyAsync(x, c->stage_y, bytes, cudaMemcpyDeviceToHost, c->stream));
            FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
        }
    } else {
        return set_err(c, FDFD_EINVAL, "fdfd_solve: bad `where`");
    }
    if (iters) *iters = it;
    if (relres) *relres = rr;
    return r;
}

int fdfd_export_pattern(fdfd_handle h, int64_t *colptr, int64_t *rowval, fdfd_c128 *nzval, int64_t *nnz_inout) {
    CHECK_H(h);
    if (c->eps_obj.set)
        return set_err(c, FDFD_ESTATE, "fdfd_export_pattern needs eps as a host array (fdfd_set_eps); this handle's eps was rasterised on the device from objects");
    return export_pattern(c, colptr, rowval, nzval, nnz_inout);
}

// The three source / post-processing operators of the reference (model.jl:251-284) are two shapes of the same
// stencils; which curl and which material each one needs depends on the handle's formulation:
//   first-curl shape   out = alpha * q .* (C1 f + sj * j)      FT_EE: h_from_e (C1 = Ce, q = 1/mu)
//                                                              FT_HH: e_from_h (C1 = Cm, q = 1/eps)
//   second-curl shape  out = (beta * C2 (q? .* f) + gamma * j) [/ md]
//                                                              FT_EE: create_b (C2 = Cm, with q), e_from_h (no q, / md)
//                                                              FT_HH: create_b (C2 = Ce, with q), h_from_e (no q, / md)
// f is required, j optional; `where` selects host or device buffers.
namespace {
struct PostArgs {
    const fdfd_c128 *f, *j;
    fdfd_c128 *out;
    int where;
};

// stages f -> stage_x, j -> tmp (host case); returns device pointers
int post_stage(Ctx *c, const PostArgs &a, const double2 *&df, const double2 *&dj, double2 *&dout, double2 *&tmp) {
    const size_t bytes = (size_t)c->nloc * sizeof(double2);
    df = reinterpret_cast<const double2 *>(a.f);
    dj = reinterpret_cast<const double2 *>(a.j);
    dout = reinterpret_cast<double2 *>(a.out);
    tmp = nullptr;
    if (a.where == FDFD_HOST) {
        int r;
        if ((r = stage_buffers(c)) != FDFD_OK) return r;
        if (a.f) {
            FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, a.f, bytes, cudaMemcpyHostToDevice, c->stream));
            df = c->stage_x;
        }
        if (a.j) {
            FDFD_CUDA(c, cudaMalloc((void **)&tmp, bytes));
            FDFD_CUDA(c, cudaMemcpyAsync(tmp, a.j, bytes, cudaMemcpyHostToDevice, c->stream));
            dj = tmp;
        }
        dout = c->stage_y;
    } else if (a.where != FDFD_DEVICE) {
        return set_err(c, FDFD_EINVAL, "bad `where`");
    }
    return FDFD_OK;
}

int post_finish(Ctx *c, const PostArgs &a, double2 *tmp, int rc) {
    const size_t bytes = (size_t)c->nloc * sizeof(double2);
    cudaError_t e = cudaSuccess;
    if (rc == FDFD_OK && a.where == FDFD_HOST)
        e = cudaMemcpyAsync(a.out, c->stage_y, bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (tmp) cudaFree(tmp);
    if (rc != FDFD_OK) return rc;
    if (e != cudaSuccess) return set_err(c, FDFD_ECUDA, cudaGetErrorString(e));
    return FDFD_OK;
}

int first_curl_post(Ctx *c, const PostArgs &a, cplx alpha, double sj) {
    const double2 *df, *dj;
    double2 *dout, *tmp;
    int r = post_stage(c, a, df, dj, dout, tmp);
    if (r == FDFD_OK) {
        ApplyParams p;
        fill_params(c, p, df, dout, false);
        if (c->d.nranks > 1) r = halo_exchange(c, df, c->halo_lo, c->halo_hi, c->stream);
        if (r == FDFD_OK) {
            cudaError_t e = launch_curl1(p, dj, make_double2(alpha.real(), alpha.imag()), sj, c->stream);
            if (e != cudaSuccess) r = set_err(c, FDFD_ECUDA, cudaGetErrorString(e));
            c->launches += 1;
        }
    }
    return post_finish(c, a, tmp, r);
}

int second_curl_post(Ctx *c, const PostArgs &a, cplx beta, cplx gamma, bool with_q, bool divide_by_md) {
    const double2 *df, *dj;
    double2 *dout, *tmp;
    int r = post_stage(c, a, df, dj, dout, tmp);
    if (r == FDFD_OK) {
        ApplyParams p;
        fill_params(c, p, df ? df : dj, dout, false);
        if (!with_q) p.has_q = 0;
        if (df && c->d.nranks > 1) r = halo_exchange(c, df, c->halo_lo, c->halo_hi, c->stream);
        if (r == FDFD_OK) {
            cudaError_t e = launch_curl2(p, dj, make_double2(beta.real(), beta.imag()),
                                         make_double2(gamma.real(), gamma.imag()), df ? 1 : 0, c->stream,
                                         divide_by_md ? 1 : 0);
            if (e != cudaSuccess) r = set_err(c, FDFD_ECUDA, cudaGetErrorString(e));
            c->launches += 1;
        }
    }
    return post_finish(c, a, tmp, r);
}
}  // namespace

int fdfd_h_from_e(fdfd_handle h, const fdfd_c128 *e, const fdfd_c128 *jm, fdfd_c128 *hout, int where) {
    CHECK_H(h);
    if (!e || !hout) return set_err(c, FDFD_EINVAL, "fdfd_h_from_e: null argument");
    if (c->omega == cplx(0.0)) return set_err(c, FDFD_EINVAL, "fdfd_h_from_e: omega == 0");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    const cplx iw = cplx(0.0, 1.0) * c->omega;
    const PostArgs a{e, jm, hout, where};
    // h = (i/w) Pmu \ (Ce e + jm)   (model.jl:276-279)
    if (c->d.field_type == FDFD_FT_EE) return first_curl_post(c, a, cplx(0.0, 1.0) / c->omega, +1.0);
    // FT_HH handle: Ce is the second curl and mu the mass parameter: (i/w)(..)/mu = (-i w)(..)/(-w^2 mu)
    if (c->mo[0]) return set_err(c, FDFD_EINVAL, "fdfd_h_from_e: Pmu must be diagonal (reference model.jl:236,279)");
    return second_curl_post(c, a, -iw, -iw, false, true);
}

int fdfd_create_b(fdfd_handle h, const fdfd_c128 *je, const fdfd_c128 *jm, fdfd_c128 *b, int where) {
    CHECK_H(h);
    if (!je || !b) return set_err(c, FDFD_EINVAL, "fdfd_create_b: null argument");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    const cplx gm = -cplx(0.0, 1.0) * c->omega;   // the -i w j term is skipped for w == 0 (model.jl:265,270)
    const bool w0 = c->omega == cplx(0.0);
    if (c->d.field_type == FDFD_FT_EE) {
        // b = -Cm (mu^-1 jm) - i w je   (model.jl:262-265)
        const PostArgs a{jm, w0 ? nullptr : je, b, where};
        return second_curl_post(c, a, cplx(-1.0), gm, true, false);
    }
    // b = Ce (eps^-1 je) - i w jm       (model.jl:267-270)
    const PostArgs a{je, w0 ? nullptr : jm, b, where};
    return second_curl_post(c, a, cplx(1.0), gm, true, false);
}

int fdfd_e_from_h(fdfd_handle h, const fdfd_c128 *hf, const fdfd_c128 *je, fdfd_c128 *eout, int where) {
    CHECK_H(h);
    if (!hf || !eout) return set_err(c, FDFD_EINVAL, "fdfd_e_from_h: null argument");
    if (c->omega == cplx(0.0)) return set_err(c, FDFD_EINVAL, "fdfd_e_from_h: omega == 0");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    const cplx iw = cplx(0.0, 1.0) * c->omega;
    const PostArgs a{hf, je, eout, where};
    // e = (-i/w) Peps \ (Cm h - je)   (model.jl:281-284)
    if (c->d.field_type == FDFD_FT_HH) return first_curl_post(c, a, -cplx(0.0, 1.0) / c->omega, -1.0);
    // FT_EE handle: Cm is the second curl and eps the mass parameter: (-i/w)(..)/eps = (i w)(..)/(-w^2 eps)
    if (c->mo[0]) return set_err(c, FDFD_EINVAL, "fdfd_e_from_h: Peps must be diagonal (reference model.jl:239,283)");
    return second_curl_post(c, a, iw, -iw, false, true);
}

int fdfd_interp_corners(fdfd_handle h, int which, const fdfd_c128 *f, fdfd_c128 *out, int where) {
    CHECK_H(h);
    if (!f || !out) return set_err(c, FDFD_EINVAL, "fdfd_interp_corners: null argument");
    if (which != FDFD_FT_EE && which != FDFD_FT_HH) return set_err(c, FDFD_EINVAL, "ft is unsupported.");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    const size_t bytes = (size_t)c->nloc * sizeof(double2);
    const double2 *df = reinterpret_cast<const double2 *>(f);
    double2 *dout = reinterpret_cast<double2 *>(out);
    if (where == FDFD_HOST) {
        if ((r = stage_buffers(c)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaMemcpyAsync(c->stage_x, f, bytes, cudaMemcpyHostToDevice, c->stream));
        df = c->stage_x;
        dout = c->stage_y;
    }
    ApplyParams p;
    fill_params(c, p, df, dout, false);
    if (c->d.nranks > 1 && (r = halo_exchange(c, df, c->halo_lo, c->halo_hi, c->stream)) != FDFD_OK) return r;
    // the handle's own field type uses the in-average tables (shift -s1), the other field the mh tables (+s1)
    FDFD_CUDA(c, launch_interp(p, which == c->d.field_type ? 0 : 1, c->stream));
    c->launches += 1;
    if (where == FDFD_HOST) FDFD_CUDA(c, cudaMemcpyAsync(out, c->stage_y, bytes, cudaMemcpyDeviceToHost, c->stream));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

int fdfd_comm_unique_id(char id[128]) {
    std::string err;
    int r = comm_unique_id(id, err);
    if (r != FDFD_OK) set_err(nullptr, r, err);
    return r;
}

int fdfd_comm_init(fdfd_handle h, const char id[128]) {
    CHECK_H(h);
    if (c->d.nranks == 1) return FDFD_OK;
    return comm_init(c, id);
}

int fdfd_bench_apply(fdfd_handle h, const fdfd_c128 *x, fdfd_c128 *y, int warmup, int iters, int flush_l2,
                     double *ms_total, double *ms_min) {
    CHECK_H(h);
    if (!x || !y || iters < 1 || warmup < 0) return set_err(c, FDFD_EINVAL, "fdfd_bench_apply: bad argument");
    { int r0 = ensure_ready(c); if (r0 != FDFD_OK) return r0; }
    const double2 *dx = reinterpret_cast<const double2 *>(x);
    double2 *dy = reinterpret_cast<double2 *>(y);
    int r;
    if (flush_l2 && !c->flush_buf) {
        c->flush_bytes = (size_t)512 << 20;
        FDFD_CUDA(c, cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    for (int i = 0; i < warmup; ++i)
        if ((r = apply_device(c, dx, dy, false)) != FDFD_OK) return r;
    cudaEvent_t e0, e1;
    FDFD_CUDA(c, cudaEventCreate(&e0));
    FDFD_CUDA(c, cudaEventCreate(&e1));
    double total = 0, mn = 1e300;
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    if (!flush_l2) {
        // back-to-back applies, one event pair around the whole run (inputs must exceed L2)
        FDFD_CUDA(c, cudaEventRecord(e0, c->stream));
        for (int i = 0; i < iters; ++i)
            if ((r = apply_device(c, dx, dy, false)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaEventRecord(e1, c->stream));
        FDFD_CUDA(c, cudaEventSynchronize(e1));
        float ms = 0;
        FDFD_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
        total = ms;
        mn = ms / iters;
    } else {
        for (int i = 0; i < iters; ++i) {
            flush_kernel<<<148 * 8, 256, 0, c->stream>>>((float4 *)c->flush_buf, (int64_t)(c->flush_bytes / 16), (float)i);
            FDFD_CUDA(c, cudaEventRecord(e0, c->stream));
            if ((r = apply_device(c, dx, dy, false)) != FDFD_OK) return r;
            FDFD_CUDA(c, cudaEventRecord(e1, c->stream));
            FDFD_CUDA(c, cudaEventSynchronize(e1));
            float ms = 0;
            FDFD_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
            total += ms;
            if (ms < mn) mn = ms;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_total) *ms_total = total;
    if (ms_min) *ms_min = mn;
    return FDFD_OK;
}

int fdfd_bench_solve(fdfd_handle h, int method, const fdfd_c128 *b, fdfd_c128 *x, int warmup, int iters,
                     double *ms_total) {
    CHECK_H(h);
    if (!b || !x || iters < 1 || warmup < 0) return set_err(c, FDFD_EINVAL, "fdfd_bench_solve: bad argument");
    { int r0 = ensure_ready(c); if (r0 != FDFD_OK) return r0; }
    int it = 0, r;
    double rr = 0;
    const double2 *db = reinterpret_cast<const double2 *>(b);
    double2 *dx = reinterpret_cast<double2 *>(x);
    if (warmup > 0) {
        r = krylov_solve(c, method, db, dx, 0.0, warmup, 1 << 30, true, &it, &rr, nullptr);
        if (r != FDFD_OK && r != FDFD_ENOCONV) return r;
    }
    cudaEvent_t e0, e1;
    FDFD_CUDA(c, cudaEventCreate(&e0));
    FDFD_CUDA(c, cudaEventCreate(&e1));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    FDFD_CUDA(c, cudaEventRecord(e0, c->stream));
    r = krylov_solve(c, method, db, dx, 0.0, iters, 1 << 30, true, &it, &rr, nullptr);
    FDFD_CUDA(c, cudaEventRecord(e1, c->stream));
    FDFD_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    FDFD_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_total) *ms_total = ms;
    if (r == FDFD_ENOCONV) r = FDFD_OK;
    return r;
}

int fdfd_bench_halo(fdfd_handle h, const fdfd_c128 *x, int warmup, int iters, double *ms_total, uint64_t *bytes_sent) {
    CHECK_H(h);
    if (!x || iters < 1 || warmup < 0) return set_err(c, FDFD_EINVAL, "fdfd_bench_halo: bad argument");
    { int r0 = ensure_ready(c); if (r0 != FDFD_OK) return r0; }
    if (ms_total) *ms_total = 0.0;
    if (bytes_sent) *bytes_sent = 0;
    if (c->d.nranks == 1) return FDFD_OK;     // a single slab exchanges nothing
    const double2 *dx = reinterpret_cast<const double2 *>(x);
    int up, dn, r;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    if (c->comm_pending) {
        FDFD_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_halo, 0));
        c->comm_pending = false;
    }
    for (int i = 0; i < warmup; ++i)
        if ((r = halo_exchange(c, dx, c->halo_lo, c->halo_hi, c->stream)) != FDFD_OK) return r;
    cudaEvent_t e0, e1;
    FDFD_CUDA(c, cudaEventCreate(&e0));
    FDFD_CUDA(c, cudaEventCreate(&e1));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    FDFD_CUDA(c, cudaEventRecord(e0, c->stream));
    for (int i = 0; i < iters; ++i)
        if ((r = halo_exchange(c, dx, c->halo_lo, c->halo_hi, c->stream)) != FDFD_OK) {
            cudaEventDestroy(e0); cudaEventDestroy(e1);
            return r;
        }
    FDFD_CUDA(c, cudaEventRecord(e1, c->stream));
    FDFD_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0;
    FDFD_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_total) *ms_total = ms;
    if (bytes_sent) *bytes_sent = (uint64_t)((up >= 0) + (dn >= 0)) * (uint64_t)c->plane * sizeof(double2);
    return FDFD_OK;
}

int fdfd_set_shared_process(fdfd_handle h, int on) {
    if (!h) return set_err(nullptr, FDFD_EINVAL, "null handle");
    static_cast<Ctx *>(h)->shared_process = on != 0;
    return FDFD_OK;
}

int fdfd_halo_data_plane(fdfd_handle h, int *kind) {
    if (!h) return set_err(nullptr, FDFD_EINVAL, "null handle");
    Ctx *c = static_cast<Ctx *>(h);
    if (!kind) return set_err(c, FDFD_EINVAL, "null argument");
    *kind = (c->d.nranks <= 1 || !c->comm) ? 0 : (c->peer.ready ? 2 : 1);
    return FDFD_OK;
}

int fdfd_offdiag_fraction(fdfd_handle h, double *frac) {
    CHECK_H(h);
    if (!frac) return set_err(c, FDFD_EINVAL, "null argument");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    *frac = c->off_frac;
    return FDFD_OK;
}

int fdfd_mass_bytes_per_dof(fdfd_handle h, double *bytes) {
    CHECK_H(h);
    if (!bytes) return set_err(c, FDFD_EINVAL, "null argument");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    // what the operator kernel streams for the diagonal mass term: nothing (omega = 0 or an identity parameter held as
    // a scalar), complex entries, or doubles when every entry is real
    double b = (c->has_mass && c->md[0]) ? (c->md_aos_r ? 8.0 : 16.0) : 0.0;
    if (c->q[0]) b += 16.0;   // inverse middle parameter (mu^-1 for FT_EE with a mu array, eps^-1 for FT_HH)
    *bytes = b;
    return FDFD_OK;
}

int fdfd_offdiag_bytes_per_dof(fdfd_handle h, double *bytes) {
    CHECK_H(h);
    if (!bytes) return set_err(c, FDFD_EINVAL, "null argument");
    int r = ensure_ready(c);
    if (r != FDFD_OK) return r;
    // the fused row-pair shape runs when its arrays exist and the occupancy mask was built on its tiles
    const bool fused_real = c->mo_aos_r != nullptr && c->offmask_ty == 16 && c->d.kernel != FDFD_KERNEL_NAIVE;
    *bytes = !c->mo[0] ? 0.0 : fused_real ? 8.0 : c->off_sym ? 16.0 : 32.0;
    return FDFD_OK;
}

int64_t fdfd_launch_count(fdfd_handle h) { return h ? static_cast<Ctx *>(h)->launches : 0; }

int fdfd_host_alloc(void **p, uint64_t bytes) {
    if (!p) return FDFD_EINVAL;
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess) return set_err(nullptr, FDFD_ENOMEM, cudaGetErrorString(e));
    return FDFD_OK;
}
int fdfd_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? FDFD_OK : FDFD_ECUDA; }

int fdfd_dev_alloc(fdfd_handle h, void **p, uint64_t bytes) {
    CHECK_H(h);
    if (c->dev == -2) return set_err(c, FDFD_ESTATE, "host-only handle");
    if (!p) return set_err(c, FDFD_EINVAL, "null argument");
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return set_err(c, FDFD_ENOMEM, cudaGetErrorString(e));
    return FDFD_OK;
}
int fdfd_dev_free(fdfd_handle h, void *p) {
    CHECK_H(h);
    if (c->dev == -2) return set_err(c, FDFD_ESTATE, "host-only handle");
    FDFD_CUDA(c, cudaFree(p));
    return FDFD_OK;
}
int fdfd_memcpy(fdfd_handle h, void *dst, const void *src, uint64_t bytes, int dst_where, int src_where) {
    CHECK_H(h);
    if (c->dev == -2) return set_err(c, FDFD_ESTATE, "host-only handle");
    cudaMemcpyKind k = dst_where == FDFD_DEVICE
                           ? (src_where == FDFD_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                           : (src_where == FDFD_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    FDFD_CUDA(c, cudaMemcpyAsync(dst, src, bytes, k, c->stream));
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    return FDFD_OK;
}

}  // extern "C"
