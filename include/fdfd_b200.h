/*
 * fdfd_b200.h - C ABI of libfdfd_b200.so: matrix-free FDFD operator A = curl mu^-1 curl - w^2 eps
 * and its Krylov solve on NVIDIA B200 (sm_100a).
 *
 * Drop-in boundary.  The reference (MaxwellFDFD.jl) has no FFI; the seam this library replaces is
 * the VALUE returned by create_linsys / create_A (reference src/model/model.jl:209-246) - a
 * SparseMatrixCSC the user multiplies by (`mul!`, `*`) or solves with (`\`).  Each entry point
 * below cites the reference interface whose role it takes.  All functions are `extern "C"`, take
 * plain pointers and sizes, return an int status (0 = ok) and never let a C++ exception escape.
 *
 * Conventions
 *   fdfd_c128      {double re, im}: layout-identical to Julia ComplexF64 / C double _Complex /
 *                  CUDA double2.
 *   DOF order      reference model.jl:75-83.  order_cmpfirst=1: r = c + 3*(i + Nx*(j + Ny*k));
 *                  order_cmpfirst=0: r = i + Nx*(j + Ny*(k + Nz*c))   (0-based here).
 *   z-slabs        one process per GPU.  Rank p of P owns the contiguous planes
 *                  [k0,k1) given by fdfd_slab_range(); every vector / eps / mu pointer passed by
 *                  that rank covers ITS OWN planes only (Nz_local = k1-k0 in the formulas above).
 *                  With nranks == 1 the slab is the whole grid.
 *   ownership      caller owns every buffer it passes; set_* copy; the library owns its device
 *                  memory until fdfd_destroy.
 *   threading      a handle is not thread-safe; every call blocks until its results are visible
 *                  to the host (the library synchronises its own CUDA stream before returning).
 */
#ifndef FDFD_B200_H
#define FDFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } fdfd_c128;
typedef struct fdfd_ctx *fdfd_handle;

/* status codes (reference: ArgumentError at source.jl:214,230 and @error at model.jl:242,270) */
enum {
    FDFD_OK = 0,
    FDFD_EINVAL = 1,   /* bad argument / unsupported configuration */
    FDFD_ECUDA = 2,    /* CUDA runtime error (message in fdfd_last_error) */
    FDFD_ENCCL = 3,    /* NCCL error or NCCL not loadable */
    FDFD_ENOMEM = 4,   /* host or device allocation failed */
    FDFD_ENOCONV = 5,  /* solver stopped at maxit; x, iters, relres are still valid */
    FDFD_ESTATE = 6    /* call sequence error (e.g. apply before set_eps) */
};

enum { FDFD_HOST = 0, FDFD_DEVICE = 1 };            /* where x / y / b live */
enum { FDFD_BICGSTAB = 0, FDFD_QMR = 1 };            /* Krylov method (BASELINE.json north_star) */
enum { FDFD_FT_EE = 0, FDFD_FT_HH = 1 };             /* reference FieldType, model.jl:235,238 */
enum { FDFD_KERNEL_AUTO = 0, FDFD_KERNEL_NAIVE = 1, FDFD_KERNEL_TILED = 2 };

/* Problem descriptor: the fields of the reference Model that the hot path consumes
 * (model.jl:32-73): grid.N, grid.isbloch (:40), boundft (:46), order_cmpfirst (:72). */
typedef struct {
    int64_t N[3];             /* global Yee grid size Nx,Ny,Nz */
    int32_t isbloch[3];       /* 1: Bloch-periodic, 0: symmetry boundary, per axis */
    int32_t boundft_is_E[3];  /* 1: boundft[w]==EE (default), 0: HH */
    int32_t order_cmpfirst;   /* DOF layout flag, model.jl:72 */
    int32_t field_type;       /* FDFD_FT_EE: A = Cm (Pmu\Ce) - w^2 Peps ; FDFD_FT_HH: A = Ce (Peps\Cm) - w^2 Pmu */
    int32_t device;           /* CUDA device ordinal for this process; -1 = current device;
                                 -2 = host-only handle: fdfd_export_pattern works, every GPU call fails */
    int32_t rank, nranks;     /* z-slab owner / number of slabs (processes) */
    int32_t weighted_out_avg; /* 0: unweighted output average in create_paramop (default), 1: weighted */
    int32_t kernel;           /* FDFD_KERNEL_*; AUTO picks the tiled kernel when it applies */
} fdfd_desc;

const char *fdfd_version(void);

/* Lifetime.  fdfd_create replaces constructing the operator value of create_A (model.jl:225). */
int fdfd_create(fdfd_handle *out, const fdfd_desc *desc);
int fdfd_destroy(fdfd_handle h);
/* Last error message of this handle (h == NULL: last error of a failed fdfd_create in this thread). */
const char *fdfd_last_error(fdfd_handle h);

/* This rank's plane range [k0,k1) of the global z axis. */
int fdfd_slab_range(fdfd_handle h, int64_t *k0, int64_t *k1);
/* Pure host helper (no GPU needed): the partition rule itself, for planning and tests. */
int fdfd_partition(int64_t Nz, int32_t nranks, int32_t rank, int64_t *k0, int64_t *k1);

/* Pure host helper: the halo-exchange plan of rank `rank`.  *up / *dn = the rank that owns plane k1
 * (above) / plane k0-1 (below), or -1 when that side is a symmetry boundary; the global z boundary
 * wraps (rank 0 <-> rank P-1) only when wrapz != 0 (Bloch).  Message order used by the library when
 * up == dn (P == 2 with wrap): first {send my LAST plane -> up, recv my LO halo <- dn}, then
 * {send my FIRST plane -> dn, recv my HI halo <- up}. */
int fdfd_halo_plan(int32_t nranks, int32_t rank, int32_t wrapz, int32_t *up, int32_t *dn);

/* Operator inputs - exactly what create_curls / create_paramops hand to MaxwellBase
 * (model.jl:152-155,171-172):
 *   sdl_e[w], sdl_m[w]: the NON-inverted stretched cell sizes s*dl centred at E-field / H-field
 *   plane locations (create_stretched_dls, model.jl:122-139), GLOBAL length N[w], host pointers. */
int fdfd_set_coeffs(fdfd_handle h, const fdfd_c128 *const sdl_e[3], const fdfd_c128 *const sdl_m[3]);
/* e^{-i k L} per axis (create_e^{-ikL}, model.jl:91). Default (1,1,1). */
int fdfd_set_bloch(fdfd_handle h, const fdfd_c128 e_mikL[3]);
/* omega of create_A (model.jl:226); omega == 0 skips the mass term (model.jl:237). */
int fdfd_set_omega(fdfd_handle h, fdfd_c128 omega);
/* eps array of the model (model.jl:51): host pointer, Julia column-major layout
 * (Nx,Ny,Nz_local,3,3); diagonal entries at the E_v locations, off-diagonal at voxel corners.
 * has_offdiag == 0 promises every off-diagonal entry is zero (they are not read). */
int fdfd_set_eps(fdfd_handle h, const fdfd_c128 *eps, int has_offdiag);
/* mu array (model.jl:52), same layout; NULL = identity.  For FT_EE mu must be diagonal
 * (model.jl:236: `Pmu \ Ce` is only defined for diagonal Pmu) else FDFD_EINVAL. */
int fdfd_set_mu(fdfd_handle h, const fdfd_c128 *mu_or_null);

/* y = A x : the per-iteration `mul!(y, A, x)` of the reference path (SparseMatrixCSC product).
 * x, y: this rank's slab of the DOF vector, `where` = FDFD_HOST or FDFD_DEVICE. x != y. */
int fdfd_apply(fdfd_handle h, const fdfd_c128 *x, fdfd_c128 *y, int where);
/* z-slab variant of the host-buffer apply for callers that hold the WHOLE vector in host memory (one process driving
 * several GPUs, fdfd_multi_* below): x, y = this slab's planes, x_below / x_above = host pointers to plane k0-1 / k1 of
 * the global vector (the wrapped plane on a Bloch z axis; NULL on a symmetry boundary).  The halo planes ride along
 * with the H2D copies - no exchange between the GPUs, and the sub-slab pipeline of fdfd_apply(FDFD_HOST) runs per slab.
 * For the component-major layout the halo planes are three (Nx*Ny)-element pieces, component after component. */
int fdfd_apply_host_halos(fdfd_handle h, const fdfd_c128 *x, const fdfd_c128 *x_below_or_null, const fdfd_c128 *x_above_or_null,
                          fdfd_c128 *y, int transpose);
/* y = A^T x (plain transpose, needed by QMR). */
int fdfd_apply_transpose(fdfd_handle h, const fdfd_c128 *x, fdfd_c128 *y, int where);

/* x = A \ b by a Krylov method (the solve the reference leaves to the user, README.md:27-33).
 * x: in = initial guess, out = solution.  Stops when ||b - A x|| <= rtol*||b|| (recurrence
 * residual) or at maxit (-> FDFD_ENOCONV).  hist_or_null: maxit+1 doubles, relative residual per
 * iteration.  check_every: residual is read back to the host every that many iterations (>=1). */
int fdfd_solve(fdfd_handle h, int method, const fdfd_c128 *b, fdfd_c128 *x, int where,
               double rtol, int maxit, int check_every,
               int *iters, double *relres, double *hist_or_null);

/* Debug export of the assembled operator in Julia CSC form (1-based Int64 colptr/rowval,
 * SparseMatrixCSC{ComplexF64,Int64} of create_A, model.jl:236-237; pattern rule set of
 * SURVEY.md A.6).  Single-slab handles only.  Call with colptr/rowval/nzval == NULL to get nnz.
 * nnz_inout: in = capacity of rowval/nzval, out = nnz.  nzval_or_null may be NULL. */
int fdfd_export_pattern(fdfd_handle h, int64_t *colptr, int64_t *rowval, fdfd_c128 *nzval_or_null,
                        int64_t *nnz_inout);

/* Post-processing next to the solve: h = (i/w) mu^-1 (Ce e + jm)  (h_from_e, model.jl:276-279).
 * jm_or_null == NULL means jm = 0.  Works on handles of either formulation (on an FT_HH handle mu is the
 * mass parameter and must be diagonal, the `Pmu \` restriction). */
int fdfd_h_from_e(fdfd_handle h, const fdfd_c128 *e, const fdfd_c128 *jm_or_null, fdfd_c128 *hout, int where);
/* e = (-i/w) eps^-1 (Cm h - je)  (e_from_h, model.jl:281-284); diagonal eps only (`Peps \` restriction),
 * je_or_null == NULL means je = 0.  Either formulation. */
int fdfd_e_from_h(fdfd_handle h, const fdfd_c128 *hfield, const fdfd_c128 *je_or_null, fdfd_c128 *eout, int where);
/* Interpolate a solution field to the voxel corners: out = Mc_e * f (which = FDFD_FT_EE) or Mc_m * f
 * (which = FDFD_FT_HH), the operators of create_Mcs (model.jl:287-306): component w is averaged along its
 * own axis w with the weighted two-point mean of create_mean. */
int fdfd_interp_corners(fdfd_handle h, int which, const fdfd_c128 *f, fdfd_c128 *out, int where);
/* RHS (create_b, model.jl:251-274): FT_EE handle: b = -Cm (mu^-1 jm) - i w je;
 * FT_HH handle: b = Ce (eps^-1 je) - i w jm.  The -i w term is skipped for w == 0. */
int fdfd_create_b(fdfd_handle h, const fdfd_c128 *je, const fdfd_c128 *jm_or_null, fdfd_c128 *b, int where);

/* Material-parameter pipeline: calc_matparams!(mdl) of the reference (src/model/full.jl:16-70:
 * assign_param! + smooth_param! of MaxwellBase over the shapes added with add_obj!, model.jl:107-120) as one GPU
 * kernel - object assignment at the Yee locations and Kottke subpixel smoothing of every voxel an interface cuts.
 * Shapes are listed in the order they were added (later shapes lie on top); every voxel corner must be covered by at
 * least one shape (reference users add a background Box first).  out: the eps (field_type FDFD_FT_EE) or mu
 * (FDFD_FT_HH) array in the layout fdfd_set_eps / fdfd_set_mu take - Julia column-major (Nx,Ny,k1-k0,3,3): diagonal
 * entries at the field-component locations, off-diagonal entries at the voxel corners - for planes [k0,k1) of the
 * global grid (a z-slab), host or device buffer.  Errors are reported through fdfd_last_error(NULL). */
enum { FDFD_SHAPE_BOX = 0, FDFD_SHAPE_BALL = 1, FDFD_SHAPE_CYLINDER = 2 };
typedef struct {
    int32_t kind;      /* FDFD_SHAPE_* */
    int32_t axis;      /* cylinder: coordinate axis 0..2 */
    int32_t pind;      /* index of the object's material in params */
    int32_t reserved;
    double c[3];       /* centre */
    double r[3];       /* box: half-widths; ball: r[0] = radius; cylinder: r[0] = radius, r[1] = half-height */
} fdfd_shape;
typedef struct {
    int64_t N[3];
    int32_t isbloch[3];
    int32_t boundft_is_E[3];
    int32_t field_type;          /* FDFD_FT_EE: eps at the E locations; FDFD_FT_HH: mu at the H locations */
    int32_t field_ortho_shape;   /* ise-perp-shp / ish-perp-shp of the model (model.jl:65-69); 0 for full 3-D models */
    const double *lprim[3];      /* N[w]+1 primal plane positions incl. the +end ghost point (Grid ctor argument) */
    int64_t k0, k1;              /* z-planes to compute */
    int32_t nshape, nparam;
    const fdfd_shape *shapes;    /* host pointer */
    const fdfd_c128 *params;     /* host pointer, nparam x 9: row-major 3x3 tensors */
    int32_t device;              /* CUDA device ordinal; -1 = current device */
} fdfd_matparams_desc;
int fdfd_calc_matparams(const fdfd_matparams_desc *desc, fdfd_c128 *out, int where);
/* eps of an FT_EE handle straight from objects: instead of filling mdl.eps_arr on the host (calc_matparams!) and passing it
 * to fdfd_set_eps, the library rasterises and smooths this rank's z-slab on the device, directly into the operator's
 * material arrays - no (Nx,Ny,Nz,3,3) host array exists (it would be 116 GB for the 1024x1024x768 configuration).
 * desc->N / isbloch / boundft_is_E must equal the handle's; k0, k1, device and field_type are taken from the handle.
 * fdfd_export_pattern is not available on such a handle (it needs the host array). */
int fdfd_set_eps_objects(fdfd_handle h, const fdfd_matparams_desc *desc);

/* Multi-GPU plumbing: NCCL communicator over the z-slab ranks (halo send/recv + allreduce).
 * Rank 0 calls fdfd_comm_unique_id, ships the 128 bytes to the other ranks by any means
 * (torch.distributed broadcast, MPI, a file), then every rank calls fdfd_comm_init. */
int fdfd_comm_unique_id(char id[128]);
int fdfd_comm_init(fdfd_handle h, const char id[128]);

/* Measurement (CUDA events on the library's own stream; device pointers only).
 * Runs `warmup` untimed + `iters` timed applies back to back; *ms_total = time of the `iters`
 * applies; if flush_l2 != 0 a >L2-sized buffer is rewritten before every timed apply and excluded
 * from the time (events bracket each apply). */
int fdfd_bench_apply(fdfd_handle h, const fdfd_c128 *x_dev, fdfd_c128 *y_dev, int warmup, int iters,
                     int flush_l2, double *ms_total, double *ms_min);
/* Fixed-iteration Krylov run without convergence exit (iterations/s): *ms_total for `iters`. */
int fdfd_bench_solve(fdfd_handle h, int method, const fdfd_c128 *b_dev, fdfd_c128 *x_dev,
                     int warmup, int iters, double *ms_total);
/* z-slabs: the halo exchange of one apply by itself (the two boundary planes of x_dev to / from the z-neighbours),
 * `iters` times back to back; *bytes_sent = bytes this rank sends per exchange (0 on a single slab). */
int fdfd_bench_halo(fdfd_handle h, const fdfd_c128 *x_dev, int warmup, int iters, double *ms_total, uint64_t *bytes_sent);
/* Tell a slab handle that it is one of several driven by host threads of ONE process (fdfd_multi_* does this itself): the
 * Krylov loops of such handles take turns issuing an iteration's launches instead of contending for the CUDA driver. */
int fdfd_set_shared_process(fdfd_handle h, int on);
/* Which data plane moves this handle's halo planes: 0 = none (single slab), 1 = grouped ncclSend / ncclRecv, 2 = copy-engine
 * peer exchange into CUDA-IPC-mapped neighbour buffers (the default when every rank can set it up). */
int fdfd_halo_data_plane(fdfd_handle h, int *kind);
/* Fraction of the (x-y tile, z-plane) blocks of this slab that hold a non-zero off-diagonal eps entry.  The
 * tiled kernel skips the six off-diagonal streams on empty blocks (subpixel smoothing puts off-diagonal
 * entries only at material interfaces), so the bytes an apply must move are (48 + 32*frac) B/DOF. */
int fdfd_offdiag_fraction(fdfd_handle h, double *frac);
/* 1 if the off-diagonal entries of the mass tensor are pointwise symmetric (P_vu == P_uv exactly - what subpixel
 * smoothing of reciprocal media produces): the library then keeps three off-diagonal arrays instead of six and an apply
 * moves (48 + 16*frac) B/DOF.  0 otherwise (and when there are no off-diagonal entries). */
int fdfd_offdiag_symmetric(fdfd_handle h, int *symmetric);
/* Bytes per DOF the operator kernel streams for the diagonal material terms of this handle: the mass entries (16, or 8
 * when every entry -w^2 P_vv is real - lossless media at a real frequency - and they travel as doubles; 0 when w = 0
 * or the parameter is the identity) plus 16 for the inverse middle parameter when one was supplied.  With the 32 B of
 * x and y and the off-diagonal streams (fdfd_offdiag_fraction) this is the algorithmic traffic of one apply. */
int fdfd_mass_bytes_per_dof(fdfd_handle h, double *bytes);
/* Bytes per DOF the off-diagonal material streams cost on a block that holds any (fdfd_offdiag_fraction of the blocks):
 * 32 for a general tensor, 16 for a pointwise symmetric one (three arrays), 8 when it is symmetric with real entries and
 * the fused row-pair kernel streams them as doubles through its TMA ring; 0 when there are none.  Algorithmic traffic
 * of one apply = 32 + fdfd_mass_bytes_per_dof + fdfd_offdiag_bytes_per_dof * fdfd_offdiag_fraction  [B/DOF]. */
int fdfd_offdiag_bytes_per_dof(fdfd_handle h, double *bytes);
/* Number of kernels this handle has launched since creation (for bench.py's gpu_launches). */
int64_t fdfd_launch_count(fdfd_handle h);

/* Pinned host memory for callers that want full PCIe speed on FDFD_HOST calls. */
int fdfd_host_alloc(void **p, uint64_t bytes);
int fdfd_host_free(void *p);
/* Raw device memory on the handle's device, for hosts without their own CUDA allocator. */
int fdfd_dev_alloc(fdfd_handle h, void **p, uint64_t bytes);
int fdfd_dev_free(fdfd_handle h, void *p);
int fdfd_memcpy(fdfd_handle h, void *dst, const void *src, uint64_t bytes, int dst_where, int src_where);

/* ---------------------------------------------------------------------------------------------------------------------
 * ONE call, N GPUs.  The reference's seam is a single value in a single process - A = create_A(...), then
 * A * x / A \ b (model.jl:209-246).  A multi handle gives a single-process host (a Julia session) every GPU of the box
 * without an MPI launcher: it owns one z-slab handle per device and one host thread per handle, builds the NCCL
 * communicator across them, takes FULL-GRID host arrays (the layouts documented above with Nz_local = Nz) and splits
 * them into slabs itself.  fdfd_multi_apply moves each slab, together with the two neighbour planes it needs, over that
 * GPU's own PCIe link (no exchange between GPUs); fdfd_multi_solve runs the slab Krylov loops with halos and inner
 * products over NCCL.  desc->device / rank / nranks are ignored; devices_or_null == NULL uses devices 0..ngpu-1.
 * Every function blocks until all slabs are done; status = the first failing slab's (message: fdfd_multi_last_error). */
typedef struct fdfd_multi_ctx *fdfd_multi;
int fdfd_multi_create(fdfd_multi *out, const fdfd_desc *desc, int32_t ngpu, const int32_t *devices_or_null);
int fdfd_multi_destroy(fdfd_multi m);
const char *fdfd_multi_last_error(fdfd_multi m);   /* m == NULL: last error of a failed fdfd_multi_create in this thread */
int fdfd_multi_ngpu(fdfd_multi m);
/* slab handle `slab` and its plane range (for the measurement / debug entry points of the slab API) */
int fdfd_multi_slab(fdfd_multi m, int32_t slab, fdfd_handle *h, int64_t *k0, int64_t *k1);
int fdfd_multi_set_coeffs(fdfd_multi m, const fdfd_c128 *const sdl_e[3], const fdfd_c128 *const sdl_m[3]);
int fdfd_multi_set_bloch(fdfd_multi m, const fdfd_c128 e_mikL[3]);
int fdfd_multi_set_omega(fdfd_multi m, fdfd_c128 omega);
int fdfd_multi_set_eps(fdfd_multi m, const fdfd_c128 *eps, int has_offdiag);   /* (Nx,Ny,Nz,3,3), the whole grid */
int fdfd_multi_set_mu(fdfd_multi m, const fdfd_c128 *mu_or_null);
int fdfd_multi_set_eps_objects(fdfd_multi m, const fdfd_matparams_desc *desc);
int fdfd_multi_apply(fdfd_multi m, const fdfd_c128 *x, fdfd_c128 *y);           /* mul!(y, A, x), full-grid host vectors */
int fdfd_multi_apply_transpose(fdfd_multi m, const fdfd_c128 *x, fdfd_c128 *y);
int fdfd_multi_solve(fdfd_multi m, int method, const fdfd_c128 *b, fdfd_c128 *x, double rtol, int maxit, int check_every,
                     int *iters, double *relres, double *hist_or_null);          /* x = A \ b */
int fdfd_multi_create_b(fdfd_multi m, const fdfd_c128 *je, const fdfd_c128 *jm_or_null, fdfd_c128 *b);
int fdfd_multi_h_from_e(fdfd_multi m, const fdfd_c128 *e, const fdfd_c128 *jm_or_null, fdfd_c128 *hout);
int fdfd_multi_e_from_h(fdfd_multi m, const fdfd_c128 *hfield, const fdfd_c128 *je_or_null, fdfd_c128 *eout);

#ifdef __cplusplus
}
#endif
#endif /* FDFD_B200_H */
