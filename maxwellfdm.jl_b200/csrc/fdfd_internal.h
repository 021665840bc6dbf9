// Internal declarations of libfdfd_b200 (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/fdfd_b200.h"

namespace fdfd {

using cplx = std::complex<double>;

// ---------------------------------------------------------------------------------------------
// Generic two-term 1-D operator:  (T f)[i] = t0[i] f[i] + t1[i] f[(i+s) mod N]
// Every difference / averaging operator of the reference (create_∂, create_mean: SURVEY App. A.2,
// A.3) with any boundary rule (Bloch phase, symmetry zeros, "0's and 2's") is of this form, so the
// device code never branches on boundary conditions: they live in the 1-D coefficient arrays.
// ---------------------------------------------------------------------------------------------
struct AxisOp {
    std::vector<cplx> t0, t1;
    int shift = +1;
};

// Device view of the coefficient set of one operator application y = C2 q C1 x + (mass) x
struct CoefDev {
    const double2 *a0[3], *a1[3];    // first curl  (shift s1[w])
    const double2 *b0[3], *b1[3];    // second curl (shift -s1[w])
    const double2 *mi0[3], *mi1[3];  // input average of the mass operator  (shift -s1[w])
    const double2 *mo0[3], *mo1[3];  // output average of the mass operator (shift +s1[w])
    const double2 *mh0[3], *mh1[3];  // corner interpolation of the OTHER field (create_Mcs), shift +s1[w]
};

// A z-plane of a DOF vector: element (c,i,j) = p[c*cs + (j*Nx+i)*es]
struct PlaneSet {
    const double2 *base;  // plane kl=0 of this rank's slab
    int64_t pstride;      // elements between consecutive planes
    int64_t cs;           // component stride
    int32_t es;           // cell stride (3 for cmp-first, 1 otherwise)
    const double2 *lo;    // plane kl=-1   (same es; component stride cs_halo)
    const double2 *hi;    // plane kl=nzl
    int64_t cs_lo, cs_hi;
};

struct ApplyParams {
    int32_t Nx, Ny, nzl;  // local slab extent
    int32_t Nz;           // global
    int32_t kz0;          // global index of local plane 0
    int32_t s1[3];        // shift of the first curl per axis (+1 forward, -1 backward)
    int32_t wrap[3];      // 1: Bloch-periodic axis (halo = wrapped cells), 0: symmetry boundary
    int32_t cmpfirst;
    int32_t has_mass, has_off, has_q;
    CoefDev c;
    // material arrays, SoA, ghosted in z: plane index kl+1, i.e. element (kl,j,i) at
    // [(kl+1)*Nx*Ny + j*Nx + i]
    const double2 *md[3];   // -w^2 * P_vv (null when the mass parameter is the identity: md_uniform = -w^2)
    double2 md_uniform;
    const double2 *md_aos;  // cmp-first layout only: the three md arrays interleaved, element ((kl+1)*Nx*Ny + j*Nx + i)*3 + v
    const double *md_aos_r; // the same, real parts only, rows of mdr_row_pitch(Nx) doubles: [(kl+1)*Ny + j][3*i + v] - present when every diagonal mass entry is real (then md_aos is null)
    const double2 *mo[6];   // -w^2 * P_vu, order (0,1),(0,2),(1,0),(1,2),(2,0),(2,1)
    const double *mo_aos_r; // cmp-first layout, symmetric tensor with real entries only: (0,1),(0,2),(1,2) interleaved, padded by one
                            // cell / row on either side in x and y (wrapped copies on Bloch axes, zeros otherwise), rows of
                            // mdr_row_pitch(Nx + 2) doubles: [((kl+1)*(Ny+2) + j+1)][3*(i+1) + e] (fused row-pair kernel)
    const double2 *q[3];    // inverse of the middle diagonal parameter (mu^-1 for FT_EE)
    PlaneSet x;
    // z-slabs, in-kernel halo wait (opt-in): the CTAs of the first / last z-chunk spin until *halo_flag ==
    // halo_expect (written by a stream memory operation behind the NCCL exchange); null = halos already in place
    const uint32_t *halo_flag;
    uint32_t halo_expect;
    int32_t halo_sm_free;          // the exchange behind halo_flag uses no SM (copy-engine peer exchange): keep the whole grid
    const unsigned char *offmask;  // occupancy mask of the off-diagonal material (tiled kernel), or null
    int32_t offmask_ty;            // tile height the mask was built for
    const int4 *corr_list;         // sparse off-diagonals: work items (tile, ks, ke) of the correction pass
    int32_t corr_count;
    // fused Krylov inner products (tiled kernel epilogue): dot_mode 2 = accumulate (y,x) and (y,y) over the outputs
    int32_t dot_mode, dot_cap;
    double *dot_out;               // 4 doubles: Re (y,x), Im (y,x), (y,y), 0
    double *dot_partial;           // per-CTA partials (4 doubles each), capacity dot_cap CTAs
    unsigned int *dot_ticket;
    double2 *y;             // output slab, same layout as x.base
    int64_t y_pstride, y_cs;
    int32_t y_es;
    // fused epilogue/aux (unused by the naive kernel)
};

// ---------------------------------------------------------------------------------------------
// Host context
// ---------------------------------------------------------------------------------------------
struct NcclApi;  // comm.cpp

// Halo exchange over NVLink peer memory without any SM-resident collective (default data plane since round 2, peer.cpp): neighbours' halo buffers and flag words are mapped with CUDA IPC; a plane
// travels by a copy-engine cudaMemcpyAsync into the neighbour's buffer, ordering is by stream memory operations.
struct PeerHalo {
    bool ready = false;
    uint32_t *flags = nullptr;       // local, written by the neighbours: [DATA_LO, DATA_HI, FREE_UP, FREE_DN]
    double2 *up_halo_lo = nullptr;   // the up neighbour's halo_lo   (receives my last plane)
    double2 *dn_halo_hi = nullptr;   // the down neighbour's halo_hi (receives my first plane)
    uint32_t *up_flags = nullptr, *dn_flags = nullptr;
    void *mapped[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // bases to cudaIpcCloseMemHandle
    uint32_t epoch = 0;
    // Peer-direct reads (opt-in, FDFD_PEER_DIRECT; ran correctly on 2x B200 in round 2, no gain): inside BiCGSTAB the neighbours'
    // Krylov workspaces are mapped too and the apply kernel reads their boundary planes IN PLACE over NVLink (its
    // bulk copies of the first / last z-chunk address peer memory) - no halo copy at all, only a flag per direction.
    bool direct = false;                 // workspaces mapped
    bool direct_pending = false;         // halo_for's planes were announced with peer_direct_signal
    double2 *up_work = nullptr, *dn_work = nullptr;
    int64_t up_nloc = 0, dn_nloc = 0, dn_nzl = 0;
    void *work_mapped[2] = {nullptr, nullptr};
    unsigned char up_handle[64] = {}, dn_handle[64] = {};   // handles the current mappings were opened from
    uint32_t depoch = 0;
};

// eps given as objects (fdfd_set_eps_objects): the material arrays are rasterised and smoothed on the device straight
// into the operator's arrays - no host array exists
struct ObjMaterial {
    bool set = false;
    std::vector<fdfd_shape> shapes;
    std::vector<cplx> params;          // nparam x 9
    std::vector<double> lprim[3];
    int ortho = 0;
    bool symmetric = false;            // every tensor symmetric: the smoothed array is too (stored once)
};

struct Ctx {
    fdfd_desc d{};
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_copy = nullptr;   // H2D leg of the pipelined host apply
    cudaStream_t stream_d2h = nullptr;    // D2H leg
    std::vector<cudaEvent_t> ev_h2d, ev_k;
    std::string err;
    int64_t launches = 0;

    int64_t k0 = 0, k1 = 0;  // slab
    int64_t nloc = 0;        // local DOFs
    int64_t plane = 0;       // DOFs per z-plane (3*Nx*Ny)

    // host copies of inputs
    std::vector<cplx> sdl_e[3], sdl_m[3];
    cplx phase[3] = {1.0, 1.0, 1.0};
    cplx omega = 0.0;
    bool have_coeffs = false, have_eps = false, have_omega = false;
    bool eps_off = false, have_mu = false, mu_off = false;
    std::vector<cplx> eps_host;  // local slab, Julia layout (kept for re-scaling by omega and export)
    std::vector<cplx> mu_host;
    ObjMaterial eps_obj;

    // device state
    bool dirty = true;               // coefficient / material device arrays need rebuilding
    double2 *coef_dev = nullptr;     // all 1-D coefficient arrays, forward and transposed sets
    size_t coef_bytes = 0;
    CoefDev cf{}, ct{};              // forward operator / transposed operator
    double2 *mat_dev = nullptr;      // md[3], mo[6], q[3] ghosted slabs
    double2 *md_aos = nullptr;       // interleaved copy of md[3] (cmp-first layout; row-pair kernel)
    double *md_aos_r = nullptr;      // real-valued interleaved copy (instead of md_aos) when every entry is real
    double *mo_aos_r = nullptr;      // real-valued, padded, interleaved copy of the three symmetric off-diagonal arrays
    size_t mat_bytes = 0;
    const double2 *md[3]{}, *mo[6]{}, *mo_t[6]{}, *q[3]{};
    bool has_mass = false;           // omega != 0
    double2 md_uniform{};            // -w^2, used when the mass parameter was not supplied (identity)
    unsigned char *offmask = nullptr;  // device, (nzl+2) x ntiles
    int offmask_ty = 0;
    int4 *corr_list = nullptr;         // device list of (tile, ks, ke) runs for the correction pass
    int corr_count = 0;
    cudaStream_t stream_comm = nullptr; // halo exchange overlapped with interior compute
    cudaStream_t stream_bnd = nullptr;  // boundary planes (high priority), concurrent with the interior kernel
    cudaEvent_t ev_x = nullptr, ev_halo = nullptr, ev_bnd = nullptr;
    const double2 *halo_for = nullptr;  // vector whose halo planes are (being) exchanged ahead of its apply
    bool comm_pending = false;          // an NCCL op may still be running on stream_comm (ordered by ev_halo)
    bool shared_process = false;        // one of several slab handles driven by host threads of ONE process (fdfd_multi_*)
    bool halo_preloaded = false;        // halo_lo / halo_hi were filled by the caller (host halos): apply_device skips the exchange
    double off_frac = 1.0;             // fraction of (tile, plane) blocks holding off-diagonal material
    bool off_sym = false;              // off-diagonal mass entries pointwise symmetric: three arrays, three aliases
    int s1[3]{+1, +1, +1};

    // halo buffers (device): 2 receive planes, 2 send staging not needed (planes are contiguous
    // or 3 contiguous pieces)
    double2 *halo_lo = nullptr, *halo_hi = nullptr;
    uint32_t *halo_flag = nullptr;      // device word the exchange stream sets to halo_epoch when the planes have landed
    uint32_t halo_epoch = 0;

    // Krylov workspace
    double2 *work = nullptr;
    size_t work_bytes = 0;
    double *dot_partial = nullptr;   // per-CTA partials of the fused apply-epilogue dots
    unsigned int *dot_ticket = nullptr;
    int dot_cap = 0;
    double *dot_req = nullptr;       // set while apply_device runs on behalf of apply_device_dots
    double *scal = nullptr;          // device scalars
    double *partial = nullptr;       // per-block partial sums
    double *scal_host = nullptr;     // pinned
    // staging for FDFD_HOST calls
    double2 *stage_x = nullptr, *stage_y = nullptr;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;

    // NCCL
    void *comm = nullptr;
    NcclApi *nccl = nullptr;
    PeerHalo peer;
};

int set_err(Ctx *c, int code, const std::string &msg);

// Slab handles of ONE process (fdfd_multi_*): every host thread issues ~35 driver calls per Krylov iteration, and N threads
// doing so at once contend for the driver's locks (measured on 8 GPUs: 50 iterations/s against 450 with one process per
// GPU).  The threads therefore take turns: a whole iteration's launches are issued under a process-wide FIFO ticket lock -
// first come, first served, so no slab runs more than one iteration ahead of one that is waiting - and nothing that blocks
// on the device is ever called while holding it.
struct BurstTurn {
    explicit BurstTurn(const Ctx *c);
    ~BurstTurn();
    bool held;
};

#define FDFD_CUDA(c, call)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fdfd::set_err((c), FDFD_ECUDA,                                                   \
                                 std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +      \
                                     __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
    } while (0)

// coeffs.cpp ---------------------------------------------------------------------------------------
struct CoefHost {
    AxisOp a[3], b[3], mi[3], mo[3];  // first curl, second curl, in-average, out-average
    AxisOp mh[3];                     // corner interpolation of the other field type (create_Mcs)
};
// Build the coefficient set of the forward operator from the reference-level inputs.
void build_coefs(const fdfd_desc &d, const std::vector<cplx> sdl_e[3], const std::vector<cplx> sdl_m[3],
                 const cplx phase[3], CoefHost &out);
// Coefficient set of the plain transpose A^T.
void transpose_coefs(const CoefHost &in, CoefHost &out);
AxisOp make_diff(bool isfwd, const std::vector<cplx> &dinv, bool isbloch, cplx ph);
AxisOp make_mean(bool isfwd, const std::vector<cplx> *dl, const std::vector<cplx> *dlo_inv, bool isbloch,
                 cplx ph, int N);
AxisOp transpose_op(const AxisOp &t);

// pattern.cpp --------------------------------------------------------------------------------------
int export_pattern(Ctx *c, int64_t *colptr, int64_t *rowval, fdfd_c128 *nzval, int64_t *nnz_inout);

// apply_naive.cu / apply_tiled.cu ------------------------------------------------------------------
cudaError_t launch_apply_naive(const ApplyParams &p, cudaStream_t s);
// returns cudaErrorNotSupported when the tiled kernel does not cover this configuration
cudaError_t launch_apply_tiled(const ApplyParams &p, int kl_begin, int kl_end, cudaStream_t s, int *nlaunch);
bool tiled_supported(const ApplyParams &p);
// apply_rowpair.cu: second-generation K1 (persistent, warp-specialised; diagonal mass parameter) over local planes
// [kl_begin, kl_end); cudaErrorNotSupported when the configuration is outside its range
bool rowpair_supported(const ApplyParams &p, int kl_begin, int kl_end);
bool rowpair_halo_overlap_ok(const ApplyParams &p);   // whole-slab apply with the in-kernel halo wait on this kernel?
bool rowpair_fused_available(const ApplyParams &p);   // fused full-tensor shape: arrays, layout and mask fit
int64_t mdr_row_pitch(int Nx);   // doubles per row of ApplyParams::md_aos_r
cudaError_t launch_apply_rowpair(const ApplyParams &p, int kl_begin, int kl_end, cudaStream_t s);
// number of z-chunks per tile column the main kernel of launch_apply_tiled(p, 0, nzl) will use
int tiled_plan_nchunk(const ApplyParams &p);
cudaError_t tiled_build_offmask(const ApplyParams &p, unsigned char **mask, int *ty_used, double *frac, int4 **corr_list,
                                int *corr_count, cudaStream_t s, double fuse_min = 0.25);
cudaError_t launch_offdiag_correction(const ApplyParams &p, const int4 *items, int count, int ntx, int kl_begin,
                                      int kl_end, cudaStream_t s);
// first-curl only: h = scale * q .* (C1 e + jm)   (h_from_e), naive kernel
cudaError_t launch_curl1(const ApplyParams &p, const double2 *jm, double2 alpha, double sj, cudaStream_t s);
// second-curl only: y = beta * C2 (q .* h) + gamma * je   (create_b), naive kernel
cudaError_t launch_curl2(const ApplyParams &p, const double2 *je, double2 beta, double2 gamma, int has_h,
                         cudaStream_t s, int divide_by_md = 0);
// out_w = mean along w of component w (create_Mcs); which_other = 0: in-average tables (mi), 1: mh tables
cudaError_t launch_interp(const ApplyParams &p, int which_other, cudaStream_t s);

// tmap.cpp -----------------------------------------------------------------------------------------
// 3-D tiled tensor map over 8-byte elements (cuTensorMapEncodeTiled through the runtime's driver entry point);
// false when the driver does not offer it or rejects the geometry - callers then use 1-D bulk copies
struct TmaMap;
bool tmap_probe(const void *dev_ptr);   // can tensor maps be encoded on this driver at all?
bool tmap_encode_f64_3d(TmaMap *out, const void *base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                        const uint32_t box[3]);

// krylov.cu ----------------------------------------------------------------------------------------
int krylov_solve(Ctx *c, int method, const double2 *b, double2 *x, double rtol, int maxit, int check_every,
                 bool fixed_iters, int *iters, double *relres, double *hist);

// comm.cpp -----------------------------------------------------------------------------------------
void halo_neighbours(int nranks, int rank, bool wrapz, int *up, int *dn);
int comm_unique_id(char id[128], std::string &err);
int comm_init(Ctx *c, const char id[128]);
void comm_destroy(Ctx *c);
// exchange the z-halo planes of a slab vector (AoS: one contiguous plane; SoA: 3 pieces)
int halo_exchange(Ctx *c, const double2 *v, double2 *lo, double2 *hi, cudaStream_t s);
int halo_exchange_ghosted(Ctx *c, double2 *arr, cudaStream_t s);
int allreduce_sum(Ctx *c, double *dev, int count, cudaStream_t s);
// stream memory operation (no SM needed): *dev_word = value once everything enqueued before it on s has completed;
// returns FDFD_ESTATE when the driver entry point is unavailable
int stream_write_u32(Ctx *c, cudaStream_t s, uint32_t *dev_word, uint32_t value);
// stream s proceeds once (int32)(*dev_word - value) >= 0
int stream_wait_geq_u32(Ctx *c, cudaStream_t s, uint32_t *dev_word, uint32_t value);
// peer.cpp: set up / tear down the IPC mappings (collective over the neighbours), and the exchange itself
int peer_halo_init(Ctx *c);
void peer_halo_destroy(Ctx *c);
int peer_halo_exchange(Ctx *c, const double2 *first_plane, const double2 *last_plane, cudaStream_t s);
// peer-direct reads: map the neighbours' workspaces (collective; call once the workspace exists), announce that the
// boundary planes of a workspace vector are final (stream s), wait for the neighbours' announcement, and the peer
// addresses of the planes below / above this slab of workspace vector x
bool peer_direct_enabled(const Ctx *c);
int peer_direct_map(Ctx *c);
void peer_direct_unmap(Ctx *c);
int peer_direct_signal(Ctx *c, cudaStream_t s);
int peer_direct_wait(Ctx *c, cudaStream_t s);
bool peer_direct_planes(const Ctx *c, const double2 *x, const double2 **lo, const double2 **hi);
// raw byte exchange with the two z-neighbours over the NCCL communicator (setup only)
int comm_exchange_bytes(Ctx *c, const void *dev_mine, void *dev_from_up, void *dev_from_dn, size_t bytes, cudaStream_t s);
bool stream_write_u32_available();

// matparams.cu -------------------------------------------------------------------------------------
int calc_matparams(const fdfd_matparams_desc *d, fdfd_c128 *out, int where, std::string &err);

// api.cu -------------------------------------------------------------------------------------------
int ensure_ready(Ctx *c);
int apply_device(Ctx *c, const double2 *x, double2 *y, bool transpose);
// y = A x and, fused into the kernel epilogue when the tiled single-launch path is taken, (y,x) -> dot_out[0..1],
// (y,y) -> dot_out[2]; *fused tells the caller whether it still has to run its own dot kernel.
int apply_device_dots(Ctx *c, const double2 *x, double2 *y, double *dot_out, bool *fused);
int ensure_dot_buffers(Ctx *c);
// Start the halo exchange of slab vector v on the comm stream (v's boundary planes must already be enqueued on the
// main stream); the next apply_device(v) then only waits for it.  No-op unless nranks > 1 and cmp-first layout.
int halo_prefetch(Ctx *c, const double2 *v);
bool halo_prefetch_usable(const Ctx *c);
void fill_params(Ctx *c, ApplyParams &p, const double2 *x, double2 *y, bool transpose);

}  // namespace fdfd

struct fdfd_ctx : fdfd::Ctx {};
