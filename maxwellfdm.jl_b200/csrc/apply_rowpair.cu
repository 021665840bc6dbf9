// K1, second generation - the "row-pair" kernel: matrix-free y = C2 (q .* C1 x) + md .* x for a diagonal mass
// parameter (the common case: off-diagonal material, where present, is added by the correction pass on the flagged
// blocks only), any Yee arrangement (boundft), any boundary condition, both DOF layouts.
//
// Replaces the per-iteration CSC SpMV `mul!(y, A, x)` on the matrix assembled by the reference's create_A
// (src/model/model.jl:225-246); stencil per SURVEY.md App. A.4-A.6.  Same arithmetic as apply_tiled.cu (K1, first
// generation, kept for the fused full tensor and as the on-device cross-check) - what changes is who waits for whom:
//
//   * PERSISTENT grid, one CTA per SM (148 on B200); a CTA walks a static round-robin list of work items
//     (x-y tile of 30 x 2*NWC output cells, z-chunk), so nothing is lost to wave quantisation and the TMA ring of
//     the next item fills while the current one drains;
//   * WARP SPECIALISATION: warp NWC is the producer - it owns the TMA unit (1-D bulk copies global -> shared, one
//     per tile row, completing on a `full` mbarrier per ring stage) and the per-item coefficient tables; warps
//     0..NWC-1 compute.  There is NO CTA-wide barrier in the plane loop: a compute warp waits on `full`, and when it
//     is done with a stage it arrives on that stage's `empty` mbarrier (count NWC), which is all the producer waits on;
//   * a compute warp owns TWO adjacent tile rows (one thread = two cells, 6 complex outputs): the y-neighbour of one
//     of its cells is its own other cell (registers), the H values of the row below are RECOMPUTED from the ring
//     (2 of 8 H components, +15 % flops) instead of being exchanged between warps, and the x-neighbour H values move
//     by warp shuffle - the intermediate field H never touches shared memory, warps never wait for each other;
//   * outputs are staged in a private per-warp buffer and leave through TMA bulk stores issued by the warp itself;
//   * x / y / z coefficient tables: x in registers (per lane), y and z in shared memory (warp-uniform -> broadcast);
//   * thread efficiency 30/32 (x halo lanes) instead of 30*6/256: 2.3x fewer instructions per output cell, and the
//     shared-memory traffic per cell drops by ~45 %.
//
// Directions: the first curl's neighbour is at +s1[w] on axis w.  Lanes and rows are numbered along that direction
// (physical column = s1x > 0 ? lane : 31 - lane, likewise rows), the z-march runs along s1z, so every arrangement
// (default, mirrored, mixed) runs the same code with run-time index maps - no per-arrangement instantiation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cplx.cuh"
#include "fdfd_internal.h"
#include "ptx_sm100.cuh"

namespace fdfd {

namespace {

#ifdef FDFD_RP_ABLATION   // timing experiments only (make abl): FDFD_RP_DEBUG bit 8 replaces the curl arithmetic by adds
constexpr bool RP_ABL = true;
#else
constexpr bool RP_ABL = false;
#endif
constexpr int RP_TX = 32;          // data columns per tile (30 outputs)
constexpr int RP_LZMAX = 62;       // max planes per z-chunk (30 for the full-tensor variant: its tables are twice as many)

struct RowPairParams {
    ApplyParams a;
    int32_t wrapx, wrapy;
    int32_t ntx, nty, nchunk;
    int32_t kl_begin, kl_end;
    int32_t nitems;
    // tensor-map TMA path (cmp-first layout): x planes [0,nzl), the plane below / above the slab, interleaved material,
    // and y (store boxes of 30 cells x 2 rows); tmap = 0: 1-D bulk row copies instead
    TmaMap mx, mlo, mhi, mmd, mmo, my;
    int32_t tmap;
    int32_t halo_last;   // z-slabs, in-kernel halo wait (a.halo_flag): the two z-chunks that touch a neighbour's plane run last
    int32_t dbg;   // timing experiments only (FDFD_RP_DEBUG bit mask; results are wrong): 1 no material loads,
                   // 2 no y stores, 4 no x loads, 8 no arithmetic; 16, 32 (results stay right): no early stage release, no table fill ahead of the item; 64: no real-coefficient fast path; 128 (results wrong): fused shape skips the off-diagonal arithmetic
};

template <int NWC, int NST, bool MDR, bool HAS_OFF, int RPW = 2>
struct RPCfg {
    static constexpr int NR = RPW * NWC + 2;          // E rows per ring stage (RPW = tile rows per compute warp)
    static constexpr int NM = RPW * NWC;              // diagonal-material rows per ring stage (output rows only)
    static constexpr int NO = HAS_OFF ? NM + 1 : 0;   // off-diagonal-material rows (output rows + the row the last G_y needs)
    static constexpr int NT = 32 * (NWC + 1);
    static constexpr int MD0 = NR * RP_TX * 3;        // offset of the material rows inside a stage
    // real material rows: 3 doubles per cell; a tensor-map box must start on a 16-byte boundary of the global array,
    // i.e. at an even cell, so the box is 34 cells wide and starts at the even cell at or below the tile origin
    static constexpr int MDCELLS = MDR ? RP_TX + 2 : RP_TX;
    static constexpr int MDROW = MDR ? MDCELLS * 3 / 2 : RP_TX * 3;          // double2 per material row
    static constexpr int MO0 = (MD0 + NM * MDROW + 7) / 8 * 8;   // offset of the off-diagonal rows (tensor-map boxes land on 128-byte boundaries)
    static constexpr int STAGE = (MO0 + NO * MDROW + 7) / 8 * 8;             // double2 per ring stage (128-byte multiple)
    static constexpr int FPAD = 8;                    // slack below stage 0 / above the last stage (halo-lane over-reads)
    static constexpr int YW = RPW * RP_TX * 3;        // per-warp y staging (RPW rows)
    // per-item tables: a0,a1,b0,b1 (+ mi0,mi1,mo0,mo1 for the full tensor) for y (NR rows) and z (planes); the
    // full-tensor variant also keeps the x tables of the two averages here (the curl's x tables live in registers)
    static constexpr int NTAB = HAS_OFF ? 8 : 4;
    static constexpr int LZP = HAS_OFF ? 32 : RP_LZMAX + 2;
    static constexpr int LZMAX = LZP - 2;
    static constexpr int XT0 = NTAB * NR + NTAB * LZP;
    static constexpr int TABS = XT0 + (HAS_OFF ? 4 * RP_TX : 0);
    static constexpr size_t smem_bytes() {
        return (size_t)(FPAD + NST * STAGE + FPAD + NWC * YW + 2 * TABS) * sizeof(double2) + 3 * NST * 8 + NST * 4 + 128;
    }
};

// real coefficient times complex value (the full-tensor variant's material entries are real)
__device__ __forceinline__ double2 r_mul(double a, double2 z) { return make_double2(a * z.x, a * z.y); }
__device__ __forceinline__ double2 r_fma(double a, double2 z, double2 acc) {
    return make_double2(fma(a, z.x, acc.x), fma(a, z.y, acc.y));
}

// coefficient times value with a compile-time choice of coefficient kind: RC = the coefficient's imaginary part is known to
// be zero (grid cells outside the PML and away from a Bloch boundary with a complex phase) - half the multiply-adds
template <bool RC> __device__ __forceinline__ double2 k_mul(double2 a, double2 z) {
    if (RC) return make_double2(a.x * z.x, a.x * z.y);
    return c_mul(a, z);
}
template <bool RC> __device__ __forceinline__ double2 k_fma(double2 a, double2 z, double2 acc) {
    if (RC) return make_double2(fma(a.x, z.x, acc.x), fma(a.x, z.y, acc.y));
    return c_fma(a, z, acc);
}
template <bool RC> __device__ __forceinline__ double2 k_fms(double2 a, double2 z, double2 acc) {
    if (RC) return make_double2(fma(-a.x, z.x, acc.x), fma(-a.x, z.y, acc.y));
    return c_fms(a, z, acc);
}
struct RealCoef { static constexpr bool value = true; };
struct CplxCoef { static constexpr bool value = false; };

// order of the z-chunks of a tile column: bottom to top, or - when the halo planes are still in flight at launch - the
// interior chunks first and the two that read a neighbour's plane last
__device__ __forceinline__ int rp_chunk(int c, int nch, int halo_last) {
    if (!halo_last || nch < 3) return c;
    return c < nch - 2 ? c + 1 : (c == nch - 2 ? 0 : nch - 1);
}

// chunk c of nch over [kb, ke): sizes differ by at most one plane
__host__ __device__ __forceinline__ int chunk_begin(int kb, int ke, int nch, int c) {
    return kb + (int)(((int64_t)(ke - kb) * c) / nch);
}

// MDR: the diagonal mass entries are REAL (real omega and real eps_vv - every lossless dielectric): the ring carries
// 8 instead of 16 bytes per entry (40 instead of 48 B/DOF of HBM traffic) and the mass term costs half the flops.
// Tensor-map path only (the rows are 24 B per cell, which 1-D bulk copies cannot always address in 16-B units).
// HAS_OFF: the fused full 3x3 tensor (symmetric, real entries): the three off-diagonal entries of every corner travel
// through the ring as a third group of rows; G = P_off (M_in x) is formed one plane ahead, its z component in registers,
// and planes whose (tile, plane) block holds no off-diagonal material are skipped (occupancy mask, as in apply_tiled.cu).
// RPW = 1 (A/B experiment, FDFD_RP_RPW1): ONE tile row per compute warp (one cell per thread): fewer registers per thread
// and more warps per scheduler, but the H values of the row below are still recomputed (2 of 6 components, +33 % flops), the
// row above is read from the ring instead of registers, and every per-step cost is paid per row - measured slower than the
// row-pair form (RP_SHAPES below).
template <bool CMPFIRST, bool HAS_Q, bool DOT, int ARR, bool MDR, bool HAS_OFF, int NWC, int NST, int RPW = 2>
__global__ void __launch_bounds__(32 * (NWC + 1), 1) apply_rowpair_kernel(const __grid_constant__ RowPairParams tp) {
    static_assert(RPW == 2 || (RPW == 1 && !HAS_OFF), "one row per warp: diagonal mass parameter only");
    constexpr bool PAIR = RPW == 2;
    static_assert(!MDR || CMPFIRST, "real material rows exist for the cmp-first layout only");
    static_assert(!HAS_OFF || MDR, "the fused full-tensor variant is built for real, symmetric material");
    static_assert(!HAS_OFF || RPCfg<NWC, NST, MDR, HAS_OFF, RPW>::LZP <= 32, "one occupancy flag per lane");
    using C = RPCfg<NWC, NST, MDR, HAS_OFF, RPW>;
    constexpr int NR = C::NR, NM = C::NM, NT = C::NT, STAGE = C::STAGE, MD0 = C::MD0, TX = RP_TX, LZP = C::LZP;
    constexpr int NTAB = C::NTAB;
    const ApplyParams &p = tp.a;
    // direction of the first curl's neighbour per axis: compile-time for the two uniform arrangements (ARR 0: the
    // reference default boundft = (EE,EE,EE) with FT_EE; ARR 1: its mirror image), run-time values for the mixed ones
    const int SGX = ARR == 0 ? 1 : ARR == 1 ? -1 : p.s1[0];
    const int SGY = ARR == 0 ? 1 : ARR == 1 ? -1 : p.s1[1];
    const int SGZ = ARR == 0 ? 1 : ARR == 1 ? -1 : p.s1[2];

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *ring = reinterpret_cast<double2 *>(smem_raw) + C::FPAD;   // NST * STAGE
    double2 *ybase = ring + NST * STAGE + C::FPAD;                     // NWC * YW
    double2 *tabs = ybase + NWC * C::YW;                               // 2 * TABS
    uint64_t *full = reinterpret_cast<uint64_t *>(tabs + 2 * C::TABS);  // NST
    uint64_t *empty = full + NST;                                      // NST
    uint64_t *aux = empty + NST;                                       // NST  (tensor-map boxes of tiles with Bloch wrap)
    volatile int *oflag = reinterpret_cast<volatile int *>(aux + NST); // NST  (does the plane in this stage carry off-diagonal rows?)

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int Nx = p.Nx, Ny = p.Ny;
    const int64_t Nxy = (int64_t)Nx * Ny;
    const bool md_tile = p.has_mass && p.md[0] != nullptr && !(tp.dbg & 1);   // material rows travel through the ring

    // ---- one-time prologue: barriers, finite ring contents --------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWC); mbar_init(&aux[s], 1); }
        fence_barrier_init();
    }
    // tile positions no copy writes (symmetry-boundary halos, overhang) only ever meet zero coefficients or masked
    // lanes, but must hold finite values: zero everything once (later items may leave stale - finite - data there)
    for (int t = tid; t < NST * STAGE + 2 * C::FPAD; t += NT) (ring - C::FPAD)[t] = c_zero();
    __syncthreads();

    // element offsets inside a ring stage / the y staging rows
    constexpr int EC = CMPFIRST ? 1 : NR * TX;        // E rows: component stride
    constexpr int EX = CMPFIRST ? 3 : 1;              //         x-neighbour stride
    constexpr int ER = CMPFIRST ? 3 * TX : TX;        //         row stride
    constexpr int MC = CMPFIRST ? 1 : NM * TX;        // material rows: component stride (row / x strides as for E)
    constexpr int YC = CMPFIRST ? 1 : RPW * TX;       // y staging: component stride (RPW rows per warp)
    constexpr int YR = CMPFIRST ? 3 * TX : TX;

    uint32_t g = 0;   // running index of plane loads of this CTA (ring stage g % NST, phase (g / NST) & 1)

    if (wid == NWC) {
        // =============================== producer warp ======================================================
        int itc = 0;
        uint32_t aux_phase = 0;   // bit s: parity of the phase of aux[s] that the next box sent there completes
        // coefficient tables of one item (double-buffered by item parity).  They are written when the buffer's previous
        // user - the item before the previous one - has been left by every compute warp: at the first load of the item
        // (the stage just waited for was filled during the previous item), or AHEAD, behind the last load of the previous
        // item, when that item is long enough for the same argument (its load NST has been waited for) - then the global
        // loads below overlap the wait for a free stage instead of delaying the item's first planes.
        bool tabs_ready = false;
        bool halo_ok = false;
        auto fill_tables = [&](int it, int itcount) {
            int b = it;
            const int tile_x = b % tp.ntx; b /= tp.ntx;
            const int tile_y = b % tp.nty;
            const int chunk = rp_chunk(b / tp.nty, tp.nchunk, tp.halo_last);
            const int ox = tile_x * (TX - 2) - 1, oy = tile_y * C::NM - 1;
            const int kc0 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk);
            const int kc1 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk + 1);
            const int nplanes = kc1 - kc0 + 2;
            double2 *tb = tabs + (itcount & 1) * C::TABS;
            auto tab = [&](int a, int w) -> const double2 * {
                return a == 0 ? p.c.a0[w] : a == 1 ? p.c.a1[w] : a == 2 ? p.c.b0[w] : a == 3 ? p.c.b1[w]
                     : a == 4 ? p.c.mi0[w] : a == 5 ? p.c.mi1[w] : a == 6 ? p.c.mo0[w] : p.c.mo1[w];
            };
            for (int t = lane; t < NTAB * NR; t += 32) {
                const int a = t / NR, r = t % NR;
                const int j = (((oy + r) % Ny) + Ny) % Ny;
                tb[t] = tab(a, 1)[j];
            }
            for (int t = lane; t < NTAB * nplanes; t += 32) {
                const int a = t / nplanes, m = t % nplanes;
                int kg = p.kz0 + (SGZ < 0 ? kc1 - m : kc0 - 1 + m);
                kg = ((kg % p.Nz) + p.Nz) % p.Nz;
                tb[NTAB * NR + a * LZP + m] = tab(a, 2)[kg];
            }
            if (HAS_OFF) {
                const int i = (((ox + lane) % Nx) + Nx) % Nx;    // x tables of the averages, by tile column
                for (int a = 0; a < 4; ++a) tb[C::XT0 + a * TX + lane] = tab(4 + a, 0)[i];
            }
            __syncwarp();
        };
        for (int item = blockIdx.x; item < tp.nitems; item += gridDim.x, ++itc) {
            int b = item;
            const int tile_x = b % tp.ntx; b /= tp.ntx;
            const int tile_y = b % tp.nty;
            const int chunk = rp_chunk(b / tp.nty, tp.nchunk, tp.halo_last);
            const int ox = tile_x * (TX - 2) - 1, oy = tile_y * C::NM - 1;
            const int kc0 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk);
            const int kc1 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk + 1);
            const int nplanes = kc1 - kc0 + 2;   // planes kc0-1 .. kc1
            auto kof = [&](int n) { return SGZ < 0 ? kc1 - n : kc0 - 1 + n; };
            // geometry of the bulk copies of one plane
            const int xlo = max(ox, 0), xhi = min(ox + TX, Nx);
            const bool lwrap = (ox < 0) && tp.wrapx;          // cell Nx-1 -> tile column 0
            const bool rwrap = (ox + TX > Nx) && tp.wrapx;    // cell 0    -> tile column Nx-ox
            auto row_src = [&](int r) -> int {
                const int j = oy + r;
                if (j >= 0 && j < Ny) return j;
                if (tp.wrapy && (j == -1 || j == Ny)) return j < 0 ? Ny - 1 : 0;
                return -1;
            };
            int nrows = 0;
            for (int r = 0; r < NR; ++r) nrows += (row_src(r) >= 0);
            const int nmrows = max(0, min(NM, Ny - (oy + 1)));   // material rows inside the domain
            const uint32_t e_bytes = (uint32_t)nrows * (uint32_t)((xhi - xlo) + (lwrap ? 1 : 0) + (rwrap ? 1 : 0)) * 48u;
            const uint32_t m_bytes = md_tile ? (uint32_t)nmrows * (uint32_t)(xhi - xlo) * 48u : 0u;
            // per-lane copy descriptors: pair index < NE -> an E row (or (component,row)), else a material row
            constexpr int NE = CMPFIRST ? NR : 3 * NR, NMP = CMPFIRST ? NM : 3 * NM;
            constexpr int NPAIR = (NE + NMP + 31) / 32;
            int cp_j[NPAIR], cp_c[NPAIR], cp_r[NPAIR];
#pragma unroll
            for (int q = 0; q < NPAIR; ++q) {
                const int idx = lane + 32 * q;
                if (idx < NE) {
                    cp_c[q] = CMPFIRST ? 0 : idx / NR;
                    cp_r[q] = CMPFIRST ? idx : idx % NR;
                    cp_j[q] = row_src(cp_r[q]);
                } else if (idx < NE + NMP) {
                    const int m = idx - NE;
                    cp_c[q] = CMPFIRST ? 0 : m / NM;
                    cp_r[q] = CMPFIRST ? m : m % NM;
                    const int j = oy + 1 + cp_r[q];
                    cp_j[q] = (md_tile && j < Ny) ? j : -1;
                } else {
                    cp_c[q] = cp_r[q] = 0;
                    cp_j[q] = -1;
                }
            }
            // tensor-map path: does this tile need wrapped cells (Bloch boundary inside the tile)?
            bool ywrap_rows = false;
            for (int r = 0; r < NR; ++r) {
                const int j = oy + r;
                ywrap_rows |= (j < 0 || j >= Ny) && row_src(r) >= 0;
            }
            const bool has_wrap = (lwrap || rwrap || ywrap_rows) && !(tp.dbg & 4);
            const uint32_t g_item = g;
            // full tensor: occupancy flags of this tile's planes, lane n <-> march step n (at most 32 planes per item)
            int mask_n = 1;
            if (HAS_OFF && p.offmask != nullptr && lane < nplanes)
                mask_n = __ldg(&p.offmask[(int64_t)(kof(lane) + 1) * (tp.ntx * tp.nty) + tile_y * tp.ntx + tile_x]);
            // bytes of the 1-D wrap pieces of one plane: in-domain rows contribute their wrapped cells, wrapped rows
            // their whole run
            uint32_t wrap_bytes = 0;
            for (int r = 0; r < NR; ++r) {
                const int j = oy + r;
                if (row_src(r) < 0) continue;
                const bool wrapped_row = j < 0 || j >= Ny;
                wrap_bytes += ((wrapped_row ? (uint32_t)(xhi - xlo) : 0u) + (lwrap ? 1u : 0u) + (rwrap ? 1u : 0u)) * 48u;
            }
            auto finish_wrap = [&](int m) {
                const uint32_t gm = g_item + (uint32_t)m;
                const int sm = gm % NST;
                mbar_wait(&aux[sm], (aux_phase >> sm) & 1u);     // the box of load m has landed
                aux_phase ^= 1u << sm;
                const int km = kof(m);
                const double2 *src = km < 0 ? p.x.lo : km >= p.nzl ? p.x.hi : p.x.base + (int64_t)km * p.x.pstride;
                double2 *dstm = ring + sm * STAGE;
                if (lane == 0) mbar_arrive_expect_tx(&full[sm], wrap_bytes);
                __syncwarp();
                if (CMPFIRST && lane < NR) {
                    const int j = row_src(lane);
                    if (j >= 0) {
                        const bool wrapped_row = (oy + lane) < 0 || (oy + lane) >= Ny;
                        const double2 *srow = src + (int64_t)j * Nx * 3;
                        double2 *drow = dstm + lane * TX * 3;
                        if (wrapped_row)
                            bulk_g2s(drow + (xlo - ox) * 3, srow + (int64_t)xlo * 3, (uint32_t)(xhi - xlo) * 48u, &full[sm]);
                        if (lwrap) bulk_g2s(drow, srow + (int64_t)(Nx - 1) * 3, 48u, &full[sm]);
                        if (rwrap) bulk_g2s(drow + (Nx - ox) * 3, srow, 48u, &full[sm]);
                    }
                }
            };
            for (int n = 0; n < nplanes; ++n, ++g) {
                const int s = g % NST;
                if (g >= NST) mbar_wait(&empty[s], ((g / NST) - 1) & 1);
                if (n == 0 && !tabs_ready) fill_tables(item, itc);
                if (n == 0) tabs_ready = false;
                const int kk = kof(n);
                if (p.halo_flag != nullptr && !halo_ok && (kk < 0 || kk >= p.nzl)) {
                    // z-slabs, exchange overlapped with the interior chunks: the neighbours' planes are final once the flag
                    // word (written by a stream memory operation behind the exchange) has reached this apply's epoch.  The
                    // spin is bounded, generously: ranks may enter an apply seconds apart (a neighbour still building its
                    // material arrays), but a transfer that never arrives (mismatched calls across ranks) must trap after a
                    // few minutes instead of hanging the GPU for ever.
                    if (lane == 0) {
                        uint32_t spins = 0;
                        while ((int32_t)(ld_acquire_sys(p.halo_flag) - p.halo_expect) < 0) {
                            __nanosleep(spins < 1000000u ? 100 : 1000);
                            if (++spins > 300000000u) __trap();
                        }
                    }
                    __syncwarp();
                    fence_proxy_async_all();   // the planes were written through the generic proxy, the TMA unit reads them
                    halo_ok = true;
                }
                const bool want_m = md_tile && n >= 1 && n + 1 < nplanes;    // output planes only
                const bool skip_x = (tp.dbg & 4) != 0;
                double2 *dst = ring + s * STAGE;
                if (CMPFIRST && tp.tmap) {
                    // ---- tensor-map path: ONE box per array and plane; parts outside the domain read as zero
                    bool want_o = false;
                    if (HAS_OFF) {
                        // off-diagonal rows of plane kk feed G(kk): needed from the first output plane on, skipped when the
                        // (tile, plane) block holds none (occupancy mask of the 30 x 14 tiles, built by tiled_build_offmask)
                        // (the item's mask bytes were fetched once, one plane per lane - no global load per plane here)
                        want_o = n >= 1 && !(tp.dbg & 1) && (__shfl_sync(0xffffffffu, mask_n, n & 31) != 0);
                        if (lane == 0) oflag[s] = want_o ? 1 : 0;
                    }
                    const uint32_t box_bytes = (skip_x ? 0u : (uint32_t)(NR * TX * 48)) + (want_m ? (uint32_t)(NM * C::MDROW * 16) : 0u) +
                                               (want_o ? (uint32_t)(C::NO * C::MDROW * 16) : 0u);
                    uint64_t *bar = has_wrap ? &aux[s] : &full[s];
                    if (lane == 0) {
                        mbar_arrive_expect_tx(bar, box_bytes);
                        // the off-diagonal array carries one wrapped (or zero) column / row on either side: cell i sits
                        // in padded column i + 1, row j in padded row j + 1, so the forward neighbours of edge tiles come
                        // with the same box
                        if (want_o) tma_load_3d(dst + C::MO0, &tp.mmo, 3 * ((ox + 1) - ((ox + 1) & 1)), oy + 1 + (SGY > 0 ? 1 : 0), kk + 1, bar);
                        if (!skip_x) {
                            if (kk < 0) tma_load_3d(dst, &tp.mlo, 6 * ox, oy, 0, bar);
                            else if (kk >= p.nzl) tma_load_3d(dst, &tp.mhi, 6 * ox, oy, 0, bar);
                            else tma_load_3d(dst, &tp.mx, 6 * ox, oy, kk, bar);
                        }
                        if (want_m) tma_load_3d(dst + MD0, &tp.mmd, MDR ? 3 * (ox - (ox & 1)) : 6 * ox, oy + 1, kk + 1, bar);
                    }
                    // tiles with Bloch wrap: the wrapped cells / rows of the PREVIOUS plane follow as 1-D pieces once
                    // its box (which zero-filled those places) has landed - it has had a whole step to do so
                    if (has_wrap && n >= 1) finish_wrap(n - 1);
                    continue;
                }
                int64_t cs;
                const double2 *src;
                if (kk < 0) { cs = p.x.cs_lo; src = p.x.lo; }
                else if (kk >= p.nzl) { cs = p.x.cs_hi; src = p.x.hi; }
                else { cs = p.x.cs; src = p.x.base + (int64_t)kk * p.x.pstride; }
                uint64_t *bar = &full[s];
                if (lane == 0) mbar_arrive_expect_tx(bar, (skip_x ? 0u : e_bytes) + (want_m ? m_bytes : 0u));
                __syncwarp();
#pragma unroll
                for (int q = 0; q < NPAIR; ++q) {
                    const int j = cp_j[q];
                    if (j < 0) continue;
                    const int idx = lane + 32 * q;
                    if (idx < NE) {
                        if (skip_x) continue;
                        if (CMPFIRST) {
                            const double2 *srow = src + (int64_t)j * Nx * 3;
                            double2 *drow = dst + cp_r[q] * TX * 3;
                            bulk_g2s(drow + (xlo - ox) * 3, srow + (int64_t)xlo * 3, (uint32_t)(xhi - xlo) * 48u, bar);
                            if (lwrap) bulk_g2s(drow, srow + (int64_t)(Nx - 1) * 3, 48u, bar);
                            if (rwrap) bulk_g2s(drow + (Nx - ox) * 3, srow, 48u, bar);
                        } else {
                            const double2 *srow = src + (int64_t)cp_c[q] * cs + (int64_t)j * Nx;
                            double2 *drow = dst + (cp_c[q] * NR + cp_r[q]) * TX;
                            bulk_g2s(drow + (xlo - ox), srow + xlo, (uint32_t)(xhi - xlo) * 16u, bar);
                            if (lwrap) bulk_g2s(drow, srow + (Nx - 1), 16u, bar);
                            if (rwrap) bulk_g2s(drow + (Nx - ox), srow, 16u, bar);
                        }
                    } else if (want_m) {
                        const int64_t mo = (int64_t)(kk + 1) * Nxy + (int64_t)j * Nx + xlo;   // ghosted material index
                        if (CMPFIRST)
                            bulk_g2s(dst + MD0 + (cp_r[q] * TX + (xlo - ox)) * 3, p.md_aos + mo * 3,
                                     (uint32_t)(xhi - xlo) * 48u, bar);
                        else
                            bulk_g2s(dst + MD0 + (cp_c[q] * NM + cp_r[q]) * TX + (xlo - ox), p.md[cp_c[q]] + mo,
                                     (uint32_t)(xhi - xlo) * 16u, bar);
                    }
                }
            }
            if (CMPFIRST && tp.tmap && has_wrap) finish_wrap(nplanes - 1);
            if (nplanes - 1 >= NST && item + (int)gridDim.x < tp.nitems && !(tp.dbg & 32)) {
                fill_tables(item + gridDim.x, itc + 1);
                tabs_ready = true;
            }
        }
    } else {
        // =============================== compute warps ======================================================
        // logical (along the first curl's direction) -> physical tile coordinates
        const int ptx = SGX > 0 ? lane : TX - 1 - lane;
        const int rA = SGY > 0 ? RPW * wid + 1 : NR - 2 - RPW * wid;  // physical tile row of cell A; B = A + s1y (pairs)
        const int rB = rA + SGY;
        const int eA = rA * ER + ptx * EX;                // E element of cell A, component 0, inside a stage
        const int dB = SGY * ER;                          // A -> B; R (row below the pair) = A - dB, F = A + 2 dB
        const int exf = SGX * EX;                         // offset to the x-neighbour of the first curl
        const int mdo = MD0 - ER;                         // E element of an output cell -> its material element (c = 0)
        const int mdr_dB = SGY * C::MDCELLS * 3;          // real material rows (doubles): cell A -> cell B
        // y staging: row slot 0 <-> the physically lower row of the pair
        double2 *yw = ybase + wid * C::YW;
        const int ysA = (PAIR && SGY < 0) ? 1 : 0;
        const bool tmap = CMPFIRST && tp.tmap != 0;
        // tensor-map stores take a dense box: 30 output cells per row (no halo columns in the staging rows)
        const int yA = tmap ? (ysA * (TX - 2) + ptx - 1) * 3 : ysA * YR + ptx * EX;
        const int yB = tmap ? ((1 - ysA) * (TX - 2) + ptx - 1) * 3 : (1 - ysA) * YR + ptx * EX;
        const int rlo = (PAIR && SGY < 0) ? rB : rA;      // physical tile row of staging slot 0
        const bool lane_out = (lane >= 1) && (lane <= TX - 2);
        constexpr int NSTORE1D = CMPFIRST ? RPW : 3 * RPW;
        const int NSTORE = tmap ? 1 : NSTORE1D;           // lanes that issue this warp's bulk stores

        const bool no_early = (tp.dbg & 16) != 0;         // A/B timing: hold every stage to the end of its step
        double ts_re = 0.0, ts_im = 0.0, tt = 0.0;
        int itc = 0;
        for (int item = blockIdx.x; item < tp.nitems; item += gridDim.x, ++itc) {
            int b = item;
            const int tile_x = b % tp.ntx; b /= tp.ntx;
            const int tile_y = b % tp.nty;
            const int chunk = rp_chunk(b / tp.nty, tp.nchunk, tp.halo_last);
            const int ox = tile_x * (TX - 2) - 1, oy = tile_y * C::NM - 1;
            const int kc0 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk);
            const int kc1 = chunk_begin(tp.kl_begin, tp.kl_end, tp.nchunk, chunk + 1);
            const int nplanes = kc1 - kc0 + 2;
            const int kfirst = SGZ < 0 ? kc1 : kc0 - 1;   // local plane of march step 0

            const int gi = ox + ptx, gjA = oy + rA, gjB = oy + rB;
            const int mdr_o = ((rA - 1) * C::MDCELLS + ptx + (ox & 1)) * 3;   // real material rows: cell A, component 0
            const int ci = ((gi % Nx) + Nx) % Nx;
            const bool okA = lane_out && gi < Nx && gjA < Ny, okB = PAIR && lane_out && gi < Nx && gjB < Ny;
            // x tables in registers
            const double2 a0x = ldg2(&p.c.a0[0][ci]), a1x = ldg2(&p.c.a1[0][ci]);
            const double2 b0x = ldg2(&p.c.b0[0][ci]), b1x = ldg2(&p.c.b1[0][ci]);
            // q (inverse middle parameter): per-thread global loads (only FT_HH and models with a mu array carry it)
            int64_t qA = 0, qB = 0, qR = 0;
            if (HAS_Q) {
                const int gjR = gjA - SGY;
                const int cjA = ((gjA % Ny) + Ny) % Ny, cjB = ((gjB % Ny) + Ny) % Ny, cjR = ((gjR % Ny) + Ny) % Ny;
                const int64_t k1 = (int64_t)(kfirst + 1) * Nxy;
                qA = k1 + (int64_t)cjA * Nx + ci; qB = k1 + (int64_t)cjB * Nx + ci; qR = k1 + (int64_t)cjR * Nx + ci;
            }
            const int64_t dN = SGZ * Nxy;
            // store geometry of this item (lanes < NSTORE): row / component handled by this lane, in-domain columns
            const int c0 = ox + 1, c1 = min(ox + TX - 1, Nx);
            const int st_slot = CMPFIRST ? lane : lane % RPW, st_c = CMPFIRST ? 0 : lane / RPW;
            const int st_j = oy + rlo + st_slot;
            const bool st_on = lane < NSTORE && (tmap || (st_j < Ny && c1 > c0)) && !(tp.dbg & 2);
            const int st_x = 6 * (ox + 1), st_y = oy + rlo;   // tensor-map store: box origin (clipped by the hardware)
            double2 *st_dst = p.y + (int64_t)kfirst * p.y_pstride +
                              (CMPFIRST ? ((int64_t)st_j * Nx + c0) * 3 : (int64_t)st_c * p.y_cs + (int64_t)st_j * Nx + c0);
            const int64_t st_step = SGZ * p.y_pstride;
            const double2 *st_src = CMPFIRST ? yw + (st_slot * TX + 1) * 3 : yw + st_c * YC + st_slot * YR + 1;
            const uint32_t st_bytes = (uint32_t)(c1 - c0) * (CMPFIRST ? 48u : 16u);

            const double2 *tb = tabs + (itc & 1) * C::TABS;
            const double2 *ty = tb + rA;                  // y tables of row A: + a * NR; rows B / R at +- s1y
            const double2 *tz = tb + NTAB * NR;           // + a * LZP + n
            // full tensor: own cells' entries inside the off-diagonal rows of a stage (doubles; 3 per corner: (0,1), (0,2),
            // (1,2)), x tables of the two averages, and the values carried from plane to plane
            const int mo_o = ((rA - (SGY > 0 ? 1 : 0)) * C::MDCELLS + ptx + ((ox + 1) & 1)) * 3;
            const double2 *txo = tb + C::XT0 + ptx;       // + a * TX: mi0, mi1, mo0, mo1
            double2 GzA = c_zero(), GzB = c_zero();       // G_z of plane k (formed one step earlier)
            double2 E2pA = c_zero(), E2pB = c_zero(), E2pF = c_zero();   // E_z of the plane before k: rows A, B, F
            bool fCz = false;

            // first plane of the item
            int s_cur = g % NST;
            mbar_wait(&full[s_cur], (g / NST) & 1);
            const double2 *es = ring + s_cur * STAGE + eA;
            double2 EA0 = es[0], EA1 = es[EC], EA2 = es[2 * EC];
            double2 EB0 = es[dB], EB1 = es[dB + EC], EB2 = es[dB + 2 * EC];
            double2 ER1 = es[EC - dB];
            double2 HpxA = c_zero(), HpyA = c_zero(), HpxB = c_zero(), HpyB = c_zero();
            // REAL-COEFFICIENT FAST PATH: outside the PML (and away from a Bloch boundary with a complex phase) every curl
            // coefficient is real, so a coefficient times a field value costs two multiply-adds instead of four.  Decided
            // per warp: x and y tables once per item (rcxy), the z tables of the plane per step - same results, since the
            // skipped terms are exact zeros times finite values.
            // (the tables become visible with the first plane of the item: read them after that wait)
            bool rcxy;
            {
                bool re = (a0x.y == 0.0) & (a1x.y == 0.0) & (b0x.y == 0.0) & (b1x.y == 0.0);
#pragma unroll
                for (int a = 0; a < 4; ++a) re &= (ty[a * NR - SGY].y == 0.0) & (ty[a * NR].y == 0.0) & (ty[a * NR + SGY].y == 0.0);
                rcxy = __all_sync(0xffffffffu, re) && !(tp.dbg & 64);
            }

#pragma unroll 2
            for (int n = 0; n + 1 < nplanes; ++n, ++g) {
                const bool do_out = n >= 1;
                double2 qA0, qA1, qA2, qB0 = c_zero(), qB1 = c_zero(), qB2 = c_zero(), qR0, qR2;
                if (HAS_Q) {
                    const int64_t o = (int64_t)n * dN;
                    qA0 = ldg2(&p.q[0][qA + o]); qA1 = ldg2(&p.q[1][qA + o]); qA2 = ldg2(&p.q[2][qA + o]);
                    if (PAIR) { qB0 = ldg2(&p.q[0][qB + o]); qB1 = ldg2(&p.q[1][qB + o]); qB2 = ldg2(&p.q[2][qB + o]); }
                    qR0 = ldg2(&p.q[0][qR + o]); qR2 = ldg2(&p.q[2][qR + o]);
                    if (n + 4 < nplanes) {
                        prefetch_l2(&p.q[0][qA + o + 3 * dN]); prefetch_l2(&p.q[1][qA + o + 3 * dN]);
                        prefetch_l2(&p.q[2][qA + o + 3 * dN]);
                        if (PAIR) {
                            prefetch_l2(&p.q[0][qB + o + 3 * dN]); prefetch_l2(&p.q[1][qB + o + 3 * dN]);
                            prefetch_l2(&p.q[2][qB + o + 3 * dN]);
                        }
                    }
                }

                // plane k (stage es, complete since the previous step): the neighbours the first curl needs
                const double2 EA1x = es[EC + exf], EA2x = es[2 * EC + exf];
                double2 EB1x = c_zero(), EB2x = c_zero(), EF0 = c_zero(), EF2 = c_zero();
                if (PAIR) {
                    EB1x = es[dB + EC + exf]; EB2x = es[dB + 2 * EC + exf];
                    EF0 = es[2 * dB]; EF2 = es[2 * dB + 2 * EC];
                } else {   // one row per warp: the row above is not in registers
                    EB0 = es[dB]; EB2 = es[dB + 2 * EC];
                }
                const double2 ER1x = es[EC - dB + exf];
                const double2 ER0 = es[-dB], ER2 = es[2 * EC - dB];
                const double2 a0yA = ty[0], a1yA = ty[NR], a0yB = ty[SGY], a1yB = ty[NR + SGY];
                const double2 a0yR = ty[-SGY], a1yR = ty[NR - SGY];
                const double2 a0z = tz[n], a1z = tz[LZP + n];
                const bool rc1 = rcxy && __all_sync(0xffffffffu, (a0z.y == 0.0) & (a1z.y == 0.0));

                // plane k + s1z
                const int s_nxt = (s_cur + 1 == NST) ? 0 : s_cur + 1;
                mbar_wait(&full[s_nxt], ((g + 1) / NST) & 1);
                const double2 *en = ring + s_nxt * STAGE + eA;
                const double2 NA0 = en[0], NA1 = en[EC], NA2 = en[2 * EC];
                double2 NB0 = c_zero(), NB1 = c_zero(), NB2 = c_zero();
                if (PAIR) { NB0 = en[dB]; NB1 = en[dB + EC]; NB2 = en[dB + 2 * EC]; }
                const double2 NR1 = en[EC - dB];

                double2 cxA = c_zero(), cxB = c_zero(), cyA = c_zero(), cyB = c_zero();
                auto off_xy = [&]() {
                    // full tensor, x / y halves of y += Mout [P_off (Min x)] on plane k (model.jl:149-153), formed FIRST so that
                    // the stage can be released before the curl arithmetic (see EARLY RELEASE below).
                    // Corner values of plane k: G_x, G_y at this pair's cells (G_y also at the row above the pair, F -
                    // recomputed like H of the row below), G_x of the forward x-neighbour by shuffle; G_z(k) was formed
                    // one step earlier.  A warp whose entries are all zero on this plane skips the block (the
                    // occupancy mask of the producer is per tile, this test is per row pair).
                    bool fC = false;
                    double o01A = 0.0, o02A = 0.0, o12A = 0.0, o01B = 0.0, o02B = 0.0, o12B = 0.0, o01F = 0.0, o12F = 0.0;
                    if (oflag[s_cur]) {
                        const double *mo = reinterpret_cast<const double *>(ring + s_cur * STAGE + C::MO0) + mo_o;
                        o01A = mo[0]; o02A = mo[1]; o12A = mo[2];
                        o01B = mo[mdr_dB]; o02B = mo[mdr_dB + 1]; o12B = mo[mdr_dB + 2];
                        o01F = mo[2 * mdr_dB]; o12F = mo[2 * mdr_dB + 2];
                        fC = __any_sync(0xffffffffu, (o01A != 0.0) | (o02A != 0.0) | (o12A != 0.0) | (o01B != 0.0) |
                                                         (o02B != 0.0) | (o12B != 0.0) | (o01F != 0.0) | (o12F != 0.0));
                    }
                    if (fC) {
                        const double2 mi0x = txo[0], mi1x = txo[TX];
                        const double2 mi0z = tz[4 * LZP + n], mi1z = tz[5 * LZP + n];
                        const double2 AxA = c_fma(mi1x, es[-exf], c_mul(mi0x, EA0));
                        const double2 AxB = c_fma(mi1x, es[dB - exf], c_mul(mi0x, EB0));
                        const double2 AxF = c_fma(mi1x, es[2 * dB - exf], c_mul(mi0x, EF0));
                        const double2 AyA = c_fma(ty[5 * NR], ER1, c_mul(ty[4 * NR], EA1));
                        const double2 AyB = c_fma(ty[5 * NR + SGY], EA1, c_mul(ty[4 * NR + SGY], EB1));
                        const double2 AzA = c_fma(mi1z, E2pA, c_mul(mi0z, EA2));
                        const double2 AzB = c_fma(mi1z, E2pB, c_mul(mi0z, EB2));
                        const double2 AzF = c_fma(mi1z, E2pF, c_mul(mi0z, EF2));
                        const double2 GxA = r_fma(o02A, AzA, r_mul(o01A, AyA));
                        const double2 GxB = r_fma(o02B, AzB, r_mul(o01B, AyB));
                        const double2 GyA = r_fma(o12A, AzA, r_mul(o01A, AxA));
                        const double2 GyB = r_fma(o12B, AzB, r_mul(o01B, AxB));
                        const double2 GyF = r_fma(o12F, AzF, r_mul(o01F, AxF));
                        double2 GxAp, GxBp;   // G_x of the forward x-neighbour: the next lane
                        GxAp.x = __shfl_down_sync(0xffffffffu, GxA.x, 1); GxAp.y = __shfl_down_sync(0xffffffffu, GxA.y, 1);
                        GxBp.x = __shfl_down_sync(0xffffffffu, GxB.x, 1); GxBp.y = __shfl_down_sync(0xffffffffu, GxB.y, 1);
                        const double2 mo0x = txo[2 * TX], mo1x = txo[3 * TX];
                        cxA = c_fma(mo1x, GxAp, c_mul(mo0x, GxA));
                        cxB = c_fma(mo1x, GxBp, c_mul(mo0x, GxB));
                        cyA = c_fma(ty[7 * NR], GyB, c_mul(ty[6 * NR], GyA));
                        cyB = c_fma(ty[7 * NR + SGY], GyF, c_mul(ty[6 * NR + SGY], GyB));
                    }
                };
                if (HAS_OFF && !HAS_Q && do_out && !(tp.dbg & 128)) off_xy();
                // material of plane k: same ring stage as E(k)
                double2 mdA0 = p.md_uniform, mdA1 = p.md_uniform, mdA2 = p.md_uniform;
                double2 mdB0 = p.md_uniform, mdB1 = p.md_uniform, mdB2 = p.md_uniform;
                auto load_md = [&]() {
                    if (!md_tile) return;
                    if (MDR) {   // 3 doubles per cell, rows of 32 cells
                        const double *mr = reinterpret_cast<const double *>(ring + s_cur * STAGE + MD0) + mdr_o;
                        mdA0 = make_double2(mr[0], 0.0); mdA1 = make_double2(mr[1], 0.0); mdA2 = make_double2(mr[2], 0.0);
                        if (PAIR) {
                            mdB0 = make_double2(mr[mdr_dB], 0.0); mdB1 = make_double2(mr[mdr_dB + 1], 0.0);
                            mdB2 = make_double2(mr[mdr_dB + 2], 0.0);
                        }
                    } else {
                        mdA0 = es[mdo]; mdA1 = es[mdo + MC]; mdA2 = es[mdo + 2 * MC];
                        if (PAIR) { mdB0 = es[mdo + dB]; mdB1 = es[mdo + dB + MC]; mdB2 = es[mdo + dB + 2 * MC]; }
                    }
                };
                // EARLY RELEASE: everything this warp needs from stage es (plane k) now sits in registers, so the producer may
                // refill the stage while the arithmetic of this step runs - the ring is effectively one stage deeper.  Only the
                // variants with a q array hold the stage to the end (their eight extra operands per step leave no registers).
                const bool early = !HAS_Q && !no_early;
                if (early) {
                    if (do_out) load_md();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s_cur]);
                }

                double2 HxA, HyA, HzA, HxB = c_zero(), HyB = c_zero(), HzB = c_zero(), HxR, HzR;
                if (!RP_ABL || !(tp.dbg & 8)) {
                    auto curl1 = [&](auto kind) {
                        constexpr bool R = decltype(kind)::value;
                        // H(k) = C1 E :  Hx = Dy Ez - Dz Ey,  Hy = Dz Ex - Dx Ez,  Hz = Dx Ey - Dy Ex
                        HxA = k_mul<R>(a0yA, EA2); HxA = k_fma<R>(a1yA, EB2, HxA); HxA = k_fms<R>(a0z, EA1, HxA); HxA = k_fms<R>(a1z, NA1, HxA);
                        HyA = k_mul<R>(a0z, EA0);  HyA = k_fma<R>(a1z, NA0, HyA);  HyA = k_fms<R>(a0x, EA2, HyA); HyA = k_fms<R>(a1x, EA2x, HyA);
                        HzA = k_mul<R>(a0x, EA1);  HzA = k_fma<R>(a1x, EA1x, HzA); HzA = k_fms<R>(a0yA, EA0, HzA); HzA = k_fms<R>(a1yA, EB0, HzA);
                        if (PAIR) {
                            HxB = k_mul<R>(a0yB, EB2); HxB = k_fma<R>(a1yB, EF2, HxB); HxB = k_fms<R>(a0z, EB1, HxB); HxB = k_fms<R>(a1z, NB1, HxB);
                            HyB = k_mul<R>(a0z, EB0);  HyB = k_fma<R>(a1z, NB0, HyB);  HyB = k_fms<R>(a0x, EB2, HyB); HyB = k_fms<R>(a1x, EB2x, HyB);
                            HzB = k_mul<R>(a0x, EB1);  HzB = k_fma<R>(a1x, EB1x, HzB); HzB = k_fms<R>(a0yB, EB0, HzB); HzB = k_fms<R>(a1yB, EF0, HzB);
                        }
                        // the two components of the row below that this pair's second curl needs (recomputed, not exchanged)
                        HxR = k_mul<R>(a0yR, ER2); HxR = k_fma<R>(a1yR, EA2, HxR); HxR = k_fms<R>(a0z, ER1, HxR); HxR = k_fms<R>(a1z, NR1, HxR);
                        HzR = k_mul<R>(a0x, ER1);  HzR = k_fma<R>(a1x, ER1x, HzR); HzR = k_fms<R>(a0yR, ER0, HzR); HzR = k_fms<R>(a1yR, EA0, HzR);
                    };
                    if (rc1) curl1(RealCoef{}); else curl1(CplxCoef{});
                } else {   // timing experiment: touch every operand, no curl arithmetic
                    HxA = c_add(EA2, EB2); HyA = c_add(NA0, EA2x); HzA = c_add(EA1x, EB0);
                    HxB = c_add(EF2, NB1); HyB = c_add(NB0, EB2x); HzB = c_add(EB1x, EF0);
                    HxR = c_add(ER2, NR1); HzR = c_add(ER1x, ER0);
                    HxA = c_add(HxA, a0yA); HyA = c_add(HyA, a1yA); HzA = c_add(HzA, a0yB); HxB = c_add(HxB, a1yB);
                    HyB = c_add(HyB, a0yR); HzB = c_add(HzB, a1yR); HxR = c_add(HxR, a0z); HzR = c_add(HzR, a1z);
                    HxA = c_add(HxA, NA1); HxB = c_add(HxB, NA2); HyB = c_add(HyB, NB2);
                }
                if (HAS_Q) {
                    HxA = c_mul(qA0, HxA); HyA = c_mul(qA1, HyA); HzA = c_mul(qA2, HzA);
                    if (PAIR) { HxB = c_mul(qB0, HxB); HyB = c_mul(qB1, HyB); HzB = c_mul(qB2, HzB); }
                    HxR = c_mul(qR0, HxR); HzR = c_mul(qR2, HzR);
                }

                // full tensor: G_z of plane k + s1z (needs only that plane's E_x, E_y and entries): the second half of this
                // plane's z out-average and, carried over, the first half of the next plane's
                double2 GzAn = c_zero(), GzBn = c_zero();
                bool fN = false;
                auto next_gz = [&]() {
                    if ((tp.dbg & 128) || !oflag[s_nxt]) return;
                    const double *mo = reinterpret_cast<const double *>(ring + s_nxt * STAGE + C::MO0) + mo_o;
                    const double o02A = mo[1], o12A = mo[2], o02B = mo[mdr_dB + 1], o12B = mo[mdr_dB + 2];
                    fN = __any_sync(0xffffffffu, (o02A != 0.0) | (o12A != 0.0) | (o02B != 0.0) | (o12B != 0.0));
                    if (!fN) return;
                    const double2 mi0x = txo[0], mi1x = txo[TX];
                    const double2 AxA = c_fma(mi1x, en[-exf], c_mul(mi0x, NA0));
                    const double2 AxB = c_fma(mi1x, en[dB - exf], c_mul(mi0x, NB0));
                    const double2 AyA = c_fma(ty[5 * NR], NR1, c_mul(ty[4 * NR], NA1));
                    const double2 AyB = c_fma(ty[5 * NR + SGY], NA1, c_mul(ty[4 * NR + SGY], NB1));
                    GzAn = r_fma(o12A, AyA, r_mul(o02A, AxA));
                    GzBn = r_fma(o12B, AyB, r_mul(o02B, AxB));
                };
                if (HAS_OFF && !do_out) next_gz();

                if (do_out) {
                    // H_y, H_z of the x-neighbour opposite to the first curl's direction: from the previous lane
                    double2 HyAm, HzAm, HyBm = c_zero(), HzBm = c_zero();
                    HyAm.x = __shfl_up_sync(0xffffffffu, HyA.x, 1); HyAm.y = __shfl_up_sync(0xffffffffu, HyA.y, 1);
                    HzAm.x = __shfl_up_sync(0xffffffffu, HzA.x, 1); HzAm.y = __shfl_up_sync(0xffffffffu, HzA.y, 1);
                    if (PAIR) {
                        HyBm.x = __shfl_up_sync(0xffffffffu, HyB.x, 1); HyBm.y = __shfl_up_sync(0xffffffffu, HyB.y, 1);
                        HzBm.x = __shfl_up_sync(0xffffffffu, HzB.x, 1); HzBm.y = __shfl_up_sync(0xffffffffu, HzB.y, 1);
                    }
                    const double2 b0yA = ty[2 * NR], b1yA = ty[3 * NR], b0yB = ty[2 * NR + SGY], b1yB = ty[3 * NR + SGY];
                    const double2 b0z = tz[2 * LZP + n], b1z = tz[3 * LZP + n];
                    const bool rc2 = rcxy && __all_sync(0xffffffffu, (b0z.y == 0.0) & (b1z.y == 0.0));
                    if (!early) load_md();
                    double2 yxA, yyA, yzA, yxB = c_zero(), yyB = c_zero(), yzB = c_zero();
                    if (!RP_ABL || !(tp.dbg & 8)) {
                        auto curl2 = [&](auto kind) {
                            constexpr bool R = decltype(kind)::value;
                            // y = C2 H :  yx = Dy Hz - Dz Hy,  yy = Dz Hx - Dx Hz,  yz = Dx Hy - Dy Hx
                            yxA = k_mul<R>(b0yA, HzA); yxA = k_fma<R>(b1yA, HzR, yxA);  yxA = k_fms<R>(b0z, HyA, yxA); yxA = k_fms<R>(b1z, HpyA, yxA);
                            yyA = k_mul<R>(b0z, HxA);  yyA = k_fma<R>(b1z, HpxA, yyA);  yyA = k_fms<R>(b0x, HzA, yyA); yyA = k_fms<R>(b1x, HzAm, yyA);
                            yzA = k_mul<R>(b0x, HyA);  yzA = k_fma<R>(b1x, HyAm, yzA);  yzA = k_fms<R>(b0yA, HxA, yzA); yzA = k_fms<R>(b1yA, HxR, yzA);
                            if (PAIR) {
                                yxB = k_mul<R>(b0yB, HzB); yxB = k_fma<R>(b1yB, HzA, yxB);  yxB = k_fms<R>(b0z, HyB, yxB); yxB = k_fms<R>(b1z, HpyB, yxB);
                                yyB = k_mul<R>(b0z, HxB);  yyB = k_fma<R>(b1z, HpxB, yyB);  yyB = k_fms<R>(b0x, HzB, yyB); yyB = k_fms<R>(b1x, HzBm, yyB);
                                yzB = k_mul<R>(b0x, HyB);  yzB = k_fma<R>(b1x, HyBm, yzB);  yzB = k_fms<R>(b0yB, HxB, yzB); yzB = k_fms<R>(b1yB, HxA, yzB);
                            }
                        };
                        if (rc2) curl2(RealCoef{}); else curl2(CplxCoef{});
                        if (p.has_mass) {
                            if (MDR && md_tile) {   // real coefficient: two fused multiply-adds per component
                                yxA.x = fma(mdA0.x, EA0.x, yxA.x); yxA.y = fma(mdA0.x, EA0.y, yxA.y);
                                yyA.x = fma(mdA1.x, EA1.x, yyA.x); yyA.y = fma(mdA1.x, EA1.y, yyA.y);
                                yzA.x = fma(mdA2.x, EA2.x, yzA.x); yzA.y = fma(mdA2.x, EA2.y, yzA.y);
                                if (PAIR) {
                                    yxB.x = fma(mdB0.x, EB0.x, yxB.x); yxB.y = fma(mdB0.x, EB0.y, yxB.y);
                                    yyB.x = fma(mdB1.x, EB1.x, yyB.x); yyB.y = fma(mdB1.x, EB1.y, yyB.y);
                                    yzB.x = fma(mdB2.x, EB2.x, yzB.x); yzB.y = fma(mdB2.x, EB2.y, yzB.y);
                                }
                            } else {
                                yxA = c_fma(mdA0, EA0, yxA); yyA = c_fma(mdA1, EA1, yyA); yzA = c_fma(mdA2, EA2, yzA);
                                if (PAIR) { yxB = c_fma(mdB0, EB0, yxB); yyB = c_fma(mdB1, EB1, yyB); yzB = c_fma(mdB2, EB2, yzB); }
                            }
                        }
                    } else {
                        yxA = c_add(c_add(HzA, HzR), c_add(HpyA, mdA0)); yyA = c_add(c_add(HxA, HpxA), c_add(HzAm, mdA1));
                        yzA = c_add(c_add(HyA, HyAm), c_add(HxR, mdA2)); yxB = c_add(c_add(HzB, HyB), c_add(HpyB, mdB0));
                        yyB = c_add(c_add(HxB, HpxB), c_add(HzBm, mdB1)); yzB = c_add(c_add(HyBm, b0yA), c_add(b1yB, mdB2));
                        yxA = c_add(yxA, b0z); yyA = c_add(yyA, b1z); yzA = c_add(yzA, b1yA); yxB = c_add(yxB, b0yB);
                    }
                    if (HAS_OFF) {
                        // ---- off-diagonal part of the mass operator: y += Mout [P_off (Min x)] (model.jl:149-153); the x / y
                        // halves were formed at the top of the step (off_xy), the z half needs G_z of both planes
                        if (HAS_Q) off_xy();   // (stage still held: no early release with a q array)
                        yxA = c_add(yxA, cxA); yxB = c_add(yxB, cxB); yyA = c_add(yyA, cyA); yyB = c_add(yyB, cyB);
                        next_gz();
                        if (fCz || fN) {
                            const double2 mo0z = tz[6 * LZP + n], mo1z = tz[7 * LZP + n];
                            yzA = c_fma(mo0z, GzA, yzA); yzA = c_fma(mo1z, GzAn, yzA);
                            yzB = c_fma(mo0z, GzB, yzB); yzB = c_fma(mo1z, GzBn, yzB);
                        }
                    }
                    if (DOT) {   // fused Krylov inner products: t = y, s = x (own cells, in registers)
                        if (okA) {
                            ts_re += yxA.x * EA0.x + yxA.y * EA0.y + yyA.x * EA1.x + yyA.y * EA1.y + yzA.x * EA2.x + yzA.y * EA2.y;
                            ts_im += yxA.x * EA0.y - yxA.y * EA0.x + yyA.x * EA1.y - yyA.y * EA1.x + yzA.x * EA2.y - yzA.y * EA2.x;
                            tt += yxA.x * yxA.x + yxA.y * yxA.y + yyA.x * yyA.x + yyA.y * yyA.y + yzA.x * yzA.x + yzA.y * yzA.y;
                        }
                        if (okB) {
                            ts_re += yxB.x * EB0.x + yxB.y * EB0.y + yyB.x * EB1.x + yyB.y * EB1.y + yzB.x * EB2.x + yzB.y * EB2.y;
                            ts_im += yxB.x * EB0.y - yxB.y * EB0.x + yyB.x * EB1.y - yyB.y * EB1.x + yzB.x * EB2.y - yzB.y * EB2.x;
                            tt += yxB.x * yxB.x + yxB.y * yxB.y + yyB.x * yyB.x + yyB.y * yyB.y + yzB.x * yzB.x + yzB.y * yzB.y;
                        }
                    }
                    // the copy engine must have finished READING the staging rows of the previous plane
                    if (lane < NSTORE) bulk_wait_read0();
                    __syncwarp();
                    if (!tmap || lane_out) {
                        yw[yA] = yxA; yw[yA + YC] = yyA; yw[yA + 2 * YC] = yzA;
                        if (PAIR) { yw[yB] = yxB; yw[yB + YC] = yyB; yw[yB + 2 * YC] = yzB; }
                    }
                    fence_proxy_async();   // generic-proxy writes -> visible to the bulk-copy (async) proxy
                }
                // a stage held to the end (see EARLY RELEASE) is released here; the same warp-wide synchronisation orders the
                // staging writes above before the bulk store below
                __syncwarp();
                if (lane == 0) {
                    if (!early) mbar_arrive(&empty[s_cur]);
                    if (n + 2 == nplanes) mbar_arrive(&empty[s_nxt]);   // last step: plane k + s1z is not revisited
                }
                if (do_out && st_on) {
                    if (tmap) tma_store_3d(&tp.my, st_x, st_y, kfirst + SGZ * n, yw);
                    else bulk_s2g(st_dst + (int64_t)n * st_step, st_src, st_bytes);
                    bulk_commit();
                }
                if (HAS_OFF) {
                    GzA = GzAn; GzB = GzBn; fCz = fN;
                    E2pA = EA2; E2pB = EB2; E2pF = EF2;
                }
                HpxA = HxA; HpyA = HyA;
                if (PAIR) { HpxB = HxB; HpyB = HyB; }
                EA0 = NA0; EA1 = NA1; EA2 = NA2;
                if (PAIR) { EB0 = NB0; EB1 = NB1; EB2 = NB2; }
                ER1 = NR1;
                es = en;
                s_cur = s_nxt;
            }
            ++g;   // the last plane of the item was consumed as `en` of the final step
        }
        if (lane < NSTORE) bulk_wait0();   // outstanding bulk stores complete before the CTA exits

        if (DOT) {
            // warp partials -> shared scratch (the y staging of this warp is idle now)
            double v3[3] = {ts_re, ts_im, tt};
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v3[q] += __shfl_xor_sync(0xffffffffu, v3[q], o);
            }
            if (lane == 0) {
                double *sc = reinterpret_cast<double *>(yw);
                sc[0] = v3[0]; sc[1] = v3[1]; sc[2] = v3[2];
            }
        }
    }
    if (DOT) {
        // one partial per CTA; the last CTA (atomic ticket) adds the partials in a fixed order -> deterministic result
        __syncthreads();
        __shared__ bool last_cta;
        if (tid == 0) {
            double a = 0, b2 = 0, c3 = 0;
            for (int w = 0; w < NWC; ++w) {
                const double *sc = reinterpret_cast<const double *>(ybase + w * C::YW);
                a += sc[0]; b2 += sc[1]; c3 += sc[2];
            }
            double *pp = p.dot_partial + (size_t)blockIdx.x * 4;
            pp[0] = a; pp[1] = b2; pp[2] = c3;
            __threadfence();
            last_cta = atomicAdd(p.dot_ticket, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (last_cta) {
            __threadfence();
            double s3[3] = {0.0, 0.0, 0.0};
            for (int bI = tid; bI < (int)gridDim.x; bI += NT) {
                const double *pp = p.dot_partial + (size_t)bI * 4;
                s3[0] += __ldcg(pp); s3[1] += __ldcg(pp + 1); s3[2] += __ldcg(pp + 2);
            }
            double *sc = reinterpret_cast<double *>(tabs);   // tables are dead now
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s3[q] += __shfl_xor_sync(0xffffffffu, s3[q], o);
                if (lane == 0) sc[wid * 3 + q] = s3[q];
            }
            __syncthreads();
            if (tid == 0) {
                double a = 0, b2 = 0, c3 = 0;
                for (int w = 0; w < NWC + 1; ++w) { a += sc[w * 3]; b2 += sc[w * 3 + 1]; c3 += sc[w * 3 + 2]; }
                p.dot_out[0] = a; p.dot_out[1] = b2; p.dot_out[2] = c3; p.dot_out[3] = 0.0;
                *p.dot_ticket = 0;
            }
        }
    }
}

// real material rows: 3 doubles per cell, row pitch padded to a whole number of 16-byte units (tensor-map strides)
}  // namespace
int64_t mdr_row_pitch(int Nx) { return ((int64_t)3 * Nx + 1) & ~(int64_t)1; }
namespace {

int sm_count() {
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!n[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        n[dev] = v;
    }
    return n[dev];
}

template <bool CMPFIRST, bool HAS_Q, bool DOT, int ARR, bool MDR, bool HAS_OFF, int NWC, int NST, int RPW>
cudaError_t launch_rp(const RowPairParams &tp, int grid, cudaStream_t s) {
    if constexpr ((MDR && !CMPFIRST) || (RPW == 1 && (HAS_Q || HAS_OFF || !CMPFIRST))) {
        return cudaErrorInvalidConfiguration;
    } else {
        auto kern = apply_rowpair_kernel<CMPFIRST, HAS_Q, DOT, ARR, MDR, HAS_OFF, NWC, NST, RPW>;
        const size_t smem = RPCfg<NWC, NST, MDR, HAS_OFF, RPW>::smem_bytes();
        static bool attr_set[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
        kern<<<grid, 32 * (NWC + 1), smem, s>>>(tp);
        return cudaGetLastError();
    }
}

// Compiled shapes.  Warps are spread over the four SM sub-partitions, so the register budget per thread steps with
// ceil(warps / 4): 8 warps (7 compute + producer) may use 255 registers.  Shape 0: complex material rows, 4 ring
// stages of 46 KB; shape 1: real material rows (MDR), 5 stages of 35 KB; shape 2: the fused full tensor (real diagonal
// and off-diagonal rows), 4 stages of 47 KB.
struct RpShape { int nwc, nst; bool mdr, off; int rpw; };
// [3], [4]: A/B timing only - the real-row shape on a 4-stage ring (FDFD_RP_NST4: same speed, the ring depth is not what
// limits the kernel) and ONE tile row per compute warp (FDFD_RP_RPW1: 10 compute warps at 166 registers - measured 108 vs
// 123 GDOF/s on C2, 121 vs 127 on C4; with 14 warps the kernel spills at 128 registers: 70 GDOF/s - so two rows per warp stay)
constexpr RpShape RP_SHAPES[] = {{7, 4, false, false, 2}, {7, 5, true, false, 2}, {7, 4, true, true, 2}, {7, 4, true, false, 2}, {10, 6, true, false, 1}};
constexpr int RP_NSHAPES = sizeof(RP_SHAPES) / sizeof(RP_SHAPES[0]);

}  // namespace

// Plan: number of z-chunks per tile column so that the per-CTA sum of (planes + 1 extra first-curl step) is smallest.
static int rp_pick_nchunk(int ncols, int nplanes, int nsm, int nst, int lzmax, double *cost_out) {
    const int lzmin = std::max(1, nst - 2);   // an item must span at least NST plane loads (table double-buffering)
    int best = 0;
    double best_cost = 1e300;
    const int nch_min = (nplanes + lzmax - 1) / lzmax;
    for (int nch = std::max(1, nch_min); nch <= nplanes; ++nch) {
        const int lz_lo = nplanes / nch, lz_hi = (nplanes + nch - 1) / nch;
        if (lz_lo < lzmin) break;
        if (lz_hi > lzmax) continue;
        const long items = (long)ncols * nch;
        const long per_cta = (items + nsm - 1) / nsm;
        const double cost = (double)per_cta * (0.5 * (lz_lo + lz_hi) + 1.6);
        if (cost < best_cost) { best_cost = cost; best = nch; }
    }
    if (cost_out) *cost_out = best_cost;
    return best;
}

static int env_int(const char *name) {
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
}

static bool want_tmap() {
    static const bool v = [] { const char *e = getenv("FDFD_RP_TMAP"); return !e || atoi(e) != 0; }();
    return v;
}

static int rp_lzmax(int shape) { return RP_SHAPES[shape].off ? RPCfg<7, 4, true, true, 2>::LZMAX : RP_LZMAX; }

// can the fused full-tensor shape run p?  (cmp-first layout, tensor maps, real diagonal and real symmetric off-diagonal
// rows built by the handle, the occupancy mask - if there is one - on this shape's 30 x 14 tiles)
static bool rp_fused_ok(const ApplyParams &p) {
    return p.cmpfirst && p.has_mass && p.has_off && p.md[0] != nullptr && p.md_aos_r != nullptr && p.mo_aos_r != nullptr &&
           (p.offmask == nullptr || p.offmask_ty == 16) && want_tmap();
}

// shape for this launch: real material rows when the handle built them (tensor-map path, cmp-first layout)
// SMs left to the exchange while an apply with the in-kernel halo wait runs (its CTAs own a whole SM each): none when the
// planes travel by copy engine (peer exchange), 8 for the NCCL send / recv kernels (they did not progress on 4)
static int rp_sm_reserve(const ApplyParams &p) {
    static const int v = [] { const char *e = getenv("FDFD_HALO_SM_RESERVE"); return e ? std::max(0, atoi(e)) : -1; }();
    return v >= 0 ? v : (p.halo_sm_free ? 0 : 8);
}
static int rp_grid_cap(const ApplyParams &p) {
    const int nsm = sm_count();
    return p.halo_flag != nullptr ? std::max(1, nsm - rp_sm_reserve(p)) : nsm;
}

static int rp_pick_shape(const ApplyParams &p, int kl_begin, int kl_end, int *nchunk) {
    static const int want_nch = env_int("FDFD_RP_NCHUNK");
    const int n = kl_end - kl_begin;
    const bool mdr = p.cmpfirst && p.has_mass && p.md[0] != nullptr && p.md_aos_r != nullptr && want_tmap();
    static const bool nst4 = getenv("FDFD_RP_NST4") != nullptr;   // A/B timing: real-row shape with a 4-stage ring
    static const bool rpw1 = [] { const char *e = getenv("FDFD_RP_RPW1"); return e && atoi(e) != 0; }();   // A/B timing
    int i = mdr ? (nst4 ? 3 : (rpw1 && !p.has_q ? 4 : 1)) : 0;
    if (p.has_off && p.has_mass) {
        if (!rp_fused_ok(p)) return -1;
        i = 2;
    }
    const int nwc = RP_SHAPES[i].nwc, nst = RP_SHAPES[i].nst, lzmax = rp_lzmax(i);
    const int nrow = RP_SHAPES[i].rpw * nwc;
    const int ntx = (p.Nx + RP_TX - 3) / (RP_TX - 2), nty = (p.Ny + nrow - 1) / nrow;
    int nch = rp_pick_nchunk(ntx * nty, n, rp_grid_cap(p), nst, lzmax, nullptr);
    if (want_nch >= 1 && n / want_nch >= std::max(1, nst - 2) && (n + want_nch - 1) / want_nch <= lzmax) nch = want_nch;
    if (nch < 1) return -1;
    *nchunk = nch;
    return i;
}

bool rowpair_fused_available(const ApplyParams &p) { return rp_fused_ok(p); }

bool rowpair_supported(const ApplyParams &p, int kl_begin, int kl_end) {
    for (int w = 0; w < 3; ++w)
        if (p.s1[w] != 1 && p.s1[w] != -1) return false;
    int nch = 0;
    if (!(kl_end > kl_begin && rp_pick_shape(p, kl_begin, kl_end, &nch) >= 0)) return false;
    // in-kernel halo wait: whole slab in one launch, tensor-map path, and interior chunks to hide the exchange behind
    if (p.halo_flag != nullptr && !(kl_begin == 0 && kl_end == p.nzl && nch >= 3 && p.cmpfirst && want_tmap())) return false;
    return true;
}

// would an apply of the whole slab with the exchange overlapped (in-kernel halo wait) run on the row-pair kernel?
bool rowpair_halo_overlap_ok(const ApplyParams &p) {
    static const bool other_kernel = [] {   // A/B overrides that send the operator to the first-generation kernel
        const char *g = getenv("FDFD_K1_GEN"), *t = getenv("FDFD_TY");
        return (g && atoi(g) == 1) || (t && atoi(t) != 0);
    }();
    if (other_kernel) return false;
    ApplyParams q = p;
    uint32_t dummy = 0;
    q.halo_flag = &dummy;                      // plan as the gated launch would
    if (q.has_off && q.has_mass && !rp_fused_ok(q)) {
        if (!(p.offmask && p.offmask_ty == 8)) return false;   // dense, not fusable: first-generation kernel
        q.has_off = 0;                                         // sparse plan: the diagonal part runs on this kernel
    }
    for (int w = 0; w < 3; ++w)
        if (q.s1[w] != 1 && q.s1[w] != -1) return false;
    return rowpair_supported(q, 0, q.nzl);
}

// Tensor maps are cached by (address, geometry): a Krylov solve applies the operator to a handful of workspace
// vectors over and over, so the driver's encode call (~1 us) is paid once per vector, not once per apply.
static bool cached_map(TmaMap *out, const void *base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                       uint64_t pitch0 = 0) {
    struct Entry { const void *base; uint64_t d0, d1, d2, pitch; uint32_t b0, b1; TmaMap m; };
    static thread_local std::vector<Entry> cache;
    const uint64_t pitch = pitch0 ? pitch0 : d0;      // elements between consecutive rows
    for (size_t i = 0; i < cache.size(); ++i) {
        const Entry &e = cache[i];
        if (e.base == base && e.d0 == d0 && e.d1 == d1 && e.d2 == d2 && e.pitch == pitch && e.b0 == b0 && e.b1 == b1) {
            *out = e.m;
            return true;
        }
    }
    const uint64_t dims[3] = {d0, d1, d2}, strides[2] = {pitch * 8, pitch * d1 * 8};
    const uint32_t box[3] = {b0, b1, 1};
    Entry e{base, d0, d1, d2, pitch, b0, b1, {}};
    if (!tmap_encode_f64_3d(&e.m, base, dims, strides, box)) return false;
    if (cache.size() >= 64) cache.erase(cache.begin());
    cache.push_back(e);
    *out = e.m;
    return true;
}

template <int NWC, int NST, bool MDR, bool HAS_OFF, int RPW>
static cudaError_t launch_rp_shape(RowPairParams &tp, const ApplyParams &p, cudaStream_t s) {
    constexpr int NROW = RPW * NWC;   // output rows of a tile
    // tensor-map TMA path (cmp-first layout): boxes of 32 cells x (E rows | material rows), y boxes of 30 cells x 2 rows
    tp.tmap = 0;
    if (want_tmap() && p.cmpfirst && p.x.base && p.y) {
        const uint64_t d0 = 6ull * p.Nx, d1 = (uint64_t)p.Ny;
        const bool md_tile = p.has_mass && p.md[0] != nullptr;
        bool ok = cached_map(&tp.mx, p.x.base, d0, d1, (uint64_t)p.nzl, 6 * RP_TX, NROW + 2) &&
                  cached_map(&tp.mlo, p.x.lo, d0, d1, 1, 6 * RP_TX, NROW + 2) &&
                  cached_map(&tp.mhi, p.x.hi, d0, d1, 1, 6 * RP_TX, NROW + 2) &&
                  cached_map(&tp.my, p.y, d0, d1, (uint64_t)p.nzl, 6 * (RP_TX - 2), RPW);
        if (ok && md_tile) {
            if (MDR) ok = p.md_aos_r != nullptr && cached_map(&tp.mmd, p.md_aos_r, d0 / 2, d1, (uint64_t)p.nzl + 2, 3 * (RP_TX + 2), NROW,
                                                                 mdr_row_pitch(p.Nx));
            else ok = p.md_aos != nullptr && cached_map(&tp.mmd, p.md_aos, d0, d1, (uint64_t)p.nzl + 2, 6 * RP_TX, NROW);
        }
        if (ok && HAS_OFF)   // off-diagonal rows: padded by one cell / row on either side (wrapped copies or zeros)
            ok = p.mo_aos_r != nullptr && cached_map(&tp.mmo, p.mo_aos_r, 3ull * (p.Nx + 2), (uint64_t)p.Ny + 2, (uint64_t)p.nzl + 2,
                                                     3 * (RP_TX + 2), NROW + 1, mdr_row_pitch(p.Nx + 2));
        tp.tmap = ok ? 1 : 0;
    }
    if (MDR && !tp.tmap) return cudaErrorNotSupported;   // the handle keeps complex rows whenever tensor maps are unavailable
    if (HAS_OFF && p.offmask && p.offmask_ty != 16) return cudaErrorInvalidConfiguration;
    tp.ntx = (p.Nx + RP_TX - 3) / (RP_TX - 2);
    tp.nty = (p.Ny + NROW - 1) / NROW;
    tp.nitems = tp.ntx * tp.nty * tp.nchunk;
    int grid = std::min(tp.nitems, rp_grid_cap(p));
    tp.halo_last = p.halo_flag != nullptr ? 1 : 0;
    static const int want_grid = env_int("FDFD_RP_GRID");   // tuning / test override of the persistent grid size
    if (want_grid >= 1) grid = std::min(tp.nitems, want_grid);
    const bool dot = p.dot_mode == 2;
    if (dot && grid > p.dot_cap) return cudaErrorInvalidConfiguration;
    const bool cf = p.cmpfirst != 0, q = p.has_q != 0;
    const int nfwd = (p.s1[0] > 0) + (p.s1[1] > 0) + (p.s1[2] > 0);
    const int arr = nfwd == 3 ? 0 : nfwd == 0 ? 1 : 2;
#define W(CF, Q, D)                                                                                          \
    (arr == 0 ? launch_rp<CF, Q, D, 0, MDR, HAS_OFF, NWC, NST, RPW>(tp, grid, s)                             \
              : arr == 1 ? launch_rp<CF, Q, D, 1, MDR, HAS_OFF, NWC, NST, RPW>(tp, grid, s)                  \
                         : launch_rp<CF, Q, D, 2, MDR, HAS_OFF, NWC, NST, RPW>(tp, grid, s))
#define V(CF, Q) (dot ? W(CF, Q, true) : W(CF, Q, false))
    if (cf) return q ? V(true, true) : V(true, false);
    if constexpr (MDR) return cudaErrorInvalidConfiguration;
    else return q ? V(false, true) : V(false, false);
#undef W
#undef V
}

// apply over local planes [kl_begin, kl_end) with the row-pair kernel (diagonal mass parameter, or the fused full tensor)
cudaError_t launch_apply_rowpair(const ApplyParams &p, int kl_begin, int kl_end, cudaStream_t s) {
    if (kl_end <= kl_begin) return cudaSuccess;
    RowPairParams tp;
    tp.a = p;
    tp.wrapx = p.wrap[0];
    tp.wrapy = p.wrap[1];
    tp.kl_begin = kl_begin;
    tp.kl_end = kl_end;
    static const int dbg = env_int("FDFD_RP_DEBUG");
    tp.dbg = dbg;
    const int shape = (p.s1[0] * p.s1[0] == 1 && p.s1[1] * p.s1[1] == 1 && p.s1[2] * p.s1[2] == 1)
                          ? rp_pick_shape(p, kl_begin, kl_end, &tp.nchunk) : -1;
    static const bool verbose = getenv("FDFD_VERBOSE") != nullptr;
    if (verbose) fprintf(stderr, "fdfd: row-pair kernel shape %d (%s), %d z-chunk(s), planes [%d, %d)\n", shape,
                         shape == 2 ? "fused full tensor" : shape == 1 ? "real diagonal mass" : shape == 0 ? "complex diagonal mass" : shape == 3 ? "real diagonal mass, 4 stages" : shape >= 4 ? "real diagonal mass, one row per warp" : "unsupported",
                         tp.nchunk, kl_begin, kl_end);
    switch (shape) {
        case 0: return launch_rp_shape<RP_SHAPES[0].nwc, RP_SHAPES[0].nst, RP_SHAPES[0].mdr, RP_SHAPES[0].off, RP_SHAPES[0].rpw>(tp, p, s);
        case 1: return launch_rp_shape<RP_SHAPES[1].nwc, RP_SHAPES[1].nst, RP_SHAPES[1].mdr, RP_SHAPES[1].off, RP_SHAPES[1].rpw>(tp, p, s);
        case 2: return launch_rp_shape<RP_SHAPES[2].nwc, RP_SHAPES[2].nst, RP_SHAPES[2].mdr, RP_SHAPES[2].off, RP_SHAPES[2].rpw>(tp, p, s);
        case 3: return launch_rp_shape<RP_SHAPES[3].nwc, RP_SHAPES[3].nst, RP_SHAPES[3].mdr, RP_SHAPES[3].off, RP_SHAPES[3].rpw>(tp, p, s);
        case 4: return launch_rp_shape<RP_SHAPES[4].nwc, RP_SHAPES[4].nst, RP_SHAPES[4].mdr, RP_SHAPES[4].off, RP_SHAPES[4].rpw>(tp, p, s);
        default: return cudaErrorNotSupported;
    }
}

}  // namespace fdfd
