// Micro-benchmark (round 2): how fast can one persistent CTA per SM stream x-y tiles of a cmp-first complex128 field
// (48 B per cell, rows of Nx cells) from HBM into a shared-memory ring, depending on HOW the rows are fetched?
//   mode 0  1-D bulk copies (cp.async.bulk), one per tile row, 32 cells = 1536 B, start 16-B aligned only (the
//           row-pair kernel's scheme)
//   mode 1  same, tile origin shifted so that every row start is 128-B aligned
//   mode 2  tensor-map TMA (cp.async.bulk.tensor.3d), one box of ROWS x 1536 B per array and stage
//   mode 3  1-D bulk copies of one contiguous block per stage (upper bound of the 1-D path; not a tile)
//   mode 4  per-thread cp.async (LDGSTS) 16 B, 4 producer warps
// Two arrays (x and material) are streamed, as in the kernel.  Consumers only wait and release.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_stream tma_stream.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t par) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tma3d(void *d, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(s32(d)), "l"(m), "r"(s32(b)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma3d_store(const CUtensorMap *m, int c0, int c1, int c2, const void *src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(m), "r"(s32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *d, const void *s, uint32_t n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s32(s)), "r"(n) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cpasync16(void *d, const void *s) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(d)), "l"(s) : "memory"); }
__device__ __forceinline__ void cpasync_arrive(uint64_t *b) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(s32(b)) : "memory"); }

__device__ __forceinline__ unsigned char *yw_of(unsigned char *ybuf, int wid) { return ybuf + wid * 3072; }
constexpr int NWC = 7, NST = 4, ROWS_E = 16, ROWS_M = 14, ROWB = 1536;
constexpr int STAGE = (ROWS_E + ROWS_M) * ROWB;

struct P {
    const char *x, *m;
    char *y;
    int Nx, Ny, Nz, ntx, nty, nchunk, nitems, mode, shift, rows_m;
};

__global__ void __launch_bounds__(32 * (NWC + 4), 1) stream_kernel(const __grid_constant__ P p, const __grid_constant__ CUtensorMap tx,
                                                                    const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap ty, const __grid_constant__ CUtensorMap ty2, unsigned long long *sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *ybuf = smem + NST * STAGE;   // NWC * 3072
    uint64_t *full = reinterpret_cast<uint64_t *>(ybuf + NWC * 3072), *empty = full + NST;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nprod = p.mode == 4 ? 4 : 1;
    const int lmode = p.mode >= 5 ? 2 : p.mode;
    const bool aligned_st = p.mode == 7;   // load path
    const int smode = (p.mode == 5 || p.mode == 7) ? 1 : p.mode == 6 ? 2 : 0;   // store path: 0 none, 1 tensor map, 2 1-D bulk rows
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], p.mode == 4 ? 128 : 1); mbar_init(&empty[s], NWC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t rowpitch = (size_t)p.Nx * 48, planepitch = rowpitch * p.Ny;
    uint32_t g = 0;
    unsigned long long acc = 0;
    if (wid >= NWC) {
        if (wid - NWC >= nprod) return;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int b = item;
            const int tile_x = b % p.ntx; b /= p.ntx;
            const int tile_y = b % p.nty;
            const int chunk = b / p.nty;
            int ox = tile_x * 30 - 1 + p.shift, oy = tile_y * 14 - 1;
            if (ox < 0) ox = 0;
            if (ox + 32 > p.Nx) ox = p.Nx - 32;
            if (oy < 0) oy = 0;
            if (oy + ROWS_E > p.Ny) oy = p.Ny - ROWS_E;
            if (p.mode == 1) ox &= ~7;
            const int k0 = (int)((long)p.Nz * chunk / p.nchunk), k1 = (int)((long)p.Nz * (chunk + 1) / p.nchunk);
            for (int k = k0; k < k1; ++k, ++g) {
                const int s = g % NST;
                if (g >= NST) mbar_wait(&empty[s], ((g / NST) - 1) & 1);
                unsigned char *dst = smem + s * STAGE;
                const uint32_t bytes = (uint32_t)(ROWS_E + p.rows_m) * ROWB;
                if (lmode == 0 || lmode == 1) {
                    if (lane == 0) mbar_expect(&full[s], bytes);
                    __syncwarp();
                    if (lane < ROWS_E)
                        bulk_g2s(dst + lane * ROWB, p.x + (size_t)k * planepitch + (size_t)(oy + lane) * rowpitch + (size_t)ox * 48, ROWB, &full[s]);
                    else if (lane < ROWS_E + p.rows_m)
                        bulk_g2s(dst + lane * ROWB, p.m + (size_t)k * planepitch + (size_t)(oy + 1 + lane - ROWS_E) * rowpitch + (size_t)ox * 48, ROWB, &full[s]);
                } else if (lmode == 2) {
                    if (lane == 0) {
                        mbar_expect(&full[s], bytes);
                        tma3d(dst, &tx, ox * 6, oy, k, &full[s]);
                        if (p.rows_m) tma3d(dst + ROWS_E * ROWB, &tm, ox * 6, oy + 1, k, &full[s]);
                    }
                } else if (lmode == 3) {
                    if (lane == 0) {
                        mbar_expect(&full[s], bytes);
                        const size_t off = ((size_t)item * 64 + (k - k0)) * (size_t)(ROWS_E * ROWB) % ((size_t)p.Nz * planepitch - STAGE);
                        bulk_g2s(dst, p.x + (off & ~(size_t)127), ROWS_E * ROWB, &full[s]);
                        if (p.rows_m) bulk_g2s(dst + ROWS_E * ROWB, p.m + (off & ~(size_t)127), p.rows_m * ROWB, &full[s]);
                    }
                } else {
                    const int t = tid - NWC * 32;   // 0..127
                    for (int r = 0; r < ROWS_E + p.rows_m; ++r) {
                        const char *src = r < ROWS_E ? p.x + (size_t)k * planepitch + (size_t)(oy + r) * rowpitch + (size_t)ox * 48
                                                     : p.m + (size_t)k * planepitch + (size_t)(oy + 1 + r - ROWS_E) * rowpitch + (size_t)ox * 48;
                        if (t < 96) cpasync16(dst + r * ROWB + t * 16, src + t * 16);
                    }
                    cpasync_arrive(&full[s]);
                }
            }
        }
    } else {
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            int b = item;
            const int tile_x = b % p.ntx; b /= p.ntx;
            const int tile_y = b % p.nty;
            b /= p.nty;
            int ox = tile_x * 30 - 1, oy = tile_y * 14 - 1;
            if (ox < 0) ox = 0;
            if (ox + 32 > p.Nx) ox = p.Nx - 32;
            if (oy < 0) oy = 0;
            if (oy + ROWS_E > p.Ny) oy = p.Ny - ROWS_E;
            const int k0 = (int)((long)p.Nz * b / p.nchunk), k1 = (int)((long)p.Nz * (b + 1) / p.nchunk);
            for (int k = k0; k < k1; ++k, ++g) {
                const int s = g % NST;
                mbar_wait(&full[s], (g / NST) & 1);
                acc += *reinterpret_cast<const unsigned long long *>(smem + s * STAGE + tid * 16);
                if (smode) {
                    unsigned char *yw = ybuf + wid * 3072;
                    if (lane < 2) bulk_wait_read0();
                    __syncwarp();
                    *reinterpret_cast<double2 *>(yw + lane * 48) = make_double2((double)acc, 1.0);
                    *reinterpret_cast<double2 *>(yw + 1440 + lane * 48) = make_double2((double)acc, 2.0);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
                if (smode == 1 && lane == 0) {
                    if (aligned_st) tma3d_store(&ty2, (ox & ~7) * 6, oy + 1 + 2 * wid, k, yw_of(ybuf, wid));   // 32 cells, 128-B aligned rows (timing only)
                    else tma3d_store(&ty, (ox + 1) * 6, oy + 1 + 2 * wid, k, yw_of(ybuf, wid));
                    bulk_commit();
                } else if (smode == 2 && lane < 2) {
                    bulk_s2g(p.y + (size_t)k * planepitch + (size_t)(oy + 1 + 2 * wid + lane) * rowpitch + (size_t)(ox + 1) * 48,
                             ybuf + wid * 3072 + lane * 1440, 1440);
                    bulk_commit();
                }
            }
        }
        if (lane < 2) bulk_wait0();
        if (acc == 0x1234567) *sink = acc;
    }
}

int main(int argc, char **argv) {
    const int Nx = argc > 1 ? atoi(argv[1]) : 200, Ny = argc > 2 ? atoi(argv[2]) : 200, Nz = argc > 3 ? atoi(argv[3]) : 200;
    const size_t bytes = (size_t)Nx * Ny * Nz * 48;
    char *x, *m, *y;
    unsigned long long *sink;
    CK(cudaMalloc(&x, bytes + (1 << 20)));
    CK(cudaMalloc(&m, bytes + (1 << 20)));
    CK(cudaMalloc(&sink, 8));
    CK(cudaMalloc(&y, bytes + (1 << 20)));
    CK(cudaMemset(x, 1, bytes));
    CK(cudaMemset(m, 1, bytes));
    // tensor maps: doubles, dims {6 Nx, Ny, Nz}
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeFn enc = (EncodeFn)fn;
    CUtensorMap tx, tm, ty, ty2;
    auto mk = [&](CUtensorMap *t, void *base, int rows, CUtensorMapL2promotion l2, int boxw = 192) {
        cuuint64_t dims[3] = {(cuuint64_t)6 * Nx, (cuuint64_t)Ny, (cuuint64_t)Nz};
        cuuint64_t strides[2] = {(cuuint64_t)Nx * 48, (cuuint64_t)Nx * 48 * Ny};
        cuuint32_t box[3] = {(cuuint32_t)boxw, (cuuint32_t)rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(t, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    };
    const size_t smem = NST * STAGE + NWC * 3072 + 2 * NST * 8 + 128;
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int l2 = 0; l2 < 2; ++l2) {
        mk(&tx, x, ROWS_E, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
        mk(&tm, m, ROWS_M, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
        mk(&ty, y, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 180);
        mk(&ty2, y, 2, CU_TENSOR_MAP_L2_PROMOTION_NONE, 192);
        for (int mode = 0; mode < 8; ++mode) {
            if (l2 && mode != 2) continue;
            for (int rows_m = 0; rows_m <= ROWS_M; rows_m += ROWS_M) {
                for (int nchunk : {8, 4}) {
                    P p{x, m, y, Nx, Ny, Nz, (Nx + 29) / 30, (Ny + 13) / 14, nchunk, 0, mode, 0, rows_m};
                    p.nitems = p.ntx * p.nty * p.nchunk;
                    for (int it = 0; it < 3; ++it) stream_kernel<<<148, 32 * (NWC + 4), smem>>>(p, tx, tm, ty, ty2, sink);
                    CK(cudaDeviceSynchronize());
                    CK(cudaEventRecord(e0));
                    const int reps = 20;
                    for (int it = 0; it < reps; ++it) stream_kernel<<<148, 32 * (NWC + 4), smem>>>(p, tx, tm, ty, ty2, sink);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    float ms;
                    CK(cudaEventElapsedTime(&ms, e0, e1));
                    ms /= reps;
                    const double moved = (double)p.ntx * p.nty * Nz * (ROWS_E + rows_m) * ROWB;
                    printf("{\"mode\": %d, \"l2promo\": %d, \"rows_m\": %d, \"nchunk\": %d, \"ms\": %.4f, \"smem_GBps\": %.0f, \"unique_GBps\": %.0f}\n", mode, l2,
                           rows_m, nchunk, ms, moved / ms / 1e6, (double)bytes * ((rows_m ? 2 : 1) + (mode >= 5 ? 1 : 0)) / ms / 1e6);
                    fflush(stdout);
                }
            }
        }
    }
    return 0;
}
