// Shared pieces of the Krylov vector kernels: deterministic block/grid reduction of complex sums.
#pragma once
#include "cplx.cuh"
#include "fdfd_internal.h"

namespace fdfd {
namespace kry {

constexpr int RB = 256;          // threads per block of the vector kernels
constexpr int MAXB = 148 * 8;    // blocks (persistent grid-stride)
constexpr int NSLOT = 24;         // complex scalar slots

struct Red {
    double2 *partial;      // [NSLOT][MAXB]
    unsigned int *ticket;  // [NSLOT]
    double2 *scal;         // [NSLOT]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-reduce NR complex accumulators and publish them into scal[slot0 .. slot0+NR).
template <int NR>
__device__ __forceinline__ void reduce_publish(double2 (&acc)[NR], const Red &rd, int slot0) {
    __shared__ double2 sm[NR][RB / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        acc[r].x = warp_sum(acc[r].x);
        acc[r].y = warp_sum(acc[r].y);
        if (lane == 0) sm[r][wid] = acc[r];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double2 v = lane < RB / 32 ? sm[r][lane] : c_zero();
            v.x = warp_sum(v.x);
            v.y = warp_sum(v.y);
            if (lane == 0) rd.partial[(slot0 + r) * MAXB + blockIdx.x] = v;
        }
        if (lane == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(&rd.ticket[slot0], 1u);
            is_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            double2 v = c_zero();
            for (int b = threadIdx.x; b < (int)gridDim.x; b += RB) {
                const double2 q = __ldcg(&rd.partial[(slot0 + r) * MAXB + b]);
                v.x += q.x;
                v.y += q.y;
            }
            v.x = warp_sum(v.x);
            v.y = warp_sum(v.y);
            __syncthreads();
            if (lane == 0) sm[r][wid] = v;
            __syncthreads();
            if (threadIdx.x == 0) {
                double2 s = c_zero();
                for (int w = 0; w < RB / 32; ++w) { s.x += sm[r][w].x; s.y += sm[r][w].y; }
                rd.scal[slot0 + r] = s;
            }
        }
        if (threadIdx.x == 0) rd.ticket[slot0] = 0;
    }
}

// conj(a) * b accumulated
__device__ __forceinline__ void dot_acc(double2 &acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(-a.y, b.x, acc.y);
}



// a^T b accumulated (bilinear form, no conjugation; used by QMR)
__device__ __forceinline__ void dotu_acc(double2 &acc, double2 a, double2 b) {
    acc.x = fma(a.x, b.x, acc.x);
    acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y);
    acc.y = fma(a.y, b.x, acc.y);
}

inline int grid_for(int64_t n) {
    int64_t b = (n + RB - 1) / RB;
    b = (b + 3) / 4;  // a few elements per thread
    if (b > MAXB) b = MAXB;
    if (b < 1) b = 1;
    return (int)b;
}

inline Red make_red(Ctx *c) {
    Red rd;
    rd.scal = reinterpret_cast<double2 *>(c->scal);
    rd.partial = rd.scal + NSLOT;
    rd.ticket = reinterpret_cast<unsigned int *>(rd.partial + NSLOT * MAXB);
    return rd;
}

// vector workspace (nvec * nloc) + scalar block [scal NSLOT][partial NSLOT*MAXB][ticket NSLOT]
inline int workspace(Ctx *c, int nvec) {
    const size_t need = (size_t)nvec * c->nloc * sizeof(double2);
    if (c->work_bytes < need) {
        if (c->work) cudaFree(c->work);
        c->work = nullptr;
        c->work_bytes = 0;
        FDFD_CUDA(c, cudaMalloc((void **)&c->work, need));
        c->work_bytes = need;
    }
    if (!c->scal) {
        const size_t bytes = sizeof(double2) * NSLOT + sizeof(double2) * NSLOT * MAXB + sizeof(unsigned int) * NSLOT;
        FDFD_CUDA(c, cudaMalloc((void **)&c->scal, bytes));
        FDFD_CUDA(c, cudaMemset(c->scal, 0, bytes));
        FDFD_CUDA(c, cudaMallocHost((void **)&c->scal_host, sizeof(double2) * NSLOT));
    }
    return FDFD_OK;
}

#define GRID_STRIDE(i, n) \
    for (int64_t i = blockIdx.x * (int64_t)fdfd::kry::RB + threadIdx.x; i < (n); i += (int64_t)gridDim.x * fdfd::kry::RB)

}  // namespace kry
}  // namespace fdfd
