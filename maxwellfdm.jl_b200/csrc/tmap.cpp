// Host side of the tensor-map TMA path: encodes 3-D tiled tensor maps (8-byte elements) for the row-pair kernel.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint), so the library still links
// cudart only.  A DOF vector in the reference's cmp-first order (model.jl:75-83) is viewed as doubles
// [Nz][Ny][6 Nx]: one box row = 32 cells x 3 components x complex128 = 1536 B.
#include <cstring>

#include "fdfd_internal.h"
#include "ptx_sm100.cuh"

namespace fdfd {

#ifdef FDFD_EMU
// logic-check build (tests/emu): the image holds plain fields that the shim's tma_load_3d / tma_store_3d interpret
bool tmap_encode_f64_3d(TmaMap *out, const void *base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                        const uint32_t box[3]) {
    std::memset(out, 0, sizeof(*out));
    out->opaque[0] = (unsigned long long)(uintptr_t)base;
    for (int i = 0; i < 3; ++i) out->opaque[1 + i] = dims[i];
    for (int i = 0; i < 2; ++i) out->opaque[4 + i] = strides_bytes[i];
    for (int i = 0; i < 3; ++i) out->opaque[6 + i] = box[i];
    return base != nullptr && ((uintptr_t)base & 15) == 0 && box[0] <= 256 && (box[0] * 8) % 16 == 0 &&
           strides_bytes[0] % 16 == 0 && strides_bytes[1] % 16 == 0;
}
#else
namespace {
// CUresult cuTensorMapEncodeTiled(CUtensorMap*, CUtensorMapDataType, cuuint32_t rank, void* gaddr, const cuuint64_t* gdim,
//     const cuuint64_t* gstride, const cuuint32_t* box, const cuuint32_t* estride, CUtensorMapInterleave,
//     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill)   - enums are C ints (cuda.h)
typedef int (*EncodeTiledFn)(void *, int, uint32_t, void *, const uint64_t *, const uint64_t *, const uint32_t *,
                             const uint32_t *, int, int, int, int);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
}  // namespace

bool tmap_encode_f64_3d(TmaMap *out, const void *base, const uint64_t dims[3], const uint64_t strides_bytes[2],
                        const uint32_t box[3]) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || !base) return false;
    const uint32_t es[3] = {1, 1, 1};
    constexpr int kFloat64 = 8 /* CU_TENSOR_MAP_DATA_TYPE_FLOAT64 */, kNone = 0;
    // interleave none, swizzle none (rows are read by 16-B lanes at a 48-B stride: conflict-free as they are),
    // L2 promotion none (measured equal to 256 B, scripts/micro/tma_stream.cu), out-of-range elements read as zero
    return fn(out, kFloat64, 3, const_cast<void *>(base), dims, strides_bytes, box, es, kNone, kNone, kNone, kNone) == 0;
}
#endif

bool tmap_probe(const void *dev_ptr) {
    TmaMap m;
    const uint64_t dims[3] = {64, 4, 1}, strides[2] = {512, 2048};
    const uint32_t box[3] = {64, 2, 1};
    return dev_ptr != nullptr && tmap_encode_f64_3d(&m, dev_ptr, dims, strides, box);
}

}  // namespace fdfd
