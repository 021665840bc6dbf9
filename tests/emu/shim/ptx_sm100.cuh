// TEST INFRASTRUCTURE ONLY - host stand-ins for maxwellfdm.jl_b200/csrc/ptx_sm100.cuh (same names, same contracts),
// used when the kernel sources are compiled for the CPU logic-check harness (tests/emu/README.md).
//   * mbarrier: phase bit + pending-arrival count + transaction-byte count, packed into the 64-bit word;
//   * cp.async.bulk global->shared: a memcpy whose completion is counted on the mbarrier - immediately
//     (FDFD_EMU_ASYNC=eager, default) or only when the CTA cannot make progress otherwise (=lazy, the latest legal
//     moment: catches reads that do not wait for the barrier);
//   * cp.async.bulk shared->global: a memcpy performed at issue (eager) or when the issuing thread waits for its
//     bulk groups / exits (lazy: catches staging buffers that are overwritten before the copy engine has read them);
//   * fences and prefetches: no-ops.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fdfd {

inline uint32_t smem_u32(const void *p) { return (uint32_t)(uintptr_t)p; }
inline void mbar_init(uint64_t *bar, uint32_t count) { emu::mbar_init(bar, count); }
inline void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) { emu::mbar_arrive_expect_tx(bar, bytes); }
inline void mbar_arrive(uint64_t *bar) { emu::mbar_arrive_expect_tx(bar, 0); }
inline void mbar_wait(uint64_t *bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) { emu::bulk_g2s(dst, src, bytes, bar); }
// tensor-map stand-in: the encode helper of csrc/tmap.cpp fills the image with plain fields under FDFD_EMU
struct alignas(64) TmaMap { unsigned long long opaque[16]; };
inline void tma_load_3d(void *dst, const TmaMap *m, int c0, int c1, int c2, uint64_t *bar) { emu::tma_load_3d(dst, m->opaque, c0, c1, c2, bar); }
inline void tma_store_3d(const TmaMap *m, int c0, int c1, int c2, const void *src) { emu::tma_store_3d(m->opaque, c0, c1, c2, src); }
inline void fence_barrier_init() {}
inline void bulk_s2g(void *dst, const void *src, uint32_t bytes) { emu::bulk_s2g(dst, src, bytes); }
inline void bulk_commit() { emu::bulk_commit(); }
inline void bulk_wait_read0() { emu::bulk_wait_read0(); }
inline void bulk_wait0() { emu::bulk_wait0(); }
inline void prefetch_l2(const void *) {}
inline void fence_proxy_async() {}
inline void fence_proxy_async_all() {}
inline uint32_t ld_acquire_sys(const uint32_t *p) { return *(const volatile uint32_t *)p; }

}  // namespace fdfd
