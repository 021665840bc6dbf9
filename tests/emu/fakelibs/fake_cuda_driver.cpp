// TEST INFRASTRUCTURE ONLY - stand-in for the two driver entry points csrc/comm.cpp resolves from libcuda.so.1 (stream
// memory operations), for the multi-rank runs of the CPU logic-check build: streams are synchronous there, so a write
// is a store and a wait is a bounded spin on memory another rank's process maps too (tests/emu/README.md).
#include <sched.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

extern "C" {

int cuStreamWriteValue32_v2(void *, unsigned long long addr, uint32_t value, unsigned int) {
    __atomic_store_n(reinterpret_cast<uint32_t *>(addr), value, __ATOMIC_RELEASE);
    return 0;
}
int cuStreamWaitValue32_v2(void *, unsigned long long addr, uint32_t value, unsigned int flags) {
    if (flags != 0) return 1;   // only CU_STREAM_WAIT_VALUE_GEQ is used
    const auto t0 = std::chrono::steady_clock::now();
    while ((int32_t)(__atomic_load_n(reinterpret_cast<uint32_t *>(addr), __ATOMIC_ACQUIRE) - value) < 0) {
        sched_yield();
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(60)) {
            fprintf(stderr, "fake libcuda: cuStreamWaitValue32 waited 60 s for %u (dead-lock)\n", value);
            abort();
        }
    }
    return 0;
}

}  // extern "C"
