// Halo data plane of the z-slabs since round 2 (default whenever every rank can set it up; FDFD_PEER_HALO=0 forces NCCL).
// First run on hardware in round 2 (2x B200: scripts/dist_check.py clean, exchange 16.9 us vs 20.6 us for NCCL send/recv,
// and - because it needs no SM - the exchange can hide behind the apply kernel: 245 vs 219 GDOF/s).
//
// z-halo exchange over NVLink peer memory with no SM-resident collective: every rank maps its two neighbours' halo
// buffers and flag words with CUDA IPC (handles travel once over the NCCL communicator).  Per exchange a plane moves
// by ONE copy-engine cudaMemcpyAsync straight into the neighbour's receive buffer; ordering uses stream memory
// operations (cuStreamWriteValue32 / cuStreamWaitValue32) on flag words that live in the RECEIVER's memory:
//     DATA_LO / DATA_HI   "your halo_lo / halo_hi holds the planes of epoch e"      (written by the sender)
//     FREE_UP / FREE_DN   "I have consumed what you sent me in epoch e"             (written by the receiver)
// Exchange e on stream s:   release(e-1) to both neighbours            [s is ordered after the consumer of e-1]
//                           wait FREE >= e-1, copy, write DATA = e     [towards each neighbour]
//                           wait DATA >= e                             [from each neighbour]
// Nothing here needs an SM, which is what the in-kernel halo wait (apply_tiled.cu, HWAIT) requires of its transfer.
// Replaces the ncclSend/ncclRecv pair of comm.cpp; the reference has no distributed code.
#include <cstdlib>
#include <cstring>

#include "fdfd_internal.h"

namespace fdfd {

namespace {
enum { DATA_LO = 0, DATA_HI = 1, FREE_UP = 2, FREE_DN = 3, DIRECT_LO = 4, DIRECT_HI = 5, NFLAGS = 6 };

struct Handles {
    cudaIpcMemHandle_t halo_lo, halo_hi, flags;
};
}  // namespace

int peer_halo_init(Ctx *c) {
    PeerHalo &ph = c->peer;
    if (ph.ready || c->d.nranks == 1) return FDFD_OK;
    if (!c->d.order_cmpfirst) return FDFD_OK;                 // component-major planes are not contiguous: NCCL path
    int up, dn;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    FDFD_CUDA(c, cudaMalloc((void **)&ph.flags, NFLAGS * sizeof(uint32_t)));
    FDFD_CUDA(c, cudaMemset(ph.flags, 0, NFLAGS * sizeof(uint32_t)));
    Handles mine;
    FDFD_CUDA(c, cudaIpcGetMemHandle(&mine.halo_lo, c->halo_lo));
    FDFD_CUDA(c, cudaIpcGetMemHandle(&mine.halo_hi, c->halo_hi));
    FDFD_CUDA(c, cudaIpcGetMemHandle(&mine.flags, ph.flags));
    // ship the handles to both neighbours (device staging: NCCL moves device memory)
    unsigned char *stage = nullptr;
    FDFD_CUDA(c, cudaMalloc((void **)&stage, 3 * sizeof(Handles)));
    FDFD_CUDA(c, cudaMemcpy(stage, &mine, sizeof(Handles), cudaMemcpyHostToDevice));
    int r = comm_exchange_bytes(c, stage, stage + sizeof(Handles), stage + 2 * sizeof(Handles), sizeof(Handles), c->stream);
    if (r != FDFD_OK) { cudaFree(stage); return r; }
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    Handles from_up, from_dn;
    FDFD_CUDA(c, cudaMemcpy(&from_up, stage + sizeof(Handles), sizeof(Handles), cudaMemcpyDeviceToHost));
    FDFD_CUDA(c, cudaMemcpy(&from_dn, stage + 2 * sizeof(Handles), sizeof(Handles), cudaMemcpyDeviceToHost));
    cudaFree(stage);
    int nm = 0;
    auto open = [&](const cudaIpcMemHandle_t &h, void **out) -> cudaError_t {
        cudaError_t e = cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) ph.mapped[nm++] = *out;
        return e;
    };
    if (up >= 0) {
        FDFD_CUDA(c, open(from_up.halo_lo, (void **)&ph.up_halo_lo));
        FDFD_CUDA(c, open(from_up.flags, (void **)&ph.up_flags));
    }
    if (dn >= 0) {
        FDFD_CUDA(c, open(from_dn.halo_hi, (void **)&ph.dn_halo_hi));
        if (dn == up) ph.dn_flags = ph.up_flags;              // two ranks with a periodic z axis: one peer, map once
        else FDFD_CUDA(c, open(from_dn.flags, (void **)&ph.dn_flags));
    }
    ph.epoch = 0;
    ph.ready = true;
    return FDFD_OK;
}

void peer_halo_destroy(Ctx *c) {
    PeerHalo &ph = c->peer;
    peer_direct_unmap(c);
    for (void *&m : ph.mapped) {
        if (m) cudaIpcCloseMemHandle(m);
        m = nullptr;
    }
    if (ph.flags) cudaFree(ph.flags);
    ph = PeerHalo();
}

int peer_halo_exchange(Ctx *c, const double2 *first_plane, const double2 *last_plane, cudaStream_t s) {
    PeerHalo &ph = c->peer;
    int up, dn, r;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    const size_t bytes = (size_t)c->plane * sizeof(double2);
    const uint32_t e = ++ph.epoch;
    // 1. the consumer of epoch e-1 is behind us on this stream: tell the senders their planes are consumed
    if (dn >= 0 && (r = stream_write_u32(c, s, &ph.dn_flags[FREE_UP], e - 1)) != FDFD_OK) return r;
    if (up >= 0 && (r = stream_write_u32(c, s, &ph.up_flags[FREE_DN], e - 1)) != FDFD_OK) return r;
    // 2. send: my last plane -> up's halo_lo, my first plane -> down's halo_hi
    if (up >= 0) {
        if ((r = stream_wait_geq_u32(c, s, &ph.flags[FREE_UP], e - 1)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaMemcpyAsync(ph.up_halo_lo, last_plane, bytes, cudaMemcpyDeviceToDevice, s));
        if ((r = stream_write_u32(c, s, &ph.up_flags[DATA_LO], e)) != FDFD_OK) return r;
    }
    if (dn >= 0) {
        if ((r = stream_wait_geq_u32(c, s, &ph.flags[FREE_DN], e - 1)) != FDFD_OK) return r;
        FDFD_CUDA(c, cudaMemcpyAsync(ph.dn_halo_hi, first_plane, bytes, cudaMemcpyDeviceToDevice, s));
        if ((r = stream_write_u32(c, s, &ph.dn_flags[DATA_HI], e)) != FDFD_OK) return r;
    }
    // 3. receive
    if (dn >= 0 && (r = stream_wait_geq_u32(c, s, &ph.flags[DATA_LO], e)) != FDFD_OK) return r;
    if (up >= 0 && (r = stream_wait_geq_u32(c, s, &ph.flags[DATA_HI], e)) != FDFD_OK) return r;
    return FDFD_OK;
}

// ---- peer-direct reads ---------------------------------------------------------------------------------------------
// Inside BiCGSTAB the vectors that are applied (p, s) live in the library's own workspace, so a neighbour can read their
// boundary planes in place.  Ordering: the producer announces "my boundary planes of epoch e are final" with a stream
// memory operation into the consumer's flag word (behind the kernel that wrote the planes); the consumer's stream
// waits for that word before the apply kernel.  The reverse hazard (a producer overwriting planes a neighbour still
// reads) cannot occur: between an apply of a vector and the next kernel that writes it lies an allreduce of every
// rank (sigma after A p, (t,s),(t,t) after A s), which completes only after every rank's apply.
namespace {
struct WorkInfo {
    cudaIpcMemHandle_t handle;
    int64_t nloc, nzl;
};
}  // namespace

bool peer_direct_enabled(const Ctx *c) {
    static const bool env = getenv("FDFD_PEER_DIRECT") != nullptr;
    return env && c->peer.ready && c->d.nranks > 1 && c->d.order_cmpfirst;
}

void peer_direct_unmap(Ctx *c) {
    PeerHalo &ph = c->peer;
    for (void *&m : ph.work_mapped) {
        if (m) cudaIpcCloseMemHandle(m);
        m = nullptr;
    }
    ph.up_work = ph.dn_work = nullptr;
    ph.direct = ph.direct_pending = false;
}

int peer_direct_map(Ctx *c) {
    PeerHalo &ph = c->peer;
    if (!peer_direct_enabled(c) || !c->work) return FDFD_OK;
    int up, dn;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    WorkInfo mine{}, from_up{}, from_dn{};
    FDFD_CUDA(c, cudaIpcGetMemHandle(&mine.handle, c->work));
    mine.nloc = c->nloc;
    mine.nzl = c->k1 - c->k0;
    unsigned char *stage = nullptr;
    FDFD_CUDA(c, cudaMalloc((void **)&stage, 3 * sizeof(WorkInfo)));
    FDFD_CUDA(c, cudaMemcpy(stage, &mine, sizeof(WorkInfo), cudaMemcpyHostToDevice));
    int r = comm_exchange_bytes(c, stage, stage + sizeof(WorkInfo), stage + 2 * sizeof(WorkInfo), sizeof(WorkInfo), c->stream);
    if (r != FDFD_OK) { cudaFree(stage); return r; }
    FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    FDFD_CUDA(c, cudaMemcpy(&from_up, stage + sizeof(WorkInfo), sizeof(WorkInfo), cudaMemcpyDeviceToHost));
    FDFD_CUDA(c, cudaMemcpy(&from_dn, stage + 2 * sizeof(WorkInfo), sizeof(WorkInfo), cudaMemcpyDeviceToHost));
    cudaFree(stage);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle cache size");
    const bool same_up = up < 0 || (ph.up_work && std::memcmp(ph.up_handle, &from_up.handle, 64) == 0);
    const bool same_dn = dn < 0 || (ph.dn_work && std::memcmp(ph.dn_handle, &from_dn.handle, 64) == 0);
    if (!(ph.direct && same_up && same_dn)) {     // first solve, or a neighbour re-allocated its workspace
        peer_direct_unmap(c);
        if (up >= 0) {
            FDFD_CUDA(c, cudaIpcOpenMemHandle(&ph.work_mapped[0], from_up.handle, cudaIpcMemLazyEnablePeerAccess));
            ph.up_work = static_cast<double2 *>(ph.work_mapped[0]);
            std::memcpy(ph.up_handle, &from_up.handle, 64);
        }
        if (dn >= 0) {
            if (dn == up) ph.dn_work = ph.up_work;            // one peer on both sides: map once
            else {
                FDFD_CUDA(c, cudaIpcOpenMemHandle(&ph.work_mapped[1], from_dn.handle, cudaIpcMemLazyEnablePeerAccess));
                ph.dn_work = static_cast<double2 *>(ph.work_mapped[1]);
            }
            std::memcpy(ph.dn_handle, &from_dn.handle, 64);
        }
    }
    ph.up_nloc = from_up.nloc;
    ph.dn_nloc = from_dn.nloc;
    ph.dn_nzl = from_dn.nzl;
    ph.direct = true;
    ph.direct_pending = false;
    return FDFD_OK;
}

int peer_direct_signal(Ctx *c, cudaStream_t s) {
    PeerHalo &ph = c->peer;
    int up, dn, r;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    const uint32_t e = ++ph.depoch;
    // my last plane is the plane below the up neighbour's slab, my first plane the one above the down neighbour's
    if (up >= 0 && (r = stream_write_u32(c, s, &ph.up_flags[DIRECT_LO], e)) != FDFD_OK) return r;
    if (dn >= 0 && (r = stream_write_u32(c, s, &ph.dn_flags[DIRECT_HI], e)) != FDFD_OK) return r;
    ph.direct_pending = true;
    return FDFD_OK;
}

int peer_direct_wait(Ctx *c, cudaStream_t s) {
    PeerHalo &ph = c->peer;
    int up, dn, r;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    if (dn >= 0 && (r = stream_wait_geq_u32(c, s, &ph.flags[DIRECT_LO], ph.depoch)) != FDFD_OK) return r;
    if (up >= 0 && (r = stream_wait_geq_u32(c, s, &ph.flags[DIRECT_HI], ph.depoch)) != FDFD_OK) return r;
    ph.direct_pending = false;
    return FDFD_OK;
}

bool peer_direct_planes(const Ctx *c, const double2 *x, const double2 **lo, const double2 **hi) {
    const PeerHalo &ph = c->peer;
    if (!ph.direct || !c->work || x < c->work || c->nloc <= 0) return false;
    const int64_t off = x - c->work;
    if (off % c->nloc != 0 || (size_t)(off + c->nloc) * sizeof(double2) > c->work_bytes) return false;
    const int64_t idx = off / c->nloc;
    // same vector of the neighbour's workspace: its last plane lies below my slab, its first plane above it
    if (ph.dn_work) *lo = ph.dn_work + idx * ph.dn_nloc + (ph.dn_nzl - 1) * c->plane;
    if (ph.up_work) *hi = ph.up_work + idx * ph.up_nloc;
    return true;
}

}  // namespace fdfd
