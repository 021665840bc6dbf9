// NCCL plumbing for the z-slab decomposition: 1-plane halo send/recv between neighbouring ranks and
// allreduce of the Krylov inner products.  The reference has no distributed code
// (README.md:29-33 defers it to PETSc); this is the B200-native replacement over NVLink/NVSwitch.
//
// NCCL is resolved at run time with dlopen so that single-GPU use has no NCCL dependency and so that
// a process which already loaded torch's bundled libnccl.so.2 shares that copy.
#include <dlfcn.h>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "fdfd_internal.h"

namespace fdfd {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, int enums.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDouble
enum { ncclSum = 0 };
enum { ncclInt8 = 0 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi *load_nccl(std::string &err) {
    static NcclApi api;
    static bool tried = false, ok = false;
    if (tried) {
        if (!ok) err = "NCCL not loadable";
        return ok ? &api : nullptr;
    }
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    // RTLD_NOLOAD first: reuse the copy the host process (e.g. torch) already mapped.
    for (int pass = 0; pass < 2 && !api.lib; ++pass)
        for (int i = 0; names[i] && !api.lib; ++i)
            api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
    if (!api.lib) {
        err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
        return nullptr;
    }
#define LOADSYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));    \
    if (!api.field) { err = std::string("missing NCCL symbol ") + name; return nullptr; }
    LOADSYM(GetUniqueId, "ncclGetUniqueId")
    LOADSYM(CommInitRank, "ncclCommInitRank")
    LOADSYM(CommDestroy, "ncclCommDestroy")
    LOADSYM(AllReduce, "ncclAllReduce")
    LOADSYM(Send, "ncclSend")
    LOADSYM(Recv, "ncclRecv")
    LOADSYM(GroupStart, "ncclGroupStart")
    LOADSYM(GroupEnd, "ncclGroupEnd")
    LOADSYM(GetErrorString, "ncclGetErrorString")
#undef LOADSYM
    ok = true;
    return &api;
}

#define FDFD_NCCL(c, call)                                                                          \
    do {                                                                                             \
        int r__ = (call);                                                                            \
        if (r__ != ncclSuccess)                                                                      \
            return set_err((c), FDFD_ENCCL, std::string(#call) + ": " + (c)->nccl->GetErrorString(r__)); \
    } while (0)

void halo_neighbours(int P, int r, bool wrapz, int *up, int *dn) {
    *up = (r + 1 < P) ? r + 1 : (wrapz ? 0 : -1);
    *dn = (r > 0) ? r - 1 : (wrapz ? P - 1 : -1);
}

int comm_unique_id(char id[128], std::string &err) {
    NcclApi *api = load_nccl(err);
    if (!api) return FDFD_ENCCL;
    ncclUniqueId uid;
    int r = api->GetUniqueId(&uid);
    if (r != ncclSuccess) {
        err = std::string("ncclGetUniqueId: ") + api->GetErrorString(r);
        return FDFD_ENCCL;
    }
    std::memcpy(id, uid.internal, 128);
    return FDFD_OK;
}

int comm_init(Ctx *c, const char id[128]) {
    std::string err;
    NcclApi *api = load_nccl(err);
    if (!api) return set_err(c, FDFD_ENCCL, err);
    c->nccl = api;
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, 128);
    ncclComm_t comm = nullptr;
    FDFD_CUDA(c, cudaSetDevice(c->dev));
    FDFD_NCCL(c, api->CommInitRank(&comm, c->d.nranks, uid, c->d.rank));
    c->comm = comm;
    c->dirty = true;  // material ghost planes must be exchanged
    // Halo data plane: the copy-engine peer exchange of peer.cpp (no SM-resident collective) when every rank can map its
    // neighbours' buffers with CUDA IPC - one process per GPU on one node - else NCCL send / recv.  FDFD_PEER_HALO=0 forces
    // NCCL, FDFD_PEER_HALO=1 (or FDFD_PEER_DIRECT) makes a failure of the peer set-up an error instead of a fallback.
    const char *pe = getenv("FDFD_PEER_HALO");
    const bool required = (pe && atoi(pe) != 0) || getenv("FDFD_PEER_DIRECT") != nullptr;
    if (pe && atoi(pe) == 0 && !required) return FDFD_OK;
    if (c->d.nranks == 1 || !c->d.order_cmpfirst) return FDFD_OK;
    const int rp = peer_halo_init(c);
    // every rank must take the same data plane: agree on the outcome
    double *flag = nullptr;
    const double bad_local = (rp != FDFD_OK || !c->peer.ready) ? 1.0 : 0.0;
    cudaGetLastError();   // a failed cudaIpcOpenMemHandle (same-process handles, no peer access) is not sticky
    FDFD_CUDA(c, cudaMalloc((void **)&flag, sizeof(double)));
    FDFD_CUDA(c, cudaMemcpy(flag, &bad_local, sizeof(double), cudaMemcpyHostToDevice));
    int ra = allreduce_sum(c, flag, 1, c->stream);
    double bad = 1.0;
    if (ra == FDFD_OK) {
        FDFD_CUDA(c, cudaMemcpyAsync(&bad, flag, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        FDFD_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    cudaFree(flag);
    if (ra != FDFD_OK) return ra;
    if (bad > 0.0) {
        const std::string why = c->err;
        peer_halo_destroy(c);
        if (required) return set_err(c, FDFD_ECUDA, "peer halo exchange requested but not available on every rank: " + why);
        c->err.clear();
        if (getenv("FDFD_VERBOSE")) fprintf(stderr, "fdfd rank %d: peer halo exchange unavailable (%s) - NCCL send/recv\n", c->d.rank, why.c_str());
    } else if (getenv("FDFD_VERBOSE")) {
        fprintf(stderr, "fdfd rank %d: halo planes by copy-engine peer exchange\n", c->d.rank);
    }
    return FDFD_OK;
}

void comm_destroy(Ctx *c) {
    peer_halo_destroy(c);
    if (c->comm && c->nccl) c->nccl->CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
}

// Exchange one z-plane with each neighbour.  `first`/`last` point at this rank's first / last own
// plane pieces; npieces pieces of `count` double2 each, separated by src_stride (own) and dst_stride
// (halo buffers) elements.  The global z boundary wraps (rank 0 <-> rank P-1) only for Bloch.
static int exchange_planes(Ctx *c, const double2 *first, const double2 *last, int64_t src_stride, double2 *lo,
                           double2 *hi, int64_t dst_stride, int npieces, int64_t count, cudaStream_t s) {
    int up, dn;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    ncclComm_t comm = (ncclComm_t)c->comm;
    NcclApi *n = c->nccl;
    // Message order matters when up == dn (P == 2 with Bloch wrap): NCCL matches several send/recv
    // between the same pair in issue order, so every rank issues "my last plane -> up / my lo <- dn"
    // first and "my first plane -> dn / my hi <- up" second.
    FDFD_NCCL(c, n->GroupStart());
    for (int k = 0; k < npieces; ++k) {
        const size_t nd = (size_t)count * 2;  // doubles
        if (up >= 0) FDFD_NCCL(c, n->Send(last + k * src_stride, nd, ncclFloat64, up, comm, s));
        if (dn >= 0) FDFD_NCCL(c, n->Recv(lo + k * dst_stride, nd, ncclFloat64, dn, comm, s));
    }
    for (int k = 0; k < npieces; ++k) {
        const size_t nd = (size_t)count * 2;
        if (dn >= 0) FDFD_NCCL(c, n->Send(first + k * src_stride, nd, ncclFloat64, dn, comm, s));
        if (up >= 0) FDFD_NCCL(c, n->Recv(hi + k * dst_stride, nd, ncclFloat64, up, comm, s));
    }
    FDFD_NCCL(c, n->GroupEnd());
    return FDFD_OK;
}

int halo_exchange(Ctx *c, const double2 *v, double2 *lo, double2 *hi, cudaStream_t s) {
    if (!c->comm) return set_err(c, FDFD_ESTATE, "nranks > 1 but fdfd_comm_init was not called");
    const int64_t Nxy = c->d.N[0] * c->d.N[1];
    const int64_t nzl = c->k1 - c->k0;
    if (c->d.order_cmpfirst) {
        const int64_t pl = 3 * Nxy;
        if (c->peer.ready && lo == c->halo_lo && hi == c->halo_hi) return peer_halo_exchange(c, v, v + (nzl - 1) * pl, s);
        return exchange_planes(c, v, v + (nzl - 1) * pl, 0, lo, hi, 0, 1, pl, s);
    }
    // component-major: 3 pieces of Nx*Ny, component stride Nxy*nzl in the slab, Nxy in the halo buffer
    return exchange_planes(c, v, v + (nzl - 1) * Nxy, Nxy * nzl, lo, hi, Nxy, 3, Nxy, s);
}

// ghost planes of a ghosted SoA material array (nzl+2 planes of Nxy): own planes are 1..nzl
int halo_exchange_ghosted(Ctx *c, double2 *arr, cudaStream_t s) {
    const int64_t Nxy = c->d.N[0] * c->d.N[1];
    const int64_t nzl = c->k1 - c->k0;
    return exchange_planes(c, arr + Nxy, arr + nzl * Nxy, 0, arr, arr + (nzl + 1) * Nxy, 0, 1, Nxy, s);
}

// cuStreamWriteValue32 resolved from the driver library at run time (the library links cudart statically only)
typedef int (*cuStreamWriteValue32_t)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
static cuStreamWriteValue32_t load_write_value() {
    static cuStreamWriteValue32_t fn = [] {
        void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return (cuStreamWriteValue32_t) nullptr;
        void *f = dlsym(lib, "cuStreamWriteValue32_v2");
        if (!f) f = dlsym(lib, "cuStreamWriteValue32");
        return reinterpret_cast<cuStreamWriteValue32_t>(f);
    }();
    return fn;
}

bool stream_write_u32_available() { return load_write_value() != nullptr; }

typedef int (*cuStreamWaitValue32_t)(cudaStream_t, unsigned long long, uint32_t, unsigned int);
int stream_wait_geq_u32(Ctx *c, cudaStream_t s, uint32_t *dev_word, uint32_t value) {
    static cuStreamWaitValue32_t fn = [] {
        void *lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) return (cuStreamWaitValue32_t) nullptr;
        void *f = dlsym(lib, "cuStreamWaitValue32_v2");
        if (!f) f = dlsym(lib, "cuStreamWaitValue32");
        return reinterpret_cast<cuStreamWaitValue32_t>(f);
    }();
    if (!fn) return set_err(c, FDFD_ESTATE, "cuStreamWaitValue32 not available");
    const int r = fn(s, (unsigned long long)(uintptr_t)dev_word, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (r != 0) return set_err(c, FDFD_ECUDA, "cuStreamWaitValue32 failed (" + std::to_string(r) + ")");
    return FDFD_OK;
}

// Setup-time byte exchange with the z-neighbours (IPC handles): same pairing and order as exchange_planes.
int comm_exchange_bytes(Ctx *c, const void *dev_mine, void *dev_from_up, void *dev_from_dn, size_t bytes, cudaStream_t s) {
    if (!c->comm) return set_err(c, FDFD_ESTATE, "comm_exchange_bytes: no communicator");
    int up, dn;
    halo_neighbours(c->d.nranks, c->d.rank, c->d.isbloch[2] != 0, &up, &dn);
    ncclComm_t comm = (ncclComm_t)c->comm;
    NcclApi *n = c->nccl;
    FDFD_NCCL(c, n->GroupStart());
    if (up >= 0) FDFD_NCCL(c, n->Send(dev_mine, bytes, ncclInt8, up, comm, s));
    if (dn >= 0) FDFD_NCCL(c, n->Recv(dev_from_dn, bytes, ncclInt8, dn, comm, s));
    if (dn >= 0) FDFD_NCCL(c, n->Send(dev_mine, bytes, ncclInt8, dn, comm, s));
    if (up >= 0) FDFD_NCCL(c, n->Recv(dev_from_up, bytes, ncclInt8, up, comm, s));
    FDFD_NCCL(c, n->GroupEnd());
    return FDFD_OK;
}

int stream_write_u32(Ctx *c, cudaStream_t s, uint32_t *dev_word, uint32_t value) {
    cuStreamWriteValue32_t fn = load_write_value();
    if (!fn) return set_err(c, FDFD_ESTATE, "cuStreamWriteValue32 not available");
    const int r = fn(s, (unsigned long long)(uintptr_t)dev_word, value, 0u);
    if (r != 0) return set_err(c, FDFD_ECUDA, "cuStreamWriteValue32 failed (" + std::to_string(r) + ")");
    return FDFD_OK;
}

int allreduce_sum(Ctx *c, double *dev, int count, cudaStream_t s) {
    if (c->d.nranks == 1) return FDFD_OK;
    if (c->comm_pending) {   // keep NCCL operations on this communicator totally ordered
        FDFD_CUDA(c, cudaStreamWaitEvent(s, c->ev_halo, 0));
        c->comm_pending = false;
    }
    if (!c->comm) return set_err(c, FDFD_ESTATE, "nranks > 1 but fdfd_comm_init was not called");
    FDFD_NCCL(c, c->nccl->AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)c->comm, s));
    return FDFD_OK;
}

}  // namespace fdfd
