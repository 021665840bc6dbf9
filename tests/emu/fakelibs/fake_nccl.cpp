// TEST INFRASTRUCTURE ONLY - stand-in for libnccl.so.2 used by the multi-rank runs of the CPU logic-check build
// (tests/emu/README.md): one PROCESS per rank as on the GPU box, messages over named FIFOs in a rendezvous
// directory whose name travels in the 128-byte unique id.  Implements exactly the entry points csrc/comm.cpp resolves.
// Grouped sends run on a helper thread while the calling thread receives, so a symmetric exchange cannot dead-lock on
// pipe capacity; messages between one ordered pair of ranks keep their issue order, as NCCL guarantees.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Comm {
    std::string dir;
    int nranks = 0, rank = 0;
    std::map<int, int> wfd, rfd;   // peer -> fd
};

struct Op {
    bool send;
    void *buf;
    size_t bytes;
    int peer;
    Comm *comm;
};

thread_local int group_depth = 0;
thread_local std::vector<Op> group_ops;

size_t dtype_size(int dt) {
    switch (dt) {
        case 0: case 1: return 1;      // int8 / uint8
        case 2: case 3: case 7: return 4;   // int32 / uint32 / float32
        case 4: case 5: case 8: return 8;   // int64 / uint64 / float64
        case 6: return 2;              // float16
        default: return 0;
    }
}

std::string fifo_path(Comm *c, int src, int dst) { return c->dir + "/" + std::to_string(src) + "_" + std::to_string(dst); }

int get_fd(Comm *c, int peer, bool send) {
    auto &m = send ? c->wfd : c->rfd;
    auto it = m.find(peer);
    if (it != m.end()) return it->second;
    const std::string p = send ? fifo_path(c, c->rank, peer) : fifo_path(c, peer, c->rank);
    if (mkfifo(p.c_str(), 0600) != 0 && errno != EEXIST) { perror("fake_nccl mkfifo"); abort(); }
    const int fd = open(p.c_str(), send ? O_WRONLY : O_RDONLY);   // blocks until the other side opens too
    if (fd < 0) { perror("fake_nccl open"); abort(); }
    m[peer] = fd;
    return fd;
}

void xfer(const Op &op) {
    const int fd = get_fd(op.comm, op.peer, op.send);
    char *p = static_cast<char *>(op.buf);
    size_t left = op.bytes;
    while (left) {
        const ssize_t n = op.send ? write(fd, p, left) : read(fd, p, left);
        if (n < 0 && errno == EINTR) continue;
        if (n <= 0) { fprintf(stderr, "fake_nccl: peer %d closed the pipe\n", op.peer); abort(); }
        p += n;
        left -= (size_t)n;
    }
}

void run_ops(std::vector<Op> ops) {
    std::vector<Op> sends, recvs;
    for (const Op &o : ops) (o.send ? sends : recvs).push_back(o);
    std::thread t([&sends] { for (const Op &o : sends) xfer(o); });
    for (const Op &o : recvs) xfer(o);
    t.join();
}

void submit(const Op &op) {
    if (group_depth > 0) group_ops.push_back(op);
    else run_ops({op});
}

}  // namespace

extern "C" {

typedef struct { char internal[128]; } ncclUniqueId;

int ncclGetUniqueId(ncclUniqueId *id) {
    std::memset(id->internal, 0, 128);
    const char *base = getenv("EMU_DIST_DIR");     // the test runner's scratch directory (removed by it afterwards)
    std::string tmpl = std::string(base ? base : "/tmp") + "/nccl_XXXXXX";
    if (tmpl.size() > 120 || !mkdtemp(&tmpl[0])) return 1;
    std::strncpy(id->internal, tmpl.c_str(), 127);
    return 0;
}
int ncclCommInitRank(void **comm, int nranks, ncclUniqueId id, int rank) {
    Comm *c = new Comm;
    c->dir = id.internal;
    c->nranks = nranks;
    c->rank = rank;
    *comm = c;
    return 0;
}
int ncclCommDestroy(void *comm) {
    Comm *c = static_cast<Comm *>(comm);
    for (auto &kv : c->wfd) close(kv.second);
    for (auto &kv : c->rfd) close(kv.second);
    delete c;
    return 0;
}
int ncclGroupStart() { ++group_depth; return 0; }
int ncclGroupEnd() {
    if (--group_depth == 0) {
        std::vector<Op> ops;
        ops.swap(group_ops);
        run_ops(ops);
    }
    return 0;
}
int ncclSend(const void *buf, size_t count, int dt, int peer, void *comm, void *) {
    submit(Op{true, const_cast<void *>(buf), count * dtype_size(dt), peer, static_cast<Comm *>(comm)});
    return 0;
}
int ncclRecv(void *buf, size_t count, int dt, int peer, void *comm, void *) {
    submit(Op{false, buf, count * dtype_size(dt), peer, static_cast<Comm *>(comm)});
    return 0;
}
// sum of doubles: gathered at rank 0, added in rank order, broadcast
int ncclAllReduce(const void *sendbuf, void *recvbuf, size_t count, int dt, int op, void *comm, void *) {
    Comm *c = static_cast<Comm *>(comm);
    if (dt != 8 || op != 0) return 4;
    const size_t bytes = count * sizeof(double);
    std::vector<double> acc(static_cast<const double *>(sendbuf), static_cast<const double *>(sendbuf) + count);
    if (c->rank == 0) {
        std::vector<double> tmp(count);
        for (int r = 1; r < c->nranks; ++r) {
            xfer(Op{false, tmp.data(), bytes, r, c});
            for (size_t i = 0; i < count; ++i) acc[i] += tmp[i];
        }
        for (int r = 1; r < c->nranks; ++r) xfer(Op{true, acc.data(), bytes, r, c});
    } else {
        xfer(Op{true, acc.data(), bytes, 0, c});
        xfer(Op{false, acc.data(), bytes, 0, c});
    }
    std::memcpy(recvbuf, acc.data(), bytes);
    return 0;
}
const char *ncclGetErrorString(int) { return "fake_nccl error"; }

}  // extern "C"
