// fake_nccl.cpp -- TEST INFRASTRUCTURE: the ten NCCL entry points slab_nccl.cu resolves with dlopen, implemented over
// POSIX shared memory between the rank PROCESSES of a CPU test (tests/test_emu_slabs.py), so that the C++ sequencer
// osph_slab_run can be executed without GPUs.  Collectives run synchronously at the call (the emulated stream is
// synchronous too); doubles and ncclMin / ncclSum only, which is all the sequencer uses.  Selected with
// OSPH_NCCL_LIB=<this library>; never part of the product.
#include <fcntl.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <string>
#include <vector>

#include <nccl.h>

namespace {
const int MAX_RANKS = 16, SLOT_DOUBLES = 64;
struct Board {                                   // one per communicator, in shared memory
    volatile unsigned long long arrive[MAX_RANKS];    // per-rank sequence number of the last collective entered
    volatile unsigned long long leave[MAX_RANKS];
    double slot[MAX_RANKS][SLOT_DOUBLES];
};
struct Comm { Board *b; int rank, world; unsigned long long seq; std::string name; unsigned long long p2p_seq[MAX_RANKS][2]; };
struct Op { bool send; void *buf; size_t bytes; int peer; Comm *c; };
std::vector<Op> g_ops;
int g_group = 0;

void wait_all(volatile unsigned long long *a, int world, unsigned long long seq)
{
    for (int r = 0; r < world; r++) while (a[r] < seq) sched_yield();
}

std::string p2p_name(const Comm *c, int src, int dst, unsigned long long k)
{
    char s[96]; snprintf(s, sizeof s, "%s_%d_%d_%llu", c->name.c_str(), src, dst, k); return s;
}

ncclResult_t do_send(const Op &o)
{
    Comm *c = o.c;
    const std::string nm = p2p_name(c, c->rank, o.peer, c->p2p_seq[o.peer][0]++);
    const std::string tmp = nm + "_w";
    int fd = shm_open(tmp.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) return ncclSystemError;
    const size_t total = o.bytes + 8;
    if (ftruncate(fd, (off_t)total) != 0) { close(fd); return ncclSystemError; }
    unsigned char *m = (unsigned char *)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return ncclSystemError;
    memcpy(m, &o.bytes, 8); memcpy(m + 8, o.buf, o.bytes);
    munmap(m, total);
    // publish atomically: the receiver looks for the final name only
    char a[128], b[128]; snprintf(a, sizeof a, "/dev/shm%s", tmp.c_str()); snprintf(b, sizeof b, "/dev/shm%s", nm.c_str());
    return rename(a, b) == 0 ? ncclSuccess : ncclSystemError;
}

ncclResult_t do_recv(const Op &o)
{
    Comm *c = o.c;
    const std::string nm = p2p_name(c, o.peer, c->rank, c->p2p_seq[o.peer][1]++);
    int fd = -1;
    while ((fd = shm_open(nm.c_str(), O_RDWR, 0600)) < 0) sched_yield();
    size_t bytes = 0;
    if (read(fd, &bytes, 8) != 8 || bytes != o.bytes) { close(fd); fprintf(stderr, "fake nccl: message size mismatch\n"); return ncclInvalidArgument; }
    unsigned char *m = (unsigned char *)mmap(nullptr, bytes + 8, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return ncclSystemError;
    memcpy(o.buf, m + 8, bytes);
    munmap(m, bytes + 8);
    shm_unlink(nm.c_str());
    return ncclSuccess;
}

ncclResult_t flush()
{
    ncclResult_t rc = ncclSuccess;
    for (const Op &o : g_ops) if (o.send && rc == ncclSuccess) rc = do_send(o);      // sends never block
    for (const Op &o : g_ops) if (!o.send && rc == ncclSuccess) rc = do_recv(o);
    g_ops.clear();
    return rc;
}
}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
    memset(id, 0, sizeof *id);
    static int counter = 0;
    snprintf(id->internal, sizeof id->internal, "/osph_fnccl_%d_%d", (int)getpid(), counter++);
    return ncclSuccess;
}

ncclResult_t ncclCommInitRank(ncclComm_t *out, int world, ncclUniqueId id, int rank)
{
    if (world > MAX_RANKS) return ncclInvalidArgument;
    Comm *c = new Comm();
    c->rank = rank; c->world = world; c->seq = 0; c->name = id.internal;
    memset(c->p2p_seq, 0, sizeof c->p2p_seq);
    int fd = shm_open(c->name.c_str(), O_CREAT | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, sizeof(Board)) != 0) return ncclSystemError;       // new objects are zero-filled
    c->b = (Board *)mmap(nullptr, sizeof(Board), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (c->b == MAP_FAILED) return ncclSystemError;
    *out = reinterpret_cast<ncclComm_t>(c);
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm)
{
    Comm *c = reinterpret_cast<Comm *>(comm);
    munmap(c->b, sizeof(Board));
    if (c->rank == 0) shm_unlink(c->name.c_str());
    delete c;
    return ncclSuccess;
}

static ncclResult_t exchange(Comm *c, const void *send, size_t count)
{
    if (count > (size_t)SLOT_DOUBLES) return ncclInvalidArgument;
    const unsigned long long s = ++c->seq;
    wait_all(c->b->leave, c->world, s - 1);                  // nobody is still reading the slots of the collective before
    memcpy((void *)c->b->slot[c->rank], send, count * sizeof(double));
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    c->b->arrive[c->rank] = s;
    wait_all(c->b->arrive, c->world, s);
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t, ncclRedOp_t op, ncclComm_t comm, cudaStream_t)
{
    Comm *c = reinterpret_cast<Comm *>(comm);
    if (t != ncclDouble || (op != ncclMin && op != ncclSum)) return ncclInvalidArgument;
    ncclResult_t rc = exchange(c, send, count);
    if (rc != ncclSuccess) return rc;
    double *out = (double *)recv;
    for (size_t k = 0; k < count; k++) {
        double v = c->b->slot[0][k];
        for (int r = 1; r < c->world; r++) v = op == ncclMin ? (c->b->slot[r][k] < v ? c->b->slot[r][k] : v) : v + c->b->slot[r][k];
        out[k] = v;
    }
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    c->b->leave[c->rank] = c->seq;
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t comm, cudaStream_t)
{
    Comm *c = reinterpret_cast<Comm *>(comm);
    if (t != ncclDouble) return ncclInvalidArgument;
    ncclResult_t rc = exchange(c, send, count);
    if (rc != ncclSuccess) return rc;
    for (int r = 0; r < c->world; r++) memcpy((double *)recv + (size_t)r * count, (const void *)c->b->slot[r], count * sizeof(double));
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    c->b->leave[c->rank] = c->seq;
    return ncclSuccess;
}

ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t)
{
    if (t != ncclDouble) return ncclInvalidArgument;
    g_ops.push_back(Op{true, const_cast<void *>(buf), count * sizeof(double), peer, reinterpret_cast<Comm *>(comm)});
    return g_group ? ncclSuccess : flush();
}
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t comm, cudaStream_t)
{
    if (t != ncclDouble) return ncclInvalidArgument;
    g_ops.push_back(Op{false, buf, count * sizeof(double), peer, reinterpret_cast<Comm *>(comm)});
    return g_group ? ncclSuccess : flush();
}
ncclResult_t ncclGroupStart() { g_group++; return ncclSuccess; }
ncclResult_t ncclGroupEnd() { return --g_group == 0 ? flush() : ncclSuccess; }
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "fake nccl error"; }

}  // extern "C"
