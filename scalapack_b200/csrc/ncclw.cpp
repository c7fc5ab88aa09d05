// ncclw.cpp -- run-time binding to NCCL (the torch-bundled 2.28.x or the system 2.27.x libnccl.so.2).
#include "ncclw.h"
#include "common.h"

#include <dlfcn.h>
#include <glob.h>

namespace slb {

typedef int ncclResult_t_;
struct ncclUniqueId_ { char internal[128]; };

static struct {
    void *h = nullptr;
    ncclResult_t_ (*GetVersion)(int *);
    ncclResult_t_ (*GetUniqueId)(ncclUniqueId_ *);
    ncclResult_t_ (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int);
    ncclResult_t_ (*CommSplit)(ncclComm_t_, int, int, ncclComm_t_ *, void *);
    ncclResult_t_ (*CommDestroy)(ncclComm_t_);
    const char *(*GetErrorString)(ncclResult_t_);
    ncclResult_t_ (*GroupStart)();
    ncclResult_t_ (*GroupEnd)();
    ncclResult_t_ (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t);
    ncclResult_t_ (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
    char version[32];
} N;

#define NCHECK(call)                                                                                   \
    do {                                                                                               \
        ncclResult_t_ r_ = (call);                                                                     \
        if (r_ != 0) fatal("NCCL error at %s:%d: %s", __FILE__, __LINE__, N.GetErrorString ? N.GetErrorString(r_) : "?"); \
    } while (0)

static void load()
{
    if (N.h) return;
    const char *cands[8]; int nc = 0;
    const char *env = getenv("SLB200_NCCL_LIB");
    if (env && *env) cands[nc++] = env;
    // an already-loaded copy first (e.g. torch's) so one process never holds two NCCLs
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) {
        for (int i = 0; i < nc && !h; ++i) h = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    }
    if (!h) {
        glob_t gl; memset(&gl, 0, sizeof(gl));
        const char *pats[] = { "/opt/prime-rl/.venv/lib/python3*/site-packages/nvidia/nccl/lib/libnccl.so.2",
                               "/usr/lib/x86_64-linux-gnu/libnccl.so.2", "/usr/local/cuda/lib64/libnccl.so.2" };
        for (const char *p : pats) {
            if (h) break;
            if (glob(p, 0, nullptr, &gl) == 0) for (size_t i = 0; i < gl.gl_pathc && !h; ++i) h = dlopen(gl.gl_pathv[i], RTLD_NOW | RTLD_GLOBAL);
            globfree(&gl);
        }
    }
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) fatal("cannot load libnccl.so.2 (%s); set SLB200_NCCL_LIB", dlerror());
    N.h = h;
#define SYM(field, name) do { *(void **)(&N.field) = dlsym(h, name); if (!N.field) fatal("NCCL symbol %s missing", name); } while (0)
    SYM(GetVersion, "ncclGetVersion"); SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommSplit, "ncclCommSplit"); SYM(CommDestroy, "ncclCommDestroy"); SYM(GetErrorString, "ncclGetErrorString");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Broadcast, "ncclBroadcast");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllGather, "ncclAllGather"); SYM(AllReduce, "ncclAllReduce");
#undef SYM
    int v = 0; N.GetVersion(&v);
    snprintf(N.version, sizeof(N.version), "%d.%d.%d", v / 10000, (v / 100) % 100, v % 100);
}

const char *nccl_version_string() { load(); return N.version; }

NcclComms *nccl_create(Grid *g)
{
    load();
    rt();   // make sure the CUDA device is bound before NCCL touches it
    NcclComms *c = new NcclComms();
    int np = g->nprow * g->npcol, me = g->myrow * g->npcol + g->mycol;
    ncclUniqueId_ id; memset(&id, 0, sizeof(id));
    if (me == 0) NCHECK(N.GetUniqueId(&id));
    std::vector<ncclUniqueId_> ids((size_t)np);
    grid_allgather(g, 'A', &id, ids.data(), sizeof(id));
    NCHECK(N.CommInitRank(&c->all, np, ids[0], me));
    // same colours / keys as blacs_map_.c:114,118
    NCHECK(N.CommSplit(c->all, g->myrow, g->mycol, &c->row, nullptr));
    NCHECK(N.CommSplit(c->all, g->mycol, g->myrow, &c->col, nullptr));
    NCHECK(N.CommSplit(c->all, g->mycol, g->myrow, &c->colp, nullptr));
    vlog(1, "NCCL %s communicators up: grid %dx%d me=(%d,%d)", N.version, g->nprow, g->npcol, g->myrow, g->mycol);
    return c;
}

void nccl_destroy(NcclComms *c)
{
    if (!c) return;
    if (c->row) N.CommDestroy(c->row);
    if (c->col) N.CommDestroy(c->col);
    if (c->colp) N.CommDestroy(c->colp);
    if (c->all) N.CommDestroy(c->all);
    delete c;
}

void nccl_group_start() { NCHECK(N.GroupStart()); }
void nccl_group_end() { NCHECK(N.GroupEnd()); }
void nccl_bcast(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int root, cudaStream_t s)
{ NCHECK(N.Broadcast(buf, buf, count, (int)t, root, comm, s)); counter_add("nccl_calls", 1); }
void nccl_send(ncclComm_t_ comm, const void *buf, size_t count, NcclType t, int peer, cudaStream_t s)
{ NCHECK(N.Send(buf, count, (int)t, peer, comm, s)); counter_add("nccl_calls", 1); }
void nccl_recv(ncclComm_t_ comm, void *buf, size_t count, NcclType t, int peer, cudaStream_t s)
{ NCHECK(N.Recv(buf, count, (int)t, peer, comm, s)); counter_add("nccl_calls", 1); }
void nccl_allgather(ncclComm_t_ comm, const void *send, void *recv, size_t sendcount, NcclType t, cudaStream_t s)
{ NCHECK(N.AllGather(send, recv, sendcount, (int)t, comm, s)); counter_add("nccl_calls", 1); }
// ncclRedOp_t: sum=0 prod=1 max=2 min=3
void nccl_allreduce_min_i32(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s)
{ NCHECK(N.AllReduce(send, recv, count, (int)NT_I32, 3, comm, s)); counter_add("nccl_calls", 1); }
void nccl_allreduce_sum_f64(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s)
{ NCHECK(N.AllReduce(send, recv, count, (int)NT_F64, 0, comm, s)); counter_add("nccl_calls", 1); }
void nccl_allreduce_max_f64(ncclComm_t_ comm, const void *send, void *recv, size_t count, cudaStream_t s)
{ NCHECK(N.AllReduce(send, recv, count, (int)NT_F64, 2, comm, s)); counter_add("nccl_calls", 1); }

void nccl_alltoallv(ncclComm_t_ comm, int np, int me, const void *send, const size_t *scount, const size_t *sdispl, void *recv,
                    const size_t *rcount, const size_t *rdispl, cudaStream_t s)
{
    if (scount[me] != rcount[me]) fatal("nccl_alltoallv: own part %zu != %zu", scount[me], rcount[me]);
    if (scount[me]) SLB_CUDA(cudaMemcpyAsync((char *)recv + rdispl[me], (const char *)send + sdispl[me], scount[me], cudaMemcpyDeviceToDevice, s));
    if (np <= 1) return;
    NCHECK(N.GroupStart());
    for (int p = 0; p < np; ++p) {
        if (p == me) continue;
        if (scount[p]) NCHECK(N.Send((const char *)send + sdispl[p], scount[p], (int)NT_U8, p, comm, s));
        if (rcount[p]) NCHECK(N.Recv((char *)recv + rdispl[p], rcount[p], (int)NT_U8, p, comm, s));
    }
    NCHECK(N.GroupEnd());
    counter_add("nccl_calls", 1);
}

}  // namespace slb
