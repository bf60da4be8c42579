// scg_comm.cu -- the library's own communicator: NCCL point-to-point calls issued from C++ on a library-owned
// communication stream.  The reference has no communication layer at all; this is the B200-side
// addition the domain decomposition (scg_md_halo_*, scg_md_march_dist) runs on: one process per GPU, paired
// ncclSend / ncclRecv inside one ncclGroupStart / ncclGroupEnd per exchange, over NVLink.
//
// libnccl is resolved at run time (dlopen): the library links neither NCCL nor torch.  Only the rendezvous of the
// 128-byte unique id is left to the caller (any out-of-band channel: a file, a socket, torch.distributed, MPI).
#include "scg_common.cuh"

#include <dlfcn.h>

#include <cstring>

namespace scg
{

namespace
{

struct NcclId // ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
{
    char internal[128];
};
using nccl_comm = void *;
constexpr int NCCL_FLOAT64 = 8; // ncclDataType_t: ncclFloat64 / ncclDouble

struct Nccl
{
    void * lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    int (*Send)(void const *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    char const * (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

Nccl g_nccl;
nccl_comm g_comm = nullptr;
int g_rank = -1, g_nranks = 0;
cudaStream_t g_comm_stream = nullptr;

template <typename F>
bool sym(F & f, char const * name)
{
    f = reinterpret_cast<F>(dlsym(g_nccl.lib, name));
    return nullptr != f;
}

int load_nccl(char const * path)
{
    if (g_nccl.lib) { return 0; }
    char const * cand[] = {path, "libnccl.so.2", "libnccl.so"};
    for (char const * c : cand)
    {
        if (!c || !c[0]) { continue; }
        // a copy already mapped into the process (e.g. by torch) is reused: same soname, same instance
        g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL | RTLD_NOLOAD);
        if (!g_nccl.lib) { g_nccl.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL); }
        if (g_nccl.lib) { break; }
    }
    if (!g_nccl.lib) { return fail("scg_comm: libnccl not found (%s); pass its path to scg_comm_unique_id / scg_comm_init", dlerror()); }
    bool const ok = sym(g_nccl.GetUniqueId, "ncclGetUniqueId") && sym(g_nccl.CommInitRank, "ncclCommInitRank") &&
                    sym(g_nccl.CommDestroy, "ncclCommDestroy") && sym(g_nccl.Send, "ncclSend") && sym(g_nccl.Recv, "ncclRecv") &&
                    sym(g_nccl.GroupStart, "ncclGroupStart") && sym(g_nccl.GroupEnd, "ncclGroupEnd") &&
                    sym(g_nccl.GetErrorString, "ncclGetErrorString") && sym(g_nccl.GetVersion, "ncclGetVersion");
    if (!ok)
    {
        g_nccl = Nccl();
        return fail("scg_comm: libnccl lacks a required symbol");
    }
    return 0;
}

} // namespace

#define SCG_NCCL(expr)                                                                                         \
    do                                                                                                         \
    {                                                                                                          \
        int const r__ = (expr);                                                                                \
        if (0 != r__) { return ::scg::fail("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); } \
    } while (0)

bool comm_ready() { return nullptr != g_comm; }
int comm_rank() { return g_rank; }
int comm_nranks() { return g_nranks; }
cudaStream_t comm_stream() { return g_comm_stream; }

int comm_group_start()
{
    SCG_NCCL(g_nccl.GroupStart());
    return 0;
}
int comm_group_end()
{
    SCG_NCCL(g_nccl.GroupEnd());
    return 0;
}
int comm_send(double const * buf, int64_t count, int peer)
{
    SCG_NCCL(g_nccl.Send(buf, static_cast<size_t>(count), NCCL_FLOAT64, peer, g_comm, g_comm_stream));
    return 0;
}
int comm_recv(double * buf, int64_t count, int peer)
{
    SCG_NCCL(g_nccl.Recv(buf, static_cast<size_t>(count), NCCL_FLOAT64, peer, g_comm, g_comm_stream));
    return 0;
}

} // namespace scg

using namespace scg;

extern "C" {

int scg_comm_unique_id(void * id, int64_t nbytes, const char * libnccl_path)
{
    SCG_REQUIRE(id && nbytes >= 128, "scg_comm_unique_id: the id buffer must hold 128 bytes");
    if (int rc = load_nccl(libnccl_path)) { return rc; }
    NcclId uid;
    SCG_NCCL(g_nccl.GetUniqueId(&uid));
    std::memcpy(id, uid.internal, 128);
    return 0;
}

int scg_comm_init(int rank, int nranks, const void * id, const char * libnccl_path)
{
    SCG_REQUIRE(id, "scg_comm_init: null id");
    SCG_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "scg_comm_init: rank %d out of [0, %d)", rank, nranks);
    SCG_REQUIRE(!g_comm, "scg_comm_init: the communicator already exists");
    if (int rc = ensure_init()) { return rc; }
    if (int rc = load_nccl(libnccl_path)) { return rc; }
    NcclId uid;
    std::memcpy(uid.internal, id, 128);
    if (!g_comm_stream) { SCG_CUDA(cudaStreamCreateWithFlags(&g_comm_stream, cudaStreamNonBlocking)); }
    SCG_NCCL(g_nccl.CommInitRank(&g_comm, nranks, uid, rank));
    g_rank = rank;
    g_nranks = nranks;
    return 0;
}

int scg_comm_info(int * rank, int * nranks, int * nccl_version)
{
    if (rank) { *rank = g_rank; }
    if (nranks) { *nranks = g_nranks; }
    if (nccl_version)
    {
        *nccl_version = 0;
        if (g_nccl.GetVersion) { g_nccl.GetVersion(nccl_version); }
    }
    return 0;
}

int scg_comm_finalize(void)
{
    if (g_comm)
    {
        if (g_comm_stream) { cudaStreamSynchronize(g_comm_stream); }
        g_nccl.CommDestroy(g_comm);
        g_comm = nullptr;
    }
    g_rank = -1;
    g_nranks = 0;
    return 0;
}

} // extern "C"
