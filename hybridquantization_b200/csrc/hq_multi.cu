// hq_multi.cu — the exchange step of the sharded path INSIDE the library (north_star subsystem 4):
//   * a native NCCL communicator per context, for one process per GPU (hq_comm_get_unique_id / hq_comm_init_rank:
//     the host only ships the 128 id bytes between its ranks) — replaces the torch.distributed hook of round 1;
//   * a single-process multi-device context (hq_create_multi): what a JVM host gets instead of the reference's
//     JavaCL.createBestContext() + one queue (ImageManipulation.java:58-59).  The leader splits the image rows over
//     its members (halo rows for the S-CIELAB stage included), every evaluation is enqueued on every member's
//     stream, and ONE grouped ncclAllReduce(int64, sum) of the result words follows.
// The sums are exact integers, so totals and the annealing trajectory do not depend on the number of GPUs.
//
// NCCL is resolved at run time (dlopen: the copy already in the process — e.g. torch's — else the system
// libnccl.so.2): the library still loads, and fails loudly only in these entries, on a machine without NCCL.
#include <dlfcn.h>

#include <mutex>

#include "hq_ctx.h"

namespace {

// the slice of nccl.h (NCCL 2.x, stable ABI) this file uses
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
static_assert(sizeof(ncclUniqueId) == HQ_COMM_ID_BYTES, "hq_b200.h states the id size");
enum { kNcclSuccess = 0, kNcclInt64 = 4, kNcclSum = 0 };

struct Nccl {
    void* handle = nullptr;
    int (*GetVersion)(int*) = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};

Nccl* nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = std::getenv("HQ_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* name : names) {
            if (!name || !*name) continue;
            n.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);             // the copy this process already holds (torch's)
            if (!n.handle) n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) { n.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* s) { void* p = dlsym(n.handle, s); if (!p && n.error.empty()) n.error = std::string("NCCL symbol missing: ") + s; return p; };
        n.GetVersion = reinterpret_cast<decltype(n.GetVersion)>(sym("ncclGetVersion"));
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(sym("ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(sym("ncclCommInitRank"));
        n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(sym("ncclAllReduce"));
        n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
        n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return &n;
}

int nccl_fail(hq_ctx* c, const char* what, int rc) {
    Nccl* n = nccl();
    return hqi::fail(c, HQ_ERR_CUDA, "%s failed: %s", what, (n->GetErrorString && rc >= 0) ? n->GetErrorString(rc) : n->error.c_str());
}
#define HQ_NCCL(c, call)                                                  \
    do {                                                                  \
        const int r__ = (call);                                           \
        if (r__ != kNcclSuccess) return nccl_fail((c), #call, r__);      \
    } while (0)

int need_nccl(hq_ctx* c) {
    Nccl* n = nccl();
    if (!n->handle || !n->error.empty()) return hqi::fail(c, HQ_ERR_UNSUPPORTED, "%s", n->error.c_str());
    return HQ_OK;
}

}  // namespace

namespace hqi {

bool reduces(const hq_ctx* c) { return c->allreduce != nullptr || c->comm != nullptr || c->is_multi(); }

// one rank per process: the caller's hook wins, else the context's own communicator
int reduce_words(hq_ctx* c, unsigned long long* d_words, size_t n_words, cudaStream_t st) {
    if (c->allreduce) {
        if (c->allreduce(c->allreduce_user, d_words, n_words, st) != 0) return fail(c, HQ_ERR_CALLBACK, "all-reduce hook failed");
        return HQ_OK;
    }
    if (c->comm) HQ_NCCL(c, nccl()->AllReduce(d_words, d_words, n_words, kNcclInt64, kNcclSum, static_cast<ncclComm_t>(c->comm), st));
    return HQ_OK;
}

// single process, several devices: bufs[i] on member i, one grouped collective, each on its member's stream
int group_reduce(hq_ctx* leader, const std::vector<unsigned long long*>& bufs, size_t n_words) {
    Nccl* n = nccl();
    HQ_NCCL(leader, n->GroupStart());
    for (size_t i = 0; i < leader->members.size(); ++i) {
        hq_ctx* m = leader->members[i];
        const int r = n->AllReduce(bufs[i], bufs[i], n_words, kNcclInt64, kNcclSum, static_cast<ncclComm_t>(m->comm), m->stream);
        if (r != kNcclSuccess) { n->GroupEnd(); return nccl_fail(leader, "ncclAllReduce", r); }
    }
    HQ_NCCL(leader, n->GroupEnd());
    return HQ_OK;
}

void comm_release(hq_ctx* c) {
    if (c->comm && nccl()->CommDestroy) nccl()->CommDestroy(static_cast<ncclComm_t>(c->comm));
    c->comm = nullptr; c->comm_rank = 0; c->comm_size = 1;
}

}  // namespace hqi

using namespace hqi;

extern "C" {

int hq_comm_get_unique_id(void* id128) {
    if (!id128) return fail(nullptr, HQ_ERR_INVALID, "id buffer is NULL");
    int rc = need_nccl(nullptr); if (rc) return rc;
    ncclUniqueId id;
    HQ_NCCL(nullptr, nccl()->GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return HQ_OK;
}

int hq_comm_init_rank(hq_ctx* c, const void* id128, int nranks, int rank) {
    if (!c || !id128) return c ? fail(c, HQ_ERR_INVALID, "id buffer is NULL") : HQ_ERR_INVALID;
    if (c->is_multi() || c->leader) return fail(c, HQ_ERR_UNSUPPORTED, "a multi-device context already owns its communicator");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, HQ_ERR_INVALID, "rank %d of %d", rank, nranks);
    int rc = need_nccl(c); if (rc) return rc;
    rc = bind_device(c); if (rc) return rc;
    comm_release(c);
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    HQ_NCCL(c, nccl()->CommInitRank(&comm, nranks, id, rank));
    c->comm = comm; c->comm_rank = rank; c->comm_size = nranks;
    return HQ_OK;
}

int hq_comm_allreduce(hq_ctx* c, void* d_words, size_t n_words, void* stream) {
    if (!c || !d_words) return c ? fail(c, HQ_ERR_INVALID, "NULL device buffer") : HQ_ERR_INVALID;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "device buffers belong to one device: a multi-device context reduces inside hq_eval_palettes");
    int rc = bind_device(c); if (rc) return rc;
    return reduce_words(c, static_cast<unsigned long long*>(d_words), n_words, stream ? static_cast<cudaStream_t>(stream) : c->stream);
}

int hq_comm_info(const hq_ctx* c, int* rank, int* size, int* nccl_version) {
    if (!c) return HQ_ERR_INVALID;
    if (rank) *rank = c->comm_rank;
    if (size) *size = c->is_multi() ? (int)c->members.size() : c->comm_size;
    if (nccl_version) {
        *nccl_version = 0;
        Nccl* n = nccl();
        if (n->handle && n->GetVersion) n->GetVersion(nccl_version);
    }
    return HQ_OK;
}

int hq_multi_device_count(const hq_ctx* c) { return c ? (c->is_multi() ? (int)c->members.size() : 1) : 0; }

int hq_create_multi(const int* devices, int ndev, hq_ctx** out) {
    if (!out) return fail(nullptr, HQ_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || ndev < 1 || ndev > 64) return fail(nullptr, HQ_ERR_INVALID, "device list of 1..64 entries required (got %d)", ndev);
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return fail(nullptr, HQ_ERR_INVALID, "device %d listed twice", devices[i]);
    if (ndev == 1) return hq_create(devices[0], out);
    {
        Nccl* n = nccl();
        if (!n->handle || !n->error.empty()) return fail(nullptr, HQ_ERR_UNSUPPORTED, "%s", n->error.c_str());
    }
    std::vector<hq_ctx*> ms;
    auto undo = [&] { for (hq_ctx* m : ms) { m->leader = nullptr; m->members.clear(); hq_destroy(m); } };
    for (int i = 0; i < ndev; ++i) {
        hq_ctx* m = nullptr;
        const int rc = hq_create(devices[i], &m);
        if (rc != HQ_OK) { undo(); return rc; }   // g_create_error holds hq_create's message
        ms.push_back(m);
    }
    std::vector<ncclComm_t> comms((size_t)ndev, nullptr);
    const int r = nccl()->CommInitAll(comms.data(), ndev, devices);
    if (r != kNcclSuccess) { undo(); return nccl_fail(nullptr, "ncclCommInitAll", r); }
    for (int i = 0; i < ndev; ++i) {
        ms[i]->comm = comms[i]; ms[i]->comm_rank = i; ms[i]->comm_size = ndev;
        if (i) ms[i]->leader = ms[0];
    }
    ms[0]->members = ms;
    *out = ms[0];
    return HQ_OK;
}

}  // extern "C"
