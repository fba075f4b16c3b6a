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
    if (peer_ready(c, n_words)) {
        HQ_CUDA(c, hq::launch_peer_allreduce(peer_next(c), d_words, n_words, nullptr, nullptr, 0, st));
        return HQ_OK;
    }
    if (c->comm) HQ_NCCL(c, nccl()->AllReduce(d_words, d_words, n_words, kNcclInt64, kNcclSum, static_cast<ncclComm_t>(c->comm), st));
    return HQ_OK;
}

bool peer_ready(const hq_ctx* c, size_t n_words) {
    if (n_words == 0 || n_words > hq::kPeerCapWords || c->allreduce) return false;
    if (c->is_multi()) {
        for (const hq_ctx* m : c->members) if (!m->peer_open) return false;
        return true;
    }
    return c->peer_open;
}

hq::PeerExchange peer_next(hq_ctx* c) {
    hq::PeerExchange px;
    for (int r = 0; r < c->comm_size && r < hq::kPeerMaxRanks; ++r) px.box[r] = c->peer_box[r];
    px.nranks = c->comm_size; px.rank = c->comm_rank;
    px.seq = ++c->peer_seq;
    px.timeout_ns = c->peer_timeout_ns;
    px.status = c->h_peer_status.p;
    return px;
}

int peer_check(hq_ctx* c) {
    std::vector<hq_ctx*> self(1, c);
    for (hq_ctx* m : (c->is_multi() ? c->members : self))
        if (m->h_peer_status.p && *static_cast<volatile unsigned long long*>(m->h_peer_status.p)) {
            const unsigned long long seq = *m->h_peer_status.p;
            *m->h_peer_status.p = 0;
            return fail(c, HQ_ERR_CUDA, "exchange %llu over peer memory timed out on rank %d of %d: a rank did not take part (all ranks must make the same calls)",
                        seq, m->comm_rank, m->comm_size);
        }
    return HQ_OK;
}

namespace {
bool peer_env_enabled() {
    const char* e = std::getenv("HQ_PEER_EXCHANGE");
    return !(e && e[0] == '0');
}
// this context's own mailbox, zeroed; its exchange sequence restarts
int peer_alloc_box(hq_ctx* c) {
    int rc = bind_device(c); if (rc) return rc;
    HQ_CUDA(c, c->d_peer_box.reserve(hq::kPeerBoxWords));
    HQ_CUDA(c, cudaMemset(c->d_peer_box.p, 0, hq::kPeerBoxWords * 8));
    HQ_CUDA(c, cudaDeviceSynchronize());
    HQ_CUDA(c, c->h_peer_status.reserve(1));
    *c->h_peer_status.p = 0;
    c->peer_seq = 0;
    if (const char* t = std::getenv("HQ_PEER_TIMEOUT_MS")) {
        const long long ms = std::atoll(t);
        if (ms > 0) c->peer_timeout_ns = (unsigned long long)ms * 1000000ull;
    }
    return HQ_OK;
}
void peer_close(hq_ctx* c) {
    if (!c->peer_open && !c->d_peer_box.p) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < hq::kPeerMaxRanks; ++r) {
        if (c->peer_ipc[r] && c->peer_box[r]) cudaIpcCloseMemHandle(c->peer_box[r]);
        c->peer_ipc[r] = false; c->peer_box[r] = nullptr;
    }
    c->peer_open = false;
}
}  // namespace

// single process, several devices: bufs[i] on member i, one grouped collective, each on its member's stream
int group_reduce(hq_ctx* leader, const std::vector<unsigned long long*>& bufs, size_t n_words) {
    if (peer_ready(leader, n_words)) {   // one tiny launch per device instead of a grouped collective
        for (size_t i = 0; i < leader->members.size(); ++i) {
            hq_ctx* m = leader->members[i];
            int rc = bind_device(m); if (rc) return rc;
            HQ_CUDA(leader, hq::launch_peer_allreduce(peer_next(m), bufs[i], n_words, nullptr, nullptr, 0, m->stream));
        }
        return bind_device(leader);
    }
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
    peer_close(c);
    c->d_peer_box.release();
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

int hq_comm_peer_handle(hq_ctx* c, void* handle64) {
    if (!c || !handle64) return c ? fail(c, HQ_ERR_INVALID, "handle buffer is NULL") : HQ_ERR_INVALID;
    if (c->is_multi() || c->leader) return fail(c, HQ_ERR_UNSUPPORTED, "a multi-device context maps its members' mailboxes itself");
    static_assert(sizeof(cudaIpcMemHandle_t) == HQ_PEER_HANDLE_BYTES, "hq_b200.h states the handle size");
    peer_close(c);
    int rc = peer_alloc_box(c); if (rc) return rc;
    cudaIpcMemHandle_t h;
    HQ_CUDA(c, cudaIpcGetMemHandle(&h, c->d_peer_box.p));
    std::memcpy(handle64, &h, sizeof h);
    return HQ_OK;
}

int hq_comm_open_peers(hq_ctx* c, const void* handles, int nranks, int rank) {
    if (!c || !handles) return c ? fail(c, HQ_ERR_INVALID, "handle list is NULL") : HQ_ERR_INVALID;
    if (c->is_multi() || c->leader) return fail(c, HQ_ERR_UNSUPPORTED, "a multi-device context maps its members' mailboxes itself");
    if (!c->comm || nranks != c->comm_size || rank != c->comm_rank) return fail(c, HQ_ERR_INVALID, "rank %d of %d does not match the context's communicator (hq_comm_init_rank first)", rank, nranks);
    if (!c->d_peer_box.p) return fail(c, HQ_ERR_INVALID, "hq_comm_peer_handle first");
    if (nranks > hq::kPeerMaxRanks) return fail(c, HQ_ERR_UNSUPPORTED, "the peer-memory exchange serves up to %d ranks", hq::kPeerMaxRanks);
    if (!peer_env_enabled()) return fail(c, HQ_ERR_UNSUPPORTED, "HQ_PEER_EXCHANGE=0");
    int rc = bind_device(c); if (rc) return rc;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { c->peer_box[r] = c->d_peer_box.p; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof h, sizeof h);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            peer_close(c);
            return fail(c, HQ_ERR_UNSUPPORTED, "cudaIpcOpenMemHandle for rank %d failed: %s (the exchange stays on NCCL; close the peers on EVERY rank)", r, cudaGetErrorString(e));
        }
        c->peer_box[r] = static_cast<unsigned long long*>(p); c->peer_ipc[r] = true;
    }
    c->peer_open = true;
    return HQ_OK;
}

void hq_comm_close_peers(hq_ctx* c) {
    if (!c) return;
    std::vector<hq_ctx*> self(1, c);
    for (hq_ctx* m : (c->is_multi() ? c->members : self)) peer_close(m);
}

int hq_comm_peers_open(const hq_ctx* c) {
    if (!c) return 0;
    if (c->is_multi()) { for (const hq_ctx* m : c->members) if (!m->peer_open) return 0; return 1; }
    return c->peer_open ? 1 : 0;
}

int hq_comm_allreduce(hq_ctx* c, void* d_words, size_t n_words, void* stream) {
    if (!c || !d_words) return c ? fail(c, HQ_ERR_INVALID, "NULL device buffer") : HQ_ERR_INVALID;
    if (c->is_multi()) return fail(c, HQ_ERR_UNSUPPORTED, "device buffers belong to one device: a multi-device context reduces inside hq_eval_palettes");
    int rc = bind_device(c); if (rc) return rc;
    rc = peer_check(c); if (rc) return rc;   // an earlier exchange over peer memory that timed out (this entry is asynchronous)
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
    // the exchange of small payloads over peer memory: every pair of devices must reach each other (NVSwitch: all do)
    bool peers = ndev <= hq::kPeerMaxRanks && peer_env_enabled();
    for (int i = 0; i < ndev && peers; ++i)
        for (int j = 0; j < ndev && peers; ++j) {
            int can = 0;
            if (i != j && (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) != cudaSuccess || !can)) peers = false;
        }
    for (int i = 0; i < ndev && peers; ++i) {
        if (cudaSetDevice(devices[i]) != cudaSuccess) { peers = false; break; }
        for (int j = 0; j < ndev && peers; ++j) {
            if (i == j) continue;
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaGetLastError(); peers = false; }
        }
        if (peers && peer_alloc_box(ms[i]) != HQ_OK) peers = false;
    }
    if (peers)
        for (int i = 0; i < ndev; ++i) {
            for (int j = 0; j < ndev; ++j) ms[i]->peer_box[j] = ms[j]->d_peer_box.p;
            ms[i]->peer_open = true;
        }
    cudaSetDevice(devices[0]);
    *out = ms[0];
    return HQ_OK;
}

}  // extern "C"
