// ppf_comm.cu -- the three small exchanges that couple the ranks of a sharded recognition (one process / host
// thread per GPU, scene reference points sharded, model table replicated):
//   all_reduce(MAX) of one u32       the global vote maximum of the threshold count > thr * max (model.cu:160-167)
//   all_gather of the survivor lists every rank's (vote code, count) records, variable length (model.cu:165-170)
//   all_reduce(SUM) of f32[K]        the clustering scores of the interleaved pose slices (model.cu:202-244)
// The reference has no multi-GPU path (ppf.cu:45 picks one device); SURVEY 8e defines this one.
//
// Two implementations behind one interface:
//   NcclComm   NCCL over NVLink / NVSwitch.  libnccl.so.2 is dlopen()ed on first use, so the library has no link-time
//              dependency on NCCL and shares the copy the host process already loaded (e.g. torch's).
//   LocalComm  `world` host threads of ONE process on one GPU, rendezvous through host memory: the vehicle of the
//              single-GPU tests of the sharded path (NCCL refuses two ranks on one device).
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/ppf_b200.h"
#include "ppf_internal.cuh"

namespace ppf {

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};
NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("PPF_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char *name) { void *p = dlsym(api.handle, name); if (!p) api.error = std::string("libnccl lacks ") + name; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}
#define PPF_NCCL_TRY(expr)                                                                          \
    do {                                                                                            \
        ncclResult_t _r = (expr);                                                                   \
        if (_r != ncclSuccess) {                                                                    \
            set_last_error(std::string(#expr) + ": " + nccl_api()->GetErrorString(_r));             \
            return PPF_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    bool owned = false;
    ~NcclComm() override { if (owned && comm) nccl_api()->CommDestroy(comm); }
    int allreduce_max_u32(uint32_t *dev, size_t n) override {
        PPF_NCCL_TRY(nccl_api()->AllReduce(dev, dev, n, ncclUint32, ncclMax, comm, cur_stream()));
        return PPF_OK;
    }
    int allreduce_sum_f32(float *dev, size_t n) override {
        PPF_NCCL_TRY(nccl_api()->AllReduce(dev, dev, n, ncclFloat32, ncclSum, comm, cur_stream()));
        return PPF_OK;
    }
    int allgather_u32(const uint32_t *dev_send, uint32_t *dev_recv, size_t n) override {
        PPF_NCCL_TRY(nccl_api()->AllGather(dev_send, dev_recv, n, ncclUint32, comm, cur_stream()));
        return PPF_OK;
    }
    // variable-length all-gather = one broadcast per rank inside one NCCL group (a single fused launch)
    int allgatherv(const void *dev_send, void *dev_recv, const size_t *offsets, const size_t *bytes) override {
        PPF_NCCL_TRY(nccl_api()->GroupStart());
        for (int r = 0; r < world; r++) {
            if (bytes[r] == 0) continue;
            ncclResult_t rr = nccl_api()->Broadcast(r == rank ? dev_send : nullptr, (char *)dev_recv + offsets[r], bytes[r],
                                                    ncclUint8, r, comm, cur_stream());
            if (rr != ncclSuccess) { nccl_api()->GroupEnd(); set_last_error(std::string("ncclBroadcast: ") + nccl_api()->GetErrorString(rr)); return PPF_ERR_CUDA; }
        }
        PPF_NCCL_TRY(nccl_api()->GroupEnd());
        return PPF_OK;
    }
};

// ---- threads of one process, one GPU ----------------------------------------------------------------------
struct LocalGroup {
    int world = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long generation = 0;
    std::vector<std::vector<char>> slot;         // one host staging buffer per rank
    void barrier() {
        std::unique_lock<std::mutex> lock(mu);
        const unsigned long long gen = generation;
        if (++arrived == world) { arrived = 0; generation++; cv.notify_all(); }
        else cv.wait(lock, [&] { return generation != gen; });
    }
};
struct LocalComm : Comm {
    std::shared_ptr<LocalGroup> g;
    // every rank parks its device data in its host slot, everybody reads everything, a second barrier frees the slots
    template <typename F> int exchange(const void *dev_send, size_t bytes, F &&consume) {
        g->slot[rank].resize(bytes);
        if (bytes && memcpy_sync(g->slot[rank].data(), dev_send, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_last_error("local comm: copy to host failed");
            g->barrier(); g->barrier();
            return PPF_ERR_CUDA;
        }
        g->barrier();
        int rc = consume();
        g->barrier();
        return rc;
    }
    int allreduce_max_u32(uint32_t *dev, size_t n) override {
        return exchange(dev, n * 4, [&]() -> int {
            std::vector<uint32_t> out(n, 0u);
            for (int r = 0; r < world; r++)
                for (size_t i = 0; i < n; i++) out[i] = std::max(out[i], ((const uint32_t *)g->slot[r].data())[i]);
            return memcpy_sync(dev, out.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess ? PPF_OK : PPF_ERR_CUDA;
        });
    }
    int allreduce_sum_f32(float *dev, size_t n) override {
        return exchange(dev, n * 4, [&]() -> int {
            std::vector<float> out(n, 0.f);
            for (int r = 0; r < world; r++)                       // rank order: every rank computes the same sum
                for (size_t i = 0; i < n; i++) out[i] += ((const float *)g->slot[r].data())[i];
            return memcpy_sync(dev, out.data(), n * 4, cudaMemcpyHostToDevice) == cudaSuccess ? PPF_OK : PPF_ERR_CUDA;
        });
    }
    int allgather_u32(const uint32_t *dev_send, uint32_t *dev_recv, size_t n) override {
        return exchange(dev_send, n * 4, [&]() -> int {
            for (int r = 0; r < world; r++)
                if (n && memcpy_sync(dev_recv + (size_t)r * n, g->slot[r].data(), n * 4, cudaMemcpyHostToDevice) != cudaSuccess) return PPF_ERR_CUDA;
            return PPF_OK;
        });
    }
    int allgatherv(const void *dev_send, void *dev_recv, const size_t *offsets, const size_t *bytes) override {
        return exchange(dev_send, bytes[rank], [&]() -> int {
            for (int r = 0; r < world; r++)
                if (bytes[r] && memcpy_sync((char *)dev_recv + offsets[r], g->slot[r].data(), bytes[r], cudaMemcpyHostToDevice) != cudaSuccess) return PPF_ERR_CUDA;
            return PPF_OK;
        });
    }
};
}  // namespace
}  // namespace ppf

using namespace ppf;

struct ppf_comm { std::unique_ptr<Comm> impl; };
Comm *comm_impl(ppf_comm *c) { return c ? c->impl.get() : nullptr; }

extern "C" {

int ppf_comm_unique_id(void *id_out) {
    if (!id_out) { set_last_error("comm: id_out is NULL"); return PPF_ERR_INVALID; }
    NcclApi *api = nccl_api();
    if (!api->error.empty()) { set_last_error(api->error); return PPF_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    PPF_NCCL_TRY(api->GetUniqueId(&id));
    static_assert(sizeof(id) == PPF_COMM_ID_BYTES, "ncclUniqueId size");
    std::memcpy(id_out, &id, sizeof(id));
    return PPF_OK;
}

int ppf_comm_create_nccl(const void *id, int rank, int world, ppf_comm_t **out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) { set_last_error("comm: bad argument"); return PPF_ERR_INVALID; }
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api->error.empty()) { set_last_error(api->error); return PPF_ERR_UNSUPPORTED; }
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    auto c = std::make_unique<NcclComm>();
    c->rank = rank; c->world = world; c->owned = true;
    PPF_NCCL_TRY(api->CommInitRank(&c->comm, world, uid, rank));
    *out = new ppf_comm{std::move(c)};
    return PPF_OK;
}

int ppf_comm_wrap_nccl(void *nccl_comm, int rank, int world, ppf_comm_t **out) {
    if (!nccl_comm || !out || world < 1 || rank < 0 || rank >= world) { set_last_error("comm: bad argument"); return PPF_ERR_INVALID; }
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api->error.empty()) { set_last_error(api->error); return PPF_ERR_UNSUPPORTED; }
    auto c = std::make_unique<NcclComm>();
    c->rank = rank; c->world = world; c->owned = false; c->comm = (ncclComm_t)nccl_comm;
    *out = new ppf_comm{std::move(c)};
    return PPF_OK;
}

int ppf_comm_create_local(int world, ppf_comm_t **out) {
    if (!out || world < 1 || world > 64) { set_last_error("comm: bad argument"); return PPF_ERR_INVALID; }
    auto g = std::make_shared<LocalGroup>();
    g->world = world;
    g->slot.resize(world);
    for (int r = 0; r < world; r++) {
        auto c = std::make_unique<LocalComm>();
        c->rank = r; c->world = world; c->g = g;
        out[r] = new ppf_comm{std::move(c)};
    }
    return PPF_OK;
}

int ppf_comm_rank(const ppf_comm_t *c) { return c && c->impl ? c->impl->rank : 0; }
int ppf_comm_size(const ppf_comm_t *c) { return c && c->impl ? c->impl->world : 1; }
void ppf_comm_destroy(ppf_comm_t *c) { delete c; }

}  // extern "C"
