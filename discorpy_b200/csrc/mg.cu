// mg.cu -- the multi-GPU plumbing of the C ABI (dcb_mg_*): one process per GPU, NCCL over
// NVLink / NVSwitch.  The path shards without a data-plane collective: what the
// ranks exchange is the <= 256-byte parameter block (one broadcast), the rows of ONE assembled
// sinogram when a caller asks for it (all-gather; the fused peer-store form needs no call here),
// and the scalars of a benchmark (barrier, max over ranks).  NCCL is resolved at run time
// (dlopen of libnccl.so.2): the library has no link-time dependency on it and single-GPU use
// never touches it.
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include "api_common.hpp"

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

const NcclApi &nccl() {
    std::call_once(g_nccl_once, [] {
        const char *names[] = {getenv("DCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (nm == nullptr || nm[0] == 0) continue;
            g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (g_nccl.lib != nullptr) break;
        }
        if (g_nccl.lib == nullptr) return;
        auto sym = [](const char *s) { return dlsym(g_nccl.lib, s); };
        g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
        g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
        g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
        g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(sym("ncclBroadcast"));
        g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
        g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
        g_nccl.GetErrorString =
            reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
        g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(sym("ncclGetVersion"));
        g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy &&
                    g_nccl.Broadcast && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.GetErrorString;
    });
    return g_nccl;
}

struct MgState {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0, device = -1;
    double *scratch = nullptr;   // 8 doubles on the device (barrier / scalar reductions)
};
MgState g_mg;
std::mutex g_mg_mu;

}  // namespace

#define NCCL_TRY(expr)                                                                        \
    do {                                                                                      \
        ncclResult_t r__ = (expr);                                                            \
        if (r__ != ncclSuccess)                                                               \
            return dcb::fail(DCB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                             nccl().GetErrorString(r__), __FILE__, __LINE__);                 \
    } while (0)

#define NEED_COMM()                                                                           \
    do {                                                                                      \
        if (g_mg.comm == nullptr)                                                             \
            return dcb::fail(DCB_ERR_ARG, "no communicator: call dcb_mg_init first");         \
    } while (0)

extern "C" {

int dcb_mg_unique_id(void *id_out, int *nccl_version) {
    REQUIRE(id_out != nullptr, "id_out is NULL");
    if (!nccl().ok)
        return dcb::fail(DCB_ERR_UNSUPPORTED, "libnccl.so.2 not found (set DCB_NCCL_LIB)");
    static_assert(sizeof(ncclUniqueId) == DCB_MG_UNIQUE_ID_BYTES, "NCCL unique id size");
    NCCL_TRY(nccl().GetUniqueId(reinterpret_cast<ncclUniqueId *>(id_out)));
    if (nccl_version != nullptr) {
        *nccl_version = 0;
        if (nccl().GetVersion) nccl().GetVersion(nccl_version);
    }
    return DCB_OK;
}

int dcb_mg_init(const void *id, int world, int rank) {
    REQUIRE(id != nullptr, "id is NULL");
    REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world size %d", rank, world);
    if (!nccl().ok)
        return dcb::fail(DCB_ERR_UNSUPPORTED, "libnccl.so.2 not found (set DCB_NCCL_LIB)");
    std::lock_guard<std::mutex> lk(g_mg_mu);
    REQUIRE(g_mg.comm == nullptr, "dcb_mg_init called twice (dcb_mg_finalize first)");
    CUDA_TRY(cudaGetDevice(&g_mg.device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NCCL_TRY(nccl().CommInitRank(&g_mg.comm, world, uid, rank));
    g_mg.world = world;
    g_mg.rank = rank;
    CUDA_TRY(cudaMalloc((void **)&g_mg.scratch, 8 * sizeof(double)));
    CUDA_TRY(cudaMemset(g_mg.scratch, 0, 8 * sizeof(double)));
    return DCB_OK;
}

int dcb_mg_info(int *world, int *rank) {
    if (world != nullptr) *world = g_mg.comm ? g_mg.world : 1;
    if (rank != nullptr) *rank = g_mg.comm ? g_mg.rank : 0;
    return DCB_OK;
}

int dcb_mg_bcast(void *dev_buf, size_t nbytes, int root, void *stream) {
    NEED_COMM();
    REQUIRE(dev_buf != nullptr && root >= 0 && root < g_mg.world, "bad broadcast arguments");
    NCCL_TRY(nccl().Broadcast(dev_buf, dev_buf, nbytes, ncclUint8, root, g_mg.comm,
                              (cudaStream_t)stream));
    return DCB_OK;
}

int dcb_mg_bcast_host(void *host_buf, size_t nbytes, int root) {
    NEED_COMM();
    REQUIRE(host_buf != nullptr && nbytes > 0 && nbytes <= 4096, "host broadcast of 1..4096 bytes");
    void *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, nbytes));
    cudaError_t ce = cudaMemcpy(d, host_buf, nbytes, cudaMemcpyHostToDevice);
    ncclResult_t nr = ncclSuccess;
    if (ce == cudaSuccess) nr = nccl().Broadcast(d, d, nbytes, ncclUint8, root, g_mg.comm, nullptr);
    if (ce == cudaSuccess && nr == ncclSuccess) ce = cudaStreamSynchronize(nullptr);
    if (ce == cudaSuccess && nr == ncclSuccess)
        ce = cudaMemcpy(host_buf, d, nbytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (nr != ncclSuccess)
        return dcb::fail(DCB_ERR_CUDA, "ncclBroadcast failed: %s", nccl().GetErrorString(nr));
    if (ce != cudaSuccess)
        return dcb::fail(DCB_ERR_CUDA, "broadcast staging failed: %s", cudaGetErrorString(ce));
    return DCB_OK;
}

int dcb_mg_allgather(const void *send_dev, void *recv_dev, size_t nbytes_per_rank, void *stream) {
    NEED_COMM();
    REQUIRE(send_dev != nullptr && recv_dev != nullptr, "null buffer");
    NCCL_TRY(nccl().AllGather(send_dev, recv_dev, nbytes_per_rank, ncclUint8, g_mg.comm,
                              (cudaStream_t)stream));
    return DCB_OK;
}

int dcb_mg_allreduce_max_f64(double *host_values, int count) {
    REQUIRE(host_values != nullptr && count >= 1 && count <= 8, "1..8 values");
    if (g_mg.comm == nullptr) return DCB_OK;   // single process: the maximum over one rank
    CUDA_TRY(cudaMemcpy(g_mg.scratch, host_values, count * sizeof(double), cudaMemcpyHostToDevice));
    NCCL_TRY(nccl().AllReduce(g_mg.scratch, g_mg.scratch, count, ncclFloat64, ncclMax, g_mg.comm,
                              nullptr));
    CUDA_TRY(cudaStreamSynchronize(nullptr));
    CUDA_TRY(cudaMemcpy(host_values, g_mg.scratch, count * sizeof(double), cudaMemcpyDeviceToHost));
    return DCB_OK;
}

int dcb_mg_barrier(void) {
    if (g_mg.comm == nullptr) return DCB_OK;
    CUDA_TRY(cudaDeviceSynchronize());   // everything this rank queued has finished ...
    NCCL_TRY(nccl().AllReduce(g_mg.scratch + 4, g_mg.scratch + 4, 1, ncclFloat64, ncclMax, g_mg.comm,
                              nullptr));
    CUDA_TRY(cudaStreamSynchronize(nullptr));   // ... and so has every other rank's
    return DCB_OK;
}

int dcb_mg_finalize(void) {
    std::lock_guard<std::mutex> lk(g_mg_mu);
    if (g_mg.comm != nullptr) {
        cudaDeviceSynchronize();
        nccl().CommDestroy(g_mg.comm);
        cudaFree(g_mg.scratch);
        g_mg = MgState();
    }
    return DCB_OK;
}

}  // extern "C"
