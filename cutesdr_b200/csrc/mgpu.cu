// mgpu.cu -- NCCL communicator handle and the chunked host->GPU0->all-GPUs block broadcast. See mgpu.cuh.
#include "mgpu.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace csdr {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

NcclApi& nccl()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        // RTLD_NOLOAD first: a host process that already carries an NCCL (e.g. the one bundled with PyTorch) shares it
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
        auto sym = [&](const char* s) { void* p = dlsym(api.lib, s); if (!p) api.err = std::string("NCCL symbol missing: ") + s; return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
    });
    return api;
}

#define CSDR_NCCL(call)                                                                                   \
    do {                                                                                                  \
        ncclResult_t r__ = (call);                                                                        \
        if (r__ != ncclSuccess) {                                                                         \
            csdr::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r__));     \
            return CUTESDR_E_CUDA;                                                                        \
        }                                                                                                 \
    } while (0)

int nccl_ready()
{
    NcclApi& a = nccl();
    if (!a.err.empty() || !a.lib) { set_error("NCCL unavailable: %s", a.err.c_str()); return CUTESDR_E_STATE; }
    return CUTESDR_OK;
}
}  // namespace

int mgpu_bcast_block(cutesdr_mgpu* m, const void* h_src, void* d_buf, size_t bytes, cudaStream_t st_copy, cudaEvent_t done)
{
    NcclApi& a = nccl();
    const size_t cb = m->chunk_bytes;
    const int nchunks = (int)((bytes + cb - 1) / cb);
    while ((int)m->ev_chunk.size() < nchunks) {
        cudaEvent_t e;
        CSDR_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->ev_chunk.push_back(e);
    }
    unsigned char* d = reinterpret_cast<unsigned char*>(d_buf);
    const unsigned char* h = reinterpret_cast<const unsigned char*>(h_src);
    for (int k = 0; k < nchunks; k++) {
        const size_t off = (size_t)k * cb, n = std::min(cb, bytes - off);
        if (m->rank == 0) {
            CSDR_CK(cudaMemcpyAsync(d + off, h + off, n, cudaMemcpyHostToDevice, st_copy));
            CSDR_CK(cudaEventRecord(m->ev_chunk[k], st_copy));
            CSDR_CK(cudaStreamWaitEvent(m->st_comm, m->ev_chunk[k], 0));
        }
        if (m->world > 1) CSDR_NCCL(a.Broadcast(d + off, d + off, n, ncclUint8, 0, (ncclComm_t)m->comm, m->st_comm));
    }
    CSDR_CK(cudaEventRecord(done, m->st_comm));
    m->blocks++;
    m->bytes_bcast += (m->world > 1) ? (long long)bytes : 0;
    return CUTESDR_OK;
}

}  // namespace csdr

using namespace csdr;

cutesdr_mgpu::~cutesdr_mgpu()
{
    cudaSetDevice(device);
    if (st_comm) cudaStreamSynchronize(st_comm);
    for (auto e : ev_chunk) cudaEventDestroy(e);
    if (comm && nccl().CommDestroy) nccl().CommDestroy((ncclComm_t)comm);
    if (st_comm) cudaStreamDestroy(st_comm);
}

extern "C" {

int cutesdr_mgpu_unique_id(void* id128)
{
    if (!id128) { set_error("mgpu_unique_id: bad argument"); return CUTESDR_E_ARG; }
    CSDR_TRY(nccl_ready());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    CSDR_NCCL(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return CUTESDR_OK;
}

int cutesdr_mgpu_init(cutesdr_mgpu** out, const void* id128, int rank, int world, int device)
{
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !id128)) { set_error("mgpu_init: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    CSDR_CK(cudaSetDevice(device));
    std::unique_ptr<cutesdr_mgpu> m(new cutesdr_mgpu());
    m->rank = rank; m->world = world; m->device = device;
    if (const char* e = getenv("CUTESDR_BCAST_CHUNK_KB")) { const long kb = atol(e); if (kb >= 64) m->chunk_bytes = (size_t)kb << 10; }
    CSDR_CK(cudaStreamCreateWithFlags(&m->st_comm, cudaStreamNonBlocking));
    if (world > 1) {
        CSDR_TRY(nccl_ready());
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclComm_t c = nullptr;
        CSDR_NCCL(nccl().CommInitRank(&c, world, id, rank));
        m->comm = c;
    }
    *out = m.release();
    return CUTESDR_OK;
}

void cutesdr_mgpu_destroy(cutesdr_mgpu* m) { delete m; }

int cutesdr_mgpu_info(cutesdr_mgpu* m, int* rank, int* world, int* nccl_version, long long* blocks, long long* bytes_bcast)
{
    if (!m) { set_error("mgpu_info: bad handle"); return CUTESDR_E_ARG; }
    if (rank) *rank = m->rank;
    if (world) *world = m->world;
    if (nccl_version) { *nccl_version = 0; if (m->world > 1 && nccl().GetVersion) nccl().GetVersion(nccl_version); }
    if (blocks) *blocks = m->blocks;
    if (bytes_bcast) *bytes_bcast = m->bytes_bcast;
    return CUTESDR_OK;
}

/* contiguous channel slice of rank `rank` out of n_channels over `world` ranks (the first n % world ranks get one more) */
int cutesdr_mgpu_channel_slice(int n_channels, int rank, int world, int* first, int* count)
{
    if (n_channels < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) { set_error("channel_slice: bad arguments"); return CUTESDR_E_ARG; }
    const int base = n_channels / world, extra = n_channels % world;
    *first = rank * base + std::min(rank, extra);
    *count = base + (rank < extra ? 1 : 0);
    return CUTESDR_OK;
}

}  // extern "C"
