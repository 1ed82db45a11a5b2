// Peer-memory communicator of the finish kernel (corr_host.cu): every rank owns an inbox in device memory that its
// peers map with CUDA IPC and write into over NVLink.  Layout of an inbox allocation:
//     [2 parities][world slots][capacity 8-byte words]   the partial sums / counts of every rank
//     [2 parities][world] 8-byte sequence flags          (starts on a 256-byte boundary after the slots)
// No NCCL call is made on the data path; torch.distributed (or anything else) only carries the 64-byte IPC handles
// once, at start-up.
#include <new>

#include "corr_core.cuh"

struct mdb_comm {
    int rank = 0, world = 1;
    size_t capacity = 0;               // words per slot
    size_t flags_offset = 0;           // bytes
    size_t bytes = 0;
    void *local = nullptr;             // my inbox (cudaMalloc)
    void *peer[MDB_COMM_MAX_RANKS] = {};
    bool opened[MDB_COMM_MAX_RANKS] = {};
    bool connected = false;
    int timeout_ms = 0;
    int device = 0;
    unsigned long long seq = 0;        // launches of finish kernels that used this communicator
};

namespace mdb {

unsigned long long *comm_sequence(mdb_comm *comm) { return &comm->seq; }

int comm_device_view(mdb_comm *comm, CommDev *out) {
    if (!comm || !comm->connected) return fail(MDB_ERR_ARGUMENT, "the communicator is not connected");
    memset(out, 0, sizeof(*out));
    out->rank = comm->rank;
    out->world = comm->world;
    out->capacity = comm->capacity;
    for (int r = 0; r < comm->world; ++r) {
        out->inbox[r] = static_cast<unsigned long long *>(comm->peer[r]);
        out->flags[r] = reinterpret_cast<unsigned long long *>(static_cast<char *>(comm->peer[r]) + comm->flags_offset);
    }
    int khz = 1965000;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, comm->device);
    out->timeout_cycles = (long long)comm->timeout_ms * (long long)khz;
    return MDB_OK;
}

}  // namespace mdb

using namespace mdb;

extern "C" int mdb_comm_create(int32_t rank, int32_t world, size_t capacity_words, int32_t timeout_ms, mdb_comm **out) {
    if (!out) return fail(MDB_ERR_ARGUMENT, "comm is NULL");
    *out = nullptr;
    if (world < 1 || world > MDB_COMM_MAX_RANKS || rank < 0 || rank >= world)
        return fail(MDB_ERR_ARGUMENT, "need 0 <= rank < world <= %d", MDB_COMM_MAX_RANKS);
    if (capacity_words == 0) return fail(MDB_ERR_ARGUMENT, "capacity_words must be positive");
    mdb_comm *c = new (std::nothrow) mdb_comm();
    if (!c) return fail(MDB_ERR_CUDA, "out of host memory");
    c->rank = rank;
    c->world = world;
    c->capacity = (capacity_words + 31) / 32 * 32;
    c->flags_offset = 2 * (size_t)world * c->capacity * 8;
    c->bytes = c->flags_offset + 2 * (size_t)world * 8 + 256;
    c->timeout_ms = timeout_ms > 0 ? timeout_ms : 20000;
    cudaGetDevice(&c->device);
    cudaError_t err = cudaMalloc(&c->local, c->bytes);
    if (err == cudaSuccess) err = cudaMemset(c->local, 0, c->bytes);
    if (err == cudaSuccess) err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        delete c;
        return fail(MDB_ERR_CUDA, "allocating the inbox (%zu bytes) failed: %s", c->bytes, cudaGetErrorString(err));
    }
    c->peer[rank] = c->local;
    if (world == 1) c->connected = true;
    *out = c;
    return MDB_OK;
}

extern "C" int mdb_comm_handle(mdb_comm *c, void *handle_out) {
    if (!c || !handle_out) return fail(MDB_ERR_ARGUMENT, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == MDB_COMM_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    MDB_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->local));
    memcpy(handle_out, &h, sizeof(h));
    return MDB_OK;
}

extern "C" int mdb_comm_connect(mdb_comm *c, const void *handles) {
    if (!c || !handles) return fail(MDB_ERR_ARGUMENT, "NULL argument");
    if (c->connected) return MDB_OK;
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char *>(handles) + (size_t)r * MDB_COMM_HANDLE_BYTES, sizeof(h));
        void *ptr = nullptr;
        const cudaError_t err = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess)
            return fail(MDB_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s", r, cudaGetErrorString(err));
        c->peer[r] = ptr;
        c->opened[r] = true;
    }
    c->connected = true;
    return MDB_OK;
}

extern "C" int mdb_comm_destroy(mdb_comm *c) {
    if (!c) return MDB_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return MDB_OK;
}
