// Multi-GPU library build: the envelope shards of all ranks are assembled with ONE ncclAllGather.
//
// Every rank holds the same slot layout (qms_set_envelope_terms) and has computed the envelopes of its shard
// [begin_r, end_r) in place (qms_build_envelopes).  The shard's seven arrays (per slot: mass f64, relabun f64, abun i32,
// pos i32, c_flag u8; per envelope: len i32, n_c i32) are packed into one fixed-stride record block -- the stride is the
// widest shard, so the blocks of all ranks have the same size --, all-gathered over NVLink / NVSwitch, and every foreign
// block is unpacked to its place in the global arrays.  NCCL is loaded at run time (libnccl.so.2: the copy the process
// already holds, e.g. PyTorch's, else the system one), so the library itself has no link-time dependency on it and loads
// on hosts without NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include "qms_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString) api.handle = h;
        }
    }
    return api.handle ? &api : nullptr;
}

struct ShardRange {
    int64_t env_begin, env_end, slot_begin, slot_end;
};

// contiguous block partition of range(n): the same rule as pyqms_b200.distributed.shard_range
ShardRange shard_of(const qms_ctx* ctx, int r, int world) {
    const int64_t base = ctx->n_env / world, extra = ctx->n_env % world;
    ShardRange s;
    s.env_begin = r * base + (r < extra ? r : extra);
    s.env_end = s.env_begin + base + (r < extra ? 1 : 0);
    s.slot_begin = ctx->h_slot_off[(size_t)s.env_begin];
    s.slot_end = ctx->h_slot_off[(size_t)s.env_end];
    return s;
}

struct BlockLayout {          // byte offsets of the fields inside one rank's record block
    int64_t mass, relabun, abun, pos, cflag, len, nc, stride;
};

BlockLayout block_layout(int64_t widest_slots, int64_t widest_envs) {
    auto up = [](int64_t v) { return (v + 15) / 16 * 16; };
    BlockLayout b;
    b.mass = 0;
    b.relabun = b.mass + up(widest_slots * 8);
    b.abun = b.relabun + up(widest_slots * 8);
    b.pos = b.abun + up(widest_slots * 4);
    b.cflag = b.pos + up(widest_slots * 4);
    b.len = b.cflag + up(widest_slots);
    b.nc = b.len + up(widest_envs * 4);
    b.stride = b.nc + up(widest_envs * 4);
    return b;
}

// pack == true: global arrays -> block of this rank; pack == false: blocks of the other ranks -> global arrays
__global__ void shard_records_kernel(bool pack, int me, int world, const ShardRange* __restrict__ ranges, BlockLayout layout,
                                     int64_t widest_slots, int64_t widest_envs, unsigned char* __restrict__ blocks,
                                     double* __restrict__ mass, double* __restrict__ relabun, int32_t* __restrict__ abun,
                                     int32_t* __restrict__ pos, uint8_t* __restrict__ cflag, int32_t* __restrict__ len,
                                     int32_t* __restrict__ nc) {
    const int r = pack ? me : (int)blockIdx.y;
    if (!pack && r == me) return;
    const ShardRange range = ranges[r];
    unsigned char* block = blocks + (pack ? 0 : (int64_t)r * layout.stride);
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < range.slot_end - range.slot_begin) {
        const int64_t s = range.slot_begin + i;
        double* b_mass = reinterpret_cast<double*>(block + layout.mass);
        double* b_relabun = reinterpret_cast<double*>(block + layout.relabun);
        int32_t* b_abun = reinterpret_cast<int32_t*>(block + layout.abun);
        int32_t* b_pos = reinterpret_cast<int32_t*>(block + layout.pos);
        uint8_t* b_cflag = block + layout.cflag;
        if (pack) {
            b_mass[i] = mass[s];
            b_relabun[i] = relabun[s];
            b_abun[i] = abun[s];
            b_pos[i] = pos[s];
            b_cflag[i] = cflag[s];
        } else {
            mass[s] = b_mass[i];
            relabun[s] = b_relabun[i];
            abun[s] = b_abun[i];
            pos[s] = b_pos[i];
            cflag[s] = b_cflag[i];
        }
    }
    if (i < range.env_end - range.env_begin) {
        const int64_t e = range.env_begin + i;
        int32_t* b_len = reinterpret_cast<int32_t*>(block + layout.len);
        int32_t* b_nc = reinterpret_cast<int32_t*>(block + layout.nc);
        if (pack) {
            b_len[i] = len[e];
            b_nc[i] = nc[e];
        } else {
            len[e] = b_len[i];
            nc[e] = b_nc[i];
        }
    }
}

}  // namespace

extern "C" int qms_comm_unique_id(void* out, size_t capacity) {
    NcclApi* api = nccl_api();
    if (!api) return qms_fail(nullptr, QMS_E_CUDA, "libnccl.so.2 not found (dlopen): %s", dlerror());
    if (!out || capacity < sizeof(ncclUniqueId)) return qms_fail(nullptr, QMS_E_ARG, "unique id needs %zu bytes", sizeof(ncclUniqueId));
    ncclUniqueId id;
    const ncclResult_t rc = api->GetUniqueId(&id);
    if (rc != ncclSuccess) return qms_fail(nullptr, QMS_E_CUDA, "ncclGetUniqueId: %s", api->GetErrorString(rc));
    memcpy(out, &id, sizeof(id));
    return QMS_OK;
}

extern "C" int qms_comm_init(qms_ctx* ctx, const void* unique_id, int rank, int nranks) {
    if (!ctx || !unique_id || nranks < 1 || rank < 0 || rank >= nranks) return qms_fail(ctx, QMS_E_ARG, "qms_comm_init: bad arguments");
    NcclApi* api = nccl_api();
    if (!api) return qms_fail(ctx, QMS_E_CUDA, "libnccl.so.2 not found (dlopen)");
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) {
        api->CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm = nullptr;
    const ncclResult_t rc = api->CommInitRank(&comm, nranks, id, rank);
    if (rc != ncclSuccess) return qms_fail(ctx, QMS_E_CUDA, "ncclCommInitRank: %s", api->GetErrorString(rc));
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_size = nranks;
    return QMS_OK;
}

extern "C" int qms_comm_destroy(qms_ctx* ctx) {
    if (!ctx) return QMS_E_ARG;
    NcclApi* api = nccl_api();
    if (api && ctx->nccl_comm) api->CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->comm_size = 0;
    return QMS_OK;
}

extern "C" int qms_allgather_library(qms_ctx* ctx, int64_t* bytes_received, double* ms) {
    if (!ctx || !ctx->have_terms) return qms_fail(ctx, QMS_E_ARG, "qms_allgather_library: envelope terms not set");
    if (!ctx->nccl_comm || ctx->comm_size < 1) return qms_fail(ctx, QMS_E_ARG, "qms_comm_init must be called first");
    NcclApi* api = nccl_api();
    QMS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int world = ctx->comm_size, me = ctx->comm_rank;
    std::vector<ShardRange> ranges((size_t)world);
    int64_t widest_slots = 0, widest_envs = 0, foreign = 0;
    for (int r = 0; r < world; ++r) {
        ranges[(size_t)r] = shard_of(ctx, r, world);
        widest_slots = std::max(widest_slots, ranges[(size_t)r].slot_end - ranges[(size_t)r].slot_begin);
        widest_envs = std::max(widest_envs, ranges[(size_t)r].env_end - ranges[(size_t)r].env_begin);
    }
    const BlockLayout layout = block_layout(widest_slots, widest_envs);
    for (int r = 0; r < world; ++r)
        if (r != me) foreign += (ranges[(size_t)r].slot_end - ranges[(size_t)r].slot_begin) * 25 + (ranges[(size_t)r].env_end - ranges[(size_t)r].env_begin) * 8;
    if (bytes_received) *bytes_received = foreign;
    if (ms) *ms = 0.0;
    if (world == 1 || widest_envs == 0) return QMS_OK;
    TmpBuf<unsigned char> send(ctx), recv(ctx);
    TmpBuf<ShardRange> d_ranges(ctx);
    QMS_TRY(qms_reserve(ctx, send.b, (size_t)layout.stride));
    QMS_TRY(qms_reserve(ctx, recv.b, (size_t)layout.stride * (size_t)world));
    QMS_TRY(qms_upload(ctx, d_ranges.b, ranges.data(), (size_t)world));
    const int threads = 256;
    const unsigned blocks = (unsigned)((std::max(widest_slots, widest_envs) + threads - 1) / threads);
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    shard_records_kernel<<<dim3(blocks, 1), threads, 0, ctx->stream>>>(true, me, world, d_ranges.b.p, layout, widest_slots, widest_envs,
                                                                      send.b.p, ctx->env_mass.p, ctx->env_relabun.p, ctx->env_abun.p,
                                                                      ctx->env_pos.p, ctx->env_cflag.p, ctx->env_len.p, ctx->env_nc.p);
    QMS_CUDA(ctx, cudaGetLastError());
    const ncclResult_t rc = api->AllGather(send.b.p, recv.b.p, (size_t)layout.stride, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (rc != ncclSuccess) return qms_fail(ctx, QMS_E_CUDA, "ncclAllGather: %s", api->GetErrorString(rc));
    shard_records_kernel<<<dim3(blocks, (unsigned)world), threads, 0, ctx->stream>>>(
        false, me, world, d_ranges.b.p, layout, widest_slots, widest_envs, recv.b.p, ctx->env_mass.p, ctx->env_relabun.p, ctx->env_abun.p,
        ctx->env_pos.p, ctx->env_cflag.p, ctx->env_len.p, ctx->env_nc.p);
    QMS_CUDA(ctx, cudaGetLastError());
    ctx->cnt.kernel_launches += 2;
    QMS_CUDA(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
    QMS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float elapsed = 0;
    cudaEventElapsedTime(&elapsed, ctx->ev[0], ctx->ev[1]);
    if (ms) *ms = elapsed;
    return QMS_OK;
}
