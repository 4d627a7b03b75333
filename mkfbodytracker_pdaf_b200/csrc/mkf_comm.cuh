// mkf_comm.cuh -- the only collective of the path: the final gather of per-track summaries over the GPUs of one box
// (kernel K8).  Tracks are independent, so nothing is exchanged per frame; at the end of a run every
// rank contributes one row {pose[D], wsum, status} per track and ncclAllGather delivers all rows to every rank over
// NVLink / NVSwitch.  Included by mkf_api.cu.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the library has no link-time dependency on it, a process that
// already carries an NCCL (torch.distributed does) shares that copy, and a single-GPU user never loads one.
#ifndef MKF_COMM_CUH
#define MKF_COMM_CUH

#include <dlfcn.h>
#include <nccl.h> // types and prototypes only

struct MkfNccl {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommCount) CommCount = nullptr;
    decltype(&ncclCommUserRank) CommUserRank = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};

static MkfNccl& mkf_nccl()
{
    static MkfNccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        // a copy already mapped into the process (torch's bundled one) wins; else the system library
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names)
            if (!n.handle) n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char* nm : names)
            if (!n.handle) n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (!n.handle) return;
#define MKF_NCCL_SYM(f) n.f = reinterpret_cast<decltype(n.f)>(dlsym(n.handle, "nccl" #f))
        MKF_NCCL_SYM(GetUniqueId);
        MKF_NCCL_SYM(CommInitRank);
        MKF_NCCL_SYM(CommDestroy);
        MKF_NCCL_SYM(AllGather);
        MKF_NCCL_SYM(GetErrorString);
        MKF_NCCL_SYM(CommCount);
        MKF_NCCL_SYM(CommUserRank);
        MKF_NCCL_SYM(GetVersion);
#undef MKF_NCCL_SYM
        n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllGather && n.GetErrorString && n.CommCount &&
               n.CommUserRank;
    });
    return n;
}

#define NCK(call)                                                                                       \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) {                                                                        \
            mkf_set_error("%s failed: %s (%s:%d)", #call, mkf_nccl().GetErrorString(r_), __FILE__, __LINE__); \
            return MKF_E_CUDA;                                                                          \
        }                                                                                               \
    } while (0)

static int mkf_nccl_required()
{
    if (!mkf_nccl().ok) {
        mkf_set_error("libnccl.so.2 could not be loaded (%s): the multi-GPU summary gather needs NCCL",
                      mkf_nccl().handle ? "symbols missing" : "dlopen failed");
        return MKF_E_UNSUPPORTED;
    }
    return MKF_OK;
}

struct mkf_comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    bool owned = false;
    DevBuf rows; // gather target when the caller's buffer is host memory
};

extern "C" int mkf_comm_unique_id(void* id128)
{
    if (!id128) {
        mkf_set_error("mkf_comm_unique_id: null argument");
        return MKF_E_INVALID;
    }
    int rc = mkf_nccl_required();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == MKF_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCK(mkf_nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return MKF_OK;
}

extern "C" int mkf_comm_create(mkf_comm** out, int nranks, int rank, const void* id128, int device)
{
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) {
        mkf_set_error("mkf_comm_create: invalid argument (nranks=%d rank=%d)", nranks, rank);
        return MKF_E_INVALID;
    }
    *out = nullptr;
    int rc = mkf_nccl_required();
    if (rc) return rc;
    const int ndev = mkf_device_count();
    if (device < 0 || device >= ndev) {
        mkf_set_error("mkf_comm_create: device %d out of range (%d visible)", device, ndev);
        return ndev > 0 ? MKF_E_INVALID : MKF_E_CUDA;
    }
    CK(cudaSetDevice(device));
    mkf_comm* c = new (std::nothrow) mkf_comm;
    if (!c) return MKF_E_NOMEM;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = mkf_nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        mkf_set_error("ncclCommInitRank failed: %s", mkf_nccl().GetErrorString(r));
        delete c;
        return MKF_E_CUDA;
    }
    c->nranks = nranks;
    c->rank = rank;
    c->device = device;
    c->owned = true;
    *out = c;
    return MKF_OK;
}

extern "C" int mkf_comm_wrap(mkf_comm** out, void* nccl_comm, int device)
{
    if (!out || !nccl_comm) {
        mkf_set_error("mkf_comm_wrap: null argument");
        return MKF_E_INVALID;
    }
    *out = nullptr;
    int rc = mkf_nccl_required();
    if (rc) return rc;
    mkf_comm* c = new (std::nothrow) mkf_comm;
    if (!c) return MKF_E_NOMEM;
    c->comm = (ncclComm_t)nccl_comm;
    c->device = device;
    ncclResult_t r = mkf_nccl().CommCount(c->comm, &c->nranks);
    if (r == ncclSuccess) r = mkf_nccl().CommUserRank(c->comm, &c->rank);
    if (r != ncclSuccess) {
        mkf_set_error("mkf_comm_wrap: %s", mkf_nccl().GetErrorString(r));
        delete c;
        return MKF_E_CUDA;
    }
    *out = c;
    return MKF_OK;
}

extern "C" void mkf_comm_destroy(mkf_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    c->rows.release();
    if (c->owned && c->comm && mkf_nccl().ok) mkf_nccl().CommDestroy(c->comm);
    delete c;
}

extern "C" int mkf_comm_info(const mkf_comm* c, int* nranks, int* rank, int* nccl_version)
{
    if (!c) {
        mkf_set_error("null comm");
        return MKF_E_INVALID;
    }
    if (nranks) *nranks = c->nranks;
    if (rank) *rank = c->rank;
    if (nccl_version) {
        *nccl_version = 0;
        if (mkf_nccl().GetVersion) mkf_nccl().GetVersion(nccl_version);
    }
    return MKF_OK;
}

// contiguous block partition of `total` tracks over `world` ranks (the first total % world ranks take one more); a
// person's two arm filters share a track id, hence a rank
extern "C" int mkf_shard_tracks(int64_t total, int world, int rank, int64_t* first, int64_t* count)
{
    if (total < 0 || world < 1 || rank < 0 || rank >= world || !first || !count) {
        mkf_set_error("mkf_shard_tracks: invalid argument");
        return MKF_E_INVALID;
    }
    const int64_t base = total / world, rem = total % world;
    *count = base + (rank < rem ? 1 : 0);
    *first = rank * base + (rank < rem ? rank : rem);
    return MKF_OK;
}

// row t = { pose[D], wsum, (double)status }; rows beyond T (padding up to rows_per_rank) are zero
__global__ void k_pack_summary(const double* __restrict__ pose, const double* __restrict__ wsum,
                               const uint32_t* __restrict__ status, long long T, long long rows, int D,
                               double* __restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int W = D + 2;
    if (i >= rows * W) return;
    const long long t = i / W;
    const int c = (int)(i - t * W);
    double v = 0.0;
    if (t < T) v = c < D ? pose[t * D + c] : (c == D ? wsum[t] : (double)status[t]);
    out[i] = v;
}

extern "C" int mkf_batch_summaries(mkf_batch* b, int64_t rows, double* out, int mem);

// local rows of this rank's batch: T (or `rows` >= T, zero-padded) x (D + 2)
extern "C" int mkf_batch_summaries(mkf_batch* b, int64_t rows, double* out, int mem)
{
    if (!b || !out || rows < b->T) {
        mkf_set_error("mkf_batch_summaries: invalid argument");
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const int D = b->m->D, W = D + 2;
    int rc;
    const double* d_pose;
    if (b->est_valid) { // the pose k_resample_runs / k_frame_small left in the batch
        d_pose = b->est[b->est_slot] + (size_t)b->T * b->m->d;
    } else {
        if ((rc = b->out_b.ensure((size_t)b->T * D * 8))) return rc;
        if ((rc = launch_estimate(b, nullptr, (double*)b->out_b.p))) return rc;
        d_pose = (const double*)b->out_b.p;
    }
    OutPtr<double> o;
    if ((rc = o.init(b, out, (size_t)rows * W, mem, b->out_a))) return rc;
    k_pack_summary<<<grid_for(rows * W, 256), 256, 0, b->stream>>>(d_pose, b->wsum, b->status, b->T, rows, D, o.devp);
    MKF_LAUNCHED();
    CK(cudaGetLastError());
    if ((rc = o.finish(b))) return rc;
    if (o.host) CK(cudaStreamSynchronize(b->stream));
    return MKF_OK;
}

extern "C" int mkf_batch_gather_summaries(mkf_batch* b, mkf_comm* c, int64_t rows_per_rank, double* out, int mem)
{
    if (!b || !c || !out) {
        mkf_set_error("mkf_batch_gather_summaries: null argument");
        return MKF_E_INVALID;
    }
    if (rows_per_rank <= 0) rows_per_rank = b->T;
    if (rows_per_rank < b->T) {
        mkf_set_error("mkf_batch_gather_summaries: rows_per_rank %lld < T %lld", (long long)rows_per_rank, b->T);
        return MKF_E_INVALID;
    }
    if (c->device != b->device) {
        mkf_set_error("mkf_batch_gather_summaries: communicator on device %d, batch on device %d", c->device, b->device);
        return MKF_E_INVALID;
    }
    CK(cudaSetDevice(b->device));
    const int W = b->m->D + 2;
    const size_t cnt = (size_t)rows_per_rank * W; // doubles per rank
    const bool host = !is_device_ptr(out, mem);
    double* d_all = out;
    int rc;
    if (host) {
        if ((rc = c->rows.ensure(cnt * c->nranks * 8))) return rc;
        d_all = (double*)c->rows.p;
    }
    // the local rows are packed straight into this rank's stretch of the gathered buffer: in-place all-gather
    double* mine = d_all + (size_t)c->rank * cnt;
    if ((rc = mkf_batch_summaries(b, rows_per_rank, mine, MKF_MEM_DEVICE))) return rc;
    NCK(mkf_nccl().AllGather(mine, d_all, cnt, ncclDouble, c->comm, b->stream)); // (a copy-free no-op for one rank)
    if (host) {
        CK(cudaMemcpyAsync(out, d_all, cnt * c->nranks * 8, cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));
    }
    return MKF_OK;
}

#endif
