// Context, device memory, HaloVector layout bridge and the multi-GPU plumbing
// (NCCL all-reduce for the Arnoldi inner products, nearest-neighbour halo rows for the
// 2-D stencils — the distributed form of bc!(u), examples/heat_2D.jl:51, on the slab
// decomposition modelled on examples/halovector.jl).
//
// NCCL is resolved at run time with dlopen/dlsym so that libariadne_b200.so has no link-time
// dependency on it (single-GPU users never load it, and inside a torch process the already
// loaded libnccl.so.2 is reused).
#include <dlfcn.h>
#include <stdarg.h>
#include <string.h>

#include "ak_internal.h"
#include "common.cuh"
#include "sweep.h"

namespace ak {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- in-stream profiler --------------------------------------------------------------------
constexpr size_t kProfMaxRecs = 1 << 16;
ProfScope::ProfScope(Ctx* ctx, int cls) : c(ctx), idx(-1) {
    if (!c->prof_on || c->prof_recs.size() >= kProfMaxRecs) return;
    auto get = [&]() -> cudaEvent_t {
        if (c->prof_pool_used == c->prof_pool.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            c->prof_pool.push_back(e);
        }
        return c->prof_pool[c->prof_pool_used++];
    };
    cudaEvent_t e0 = get(), e1 = get();
    if (!e0 || !e1) return;
    cudaEventRecord(e0, c->stream);
    c->prof_recs.push_back({cls, e0, e1});
    idx = (int)c->prof_recs.size() - 1;
}
ProfScope::~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->prof_recs[(size_t)idx].e1, c->stream);
}

// ---- minimal NCCL surface (ABI-stable since NCCL 2.x) -----------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.lib) return AK_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return AK_ERR_NCCL;
    }
#define AK_SYM(field, name)                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                     \
    if (!g_nccl.field) {                                           \
        set_error("libnccl: missing symbol %s", name);             \
        return AK_ERR_NCCL;                                        \
    }
    AK_SYM(GetUniqueId, "ncclGetUniqueId");
    AK_SYM(CommInitRank, "ncclCommInitRank");
    AK_SYM(CommDestroy, "ncclCommDestroy");
    AK_SYM(AllReduce, "ncclAllReduce");
    AK_SYM(AllGather, "ncclAllGather");
    AK_SYM(Send, "ncclSend");
    AK_SYM(Recv, "ncclRecv");
    AK_SYM(GroupStart, "ncclGroupStart");
    AK_SYM(GroupEnd, "ncclGroupEnd");
    AK_SYM(GetErrorString, "ncclGetErrorString");
#undef AK_SYM
    g_nccl.lib = h;
    return AK_OK;
}

#define AK_NCCL(expr)                                                                                  \
    do {                                                                                               \
        ncclResult_t _r = (expr);                                                                      \
        if (_r != 0) {                                                                                 \
            ak::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, ak::g_nccl.GetErrorString(_r)); \
            return AK_ERR_NCCL;                                                                        \
        }                                                                                              \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
};

// ---- stand-alone collectives over NVLink peer memory -----------------------------------------------------------
// All-reduce of up to kBlkSums doubles through the mailboxes: thread 0 deposits this rank's values in every rank's
// mailbox (st.global over NVLink + release of the sequence tag), then the block waits for every rank's record and adds
// them in rank order (bit-identical on all ranks).  Replaces an in-stream ncclAllReduce (a kernel launch of its own
// plus the proxy/LL protocol latency) by ~2 us of peer stores for the scalars of CG, the step-wise Gram-Schmidt kernels
// and the norms at the cycle boundaries.
__global__ void __launch_bounds__(32) k_mail_allreduce(double* vals, int count, const P2PDev pd, unsigned long long seq) {
    __shared__ double shm[kBlkSums * kMaxPeers];
    __shared__ double tot[kBlkSums];
    if (threadIdx.x == 0) mail_post(pd, seq, vals, count);
    __syncwarp();
    mail_wait_sum(pd, seq, tot, count, shm);
    if ((int)threadIdx.x < count) vals[threadIdx.x] = tot[threadIdx.x];
}

// Ghost exchange of a slab / segment: the last `cnt_up` values of v go into the up (right) neighbour's "lo" slot, the
// first `cnt_down` values into the down (left) neighbour's "hi" slot (peer stores), then an all-to-all signal through
// the mailboxes: when the kernel ends, this rank's own slots hold its neighbours' values.
__global__ void __launch_bounds__(1024) k_push_ghost(const double* __restrict__ v, int64_t n, int cnt_up, int cnt_down,
                                                     double* up_lo, double* down_hi, const P2PDev pd,
                                                     unsigned long long seq) {
    __shared__ double shm[kBlkSums * kMaxPeers];
    if (up_lo != nullptr)
        for (int i = threadIdx.x; i < cnt_up; i += blockDim.x) up_lo[i] = v[n - cnt_up + i];
    if (down_hi != nullptr)
        for (int i = threadIdx.x; i < cnt_down; i += blockDim.x) down_hi[i] = v[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) mail_post(pd, seq, nullptr, 0);
    mail_wait_sum(pd, seq, nullptr, 0, shm);
}

static int p2p_allreduce(Ctx* ctx, double* dev, int count) {
    const unsigned long long seq = ++ctx->p2p_seq;
    ProfScope prof(ctx, PK_SCALAR);
    k_mail_allreduce<<<1, 32, 0, ctx->stream>>>(dev, count, ctx->p2p_dev(), seq);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

int allreduce_sum(Ctx* ctx, double* dev, int count) {
    if (ctx->nranks <= 1) return AK_OK;
    if (ctx->p2p_on && count <= kBlkSums) return p2p_allreduce(ctx, dev, count);
    AK_NCCL(g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, ctx->comm->comm, ctx->stream));
    return AK_OK;
}

// peer-memory form of the two exchanges below; cnt_* <= p2p_halo_cap
static int p2p_exchange(Ctx* ctx, const double* v, int64_t n, int cnt_up, int cnt_down, int up, int down,
                        const double** lo, const double** hi) {
    const int par = (int)(ctx->p2p_xchg++ & 1);
    const unsigned long long seq = ++ctx->p2p_seq;
    k_push_ghost<<<1, 1024, 0, ctx->stream>>>(v, n, cnt_up, cnt_down, up >= 0 ? ctx->p2p_ghost_of(up, par, 0) : nullptr,
                                              down >= 0 ? ctx->p2p_ghost_of(down, par, 1) : nullptr, ctx->p2p_dev(), seq);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    *lo = down >= 0 ? ctx->p2p_ghost_local(par, 0) : nullptr;
    *hi = up >= 0 ? ctx->p2p_ghost_local(par, 1) : nullptr;
    return AK_OK;
}

// One-sweep GMRES on slabs: the opening vector of a cycle (r0) did not come out of a sweep, so its boundary rows are
// pushed into the neighbours' ghost rows of sweep slot `slot` here (peer stores + the all-to-all signal of k_push_ghost).
int sweep_push_rows(Ctx* ctx, const double* v, int64_t n, int count, bool periodic, int slot) {
    const int P = ctx->nranks, r = ctx->rank;
    if (P <= 1) return AK_OK;
    AK_REQUIRE(ctx->p2p_on && count <= ctx->p2p_halo_cap && slot >= 0 && slot < kSwGhostSlots, "sweep_push_rows: peer memory not set up for this row length");
    const int down = (r > 0) ? r - 1 : (periodic ? P - 1 : -1);
    const int up = (r < P - 1) ? r + 1 : (periodic ? 0 : -1);
    const unsigned long long seq = ++ctx->p2p_seq;
    k_push_ghost<<<1, 1024, 0, ctx->stream>>>(v, n, count, count, up >= 0 ? ctx->p2p_swghost_of(up, slot, 0) : nullptr,
                                              down >= 0 ? ctx->p2p_swghost_of(down, slot, 1) : nullptr, ctx->p2p_dev(), seq);
    ctx->launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

int collective_verdict(Ctx* ctx, int* rc) {
    if (ctx->nranks <= 1) return AK_OK;
    const double mine = (*rc != AK_OK) ? 1.0 : 0.0;
    double any = 0.0;
    AK_CUDA(cudaMemcpyAsync(ctx->dscal + 61, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    AK_TRY(allreduce_sum(ctx, ctx->dscal + 61, 1));
    AK_CUDA(cudaMemcpyAsync(&any, ctx->dscal + 61, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (any > 0.0 && *rc == AK_OK) {
        set_error("another rank could not grow its Krylov basis");
        *rc = AK_ERR_NOMEM;
    }
    return AK_OK;
}

int exchange_halo_rows(Ctx* ctx, const double* v, int64_t nx, int64_t ny, int32_t bc, const double** lo,
                       const double** hi) {
    const int P = ctx->nranks, r = ctx->rank;
    const bool periodic = (bc == AK_BC_PERIODIC);
    const int down = (r > 0) ? r - 1 : (periodic ? P - 1 : -1);  // owner of row gy0-1
    const int up = (r < P - 1) ? r + 1 : (periodic ? 0 : -1);    // owner of row gy0+ny
    if (ctx->p2p_on && nx <= ctx->p2p_halo_cap)  // boundary rows pushed into the neighbours' ghost slots over NVLink
        return p2p_exchange(ctx, v, nx * ny, (int)nx, (int)nx, up, down, lo, hi);
    if (ctx->halo_cap < nx) {
        AK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->halo_lo) AK_CUDA(cudaFree(ctx->halo_lo));
        if (ctx->halo_hi) AK_CUDA(cudaFree(ctx->halo_hi));
        AK_CUDA(cudaMalloc(&ctx->halo_lo, sizeof(double) * (size_t)nx));
        AK_CUDA(cudaMalloc(&ctx->halo_hi, sizeof(double) * (size_t)nx));
        ctx->halo_cap = nx;
    }
    // Per peer, NCCL matches sends and receives in posting order; with P == 2 and periodic
    // wrap both neighbours are the same rank, so "last row up" is posted before "first row down"
    // and "lo from down" before "hi from up".
    AK_NCCL(g_nccl.GroupStart());
    if (up >= 0) AK_NCCL(g_nccl.Send(v + (ny - 1) * nx, (size_t)nx, ncclFloat64, up, ctx->comm->comm, ctx->stream));
    if (down >= 0) AK_NCCL(g_nccl.Send(v, (size_t)nx, ncclFloat64, down, ctx->comm->comm, ctx->stream));
    if (down >= 0) AK_NCCL(g_nccl.Recv(ctx->halo_lo, (size_t)nx, ncclFloat64, down, ctx->comm->comm, ctx->stream));
    if (up >= 0) AK_NCCL(g_nccl.Recv(ctx->halo_hi, (size_t)nx, ncclFloat64, up, ctx->comm->comm, ctx->stream));
    AK_NCCL(g_nccl.GroupEnd());
    *lo = down >= 0 ? ctx->halo_lo : nullptr;
    *hi = up >= 0 ? ctx->halo_hi : nullptr;
    return AK_OK;
}

int exchange_halo_1d(Ctx* ctx, const double* v, int64_t n, int nlo, int nhi, bool periodic, const double** lo,
                     const double** hi) {
    const int P = ctx->nranks, r = ctx->rank;
    if (ctx->halo_cap < 8) {
        AK_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->halo_lo) AK_CUDA(cudaFree(ctx->halo_lo));
        if (ctx->halo_hi) AK_CUDA(cudaFree(ctx->halo_hi));
        AK_CUDA(cudaMalloc(&ctx->halo_lo, sizeof(double) * 8));
        AK_CUDA(cudaMalloc(&ctx->halo_hi, sizeof(double) * 8));
        ctx->halo_cap = 8;
    }
    AK_REQUIRE(nlo <= 8 && nhi <= 8 && n >= nlo && n >= nhi, "exchange_halo_1d: bad ghost width");
    const int left = (r > 0) ? r - 1 : (periodic ? P - 1 : -1);
    const int right = (r < P - 1) ? r + 1 : (periodic ? 0 : -1);
    if (ctx->p2p_on && ctx->p2p_halo_cap >= 8)  // this rank's last nlo values -> right neighbour's lo, first nhi -> left's hi
        return p2p_exchange(ctx, v, n, nlo, nhi, right, left, lo, hi);
    // posting order per peer as in exchange_halo_rows (P == 2 with wrap: both neighbours are the same rank)
    AK_NCCL(g_nccl.GroupStart());
    if (right >= 0) AK_NCCL(g_nccl.Send(v + (n - nlo), (size_t)nlo, ncclFloat64, right, ctx->comm->comm, ctx->stream));
    if (left >= 0) AK_NCCL(g_nccl.Send(v, (size_t)nhi, ncclFloat64, left, ctx->comm->comm, ctx->stream));
    if (left >= 0) AK_NCCL(g_nccl.Recv(ctx->halo_lo, (size_t)nlo, ncclFloat64, left, ctx->comm->comm, ctx->stream));
    if (right >= 0) AK_NCCL(g_nccl.Recv(ctx->halo_hi, (size_t)nhi, ncclFloat64, right, ctx->comm->comm, ctx->stream));
    AK_NCCL(g_nccl.GroupEnd());
    *lo = left >= 0 ? ctx->halo_lo : nullptr;
    *hi = right >= 0 ? ctx->halo_hi : nullptr;
    return AK_OK;
}

// ---- peer memory -------------------------------------------------------------------------
static inline size_t p2p_mail_doubles(int nranks) { return (size_t)kMailSlots * nranks * kMailRec; }
P2PDev Ctx::p2p_dev() const {
    P2PDev d{};
    d.nranks = nranks;
    d.rank = rank;
    d.mail_local = p2p_block;
    for (int q = 0; q < nranks && q < kMaxPeers; ++q) d.mail_peer[q] = (double*)p2p_peer_block[q];
    d.err = p2p_err;
    d.spin_cycles = 6000000000ll;  // ~3 s at 1.9 GHz: a peer that far behind means a bug, not load imbalance
    return d;
}
double* Ctx::p2p_halo_local(int parity, int hi) const {
    return p2p_block + p2p_mail_doubles(nranks) + ((size_t)parity * 2 + hi) * p2p_halo_cap;
}
double* Ctx::p2p_halo_of(int peer, int parity, int hi) const {
    return (double*)p2p_peer_block[peer] + p2p_mail_doubles(nranks) + ((size_t)parity * 2 + hi) * p2p_halo_cap;
}
double* Ctx::p2p_ghost_local(int parity, int hi) const { return p2p_halo_local(parity, hi) + (size_t)4 * p2p_halo_cap; }
double* Ctx::p2p_ghost_of(int peer, int parity, int hi) const { return p2p_halo_of(peer, parity, hi) + (size_t)4 * p2p_halo_cap; }
// one-sweep GMRES: [sweep mailboxes | kSwGhostSlots x {lo, hi} ghost rows] behind the 8 rows above
static inline size_t p2p_swmail_doubles(int nranks) { return (size_t)kMailSlots * nranks * kSwMailRec; }
static inline size_t p2p_sw_offset(int nranks, int64_t hcap) { return p2p_mail_doubles(nranks) + (size_t)8 * hcap; }
double* Ctx::p2p_swmail_of(int peer) const { return (double*)p2p_peer_block[peer] + p2p_sw_offset(nranks, p2p_halo_cap); }
double* Ctx::p2p_swghost_local(int slot, int hi) const {
    return p2p_block + p2p_sw_offset(nranks, p2p_halo_cap) + p2p_swmail_doubles(nranks) + ((size_t)slot * 2 + hi) * p2p_halo_cap;
}
double* Ctx::p2p_swghost_of(int peer, int slot, int hi) const {
    return (double*)p2p_peer_block[peer] + p2p_sw_offset(nranks, p2p_halo_cap) + p2p_swmail_doubles(nranks) +
           ((size_t)slot * 2 + hi) * p2p_halo_cap;
}

// ---- HaloVector layout bridge -----------------------------------------------------------
__global__ void k_halo_pack(double* __restrict__ compact, const double* __restrict__ padded, int64_t nx, int64_t ny) {
    const int64_t n = nx * ny;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / nx, i = t - j * nx;
        compact[t] = padded[(j + 1) * (nx + 2) + (i + 1)];
    }
}
__global__ void k_halo_unpack(double* __restrict__ padded, const double* __restrict__ compact, int64_t nx, int64_t ny,
                              int bc) {
    const int64_t px = nx + 2, py = ny + 2, n = px * py;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t J = t / px, I = t - J * px;
        int64_t i = I - 1, j = J - 1;
        double v = 0.0;
        bool ghost = (i < 0 || i >= nx || j < 0 || j >= ny);
        if (!ghost) {
            v = compact[j * nx + i];
        } else if (bc == AK_BC_PERIODIC) {
            // bc_periodic!: heat_2D.jl:15-26 (x ghosts first, then whole ghost rows incl. corners)
            if (i < 0) i = nx - 1; else if (i >= nx) i = 0;
            if (j < 0) j = ny - 1; else if (j >= ny) j = 0;
            v = compact[j * nx + i];
        }
        padded[t] = v;
    }
}

}  // namespace ak

using namespace ak;

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
AK_API int ak_abi_version(void) { return AK_ABI_VERSION; }
AK_API const char* ak_last_error(void) { return g_err; }

AK_API int ak_ctx_create(int device, ak_ctx** out) {
    AK_REQUIRE(out != nullptr, "ak_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("ak_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        (void)cudaGetLastError();
        return AK_ERR_CUDA;
    }
    AK_REQUIRE(device >= 0 && device < ndev, "ak_ctx_create: device index out of range");
    AK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("ak_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return AK_ERR_UNSUPPORTED;
    }
    ak_ctx* ctx = new ak_ctx();
    Ctx* c = &ctx->c;
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    // every failure below releases what was created so far (ak_ctx_destroy copes with a partially built context)
    auto build = [&]() -> int {
        AK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        cudaMemPoolProps pp;
        memset(&pp, 0, sizeof(pp));
        pp.allocType = cudaMemAllocationTypePinned;
        pp.handleTypes = cudaMemHandleTypeNone;
        pp.location.type = cudaMemLocationTypeDevice;
        pp.location.id = device;
        AK_CUDA(cudaMemPoolCreate(&c->pool, &pp));
        uint64_t keep = UINT64_MAX;
        AK_CUDA(cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep));
        AK_CUDA(cudaMalloc(&c->partials, sizeof(double) * kMaxPartials));
        AK_CUDA(cudaMalloc(&c->ticket, sizeof(unsigned int)));
        AK_CUDA(cudaMemset(c->ticket, 0, sizeof(unsigned int)));
        AK_CUDA(cudaMalloc(&c->dscal, sizeof(double) * 64));
        AK_CUDA(cudaMemset(c->dscal, 0, sizeof(double) * 64));
        AK_CUDA(cudaHostAlloc((void**)&c->hscal, sizeof(double) * 64, cudaHostAllocDefault));
        AK_CUDA(cudaEventCreate(&c->ev_t0));
        AK_CUDA(cudaEventCreate(&c->ev_t1));
        return AK_OK;
    };
    const int rc = build();
    if (rc != AK_OK) {
        ak_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return AK_OK;
}

AK_API int ak_ctx_destroy(ak_ctx* ctx) {
    if (!ctx) return AK_OK;
    Ctx* c = &ctx->c;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->p2p_block) {
        for (int q = 0; q < c->nranks && q < kMaxPeers; ++q)
            if (q != c->rank && c->p2p_peer_block[q]) cudaIpcCloseMemHandle(c->p2p_peer_block[q]);
        cudaFree(c->p2p_block);
        if (c->p2p_err) cudaFreeHost(c->p2p_err);
    }
    if (c->comm) {
        if (c->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm->comm);
        delete c->comm;
    }
    cudaFree(c->partials); cudaFree(c->ticket); cudaFree(c->dscal);
    cudaFree(c->halo_lo); cudaFree(c->halo_hi); cudaFree(c->halo_send);
    if (c->hscal) cudaFreeHost(c->hscal);
    if (c->ev_t0) cudaEventDestroy(c->ev_t0);
    if (c->ev_t1) cudaEventDestroy(c->ev_t1);
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    if (c->pool) cudaMemPoolDestroy(c->pool);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete ctx;
    return AK_OK;
}

AK_API int ak_ctx_sync(ak_ctx* ctx) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_ctx_sync: NULL ctx");
    AK_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return AK_OK;
}
AK_API uint64_t ak_ctx_stream(ak_ctx* ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->c.stream : 0; }
AK_API int64_t ak_ctx_launch_count(ak_ctx* ctx, int reset) {
    if (!ctx) return 0;
    int64_t v = ctx->c.launches;
    if (reset) ctx->c.launches = 0;
    return v;
}
AK_API int ak_timer_start(ak_ctx* ctx) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_timer_start: NULL ctx");
    AK_CUDA(cudaEventRecord(ctx->c.ev_t0, ctx->c.stream));
    return AK_OK;
}
AK_API int ak_timer_stop(ak_ctx* ctx, double* ms_out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && ms_out, "ak_timer_stop: NULL argument");
    AK_CUDA(cudaEventRecord(ctx->c.ev_t1, ctx->c.stream));
    AK_CUDA(cudaEventSynchronize(ctx->c.ev_t1));
    float ms = 0.f;
    AK_CUDA(cudaEventElapsedTime(&ms, ctx->c.ev_t0, ctx->c.ev_t1));
    *ms_out = (double)ms;
    return AK_OK;
}

AK_API int ak_profile_enable(ak_ctx* ctx, int on) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_profile_enable: NULL ctx");
    Ctx* c = &ctx->c;
    AK_CUDA(cudaStreamSynchronize(c->stream));
    c->prof_on = on != 0;
    c->prof_recs.clear();
    c->prof_pool_used = 0;
    return AK_OK;
}
AK_API int ak_profile_read(ak_ctx* ctx, int cls, int64_t* count_out, double* ms_total_out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && cls >= 0 && cls < PK_NUM, "ak_profile_read: bad argument");
    Ctx* c = &ctx->c;
    AK_CUDA(cudaStreamSynchronize(c->stream));
    int64_t cnt = 0;
    double tot = 0.0;
    for (const auto& r : c->prof_recs) {
        if (r.cls != cls) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { cnt++; tot += ms; }
    }
    if (count_out) *count_out = cnt;
    if (ms_total_out) *ms_total_out = tot;
    return AK_OK;
}

AK_API int ak_malloc(ak_ctx* ctx, int64_t n, double** out) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && out && n >= 0, "ak_malloc: bad argument");
    AK_CUDA(cudaSetDevice(ctx->c.device));
    cudaError_t e = pool_alloc(&ctx->c, (void**)out, sizeof(double) * (size_t)(n > 0 ? n : 1));
    if (e != cudaSuccess) {
        set_error("ak_malloc: %lld doubles: %s", (long long)n, cudaGetErrorString(e));
        (void)cudaGetLastError();
        return AK_ERR_NOMEM;
    }
    return AK_OK;
}
AK_API int ak_free(ak_ctx* ctx, double* p) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_free: NULL ctx");
    if (!p) return AK_OK;
    AK_CUDA(cudaFreeAsync(p, ctx->c.stream));  // stream-ordered: safe behind every kernel already enqueued
    return AK_OK;
}
AK_API int ak_upload(ak_ctx* ctx, double* dst_dev, const double* src_host, int64_t n) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_upload: bad argument");
    AK_CUDA(cudaMemcpyAsync(dst_dev, src_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->c.stream));
    AK_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return AK_OK;
}
AK_API int ak_download(ak_ctx* ctx, double* dst_host, const double* src_dev, int64_t n) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && n >= 0, "ak_download: bad argument");
    AK_CUDA(cudaMemcpyAsync(dst_host, src_dev, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->c.stream));
    AK_CUDA(cudaStreamSynchronize(ctx->c.stream));
    return AK_OK;
}
AK_API int ak_host_alloc(int64_t n, double** out) {
    AK_REQUIRE(out && n >= 0, "ak_host_alloc: bad argument");
    cudaError_t e = cudaHostAlloc((void**)out, sizeof(double) * (size_t)(n > 0 ? n : 1), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        set_error("ak_host_alloc: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return AK_ERR_NOMEM;
    }
    return AK_OK;
}
AK_API int ak_host_free(double* p) {
    if (p) AK_CUDA(cudaFreeHost(p));
    return AK_OK;
}

AK_API int ak_halo_pack(ak_ctx* ctx, double* compact, const double* padded, int64_t nx, int64_t ny) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && compact && padded && nx >= 1 && ny >= 1, "ak_halo_pack: bad argument");
    int64_t n = nx * ny;
    int blocks = (int)((n + 255) / 256 < ctx->c.num_sms * 8 ? (n + 255) / 256 : ctx->c.num_sms * 8);
    k_halo_pack<<<blocks, 256, 0, ctx->c.stream>>>(compact, padded, nx, ny);
    ctx->c.launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}
AK_API int ak_halo_unpack(ak_ctx* ctx, double* padded, const double* compact, int64_t nx, int64_t ny, int32_t bc) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && compact && padded && nx >= 1 && ny >= 1, "ak_halo_unpack: bad argument");
    int64_t n = (nx + 2) * (ny + 2);
    int blocks = (int)((n + 255) / 256 < ctx->c.num_sms * 8 ? (n + 255) / 256 : ctx->c.num_sms * 8);
    k_halo_unpack<<<blocks, 256, 0, ctx->c.stream>>>(padded, compact, nx, ny, bc);
    ctx->c.launches++;
    AK_CUDA(cudaGetLastError());
    return AK_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------
AK_API int ak_comm_unique_id(char id_out[128]) {
    AK_REQUIRE(id_out, "ak_comm_unique_id: NULL");
    AK_TRY(load_nccl());
    ncclUniqueId id;
    AK_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return AK_OK;
}
AK_API int ak_comm_init(ak_ctx* ctx, int nranks, int rank, const char id_in[128]) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && id_in && nranks >= 1 && rank >= 0 && rank < nranks, "ak_comm_init: bad argument");
    AK_REQUIRE(ctx->c.comm == nullptr, "ak_comm_init: communicator already initialised");
    if (nranks == 1) {
        ctx->c.rank = 0;
        ctx->c.nranks = 1;
        return AK_OK;
    }
    AK_TRY(load_nccl());
    AK_CUDA(cudaSetDevice(ctx->c.device));
    ncclUniqueId id;
    memcpy(id.internal, id_in, 128);
    Comm* cm = new Comm();
    ncclResult_t r = g_nccl.CommInitRank(&cm->comm, nranks, id, rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete cm;
        return AK_ERR_NCCL;
    }
    ctx->c.comm = cm;
    ctx->c.rank = rank;
    ctx->c.nranks = nranks;
    return AK_OK;
}
AK_API int ak_comm_enable_p2p(ak_ctx* ctx, int64_t halo_doubles) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx && halo_doubles >= 0, "ak_comm_enable_p2p: bad argument");
    Ctx* c = &ctx->c;
    if (c->nranks <= 1) return AK_OK;
    AK_REQUIRE(c->comm != nullptr, "ak_comm_enable_p2p: call ak_comm_init first");
    AK_REQUIRE(!c->p2p_on, "ak_comm_enable_p2p: already enabled");
    if (c->nranks > kMaxPeers) {
        set_error("ak_comm_enable_p2p: at most %d ranks per node", kMaxPeers);
        return AK_ERR_UNSUPPORTED;
    }
    AK_CUDA(cudaSetDevice(c->device));
    const int P = c->nranks;
    int64_t hcap = (halo_doubles + 3) & ~int64_t(3);
    if (hcap < 8) hcap = 8;  // the 1-D ghost exchanges move up to 8 values
    // mailboxes, 4 halo rows, 4 stand-alone ghost slots; sweep mailboxes and one pair of ghost rows per basis vector
    const size_t doubles = p2p_mail_doubles(P) + (size_t)8 * hcap + p2p_swmail_doubles(P) + (size_t)kSwGhostSlots * 2 * hcap;
    // cudaMalloc (not the stream-ordered pool): IPC handles exist only for plain allocations
    AK_CUDA(cudaMalloc(&c->p2p_block, sizeof(double) * doubles));
    AK_CUDA(cudaMemset(c->p2p_block, 0, sizeof(double) * doubles));
    AK_CUDA(cudaHostAlloc((void**)&c->p2p_err, sizeof(int), cudaHostAllocMapped));
    *c->p2p_err = 0;
    // Every rank goes through the same collectives even when a local step fails, and the verdict is
    // all-reduced, so that either all ranks switch to the peer-memory path or none does.
    int failed = 0;
    char why[256] = "";
    struct Card {  // what every rank tells the others: the IPC handle of its block and which physical GPU it sits on
        cudaIpcMemHandle_t handle;
        char busid[32];
    } mine;
    memset(&mine, 0, sizeof(mine));
    cudaError_t e = cudaIpcGetMemHandle(&mine.handle, c->p2p_block);
    if (e != cudaSuccess) {
        failed = 1;
        snprintf(why, sizeof(why), "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
    }
    if (cudaDeviceGetPCIBusId(mine.busid, (int)sizeof(mine.busid), c->device) != cudaSuccess) {
        (void)cudaGetLastError();
        snprintf(mine.busid, sizeof(mine.busid), "unknown-%d", c->rank);
    }
    // exchange the cards with an all-gather on the library's own communicator
    char *dsend = nullptr, *drecv = nullptr;
    AK_CUDA(cudaMalloc(&dsend, sizeof(mine)));
    AK_CUDA(cudaMalloc(&drecv, sizeof(mine) * P));
    AK_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(mine), cudaMemcpyHostToDevice, c->stream));
    AK_NCCL(g_nccl.AllGather(dsend, drecv, sizeof(mine), /*ncclChar*/ 0, c->comm->comm, c->stream));
    std::vector<Card> all((size_t)P);
    AK_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(mine) * P, cudaMemcpyDeviceToHost, c->stream));
    AK_CUDA(cudaStreamSynchronize(c->stream));
    AK_CUDA(cudaFree(dsend));
    AK_CUDA(cudaFree(drecv));
    // The mailbox kernels of different ranks wait for one another.  Two ranks on ONE GPU would be kernels that spin on
    // each other's flags on the same device: nothing guarantees that they run at the same time (Xid 109, context-switch
    // time-out).  Every rank sees the same cards, so every rank refuses alike.
    for (int q = 0; q < P && !failed; ++q)
        for (int q2 = q + 1; q2 < P && !failed; ++q2)
            if (strncmp(all[(size_t)q].busid, all[(size_t)q2].busid, sizeof(mine.busid)) == 0) {
                failed = 1;
                snprintf(why, sizeof(why), "ranks %d and %d share the GPU %s: the peer-memory path needs one GPU per rank "
                         "(use the NCCL path)", q, q2, all[(size_t)q].busid);
            }
    for (int q = 0; q < P && !failed; ++q) {
        if (q == c->rank) { c->p2p_peer_block[q] = c->p2p_block; continue; }
        e = cudaIpcOpenMemHandle(&c->p2p_peer_block[q], all[(size_t)q].handle, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            failed = 1;
            c->p2p_peer_block[q] = nullptr;
            snprintf(why, sizeof(why), "cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink is required)", q,
                     cudaGetErrorString(e));
            (void)cudaGetLastError();
        }
    }
    c->p2p_halo_cap = hcap;
    c->p2p_seq = 0;
    // verdict + barrier: nobody may write into a mailbox before its owner has zeroed it
    const double mine_failed = failed ? 1.0 : 0.0;
    AK_CUDA(cudaMemcpyAsync(c->dscal + 62, &mine_failed, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    AK_TRY(allreduce_sum(c, c->dscal + 62, 1));
    double any_failed = 0.0;
    AK_CUDA(cudaMemcpyAsync(&any_failed, c->dscal + 62, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    AK_CUDA(cudaStreamSynchronize(c->stream));
    if (any_failed > 0.0) {
        set_error("ak_comm_enable_p2p: %s", failed ? why : "another rank could not map its peers");
        return AK_ERR_CUDA;
    }
    c->p2p_on = true;
    return AK_OK;
}
AK_API int ak_comm_p2p_enabled(ak_ctx* ctx) {
    AK_ENTER(ctx); return (ctx && ctx->c.p2p_on) ? 1 : 0; }
AK_API int ak_comm_use_p2p(ak_ctx* ctx, int on) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_comm_use_p2p: NULL ctx");
    AK_REQUIRE(!on || ctx->c.p2p_block != nullptr, "ak_comm_use_p2p: peer memory was never mapped");
    AK_CUDA(cudaStreamSynchronize(ctx->c.stream));
    ctx->c.p2p_on = (on != 0);
    if (ctx->c.p2p_err) *ctx->c.p2p_err = 0;  // clears the time-out latch (collective by contract: all ranks call this)
    return AK_OK;
}

AK_API int ak_comm_rank(ak_ctx* ctx, int* rank, int* nranks) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_comm_rank: NULL ctx");
    if (rank) *rank = ctx->c.rank;
    if (nranks) *nranks = ctx->c.nranks;
    return AK_OK;
}
AK_API int ak_comm_barrier(ak_ctx* ctx) {
    AK_ENTER(ctx);
    AK_REQUIRE(ctx, "ak_comm_barrier: NULL ctx");
    Ctx* c = &ctx->c;
    if (c->nranks > 1) {
        AK_CUDA(cudaMemsetAsync(c->dscal + 63, 0, sizeof(double), c->stream));
        AK_TRY(allreduce_sum(c, c->dscal + 63, 1));
    }
    AK_CUDA(cudaStreamSynchronize(c->stream));
    return AK_OK;
}
