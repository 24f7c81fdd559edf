// Multi-GPU: one process per GPU, element-partitioned mesh (oofem2part-style node-cut: each
// rank owns a set of elements, nodes on partition boundaries are replicated).  Each rank
// assembles only its own elements; shared dofs are completed by summing the neighbours'
// contributions (what OOFEM does in EngngModel::updateSharedDofManagers, src/core/engngm.C)
// with ncclSend/ncclRecv over NVLink, dot products by ncclAllReduce.
//
// NCCL is resolved at run time with dlopen so that the single-GPU library has no hard
// dependency on it; if torch already loaded its bundled libnccl.so.2 the same copy is used.
#include "comm.h"
#include <dlfcn.h>
#include <nccl.h>
#include <algorithm>
#include <map>

namespace ob200 {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t ( *GetUniqueId )(ncclUniqueId *) = nullptr;
    ncclResult_t ( *CommInitRank )(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t ( *CommDestroy )(ncclComm_t) = nullptr;
    ncclResult_t ( *AllReduce )(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t ( *Send )(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t ( *Recv )(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t ( *GroupStart )() = nullptr;
    ncclResult_t ( *GroupEnd )() = nullptr;
    const char *( *GetErrorString )(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
    if ( g_nccl.handle ) return OB200_OK;
    const char *names[] = { "libnccl.so.2", "libnccl.so" };
    void *h = nullptr;
    for ( const char *n : names ) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if ( h ) break;
    }
    OB_REQUIRE(h, OB200_ENCCL, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define SYM(field, name)                                                             \
    g_nccl.field = reinterpret_cast< decltype( g_nccl.field ) >( dlsym(h, name) );   \
    OB_REQUIRE(g_nccl.field, OB200_ENCCL, "NCCL symbol %s missing", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.handle = h;
    return OB200_OK;
}

#define OB_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if ( r__ != ncclSuccess ) {                                                            \
            ob200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r__)); \
            return OB200_ENCCL;                                                                \
        }                                                                                      \
    } while ( 0 )

__global__ void halo_pack_kernel(const double *__restrict__ y, const int32_t *__restrict__ eq, int64_t n, double *__restrict__ buf)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) buf[t] = y[eq[t]];
}

// y[eq] = sum over sharers in ascending rank order (own value inserted at its rank position):
// every sharer computes the same floating-point sum, so replicated dofs stay bit-identical.
__global__ void halo_unpack_kernel(double *__restrict__ y, const int32_t *__restrict__ ueq, const int32_t *__restrict__ uptr,
                                   const int32_t *__restrict__ uidx, const int32_t *__restrict__ ubefore,
                                   const double *__restrict__ recv, int64_t nuniq)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t u = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; u < nuniq; u += stride ) {
        const int b = uptr[u], e = uptr[u + 1], nb = ubefore[u];
        const double own = y[ueq[u]];
        double acc = 0.0;
        for ( int t = b; t < e; t++ ) {
            if ( t - b == nb ) acc += own;
            acc += recv[uidx[t]];
        }
        if ( nb == e - b ) acc += own;
        y[ueq[u]] = acc;
    }
}

int comm_allreduce_sum(ob200_comm *c, double *dev, int n)
{
    if ( c->nranks == 1 ) return OB200_OK;
    OB_NCCL( g_nccl.AllReduce(dev, dev, (size_t) n, ncclFloat64, ncclSum, (ncclComm_t) c->nccl, c->ctx->stream) );
    return OB200_OK;
}

int comm_exchange_add(ob200_comm *c, double *y, const int *done)
{
    if ( c->p2p ) return comm_p2p_exchange_add(c, y, done);
    if ( c->nshared == 0 ) return OB200_OK;
    ob200_context *ctx = c->ctx;
    int grid = ctx->shape.grid(c->nshared, 256, 4);
    OB_LAUNCH(ctx, halo_pack_kernel, grid, 256, 0, y, c->shared_eq.p, c->nshared, c->sendbuf.p);
    OB_NCCL( g_nccl.GroupStart() );
    for ( int k = 0; k < c->nneigh; k++ ) {
        size_t cnt = (size_t)( c->neigh_offset[k + 1] - c->neigh_offset[k] );
        OB_NCCL( g_nccl.Send(c->sendbuf.p + c->neigh_offset[k], cnt, ncclFloat64, c->neigh_rank[k], (ncclComm_t) c->nccl, ctx->stream) );
        OB_NCCL( g_nccl.Recv(c->recvbuf.p + c->neigh_offset[k], cnt, ncclFloat64, c->neigh_rank[k], (ncclComm_t) c->nccl, ctx->stream) );
    }
    OB_NCCL( g_nccl.GroupEnd() );
    grid = ctx->shape.grid(c->nuniq, 256, 4);
    OB_LAUNCH(ctx, halo_unpack_kernel, grid, 256, 0, y, c->uniq_eq.p, c->uniq_ptr.p, c->uniq_idx.p, c->uniq_before.p,
              c->recvbuf.p, c->nuniq);
    return OB200_OK;
}

} // namespace ob200

using namespace ob200;

extern "C" {

int ob200_comm_unique_id(void *id128)
{
    OB_REQUIRE(id128, OB200_EINVAL, "comm_unique_id: null buffer");
    OB_CHECK( nccl_load() );
    static_assert( sizeof( ncclUniqueId ) == 128, "ncclUniqueId size" );
    OB_NCCL( g_nccl.GetUniqueId(reinterpret_cast< ncclUniqueId * >( id128 )) );
    return OB200_OK;
}

int ob200_comm_create(ob200_context *ctx, int nranks, int rank, const void *id128, ob200_comm **out)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && out && nranks >= 1 && rank >= 0 && rank < nranks, OB200_EINVAL, "comm_create: bad argument");
    ob200_comm *c = new ob200_comm();
    c->ctx = ctx;
    c->nranks = nranks;
    c->rank = rank;
    if ( nranks > 1 ) {
        if ( !id128 ) { delete c; set_error("comm_create: unique id required for nranks > 1"); return OB200_EINVAL; }
        int rc = nccl_load();
        if ( rc < 0 ) { delete c; return rc; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof( id ));
        cudaSetDevice(ctx->device);
        ncclComm_t comm;
        ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
        if ( r != ncclSuccess ) {
            set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
            delete c;
            return OB200_ENCCL;
        }
        c->nccl = comm;
    }
    *out = c;
    return OB200_OK;
}

void ob200_comm_destroy(ob200_comm *c)
{
    if ( !c ) return;
    if ( c->nccl ) {
        cudaStreamSynchronize(c->ctx->stream);
        g_nccl.CommDestroy((ncclComm_t) c->nccl);
    }
    for ( size_t r = 0; r < c->peer_base.size(); r++ )
        if ( c->peer_base[r] && c->peer_base[r] != c->mailbox ) cudaIpcCloseMemHandle(c->peer_base[r]);
    if ( c->mailbox ) cudaFree(c->mailbox);
    delete c;
}

int ob200_comm_set_halo(ob200_comm *c, int32_t neq, int nneigh, const int32_t *neigh_rank, const int64_t *neigh_offset,
                        const int32_t *shared_eq, const uint8_t *owned)
{
    if ( c ) ob200::bind_stream(c->ctx);
    OB_REQUIRE(c && neq >= 0 && nneigh >= 0, OB200_EINVAL, "comm_set_halo: bad argument");
    OB_REQUIRE(nneigh == 0 || ( neigh_rank && neigh_offset && shared_eq ), OB200_EINVAL, "comm_set_halo: null halo arrays");
    ob200_context *ctx = c->ctx;
    c->neq = neq;
    c->nneigh = nneigh;
    c->neigh_rank.assign(neigh_rank, neigh_rank + nneigh);
    c->neigh_offset.assign(1, 0);
    if ( nneigh ) c->neigh_offset.assign(neigh_offset, neigh_offset + nneigh + 1);
    c->nshared = nneigh ? neigh_offset[nneigh] : 0;
    for ( int k = 0; k < nneigh; k++ ) {
        OB_REQUIRE(neigh_rank[k] >= 0 && neigh_rank[k] < c->nranks && neigh_rank[k] != c->rank, OB200_EINVAL,
                   "comm_set_halo: neighbour rank %d invalid", neigh_rank[k]);
        OB_REQUIRE(neigh_offset[k + 1] >= neigh_offset[k], OB200_EINVAL, "comm_set_halo: offsets not monotone");
    }
    for ( int64_t t = 0; t < c->nshared; t++ )
        OB_REQUIRE(shared_eq[t] >= 0 && shared_eq[t] < neq, OB200_EINVAL, "comm_set_halo: shared equation %d out of range", shared_eq[t]);
    // canonical accumulation order: per unique shared equation, buffer slots by ascending neighbour rank
    std::vector< int > order(nneigh);
    for ( int k = 0; k < nneigh; k++ ) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return neigh_rank[a] < neigh_rank[b]; });
    std::map< int32_t, std::vector< int32_t > > slots;
    std::map< int32_t, int32_t > before;
    for ( int k : order )
        for ( int64_t t = neigh_offset[k]; t < neigh_offset[k + 1]; t++ ) {
            slots[shared_eq[t]].push_back((int32_t) t);
            if ( neigh_rank[k] < c->rank ) before[shared_eq[t]]++;
        }
    std::vector< int32_t > ueq, uptr(1, 0), uidx, ubef;
    for ( auto &kv : slots ) {
        ueq.push_back(kv.first);
        for ( int32_t s : kv.second ) uidx.push_back(s);
        uptr.push_back((int32_t) uidx.size());
        ubef.push_back(before.count(kv.first) ? before[kv.first] : 0);
    }
    c->nuniq = (int64_t) ueq.size();
    auto put = [&](auto &buf, const auto *src, int64_t n) -> int {
        OB_CHECK( buf.alloc(n > 0 ? n : 1) );
        if ( n ) OB_CUDA( cudaMemcpyAsync(buf.p, src, sizeof( *src ) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream) );
        return OB200_OK;
    };
    OB_CHECK( put(c->shared_eq, shared_eq, c->nshared) );
    OB_CHECK( put(c->uniq_eq, ueq.data(), c->nuniq) );
    OB_CHECK( put(c->uniq_ptr, uptr.data(), (int64_t) uptr.size()) );
    OB_CHECK( put(c->uniq_idx, uidx.data(), (int64_t) uidx.size()) );
    OB_CHECK( put(c->uniq_before, ubef.data(), c->nuniq) );
    OB_CHECK( c->sendbuf.alloc(c->nshared > 0 ? c->nshared : 1) );
    OB_CHECK( c->recvbuf.alloc(c->nshared > 0 ? c->nshared : 1) );
    std::vector< unsigned char > own(neq > 0 ? neq : 1, 1);
    if ( owned ) for ( int32_t i = 0; i < neq; i++ ) own[i] = owned[i] ? 1 : 0;
    OB_CHECK( put(c->owned, own.data(), neq) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    if ( c->p2p ) OB_CHECK( comm_p2p_prepare(c) );
    return OB200_OK;
}

int ob200_comm_exchange_add(ob200_comm *c, double *y_dev)
{
    if ( c ) ob200::bind_stream(c->ctx);
    OB_REQUIRE(c && y_dev, OB200_EINVAL, "comm_exchange_add: null argument");
    return comm_exchange_add(c, y_dev);
}

} // extern "C"
