// Peer-memory transport of the multi-GPU path (one process per GPU on one NVSwitch node).
//
// Every rank owns a mailbox in its own HBM; the other ranks map it through CUDA IPC and write into
// it with plain stores over NVLink.  Two exchanges of the distributed CG use it:
//   * the shared-dof halo sum after the SpMV (OOFEM: EngngModel::updateSharedDofManagers,
//     src/core/engngm.C): the SpMV epilogue (spmv.cuh, MODE 2) -- or halo_push_kernel for a vector
//     that does not come out of a product -- stores this rank's contributions straight into the
//     sharers' mailboxes; halo_pull_kernel waits for the sharers' values and sums them in
//     ascending rank order (bit-identical on all sharers);
//   * the dot-product reductions (cg.cu: cg_scalars_p2p_kernel): an all-gather of the local sums
//     into every mailbox, summed in rank order by every rank.
// A value travels as two 8-byte words {32 data bits | sequence number}; a word is valid once it
// carries the sequence number of the exchange, so the path needs no fence, no flag, no packing
// buffer, no NCCL call and no host round trip.  The NCCL path (comm.cu) stays for communicators
// whose mailboxes cannot be mapped.
//
// Flow control: the buffer half is selected by the parity of the exchange counter.  A rank cannot
// start exchange e+2 with a sharer before that sharer has consumed exchange e, because completing
// e+1 requires the sharer's push of e+1, which it issues after its pull of e.
// Every wait is bounded (kWaitNs); a timeout raises the mailbox error word instead of hanging.
#include "comm.h"
#include "spmv_halo.h"
#include <string.h>

namespace ob200 {

constexpr unsigned long long kWaitNs = 4000000000ull;      // 4 s

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile( "mov.u64 %0, %%globaltimer;" : "=l"( t ) );
    return t;
}

// wait for the two words of an entry; false on timeout
__device__ __forceinline__ bool ll_load(const unsigned long long *slot, unsigned int seq, double &v)
{
    unsigned long long w0, w1, t0 = 0;
    for ( int spin = 0;; spin++ ) {
        asm volatile( "ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"( w0 ), "=l"( w1 ) : "l"( slot ) : "memory" );
        if ( (unsigned int)( w0 >> 32 ) == seq && (unsigned int)( w1 >> 32 ) == seq ) break;
        if ( spin == 64 ) t0 = global_ns();
        if ( spin > 64 ) {
            if ( global_ns() - t0 > kWaitNs ) return false;
            __nanosleep(32);
        }
    }
    v = __longlong_as_double((long long)( ( w0 & 0xffffffffull ) | ( w1 << 32 ) ));
    return true;
}

// y[eq[t]] -> the sharer's mailbox (vectors that do not come out of the fused SpMV)
__global__ void __launch_bounds__(256)
halo_push_kernel(const double *__restrict__ y, const int32_t *__restrict__ eq, int64_t nshared, int nneigh,
                 const int64_t *__restrict__ off, unsigned long long *const *__restrict__ peer_data, int64_t half_words,
                 unsigned int seq, const int *done)
{
    if ( done && *done ) return;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < nshared; t += stride ) {
        int k = 0;
        while ( k + 1 < nneigh && t >= off[k + 1] ) k++;
        ll_store(peer_data[k] + half_words + 2 * ( t - off[k] ), y[eq[t]], seq);
    }
}

// y[eq] = sum over sharers in ascending rank order (own value inserted at its rank position), the
// sharers' values taken from this rank's mailbox as they arrive.  dot_p != null: also the per-CTA
// partial sums of dot_p[eq] * y[eq] over the shared dofs this rank owns.
__global__ void __launch_bounds__(256)
halo_pull_kernel(double *__restrict__ y, const int32_t *__restrict__ ueq, const int32_t *__restrict__ uptr,
                 const int32_t *__restrict__ umb, const int32_t *__restrict__ ubefore, const unsigned long long *mail,
                 unsigned int seq, int64_t nuniq, const double *__restrict__ dot_p, const unsigned char *__restrict__ owned,
                 double *__restrict__ partials, int *error, const int *done)
{
    if ( done && *done ) return;
    __shared__ double scratch[8];
    double dot = 0.0;
    bool ok = true;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t u = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; u < nuniq; u += stride ) {
        const int b = uptr[u], e = uptr[u + 1], nb = ubefore[u];
        const int row = ueq[u];
        const double own = y[row];
        double acc = 0.0;
        for ( int t = b; t < e; t++ ) {
            if ( t - b == nb ) acc += own;
            double v = 0.0;
            ok = ll_load(mail + 2 * (int64_t) umb[t], seq, v) && ok;
            acc += v;
        }
        if ( nb == e - b ) acc += own;
        y[row] = acc;
        if ( dot_p && owned[row] ) dot += dot_p[row] * acc;
    }
    if ( !ok ) atomicExch(error, 1);
    if ( dot_p ) {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if ( lane == 0 ) scratch[wid] = dot;
        __syncthreads();
        if ( threadIdx.x == 0 ) {
            double t = 0.0;
            for ( int w = 0; w < (int)( blockDim.x >> 5 ); w++ ) t += scratch[w];
            partials[blockIdx.x] = t;
        }
    }
}

int comm_p2p_pull(ob200_comm *c, double *y, unsigned int seq, const double *dot_p, const int *done)
{
    ob200_context *ctx = c->ctx;
    const ob200_mailbox_layout L = mailbox_layout(c->nranks, c->cap);
    const int64_t par = seq & 1u;
    OB_LAUNCH(ctx, halo_pull_kernel, c->pull_grid, 256, 0, y, c->uniq_eq.p, c->uniq_ptr.p, c->uniq_mb.p, c->uniq_before.p,
              reinterpret_cast< const unsigned long long * >( c->mailbox + L.data ) + 2 * par * L.data_half, seq, c->nuniq,
              dot_p, c->owned.p, c->pull_partials.p, reinterpret_cast< int * >( c->mailbox + L.error ), done);
    return OB200_OK;
}

int comm_p2p_push(ob200_comm *c, const double *y, unsigned int seq, const int *done)
{
    ob200_context *ctx = c->ctx;
    const ob200_mailbox_layout L = mailbox_layout(c->nranks, c->cap);
    const int64_t par = seq & 1u;
    const int grid = ctx->shape.grid(c->nshared, 256, 2);
    OB_LAUNCH(ctx, halo_push_kernel, grid, 256, 0, y, c->shared_eq.p, c->nshared, c->nneigh, c->p_off.p, c->p_data.p,
              2 * par * L.data_half, seq, done);
    return OB200_OK;
}

int comm_p2p_exchange_add(ob200_comm *c, double *y, const int *done)
{
    if ( c->nranks == 1 || c->nneigh == 0 ) return OB200_OK;
    OB_REQUIRE(c->uniq_mb.p, OB200_EINVAL, "comm: peer-memory transport enabled but the halo has not been described");
    const unsigned int seq = ++c->halo_seq;
    OB_CHECK( comm_p2p_push(c, y, seq, done) );
    return comm_p2p_pull(c, y, seq, nullptr, done);
}

int comm_p2p_check(ob200_comm *c)
{
    if ( !c->p2p || !c->mailbox ) return OB200_OK;
    const ob200_mailbox_layout L = mailbox_layout(c->nranks, c->cap);
    int h = 0;
    OB_CUDA( cudaMemcpyAsync(&h, c->mailbox + L.error, sizeof( int ), cudaMemcpyDeviceToHost, c->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(c->ctx->stream) );
    OB_REQUIRE(h == 0, OB200_ENCCL, "comm: a wait on a peer GPU timed out (rank %d); a rank of the job is missing or failed", c->rank);
    return OB200_OK;
}

// device-side tables that depend on both the halo description and the mapped mailboxes
int comm_p2p_prepare(ob200_comm *c)
{
    ob200_context *ctx = c->ctx;
    if ( !c->p2p || (int) c->peer_base.size() != c->nranks ) return OB200_OK;
    const ob200_mailbox_layout L = mailbox_layout(c->nranks, c->cap);
    auto put = [&](auto &buf, const auto *src, int64_t n) -> int {
        OB_CHECK( buf.alloc(n > 0 ? n : 1) );
        if ( n ) OB_CUDA( cudaMemcpyAsync(buf.p, src, sizeof( *src ) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream) );
        return OB200_OK;
    };
    // scalar entries of this rank in every mailbox
    std::vector< unsigned long long * > sd(c->nranks);
    for ( int r = 0; r < c->nranks; r++ )
        sd[r] = reinterpret_cast< unsigned long long * >( c->peer_base[r] + L.sdata ) + 2 * (int64_t) c->rank * kScalSlots;
    OB_CHECK( put(c->p_sdata, sd.data(), c->nranks) );
    if ( c->nneigh > 0 && c->shared_eq.p ) {
        std::vector< unsigned long long * > pd(c->nneigh);
        for ( int k = 0; k < c->nneigh; k++ ) {
            const int r = c->neigh_rank[k];
            OB_REQUIRE(c->neigh_offset[k + 1] - c->neigh_offset[k] <= c->cap, OB200_ECAPACITY,
                       "comm: %lld dofs shared with rank %d exceed the mailbox capacity %lld",
                       (long long)( c->neigh_offset[k + 1] - c->neigh_offset[k] ), r, (long long) c->cap);
            pd[k] = reinterpret_cast< unsigned long long * >( c->peer_base[r] + L.data ) + 2 * (int64_t) c->rank * c->cap;
        }
        OB_CHECK( put(c->p_data, pd.data(), c->nneigh) );
        OB_CHECK( put(c->p_off, c->neigh_offset.data(), c->nneigh + 1) );
        // per (shared dof, sharer): where the sharer expects this rank's value, where this rank finds the sharer's
        const int64_t nidx = c->nshared;
        std::vector< int32_t > uidx(nidx), umb(nidx), ueq(c->nuniq);
        std::vector< unsigned long long * > dst(nidx);
        OB_CUDA( cudaMemcpyAsync(uidx.data(), c->uniq_idx.p, sizeof( int32_t ) * (size_t) nidx, cudaMemcpyDeviceToHost, ctx->stream) );
        OB_CUDA( cudaMemcpyAsync(ueq.data(), c->uniq_eq.p, sizeof( int32_t ) * (size_t) c->nuniq, cudaMemcpyDeviceToHost, ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        for ( int64_t i = 0; i < nidx; i++ ) {
            const int64_t t = uidx[i];
            int k = 0;
            while ( k + 1 < c->nneigh && t >= c->neigh_offset[k + 1] ) k++;
            umb[i] = (int32_t)( (int64_t) c->neigh_rank[k] * c->cap + ( t - c->neigh_offset[k] ) );
            dst[i] = pd[k] + 2 * ( t - c->neigh_offset[k] );
        }
        OB_CHECK( put(c->uniq_mb, umb.data(), nidx) );
        OB_CHECK( put(c->push_dst, dst.data(), nidx) );
        std::vector< int32_t > route(c->neq > 0 ? c->neq : 1, -1);
        for ( int64_t u = 0; u < c->nuniq; u++ ) route[ueq[u]] = (int32_t) u;
        OB_CHECK( put(c->route, route.data(), c->neq) );
        c->pull_grid = ctx->shape.grid(c->nuniq, 256, 2);
        OB_CHECK( c->pull_partials.alloc(c->pull_grid) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    }
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

} // namespace ob200

using namespace ob200;

extern "C" {

int ob200_comm_p2p_export(ob200_comm *c, int64_t cap, void *handle64)
{
    if ( c ) ob200::bind_stream(c->ctx);
    OB_REQUIRE(c && handle64 && cap >= 0, OB200_EINVAL, "comm_p2p_export: bad argument");
    OB_REQUIRE(c->nranks <= 32, OB200_ECAPACITY, "comm_p2p_export: %d ranks exceed the peer-memory transport (32)", c->nranks);
    OB_REQUIRE((int64_t) c->nranks * ( cap > 0 ? cap : 1 ) < ( (int64_t) 1 << 30 ), OB200_ECAPACITY, "comm_p2p_export: mailbox too large");
    static_assert( sizeof( cudaIpcMemHandle_t ) == 64, "cudaIpcMemHandle_t size" );
    OB_CUDA( cudaSetDevice(c->ctx->device) );
    if ( c->mailbox ) { cudaFree(c->mailbox); c->mailbox = nullptr; }
    c->cap = cap > 0 ? cap : 1;
    const ob200_mailbox_layout L = mailbox_layout(c->nranks, c->cap);
    OB_CUDA( cudaMalloc(&c->mailbox, (size_t) L.bytes) );
    OB_CUDA( cudaMemset(c->mailbox, 0, (size_t) L.bytes) );
    OB_CUDA( cudaDeviceSynchronize() );
    cudaIpcMemHandle_t h;
    OB_CUDA( cudaIpcGetMemHandle(&h, c->mailbox) );
    memcpy(handle64, &h, sizeof( h ));
    return OB200_OK;
}

int ob200_comm_p2p_open(ob200_comm *c, const void *handles)
{
    if ( c ) ob200::bind_stream(c->ctx);
    OB_REQUIRE(c && handles && c->mailbox, OB200_EINVAL, "comm_p2p_open: export the mailbox first");
    OB_CUDA( cudaSetDevice(c->ctx->device) );
    c->peer_base.assign(c->nranks, nullptr);
    for ( int r = 0; r < c->nranks; r++ ) {
        if ( r == c->rank ) { c->peer_base[r] = c->mailbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast< const char * >( handles ) + (size_t) r * sizeof( h ), sizeof( h ));
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if ( e != cudaSuccess ) {
            cudaGetLastError();
            for ( int q = 0; q < r; q++ )
                if ( q != c->rank && c->peer_base[q] ) cudaIpcCloseMemHandle(c->peer_base[q]);
            c->peer_base.clear();
            set_error("comm_p2p_open: cannot map the mailbox of rank %d (%s)", r, cudaGetErrorString(e));
            return OB200_ENCCL;
        }
        c->peer_base[r] = static_cast< char * >( p );
    }
    c->p2p = true;
    return comm_p2p_prepare(c);
}

int ob200_comm_p2p_enabled(const ob200_comm *c) { return c && c->p2p ? 1 : 0; }

/* collective decision of the caller: some rank could not map -> every rank goes back to the NCCL transport */
int ob200_comm_p2p_disable(ob200_comm *c)
{
    OB_REQUIRE(c, OB200_EINVAL, "comm_p2p_disable: null communicator");
    c->p2p = false;
    return OB200_OK;
}

} // extern "C"
