// Internal definitions shared by the translation units of liboofem_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/oofem_b200.h"

namespace ob200 {

void set_error(const char *fmt, ...);

#define OB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if ( e__ != cudaSuccess ) {                                                            \
            ob200::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return OB200_ECUDA;                                                                \
        }                                                                                      \
    } while ( 0 )

#define OB_CHECK(expr)                      \
    do {                                    \
        int rc__ = (expr);                  \
        if ( rc__ < 0 ) return rc__;        \
    } while ( 0 )

#define OB_REQUIRE(cond, code, ...)         \
    do {                                    \
        if ( !( cond ) ) {                  \
            ob200::set_error(__VA_ARGS__);  \
            return code;                    \
        }                                   \
    } while ( 0 )

constexpr int kWarp = 32;

inline int64_t ceil_div(int64_t a, int64_t b) { return ( a + b - 1 ) / b; }

// grid for a grid-stride kernel: a whole number of waves on the device
struct LaunchShape {
    int sms = 148;
    int grid(int64_t work_items, int block, int blocks_per_sm) const
    {
        int64_t need = ceil_div(work_items, block);
        int64_t cap = (int64_t) sms * blocks_per_sm;
        if ( need < 1 ) need = 1;
        return (int)( need < cap ? need : cap );
    }
};

// cudaFuncSetAttribute is per device: `mask` (one static word per call site) remembers the devices it was done for
inline bool attr_needed(unsigned long long &mask, int device)
{
    const unsigned long long bit = 1ull << ( device & 63 );
    if ( mask & bit ) return false;
    mask |= bit;
    return true;
}

// The stream of the context the calling thread is working for.  Every C-ABI entry point binds it
// (bind_stream) before touching device memory: DevBuf allocations are stream-ordered on it.
struct StreamSlot { cudaStream_t stream = nullptr; };
StreamSlot &current_stream();
bool stream_alive(cudaStream_t s);          // false once the owning context has been destroyed

// simple owning device buffer.  Stream-ordered (cudaMallocAsync from the device's default pool, whose
// release threshold context_create raises so that freed blocks stay cached): creating and dropping
// work buffers inside a call costs no cudaMalloc/cudaFree device synchronisation.
template< class T >
struct DevBuf {
    T *p = nullptr;
    int64_t n = 0;
    cudaStream_t s = nullptr;
    int alloc(int64_t count)
    {
        release();
        n = count;
        s = current_stream().stream;
        if ( count > 0 ) {
            if ( stream_alive(s) ) OB_CUDA( cudaMallocAsync(&p, sizeof( T ) * (size_t) count, s) );
            else OB_CUDA( cudaMalloc(&p, sizeof( T ) * (size_t) count) );
        }
        return OB200_OK;
    }
    void release()
    {
        if ( p ) {
            // an object that outlives its context frees synchronously
            if ( !stream_alive(s) || cudaFreeAsync(p, s) != cudaSuccess ) {
                cudaGetLastError();
                cudaFree(p);
            }
        }
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), s(o.s) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if ( this != &o ) {
            release();
            p = o.p; n = o.n; s = o.s;
            o.p = nullptr; o.n = 0;
        }
        return *this;
    }
};

} // namespace ob200

struct ob200_prof_rec {
    const char *name;
    cudaEvent_t start, stop;
};

struct ob200_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    // second stream + event for uploads that may overlap kernels of the main stream (elemset_create: the
    // location arrays travel while the node incidence is being built from the connectivity)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_event = nullptr;
    cudaDeviceProp prop;
    ob200::LaunchShape shape;
    int64_t launches = 0;
    // optional per-kernel CUDA-event timing (bench.py roofline): every launch is bracketed by
    // two events on the launching stream; ob200_context_profile_collect folds them into totals
    bool profiling = false;
    std::vector< ob200_prof_rec > prof_pending;
    std::vector< cudaEvent_t > prof_pool;
    std::map< std::string, std::pair< double, int64_t > > prof_total;
    cudaEvent_t prof_event()
    {
        cudaEvent_t e;
        if ( !prof_pool.empty() ) { e = prof_pool.back(); prof_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    // L2 flush scratch
    ob200::DevBuf< char > flush;
    // reduction scratch (partials) and small device scalars, used by CG
    ob200::DevBuf< double > partials;
    // error word of the kernels whose waits are bounded (assemble_cluster.cu): device word + pinned host copy that follows
    // every such launch on the stream; looked at by the next call of that kind and by ob200_context_sync
    int *kerr_dev = nullptr;
    volatile int *kerr_host = nullptr;
    // assembly schedule of the last element set that was destroyed (element_kernels.cu): an element set created from the
    // same mesh, numbering and materials adopts it instead of rebuilding it
    void *sched_cache = nullptr;
    void ( *sched_cache_free )(void *) = nullptr;
};

namespace ob200 {
// Also makes the context's device current when another one is (a process that drives several devices; with one process
// per GPU the check never fires): streams, stream-ordered allocations and launches below all belong to ctx->device.
inline void bind_stream(ob200_context *ctx)
{
    current_stream().stream = ctx->stream;
    int dev = -1;
    if ( cudaGetDevice(&dev) == cudaSuccess && dev != ctx->device ) cudaSetDevice(ctx->device);
}
}

#define OB_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
    do {                                                                                 \
        ob200_prof_rec pr__ = { #kernel, nullptr, nullptr };                             \
        if ( ( ctx )->profiling ) {                                                      \
            pr__.start = ( ctx )->prof_event();                                          \
            pr__.stop = ( ctx )->prof_event();                                           \
            cudaEventRecord(pr__.start, ( ctx )->stream);                                \
        }                                                                                \
        kernel<<< ( grid ), ( block ), ( smem ), ( ctx )->stream >>>(__VA_ARGS__);       \
        if ( ( ctx )->profiling ) {                                                      \
            cudaEventRecord(pr__.stop, ( ctx )->stream);                                 \
            ( ctx )->prof_pending.push_back(pr__);                                       \
        }                                                                                \
        ( ctx )->launches++;                                                             \
        OB_CUDA( cudaGetLastError() );                                                   \
    } while ( 0 )

// stage a host or device input in device memory; owns the copy when it came from the host
template< class T >
struct Staged {
    const T *d = nullptr;
    ob200::DevBuf< T > own;
    int stage(ob200_context *ctx, const T *src, int64_t n, int on_device)
    {
        if ( on_device || n == 0 ) {
            d = src;
            return OB200_OK;
        }
        OB_CHECK( own.alloc(n) );
        OB_CUDA( cudaMemcpyAsync(own.p, src, sizeof( T ) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream) );
        d = own.p;
        return OB200_OK;
    }
};

// an output that lives on the device during the call and is copied back if the caller gave a host pointer
template< class T >
struct StagedOut {
    T *d = nullptr;
    T *host = nullptr;
    int64_t n = 0;
    ob200::DevBuf< T > own;
    int stage(ob200_context *ctx, T *dst, int64_t count, int on_device, bool copy_in = false)
    {
        n = count;
        if ( on_device || count == 0 ) {
            d = dst;
            return OB200_OK;
        }
        host = dst;
        OB_CHECK( own.alloc(count) );
        d = own.p;
        if ( copy_in ) OB_CUDA( cudaMemcpyAsync(d, dst, sizeof( T ) * (size_t) count, cudaMemcpyHostToDevice, ctx->stream) );
        return OB200_OK;
    }
    int finish(ob200_context *ctx)
    {
        if ( host && n > 0 ) {
            OB_CUDA( cudaMemcpyAsync(host, d, sizeof( T ) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream) );
            OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        }
        return OB200_OK;
    }
};
