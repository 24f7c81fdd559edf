// Context, error reporting and raw device-memory helpers of the C ABI.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <set>

namespace ob200 {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof( g_err ), fmt, ap);
    va_end(ap);
}

StreamSlot &current_stream()
{
    static thread_local StreamSlot slot;
    return slot;
}

static std::mutex g_streams_mu;
static std::set< cudaStream_t > g_streams;
bool stream_alive(cudaStream_t s)
{
    std::lock_guard< std::mutex > lk(g_streams_mu);
    return g_streams.count(s) != 0;
}
static void stream_register(cudaStream_t s, bool alive)
{
    std::lock_guard< std::mutex > lk(g_streams_mu);
    if ( alive ) g_streams.insert(s);
    else g_streams.erase(s);
}

__global__ void flush_kernel(char *p, int64_t n, char v)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x * 16;
    for ( int64_t t = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) * 16; t + 16 <= n; t += stride )
        *reinterpret_cast< int4 * >( p + t ) = make_int4(v, v, v, v);
}
} // namespace ob200

using namespace ob200;

extern "C" {

const char *ob200_last_error(void) { return g_err; }
const char *ob200_version(void) { return "oofem_b200 0.1 (sm_100a)"; }

int ob200_context_create(int device, ob200_context **out)
{
    OB_REQUIRE(out, OB200_EINVAL, "context_create: null out");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if ( e != cudaSuccess || n == 0 ) {
        set_error("context_create: no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return OB200_ENODEVICE;
    }
    OB_REQUIRE(device >= 0 && device < n, OB200_EINVAL, "context_create: device %d out of range [0,%d)", device, n);
    OB_CUDA( cudaSetDevice(device) );
    ob200_context *ctx = new ob200_context();
    ctx->device = device;
    if ( cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess ||
         cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ) {
        set_error("context_create: device query / stream creation failed");
        delete ctx;
        return OB200_ECUDA;
    }
    if ( cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
         cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming) != cudaSuccess ) {
        set_error("context_create: copy stream creation failed");
        delete ctx;
        return OB200_ECUDA;
    }
    ctx->shape.sms = ctx->prop.multiProcessorCount;
    // keep freed work buffers cached in the pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if ( cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess ) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    stream_register(ctx->stream, true);
    ob200::bind_stream(ctx);
    *out = ctx;
    return OB200_OK;
}

void ob200_context_destroy(ob200_context *ctx)
{
    if ( !ctx ) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->flush.release();
    ctx->partials.release();
    if ( ctx->sched_cache && ctx->sched_cache_free ) {
        ob200::bind_stream(ctx);
        ctx->sched_cache_free(ctx->sched_cache);
        ctx->sched_cache = nullptr;
    }
    if ( ctx->kerr_dev ) cudaFree(ctx->kerr_dev);
    if ( ctx->kerr_host ) cudaFreeHost((void *) ctx->kerr_host);
    cudaStreamSynchronize(ctx->stream);
    stream_register(ctx->stream, false);
    if ( current_stream().stream == ctx->stream ) current_stream().stream = nullptr;
    cudaStreamSynchronize(ctx->copy_stream);
    cudaStreamDestroy(ctx->copy_stream);
    cudaEventDestroy(ctx->copy_event);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int ob200_context_sync(ob200_context *ctx)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx, OB200_EINVAL, "context_sync: null context");
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    if ( ctx->kerr_host && *ctx->kerr_host ) {
        const int code = *ctx->kerr_host;
        *ctx->kerr_host = 0;
        cudaMemsetAsync(ctx->kerr_dev, 0, sizeof( int ), ctx->stream);
        OB_REQUIRE(false, OB200_ECUDA, "a wait inside an assembly kernel timed out (code %d); the matrix values are invalid", code);
    }
    return OB200_OK;
}

void *ob200_context_stream(ob200_context *ctx) { return ctx ? (void *) ctx->stream : nullptr; }
int64_t ob200_context_launch_count(ob200_context *ctx) { return ctx ? ctx->launches : 0; }

int ob200_context_set_profiling(ob200_context *ctx, int enable)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx, OB200_EINVAL, "set_profiling: null context");
    ctx->profiling = enable != 0;
    return OB200_OK;
}

static int profile_fold(ob200_context *ctx)
{
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    for ( auto &r : ctx->prof_pending ) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        auto &t = ctx->prof_total[r.name];
        t.first += ms;
        t.second += 1;
        ctx->prof_pool.push_back(r.start);
        ctx->prof_pool.push_back(r.stop);
    }
    ctx->prof_pending.clear();
    return OB200_OK;
}

int ob200_context_profile_reset(ob200_context *ctx)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx, OB200_EINVAL, "profile_reset: null context");
    OB_CHECK( profile_fold(ctx) );
    ctx->prof_total.clear();
    return OB200_OK;
}

int ob200_context_profile_report(ob200_context *ctx, char *buf, int64_t buflen)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && buf && buflen > 0, OB200_EINVAL, "profile_report: bad argument");
    OB_CHECK( profile_fold(ctx) );
    std::string out;
    for ( auto &kv : ctx->prof_total ) {
        char line[512];
        snprintf(line, sizeof( line ), "%s\t%.6f\t%lld\n", kv.first.c_str(), kv.second.first, (long long) kv.second.second);
        out += line;
    }
    OB_REQUIRE((int64_t) out.size() < buflen, OB200_EINVAL, "profile_report: buffer too small (%zu needed)", out.size() + 1);
    memcpy(buf, out.c_str(), out.size() + 1);
    return OB200_OK;
}

int ob200_malloc(ob200_context *ctx, int64_t bytes, void **dptr)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && dptr && bytes >= 0, OB200_EINVAL, "malloc: bad argument");
    OB_CUDA( cudaSetDevice(ctx->device) );
    *dptr = nullptr;
    if ( bytes ) OB_CUDA( cudaMalloc(dptr, (size_t) bytes) );
    return OB200_OK;
}

int ob200_free(ob200_context *ctx, void *dptr)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx, OB200_EINVAL, "free: null context");
    if ( dptr ) {
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        OB_CUDA( cudaFree(dptr) );
    }
    return OB200_OK;
}

int ob200_memcpy_h2d(ob200_context *ctx, void *dst, const void *src, int64_t bytes)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && ( bytes == 0 || ( dst && src ) ), OB200_EINVAL, "memcpy_h2d: bad argument");
    if ( bytes ) OB_CUDA( cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyHostToDevice, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

int ob200_memcpy_d2h(ob200_context *ctx, void *dst, const void *src, int64_t bytes)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && ( bytes == 0 || ( dst && src ) ), OB200_EINVAL, "memcpy_d2h: bad argument");
    if ( bytes ) OB_CUDA( cudaMemcpyAsync(dst, src, (size_t) bytes, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

int ob200_memset(ob200_context *ctx, void *dst, int value, int64_t bytes)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && ( bytes == 0 || dst ), OB200_EINVAL, "memset: bad argument");
    if ( bytes ) OB_CUDA( cudaMemsetAsync(dst, value, (size_t) bytes, ctx->stream) );
    return OB200_OK;
}

int ob200_flush_l2(ob200_context *ctx)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx, OB200_EINVAL, "flush_l2: null context");
    const int64_t bytes = (int64_t) 256 << 20;        // 256 MiB > 126 MB L2
    if ( !ctx->flush.p ) OB_CHECK( ctx->flush.alloc(bytes) );
    static char v = 0;
    v++;
    flush_kernel<<< ctx->shape.sms * 8, 256, 0, ctx->stream >>>(ctx->flush.p, bytes, v);
    OB_CUDA( cudaGetLastError() );
    return OB200_OK;
}

} // extern "C"
