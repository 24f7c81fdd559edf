// Small device-wide primitives for the symbolic phase: exclusive scan (int32/int64 -> int64),
// max reduction, narrowing copy.  Hand-written; no CUB/Thrust.
#pragma once
#include "common.cuh"

namespace ob200 {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;                       // per thread
constexpr int kScanTile = kScanThreads * kScanItems;

template< class TIn >
__global__ void scan_tile_kernel(const TIn *__restrict__ in, int64_t *__restrict__ out, int64_t n, int64_t *__restrict__ tile_sums)
{
    __shared__ int64_t warp_sums[kScanThreads / 32];
    const int64_t base = (int64_t) blockIdx.x * kScanTile + (int64_t) threadIdx.x * kScanItems;
    int64_t v[kScanItems], s = 0;
#pragma unroll
    for ( int i = 0; i < kScanItems; i++ ) {
        v[i] = ( base + i < n ) ? (int64_t) in[base + i] : 0;
        s += v[i];
    }
    // inclusive scan of per-thread sums across the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t incl = s;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 ) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ( lane >= o ) incl += t;
    }
    if ( lane == 31 ) warp_sums[wid] = incl;
    __syncthreads();
    if ( wid == 0 ) {
        int64_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for ( int o = 1; o < 32; o <<= 1 ) {
            int64_t t = __shfl_up_sync(0xffffffffu, w, o);
            if ( lane >= o ) w += t;
        }
        if ( lane < kScanThreads / 32 ) warp_sums[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    int64_t excl = incl - s + ( wid > 0 ? warp_sums[wid - 1] : 0 );
#pragma unroll
    for ( int i = 0; i < kScanItems; i++ ) {
        if ( base + i < n ) out[base + i] = excl;
        excl += v[i];
    }
    if ( threadIdx.x == kScanThreads - 1 ) tile_sums[blockIdx.x] = warp_sums[kScanThreads / 32 - 1];
}

static __global__ void scan_add_kernel(int64_t *__restrict__ out, int64_t n, const int64_t *__restrict__ tile_offsets)
{
    const int64_t off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t) blockIdx.x * kScanTile;
    for ( int i = threadIdx.x; i < kScanTile; i += kScanThreads )
        if ( base + i < n ) out[base + i] += off;
}

template< class TIn >
inline int exclusive_scan_rec(ob200_context *ctx, const TIn *in, int64_t *out, int64_t n, int64_t *total_dev)
{
    // total_dev: device int64 receiving the grand total
    const int64_t tiles = ceil_div(n, kScanTile);
    DevBuf< int64_t > sums, offs;
    OB_CHECK( sums.alloc(tiles) );
    OB_LAUNCH(ctx, scan_tile_kernel< TIn >, (int) tiles, kScanThreads, 0, in, out, n, sums.p);
    if ( tiles == 1 ) {
        OB_CUDA( cudaMemcpyAsync(total_dev, sums.p, sizeof( int64_t ), cudaMemcpyDeviceToDevice, ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );   // sums is freed on return
        return OB200_OK;
    }
    OB_CHECK( offs.alloc(tiles) );
    OB_CHECK( exclusive_scan_rec< int64_t >(ctx, sums.p, offs.p, tiles, total_dev) );
    OB_LAUNCH(ctx, scan_add_kernel, (int) tiles, kScanThreads, 0, out, n, offs.p);
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

// out[i] = sum_{k<i} in[k]; *total = sum of all (host value)
inline int exclusive_scan(ob200_context *ctx, const int32_t *in, int64_t *out, int64_t n, int64_t *total)
{
    *total = 0;
    if ( n <= 0 ) return OB200_OK;
    DevBuf< int64_t > tot;
    OB_CHECK( tot.alloc(1) );
    OB_CHECK( exclusive_scan_rec< int32_t >(ctx, in, out, n, tot.p) );
    OB_CUDA( cudaMemcpyAsync(total, tot.p, sizeof( int64_t ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

static __global__ void max_reduce_kernel(const int32_t *__restrict__ in, int64_t n, int32_t *__restrict__ out)
{
    int32_t m = 0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) m = max(m, in[t]);
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ( ( threadIdx.x & 31 ) == 0 ) atomicMax(out, m);
}

inline int max_reduce(ob200_context *ctx, const int32_t *in, int64_t n, int32_t *result)
{
    *result = 0;
    if ( n <= 0 ) return OB200_OK;
    DevBuf< int32_t > r;
    OB_CHECK( r.alloc(1) );
    OB_CUDA( cudaMemsetAsync(r.p, 0, sizeof( int32_t ), ctx->stream) );
    OB_LAUNCH(ctx, max_reduce_kernel, ctx->shape.grid(n, 256, 4), 256, 0, in, n, r.p);
    OB_CUDA( cudaMemcpyAsync(result, r.p, sizeof( int32_t ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    return OB200_OK;
}

static __global__ void narrow_kernel(const int64_t *__restrict__ in, int32_t *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) out[t] = (int32_t) in[t];
}

inline int narrow_i64_to_i32(ob200_context *ctx, const int64_t *in, int32_t *out, int64_t n)
{
    if ( n <= 0 ) return OB200_OK;
    OB_LAUNCH(ctx, narrow_kernel, ctx->shape.grid(n, 256, 4), 256, 0, in, out, n);
    return OB200_OK;
}

} // namespace ob200
