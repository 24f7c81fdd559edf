// CSR SpMV for sm_100a: the matrix streams through shared memory in fixed-size chunks of
// non-zeros with 1-D TMA bulk copies (cp.async.bulk + mbarrier, 3 stages in flight per CTA),
// every thread multiplies its share against gathered x, and each row of the chunk is summed
// from shared memory by 8 lanes in a fixed order (deterministic run to run).
//
// reference semantics: CompCol::times (src/core/compcol.C:119-134), y = A x.
//
// A "chunk" c is the set of rows whose first entry lies in [c*kChunk, (c+1)*kChunk); its
// entries are contiguous, at most kChunk + (longest row) - 1 of them.  The per-chunk table
// {first row, first entry} is built once per structure (csr_build_chunks).
#pragma once
#include "common.cuh"

#include "spmv_halo.h"

namespace ob200 {

constexpr int kSpmvConsumers = 256;                  // 8 consumer warps ...
constexpr int kSpmvThreads = kSpmvConsumers + 32;    // ... + 1 producer warp (one elected lane issues the TMA copies)
constexpr int kSpmvChunk = 2048;                     // non-zeros per chunk
constexpr int kSpmvSlack = 512;                      // longest row the streamed kernel accepts
constexpr int kSpmvStages = 3;
constexpr int kSpmvCap = kSpmvChunk + kSpmvSlack + 8; // entries per stage (alignment padding included)
constexpr int kSpmvRowCap = 1024;                    // row pointers staged per chunk (more rows: read from global)
constexpr int kSpmvLanesPerRow = 8;

struct SpmvStage {
    double val[kSpmvCap];
    int32_t col[kSpmvCap];
    int32_t rp[kSpmvRowCap + 8];
};
struct SpmvMeta {
    int row0, nz0, row1, nz1;
};
struct SpmvShared {
    SpmvStage st[kSpmvStages];
    unsigned long long full[kSpmvStages], empty[kSpmvStages];
    SpmvMeta meta[kSpmvStages];
    double scratch[32];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32(bar) ), "r"( count ) );
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32(bar) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"( smem_u32(bar) ), "r"( parity ) : "memory" );
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier; streamed data
// is marked evict-first in L2 so that it does not displace the gathered vector
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar, uint64_t policy)
{
    asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                  ::"r"( smem_u32(dst) ), "l"( src ), "r"( bytes ), "r"( smem_u32(bar) ), "l"( policy ) : "memory" );
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile( "createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"( p ) );
    return p;
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" );
}

// chunk table: chunk_row[c] = first row r with rowptr[r] >= c*kChunk, chunk_nz[c] = rowptr[chunk_row[c]]
__global__ void spmv_chunk_table_kernel(int32_t neq, const int32_t *__restrict__ rowptr, int32_t nchunks,
                                        int2 *__restrict__ table)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r <= neq; r += stride ) {
        // chunks whose first row is r: those c with rowptr[r-1] < c*kChunk <= rowptr[r]
        int hi = rowptr[r] / kSpmvChunk;
        const int lo = r == 0 ? 0 : rowptr[r - 1] / kSpmvChunk + 1;
        if ( hi > nchunks - 1 ) hi = nchunks - 1;
        for ( int c = lo; c <= hi; c++ ) table[c] = make_int2((int) r, rowptr[r]);
        // sentinel (and chunks past the last entry): trailing empty rows stay in the last chunk
        if ( r == neq )
            for ( int c = ( hi + 1 > lo ? hi + 1 : lo ); c <= nchunks; c++ ) table[c] = make_int2(neq, rowptr[neq]);
    }
}

// y = A x.  MODE 1: also partials[blockIdx.x] = sum over this CTA's rows of y[r]*x[r].
// MODE 2 (distributed): the partial sums cover the rows only this rank holds, the local sums of shared
// rows are pushed to the sharers' mailboxes (the halo exchange is fused into the epilogue).
// `done` (may be null): device flag, the kernel is a no-op once it is set (CG early exit).
//
// Warp-specialised: warp 8 (one lane) is the producer -- it walks this CTA's chunks, waits for a
// free stage, and issues three bulk copies (values, column indices, row pointers) that complete on
// the stage's "full" mbarrier; warps 0-7 consume: products in place, then row sums, then release
// the stage through its "empty" mbarrier.
template< int MODE >
__global__ void __launch_bounds__(kSpmvThreads, 2)
spmv_stream_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                   const double *__restrict__ val, const int2 *__restrict__ table, int32_t nchunks,
                   const double *__restrict__ x, double *__restrict__ y, double *__restrict__ partials,
                   const int *__restrict__ done, SpmvHalo halo)
{
    constexpr bool FUSE_DOT = MODE != 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvShared &sh = *reinterpret_cast< SpmvShared * >( smem_raw );
    if ( done && *done ) return;
    const int tid = threadIdx.x;
    if ( tid == 0 ) {
#pragma unroll
        for ( int s = 0; s < kSpmvStages; s++ ) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kSpmvConsumers / 32);
        }
        asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    __syncthreads();
    const int nmine = ( nchunks > (int) blockIdx.x ) ? ( nchunks - 1 - (int) blockIdx.x ) / (int) gridDim.x + 1 : 0;

    if ( tid >= kSpmvConsumers ) {
        // ---- producer ----
        if ( tid != kSpmvConsumers ) return;
        const uint64_t policy = l2_policy_evict_first();
        int2 t0 = make_int2(0, 0), t1 = t0;
        if ( nmine > 0 ) {
            t0 = table[blockIdx.x];
            t1 = table[blockIdx.x + 1];
        }
        for ( int k = 0; k < nmine; k++ ) {
            const int s = k % kSpmvStages;
            // table entries of the next chunk: in flight while this one is issued
            int2 n0 = t0, n1 = t1;
            if ( k + 1 < nmine ) {
                const int c = blockIdx.x + ( k + 1 ) * gridDim.x;
                n0 = table[c];
                n1 = table[c + 1];
            }
            if ( k >= kSpmvStages ) mbar_wait(&sh.empty[s], ( ( k / kSpmvStages ) - 1 ) & 1);
            SpmvStage &st = sh.st[s];
            sh.meta[s] = SpmvMeta{ t0.x, t0.y, t1.x, t1.y };
            const int a0 = t0.y & ~3;
            const int n = ( ( t1.y - a0 ) + 3 ) & ~3;
            const int ra0 = t0.x & ~3;
            const int rn = ( ( t1.x + 1 - ra0 ) + 3 ) & ~3;
            if ( t1.x > t0.x && n > 0 ) {
                const bool rows = rn <= kSpmvRowCap + 8;
                mbar_expect_tx(&sh.full[s], (uint32_t) n * 12u + ( rows ? (uint32_t) rn * 4u : 0u ));
                tma_load_1d(st.val, val + a0, (uint32_t) n * 8u, &sh.full[s], policy);
                tma_load_1d(st.col, colind + a0, (uint32_t) n * 4u, &sh.full[s], policy);
                if ( rows ) tma_load_1d(st.rp, rowptr + ra0, (uint32_t) rn * 4u, &sh.full[s], policy);
            } else {
                mbar_expect_tx(&sh.full[s], 0);      // chunk without rows: just complete the phase
            }
            t0 = n0;
            t1 = n1;
        }
        return;
    }

    // ---- consumers ----
    // Each group of kSpmvLanesPerRow lanes owns one row at a time and runs val * x[col] straight out of
    // the staged arrays (no second pass through shared memory, no CTA-wide barrier: a warp only
    // synchronises with the producer through the stage's mbarriers).  The four groups of a warp take
    // rows 8 apart so that their 64-byte windows do not pile onto the same banks.
    double pq = 0.0;
    const int wid = tid >> 5, gw = ( tid >> 3 ) & 3, gl = tid & 7;
    constexpr int kWarps = kSpmvConsumers / 32;
    constexpr int kRowsPerPass = kWarps * 4;           // 32 rows per pass of the CTA
    for ( int k = 0; k < nmine; k++ ) {
        const int s = k % kSpmvStages;
        const SpmvStage &st = sh.st[s];
        mbar_wait(&sh.full[s], ( k / kSpmvStages ) & 1);
        const SpmvMeta m = sh.meta[s];
        const int a0 = m.nz0 & ~3;
        const int ra0 = m.row0 & ~3;
        const bool rows_staged = ( ( ( m.row1 + 1 - ra0 ) + 3 ) & ~3 ) <= kSpmvRowCap + 8;
        for ( int rb = m.row0; rb < m.row1; rb += kRowsPerPass ) {
            const int r = rb + gw * kWarps + wid;
            double s0 = 0.0, s1 = 0.0;
            // what the epilogue needs, requested before the row is summed
            double xr = 0.0;
            bool shared_row = false;
            if ( MODE != 0 && r < m.row1 ) xr = __ldg(x + r);
            if ( r < m.row1 ) {
                int b, e;
                if ( rows_staged ) {
                    b = st.rp[r - ra0];
                    e = st.rp[r + 1 - ra0];
                } else {
                    b = rowptr[r];
                    e = rowptr[r + 1];
                }
                if ( MODE == 2 ) {
                    // the distributed product runs on a copy of the row pointers whose sign bit marks the shared rows
                    shared_row = b < 0;
                    b &= 0x7fffffff;
                    e &= 0x7fffffff;
                }
                b -= a0;
                e -= a0;
                int i = b + gl;
                for ( ; i + 3 * kSpmvLanesPerRow < e; i += 4 * kSpmvLanesPerRow ) {
                    const int c0 = st.col[i], c1 = st.col[i + 8], c2 = st.col[i + 16], c3 = st.col[i + 24];
                    const double v0 = st.val[i], v1 = st.val[i + 8], v2 = st.val[i + 16], v3 = st.val[i + 24];
                    const double x0 = __ldg(x + c0), x1 = __ldg(x + c1), x2 = __ldg(x + c2), x3 = __ldg(x + c3);
                    s0 += v0 * x0;
                    s1 += v1 * x1;
                    s0 += v2 * x2;
                    s1 += v3 * x3;
                }
                for ( ; i < e; i += kSpmvLanesPerRow ) s0 += st.val[i] * __ldg(x + st.col[i]);
                s0 += s1;
            }
#pragma unroll
            for ( int o = kSpmvLanesPerRow / 2; o > 0; o >>= 1 ) s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            if ( r < m.row1 && gl == 0 ) {
                y[r] = s0;
                if ( MODE == 1 ) pq += s0 * xr;
                if ( MODE == 2 ) {
                    if ( !shared_row ) pq += s0 * xr;
                    else if ( halo.dst ) {
                        const int u = __ldg(halo.route + r);
                        for ( int i = __ldg(halo.uptr + u), e2 = __ldg(halo.uptr + u + 1); i < e2; i++ )
                            ll_store(halo.dst[i] + halo.half_words, s0, halo.seq);
                    }
                }
            }
        }
        __syncwarp();
        if ( ( tid & 31 ) == 0 ) asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( smem_u32(&sh.empty[s]) ) : "memory" );
    }
    if ( FUSE_DOT ) {
        // block sum over the consumer warps in a fixed order
        const int lane = tid & 31;
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) pq += __shfl_xor_sync(0xffffffffu, pq, o);
        if ( lane == 0 ) sh.scratch[wid] = pq;
        asm volatile( "bar.sync 1, %0;" ::"n"( kSpmvConsumers ) : "memory" );
        if ( wid == 0 ) {
            double t = lane < kSpmvConsumers / 32 ? sh.scratch[lane] : 0.0;
#pragma unroll
            for ( int o = 16; o > 0; o >>= 1 ) t += __shfl_xor_sync(0xffffffffu, t, o);
            if ( lane == 0 ) partials[blockIdx.x] = t;
        }
    }
}

} // namespace ob200
