// SparseMtrx "cudacsr": symbolic assembly (CompCol::buildInternalStructure), value assembly
// (CompCol::assemble), SpMV (CompCol::times) -- reference: src/core/compcol.C.
//
// The pattern of a finite element matrix is structurally symmetric, so the reference's
// compressed-column arrays (colptr, rowind) and the compressed-row arrays kept here
// (rowptr, colind) are the same integers; values are stored by rows, A(i,j) at
// val[rowptr[i] + k] with colind[rowptr[i] + k] == j.
#include "common.cuh"
#include "elemset.h"
#include "scan.cuh"
#include "spmv.cuh"
#include "spmv_block.cuh"
#include <stdlib.h>
#include <vector>
#include <limits.h>

namespace ob200 {

// ---- symbolic phase ---------------------------------------------------------------------

// number of (element, local dof) incidences per equation
__global__ void count_incidence_kernel(const int32_t *__restrict__ loc, int64_t n, int32_t neq,
                                       int32_t *__restrict__ cnt, int *__restrict__ bad)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        int r = loc[t];
        if ( r < 0 || r > neq ) atomicAdd(bad, 1);
        else if ( r > 0 ) atomicAdd(cnt + r - 1, 1);
    }
}

__global__ void fill_incidence_kernel(const int32_t *__restrict__ loc, int64_t n, int nd,
                                      const int64_t *__restrict__ start, int32_t *__restrict__ fill,
                                      int32_t *__restrict__ elems)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        int r = loc[t];
        if ( r > 0 ) {
            int k = atomicAdd(fill + r - 1, 1);
            elems[start[r - 1] + k] = (int32_t)( t / nd );
        }
    }
}

// One warp per matrix row: gather the equation numbers of all elements touching the row into
// shared memory, bitonic-sort, drop duplicates.  This is std::set<int> columns[jj-1].insert(ii-1)
// of compcol.C:178-191 done row-parallel.  MODE 0 counts, MODE 1 writes colind (the row is sorted again), MODE 2 counts and keeps
// the sorted row in a scratch array [neq][tstride] from which row_copy_kernel fills colind once the row starts are known -- the
// sort, which is the whole cost, then runs once (used when the scratch array is affordable).
enum { PATTERN_COUNT = 0, PATTERN_FILL = 1, PATTERN_COUNT_KEEP = 2 };
template< int MODE >
__global__ void row_pattern_kernel(const int32_t *__restrict__ loc, int nd, int32_t neq,
                                   const int64_t *__restrict__ estart, const int32_t *__restrict__ elems,
                                   int cap, int32_t *__restrict__ rowcount, const int32_t *__restrict__ rowptr,
                                   int32_t *__restrict__ colind, int *__restrict__ overflow, int32_t *__restrict__ scratch, int tstride)
{
    constexpr bool FILL = MODE == PATTERN_FILL;
    extern __shared__ int32_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int32_t *buf = smem + (size_t) wid * cap;
    for ( int64_t row = (int64_t) blockIdx.x * nw + wid; row < neq; row += (int64_t) gridDim.x * nw ) {
        const int64_t e0 = estart[row], e1 = estart[row + 1];
        const int64_t ncand = ( e1 - e0 ) * nd;
        if ( ncand > cap ) {
            if ( lane == 0 ) atomicAdd(overflow, 1);
            continue;
        }
        int n2 = 32;
        while ( n2 < ncand ) n2 <<= 1;
        for ( int t = lane; t < n2; t += 32 ) {
            int32_t v = INT_MAX;
            if ( t < ncand ) {
                int32_t e = elems[e0 + t / nd];
                int32_t c = loc[(int64_t) e * nd + t % nd];
                if ( c > 0 ) v = c - 1;
            }
            buf[t] = v;
        }
        __syncwarp();
        for ( int k = 2; k <= n2; k <<= 1 )
            for ( int j = k >> 1; j > 0; j >>= 1 ) {
                for ( int t = lane; t < n2; t += 32 ) {
                    int p = t ^ j;
                    if ( p > t ) {
                        int32_t a = buf[t], b = buf[p];
                        bool up = ( ( t & k ) == 0 );
                        if ( ( a > b ) == up ) { buf[t] = b; buf[p] = a; }
                    }
                }
                __syncwarp();
            }
        // unique count / compaction (buf sorted ascending, INT_MAX padding last)
        int base = 0;
        const int32_t out0 = FILL ? rowptr[row] : 0;
        for ( int t0 = 0; t0 < n2; t0 += 32 ) {
            int t = t0 + lane;
            int32_t v = buf[t];
            bool keep = ( v != INT_MAX ) && ( t == 0 || buf[t - 1] != v );
            unsigned m = __ballot_sync(0xffffffffu, keep);
            if ( FILL && keep ) colind[out0 + base + __popc(m & ( ( 1u << lane ) - 1 ))] = v;
            if ( MODE == PATTERN_COUNT_KEEP && keep ) scratch[row * tstride + base + __popc(m & ( ( 1u << lane ) - 1 ))] = v;
            base += __popc(m);
        }
        if ( !FILL && lane == 0 ) rowcount[row] = base;
        __syncwarp();
    }
}

// colind[rowptr[row] + k] = scratch[row][k]: one warp per row
__global__ void row_copy_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ scratch, int tstride,
                                int32_t *__restrict__ colind)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    for ( int64_t row = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5; row < neq; row += nwarps ) {
        const int p0 = rowptr[row], n = rowptr[row + 1] - p0;
        for ( int k = lane; k < n; k += 32 ) colind[p0 + k] = scratch[row * tstride + k];
    }
}

// ---- numeric phase ----------------------------------------------------------------------

__device__ __forceinline__ int find_slot(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int r, int c)
{
    int lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while ( lo <= hi ) {
        int mid = ( lo + hi ) >> 1;
        int v = colind[mid];
        if ( v == c ) return mid;
        if ( v < c ) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

// CompCol::assemble(loc, mat) (compcol.C:263-299) for a batch of element matrices
__global__ void csr_assemble_kernel(const int32_t *__restrict__ loc, const double *__restrict__ mat, int nd,
                                    int64_t nelem, const int32_t *__restrict__ rowptr,
                                    const int32_t *__restrict__ colind, double *__restrict__ val, int *__restrict__ missing)
{
    const int64_t total = nelem * nd * nd;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride ) {
        int64_t e = t / ( nd * nd );
        int ij = (int)( t - e * nd * nd );
        int i = ij / nd, j = ij - i * nd;
        int r = loc[e * nd + i], c = loc[e * nd + j];
        if ( r > 0 && c > 0 ) {
            int p = find_slot(rowptr, colind, r - 1, c - 1);
            if ( p < 0 ) atomicAdd(missing, 1);
            else atomicAdd(val + p, mat[t]);
        }
    }
}

// CompCol::assemble(rloc, cloc, mat) (compcol.C:301-336): only the rloc x cloc entries are touched
__global__ void csr_assemble_rect_kernel(const int32_t *__restrict__ rloc, const int32_t *__restrict__ cloc, const double *__restrict__ mat,
                                         int nr, int nc, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                         double *__restrict__ val, int *__restrict__ missing)
{
    for ( int t = blockIdx.x * blockDim.x + threadIdx.x; t < nr * nc; t += gridDim.x * blockDim.x ) {
        const int i = t / nc, j = t - i * nc;
        const int r = rloc[i], c = cloc[j];
        if ( r > 0 && c > 0 ) {
            const int p = find_slot(rowptr, colind, r - 1, c - 1);
            if ( p < 0 ) atomicAdd(missing, 1);
            else atomicAdd(val + p, mat[t]);
        }
    }
}

__global__ void scale_kernel(double *__restrict__ v, int64_t n, double s)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) v[t] *= s;
}

// y = A x, one warp per row straight from global memory (fallback for very long rows)
template< bool FUSE_DOT >
__global__ void __launch_bounds__(256)
spmv_rowwarp_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                    const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
                    double *__restrict__ partials, const int *__restrict__ done)
{
    __shared__ double scratch[8];
    if ( done && *done ) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    double pq = 0.0;
    for ( int64_t row = warp0; row < neq; row += nwarps ) {
        const int b = rowptr[row], e = rowptr[row + 1];
        double s = 0.0;
        for ( int t = b + lane; t < e; t += 32 ) s += val[t] * x[colind[t]];
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ( lane == 0 ) {
            y[row] = s;
            if ( FUSE_DOT ) pq += s * x[row];
        }
    }
    if ( FUSE_DOT ) {
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) pq += __shfl_xor_sync(0xffffffffu, pq, o);
        if ( lane == 0 ) scratch[threadIdx.x >> 5] = pq;
        __syncthreads();
        if ( threadIdx.x == 0 ) {
            double t = 0.0;
            for ( int w = 0; w < 8; w++ ) t += scratch[w];
            partials[blockIdx.x] = t;
        }
    }
}

} // namespace ob200

using namespace ob200;

namespace ob200 {
// SparseMtrx::timesT (compcol.C:146-163): y = A^T x.  Eight lanes per row scatter val * x[row] into y[col].
__global__ void csr_times_t_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                   const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x, stride = ( (int64_t) gridDim.x * blockDim.x ) >> 3;
    const int sub = (int)( t & 7 );
    for ( int64_t r = t >> 3; r < neq; r += stride ) {
        const double xr = x[r];
        for ( int k = rowptr[r] + sub; k < rowptr[r + 1]; k += 8 ) atomicAdd(y + colind[k], val[k] * xr);
    }
}
}
void ob200_csr_touch(ob200_csr *A) { A->version++; }

int ob200_csr_materialize(ob200_csr *A)
{
    if ( A ) ob200::bind_stream(A->ctx);
    if ( A->zero_pending ) {
        if ( A->nnz ) OB_CUDA( cudaMemsetAsync(A->val.p, 0, sizeof( double ) * (size_t) A->nnz, A->ctx->stream) );
        A->zero_pending = false;
    }
    return OB200_OK;
}

namespace ob200 {

// Derive the blocked index (spmv_block.cuh) from rowptr/colind; leaves blk_ok = false when the matrix has no
// node-block structure worth using or a chunk would not fit the stage.
static int csr_build_blocks(ob200_csr *A)
{
    ob200_context *ctx = A->ctx;
    A->blk_tried = true;
    A->blk_ok = false;
    if ( getenv("OB200_SPMV_BLOCKED") && getenv("OB200_SPMV_BLOCKED")[0] == '0' ) return OB200_OK;
    const int32_t neq = A->neq;
    if ( neq == 0 || A->nnz == 0 || A->maxrow > kSpmvSlack ) return OB200_OK;
    // rows that repeat the pattern of the row before them -> row blocks of up to 3 rows (host pass over neq flags)
    DevBuf< unsigned char > same;
    OB_CHECK( same.alloc(neq) );
    OB_LAUNCH(ctx, blk_row_same_kernel, ctx->shape.grid((int64_t) neq * 32, 256, 8), 256, 0, neq, A->rowptr.p, A->colind.p, same.p);
    std::vector< unsigned char > hs(neq);
    OB_CUDA( cudaMemcpyAsync(hs.data(), same.p, (size_t) neq, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    std::vector< int32_t > rbrow;
    rbrow.reserve(neq / 2 + 2);
    int run = 0;
    for ( int32_t r = 0; r < neq; r++ ) {
        if ( r > 0 && hs[r] && run < 3 ) run++;
        else { rbrow.push_back(r); run = 1; }
    }
    const int32_t nrb = (int32_t) rbrow.size();
    rbrow.push_back(neq);
    DevBuf< int32_t > d_rbrow, cnt;
    DevBuf< int64_t > b0;
    OB_CHECK( d_rbrow.alloc(nrb + 1) );
    OB_CHECK( cnt.alloc(nrb + 1) );
    OB_CHECK( b0.alloc(nrb + 1) );
    OB_CUDA( cudaMemcpyAsync(d_rbrow.p, rbrow.data(), sizeof( int32_t ) * (size_t)( nrb + 1 ), cudaMemcpyHostToDevice, ctx->stream) );
    OB_CUDA( cudaMemsetAsync(cnt.p, 0, sizeof( int32_t ) * (size_t)( nrb + 1 ), ctx->stream) );
    const int grid = ctx->shape.grid((int64_t) nrb * 32, 256, 8);
    OB_LAUNCH(ctx, blk_columns_kernel< false >, grid, 256, 0, nrb, d_rbrow.p, A->rowptr.p, A->colind.p, cnt.p, nullptr, nullptr, nullptr);
    int64_t nblk = 0;
    OB_CHECK( exclusive_scan(ctx, cnt.p, b0.p, nrb + 1, &nblk) );
    // worth it?  index bytes per non-zero must drop well below the 4 B of colind
    if ( nblk * kBlkMinRatio > A->nnz || nblk >= ( (int64_t) 1 << 31 ) ) return OB200_OK;
    OB_CHECK( A->bw.alloc(nblk + 8) );
    OB_CHECK( A->rbdesc.alloc(nrb + 1) );
    OB_CUDA( cudaMemsetAsync(A->bw.p + nblk, 0, sizeof( uint2 ) * 8, ctx->stream) );
    OB_LAUNCH(ctx, blk_columns_kernel< true >, grid, 256, 0, nrb, d_rbrow.p, A->rowptr.p, A->colind.p, cnt.p, b0.p, A->bw.p, A->rbdesc.p);
    const int4 sentinel = make_int4(neq, (int) A->nnz, (int) nblk, 0);
    OB_CUDA( cudaMemcpyAsync(A->rbdesc.p + nrb, &sentinel, sizeof( int4 ), cudaMemcpyHostToDevice, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    const int32_t nchunks = (int32_t)( ( A->nnz - 1 ) / kBlkChunk + 1 );
    OB_CHECK( A->bchunks.alloc(nchunks + 1) );
    OB_LAUNCH(ctx, blk_chunk_table_kernel, ctx->shape.grid((int64_t) nrb + 1, 256, 8), 256, 0, nrb, A->rbdesc.p, nchunks, A->bchunks.p);
    std::vector< int4 > ht(nchunks + 1);
    OB_CUDA( cudaMemcpyAsync(ht.data(), A->bchunks.p, sizeof( int4 ) * (size_t)( nchunks + 1 ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    for ( int32_t c = 0; c < nchunks; c++ ) {
        const int4 t0 = ht[c], t1 = ht[c + 1];
        if ( t1.y - ( t0.y & ~1 ) + 1 > kBlkValCap - 8 || t1.z - ( t0.z & ~1 ) + 1 > kBlkCap || t1.x - t0.x > kRbCap ) return OB200_OK;
    }
    A->nrb = nrb;
    A->nblk = nblk;
    A->nbchunks = nchunks;
    A->blk_ok = true;
    return OB200_OK;
}

template< int MODE >
static int spmv_block_launch(ob200_csr *A, const int4 *desc, const double *x, double *y, double *partials, int *nblocks,
                             const int *done, const SpmvHalo &hv)
{
    ob200_context *ctx = A->ctx;
    static unsigned long long attr_set = 0;
    const int smem = (int) sizeof( SpmvBlkShared );
    if ( attr_needed(attr_set, ctx->device) ) {
        OB_CUDA( cudaFuncSetAttribute(spmv_block_kernel< MODE >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
    }
    int grid = ctx->shape.sms * kBlkCtas;
    if ( grid > A->nbchunks ) grid = A->nbchunks;
    OB_LAUNCH(ctx, spmv_block_kernel< MODE >, grid, kSpmvThreads, smem, A->val.p, A->bw.p, desc, A->bchunks.p, A->nbchunks, x, y,
              partials, done, hv);
    if ( nblocks && MODE != 0 ) *nblocks = grid;
    return OB200_OK;
}

static int spmv_launch(ob200_csr *A, const double *x, double *y, double *partials, int *nblocks, const int *done,
                       const SpmvHalo *halo = nullptr)
{
    ob200_context *ctx = A->ctx;
    if ( nblocks ) *nblocks = 0;
    if ( A->neq == 0 ) return OB200_OK;
    OB_CHECK( ob200_csr_materialize(A) );
    if ( A->nnz == 0 ) {
        OB_CUDA( cudaMemsetAsync(y, 0, sizeof( double ) * (size_t) A->neq, ctx->stream) );
        return OB200_OK;
    }
    if ( !A->blk_tried ) OB_CHECK( csr_build_blocks(A) );
    if ( A->blk_ok ) {
        const SpmvHalo none{ nullptr, nullptr, nullptr, 0, 0 };
        if ( halo ) return spmv_block_launch< 2 >(A, A->rbdesc_flag.p, x, y, partials, nblocks, done, *halo);
        if ( partials ) return spmv_block_launch< 1 >(A, A->rbdesc.p, x, y, partials, nblocks, done, none);
        return spmv_block_launch< 0 >(A, A->rbdesc.p, x, y, partials, nblocks, done, none);
    }
    if ( A->maxrow <= kSpmvSlack && A->chunks.p ) {
        static unsigned long long attr_set = 0;
        const int smem = (int) sizeof( SpmvShared );
        if ( attr_needed(attr_set, ctx->device) ) {
            OB_CUDA( cudaFuncSetAttribute(spmv_stream_kernel< 0 >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
            OB_CUDA( cudaFuncSetAttribute(spmv_stream_kernel< 1 >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
            OB_CUDA( cudaFuncSetAttribute(spmv_stream_kernel< 2 >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
        }
        int grid = ctx->shape.sms * 2;                 // persistent: 2 CTAs per SM, 3 stages each in flight
        if ( grid > A->nchunks ) grid = A->nchunks;
        const SpmvHalo none{ nullptr, nullptr, nullptr, 0, 0 };
        if ( halo ) {
            OB_LAUNCH(ctx, spmv_stream_kernel< 2 >, grid, kSpmvThreads, smem, A->neq, A->rowptr_flag.p, A->colind.p, A->val.p,
                      A->chunks.p, A->nchunks, x, y, partials, done, *halo);
            if ( nblocks ) *nblocks = grid;
        } else if ( partials ) {
            OB_LAUNCH(ctx, spmv_stream_kernel< 1 >, grid, kSpmvThreads, smem, A->neq, A->rowptr.p, A->colind.p, A->val.p,
                      A->chunks.p, A->nchunks, x, y, partials, done, none);
            if ( nblocks ) *nblocks = grid;
        } else {
            OB_LAUNCH(ctx, spmv_stream_kernel< 0 >, grid, kSpmvThreads, smem, A->neq, A->rowptr.p, A->colind.p, A->val.p,
                      A->chunks.p, A->nchunks, x, y, partials, done, none);
        }
        return OB200_OK;
    }
    OB_REQUIRE(!halo, OB200_ECAPACITY, "spmv: rows longer than %d entries are not supported by the fused halo product", kSpmvSlack);
    // rows longer than the streamed kernel's stage: warp per row straight from global memory
    int grid = ctx->shape.grid((int64_t) A->neq * 32, 256, 8);
    if ( partials ) {
        OB_LAUNCH(ctx, spmv_rowwarp_kernel< true >, grid, 256, 0, A->neq, A->rowptr.p, A->colind.p, A->val.p, x, y, partials, done);
        if ( nblocks ) *nblocks = grid;
    } else {
        OB_LAUNCH(ctx, spmv_rowwarp_kernel< false >, grid, 256, 0, A->neq, A->rowptr.p, A->colind.p, A->val.p, x, y, partials, done);
    }
    return OB200_OK;
}

int spmv(ob200_csr *A, const double *x, double *y) { return spmv_launch(A, x, y, nullptr, nullptr, nullptr); }

// q = A p with per-CTA partial sums of p.q in partials[0 .. *nblocks) (partials may be null: plain product)
int spmv_fused_dot(ob200_csr *A, const double *x, double *y, double *partials, int *nblocks, const int *done)
{
    return spmv_launch(A, x, y, partials, nblocks, done);
}

// distributed: q = A p, partial sums of p.q over the rows only this rank holds, shared rows pushed to the sharers
bool spmv_halo_supported(const ob200_csr *A) { return A->maxrow <= kSpmvSlack && A->chunks.p && A->nnz > 0; }
__global__ void flag_rowptr_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ route, int32_t *__restrict__ out)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i <= neq; i += stride )
        out[i] = rowptr[i] | ( i < neq && route[i] >= 0 ? (int32_t) 0x80000000 : 0 );
}

int spmv_fused_halo(ob200_csr *A, const double *x, double *y, double *partials, int *nblocks, const int *done, const SpmvHalo &halo)
{
    ob200_context *ctx = A->ctx;
    if ( !A->blk_tried ) OB_CHECK( csr_build_blocks(A) );
    if ( A->blk_ok ) {
        if ( A->flag_route != halo.route || A->rbdesc_flag.n != (int64_t) A->nrb + 1 ) {
            OB_CHECK( A->rbdesc_flag.alloc((int64_t) A->nrb + 1) );
            OB_LAUNCH(ctx, blk_flag_desc_kernel, ctx->shape.grid((int64_t) A->nrb + 1, 256, 8), 256, 0, A->nrb, A->rbdesc.p, halo.route, A->rbdesc_flag.p);
            A->flag_route = halo.route;
        }
        return spmv_launch(A, x, y, partials, nblocks, done, &halo);
    }
    if ( A->flag_route != halo.route || A->rowptr_flag.n != (int64_t) A->neq + 1 + kCsrPad ) {
        OB_CHECK( A->rowptr_flag.alloc((int64_t) A->neq + 1 + kCsrPad) );
        OB_CUDA( cudaMemsetAsync(A->rowptr_flag.p, 0, sizeof( int32_t ) * (size_t)( A->neq + 1 + kCsrPad ), ctx->stream) );
        OB_LAUNCH(ctx, flag_rowptr_kernel, ctx->shape.grid(A->neq + 1, 256, 8), 256, 0, A->neq, A->rowptr.p, halo.route, A->rowptr_flag.p);
        A->flag_route = halo.route;
    }
    return spmv_launch(A, x, y, partials, nblocks, done, &halo);
}
} // namespace ob200

extern "C" {

int ob200_csr_create(ob200_context *ctx, ob200_csr **out)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && out, OB200_EINVAL, "csr_create: null argument");
    ob200_csr *A = new ob200_csr();
    A->ctx = ctx;
    *out = A;
    return OB200_OK;
}

void ob200_csr_destroy(ob200_csr *A) { delete A; }

int32_t ob200_csr_rows(const ob200_csr *A) { return A ? A->neq : 0; }
int ob200_csr_spmv_layout(ob200_csr *A, int64_t *info)
{
    OB_REQUIRE(A && info, OB200_EINVAL, "csr_spmv_layout: null argument");
    ob200::bind_stream(A->ctx);
    if ( !A->blk_tried && A->rowptr.p ) OB_CHECK( ob200::csr_build_blocks(A) );
    info[0] = A->blk_ok ? 1 : 0;
    info[1] = A->blk_ok ? A->nrb : 0;
    info[2] = A->blk_ok ? A->nblk : 0;
    return OB200_OK;
}
int64_t ob200_csr_nnz(const ob200_csr *A) { return A ? A->nnz : 0; }
int64_t ob200_csr_version(const ob200_csr *A) { return A ? A->version : 0; }

int ob200_csr_build_structure(ob200_csr *A, int32_t neq, int64_t nelem, int32_t nd, const int32_t *loc, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && ( loc || nelem == 0 ), OB200_EINVAL, "csr_build_structure: null argument");
    OB_REQUIRE(neq >= 0 && nelem >= 0 && nd > 0, OB200_EINVAL, "csr_build_structure: bad sizes neq=%d nelem=%lld nd=%d", neq, (long long) nelem, nd);
    ob200_context *ctx = A->ctx;
    Staged< int32_t > L;
    OB_CHECK( L.stage(ctx, loc, nelem * nd, on_device) );
    const int64_t ninc = nelem * nd;
    OB_REQUIRE(ninc < (int64_t) INT_MAX, OB200_ECAPACITY, "csr_build_structure: %lld element dofs exceed the 32-bit index range", (long long) ninc);

    DevBuf< int32_t > cnt, fill, elems, rowcount;
    DevBuf< int64_t > estart, rp64;
    DevBuf< int > flag;
    OB_CHECK( cnt.alloc(neq + 1) );
    OB_CHECK( fill.alloc(neq + 1) );
    OB_CHECK( estart.alloc(neq + 1) );
    OB_CHECK( flag.alloc(2) );
    OB_CUDA( cudaMemsetAsync(cnt.p, 0, sizeof( int32_t ) * ( neq + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(fill.p, 0, sizeof( int32_t ) * ( neq + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(flag.p, 0, sizeof( int ) * 2, ctx->stream) );
    if ( ninc ) {
        int grid = ctx->shape.grid(ninc, 256, 8);
        OB_LAUNCH(ctx, count_incidence_kernel, grid, 256, 0, L.d, ninc, neq, cnt.p, flag.p);
    }
    int hflag[2] = { 0, 0 };
    OB_CUDA( cudaMemcpyAsync(hflag, flag.p, sizeof( int ) * 2, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    OB_REQUIRE(hflag[0] == 0, OB200_EINVAL, "csr_build_structure: %d location entries outside [0, neq=%d]", hflag[0], neq);

    int64_t total_inc = 0;
    OB_CHECK( exclusive_scan(ctx, cnt.p, estart.p, (int64_t) neq + 1, &total_inc) );
    OB_CHECK( elems.alloc(total_inc > 0 ? total_inc : 1) );
    if ( ninc ) {
        int grid = ctx->shape.grid(ninc, 256, 8);
        OB_LAUNCH(ctx, fill_incidence_kernel, grid, 256, 0, L.d, ninc, nd, estart.p, fill.p, elems.p);
    }
    // largest candidate list decides the shared-memory capacity per warp
    int32_t maxcnt = 0;
    OB_CHECK( max_reduce(ctx, cnt.p, neq, &maxcnt) );
    int64_t maxcand = (int64_t) maxcnt * nd;
    int cap = 256;
    while ( cap < maxcand ) cap <<= 1;
    OB_REQUIRE(cap <= 8192, OB200_ECAPACITY,
               "csr_build_structure: an equation couples to %lld candidate entries (> 8192 supported)", (long long) maxcand);
    int warps = cap <= 1024 ? 8 : ( cap <= 4096 ? 4 : 2 );
    size_t smem = (size_t) warps * cap * sizeof( int32_t );
    OB_CUDA( cudaFuncSetAttribute(row_pattern_kernel< PATTERN_COUNT >, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) );
    OB_CUDA( cudaFuncSetAttribute(row_pattern_kernel< PATTERN_FILL >, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) );
    OB_CUDA( cudaFuncSetAttribute(row_pattern_kernel< PATTERN_COUNT_KEEP >, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) );
    // keep the sorted rows of the count pass when the scratch array [neq][maxcand] stays below 6 GB (OB200_PATTERN_KEEP=0: never)
    DevBuf< int32_t > scratch;
    const int tstride = (int) maxcand;
    const char *pk = getenv("OB200_PATTERN_KEEP");
    const bool keep = neq > 0 && maxcand > 0 && (int64_t) neq * maxcand * 4 <= ( 6ll << 30 ) && !( pk && !strcmp(pk, "0") );
    if ( keep ) OB_CHECK( scratch.alloc((int64_t) neq * maxcand) );

    OB_CHECK( rowcount.alloc(neq + 1) );
    OB_CUDA( cudaMemsetAsync(rowcount.p, 0, sizeof( int32_t ) * ( neq + 1 ), ctx->stream) );
    int pgrid = ctx->shape.grid((int64_t) neq * 32, warps * 32, 4);
    if ( neq && keep )
        OB_LAUNCH(ctx, row_pattern_kernel< PATTERN_COUNT_KEEP >, pgrid, warps * 32, smem, L.d, nd, neq, estart.p, elems.p, cap,
                  rowcount.p, (const int32_t *) nullptr, (int32_t *) nullptr, flag.p + 1, scratch.p, tstride);
    else if ( neq )
        OB_LAUNCH(ctx, row_pattern_kernel< PATTERN_COUNT >, pgrid, warps * 32, smem, L.d, nd, neq, estart.p, elems.p, cap,
                  rowcount.p, (const int32_t *) nullptr, (int32_t *) nullptr, flag.p + 1, (int32_t *) nullptr, 0);
    OB_CHECK( rp64.alloc(neq + 1) );
    int64_t nnz = 0;
    OB_CHECK( exclusive_scan(ctx, rowcount.p, rp64.p, (int64_t) neq + 1, &nnz) );
    // CompCol stores colptr / rowind in IntArray (32-bit): keep the same limit and say so
    OB_REQUIRE(nnz < (int64_t) INT_MAX, OB200_ECAPACITY, "csr_build_structure: nnz=%lld exceeds the 32-bit range of the reference's IntArray", (long long) nnz);
    OB_CHECK( A->rowptr.alloc(neq + 1 + kCsrPad) );
    A->flag_route = nullptr;
    A->blk_tried = A->blk_ok = false;
    OB_CHECK( narrow_i64_to_i32(ctx, rp64.p, A->rowptr.p, (int64_t) neq + 1) );
    OB_CHECK( A->colind.alloc(nnz + kCsrPad) );
    OB_CHECK( A->val.alloc(nnz + kCsrPad) );
    OB_CUDA( cudaMemsetAsync(A->colind.p + nnz, 0, sizeof( int32_t ) * kCsrPad, ctx->stream) );
    OB_CHECK( max_reduce(ctx, rowcount.p, neq, &A->maxrow) );
    A->nchunks = nnz > 0 ? (int32_t)( ( nnz - 1 ) / kSpmvChunk + 1 ) : 0;
    OB_CHECK( A->chunks.alloc(A->nchunks + 1) );
    OB_LAUNCH(ctx, spmv_chunk_table_kernel, ctx->shape.grid((int64_t) neq + 1, 256, 8), 256, 0, neq, A->rowptr.p, A->nchunks, A->chunks.p);
    if ( neq && keep )
        OB_LAUNCH(ctx, row_copy_kernel, ctx->shape.grid((int64_t) neq * 32, 256, 8), 256, 0, neq, A->rowptr.p, scratch.p, tstride, A->colind.p);
    else if ( neq )
        OB_LAUNCH(ctx, row_pattern_kernel< PATTERN_FILL >, pgrid, warps * 32, smem, L.d, nd, neq, estart.p, elems.p, cap,
                  (int32_t *) nullptr, A->rowptr.p, A->colind.p, flag.p + 1, (int32_t *) nullptr, 0);
    OB_CUDA( cudaMemsetAsync(A->val.p, 0, sizeof( double ) * (size_t)( nnz + kCsrPad ), ctx->stream) );
    OB_CUDA( cudaMemcpyAsync(hflag, flag.p, sizeof( int ) * 2, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    OB_REQUIRE(hflag[1] == 0, OB200_ECAPACITY, "csr_build_structure: internal row buffer overflow");
    A->neq = neq;
    A->nnz = nnz;
    {
        static int64_t structure_counter = 0;
        A->structure_version = __atomic_add_fetch(&structure_counter, 1, __ATOMIC_RELAXED);
    }
    A->version++;
    A->diag_version = -1;
    A->zero_pending = false;
    return OB200_OK;
}

int ob200_csr_get_structure(const ob200_csr *A, int32_t *rowptr, int32_t *colind, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && rowptr && colind, OB200_EINVAL, "csr_get_structure: null argument");
    OB_REQUIRE(A->rowptr.p, OB200_EINVAL, "csr_get_structure: matrix has no structure");
    cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    OB_CUDA( cudaMemcpyAsync(rowptr, A->rowptr.p, sizeof( int32_t ) * ( A->neq + 1 ), k, A->ctx->stream) );
    if ( A->nnz ) OB_CUDA( cudaMemcpyAsync(colind, A->colind.p, sizeof( int32_t ) * (size_t) A->nnz, k, A->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
    return OB200_OK;
}

int ob200_csr_get_values(const ob200_csr *A, double *val, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && val, OB200_EINVAL, "csr_get_values: null argument");
    OB_CHECK( ob200_csr_materialize(const_cast< ob200_csr * >( A )) );
    if ( A->nnz )
        OB_CUDA( cudaMemcpyAsync(val, A->val.p, sizeof( double ) * (size_t) A->nnz,
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, A->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
    return OB200_OK;
}

int ob200_csr_set_values(ob200_csr *A, const double *val, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && val, OB200_EINVAL, "csr_set_values: null argument");
    A->zero_pending = false;
    if ( A->nnz )
        OB_CUDA( cudaMemcpyAsync(A->val.p, val, sizeof( double ) * (size_t) A->nnz,
                                 on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, A->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
    A->version++;
    return OB200_OK;
}

int ob200_csr_device_arrays(ob200_csr *A, const int32_t **rowptr, const int32_t **colind, double **val)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A, OB200_EINVAL, "csr_device_arrays: null matrix");
    OB_CHECK( ob200_csr_materialize(A) );
    if ( rowptr ) *rowptr = A->rowptr.p;
    if ( colind ) *colind = A->colind.p;
    if ( val ) *val = A->val.p;
    return OB200_OK;
}

int ob200_csr_zero(ob200_csr *A)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A, OB200_EINVAL, "csr_zero: null matrix");
    A->zero_pending = true;         // performed by the next reader / partial writer, absorbed by a full overwrite
    A->version++;
    return OB200_OK;
}

int ob200_csr_scale(ob200_csr *A, double s)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A, OB200_EINVAL, "csr_scale: null matrix");
    OB_CHECK( ob200_csr_materialize(A) );
    if ( A->nnz ) {
        int grid = A->ctx->shape.grid(A->nnz, 256, 8);
        OB_LAUNCH(A->ctx, scale_kernel, grid, 256, 0, A->val.p, A->nnz, s);
    }
    A->version++;
    return OB200_OK;
}

int ob200_csr_assemble(ob200_csr *A, int64_t nelem, int32_t nd, const int32_t *loc, const double *mat, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && ( nelem == 0 || ( loc && mat ) ), OB200_EINVAL, "csr_assemble: null argument");
    OB_REQUIRE(A->rowptr.p, OB200_EINVAL, "csr_assemble: matrix has no structure");
    OB_REQUIRE(nd > 0 && nelem >= 0, OB200_EINVAL, "csr_assemble: dimension of 'k' and 'loc' mismatch");
    if ( nelem == 0 ) return OB200_OK;
    OB_CHECK( ob200_csr_materialize(A) );
    ob200_context *ctx = A->ctx;
    Staged< int32_t > L;
    Staged< double > M;
    OB_CHECK( L.stage(ctx, loc, nelem * nd, on_device) );
    OB_CHECK( M.stage(ctx, mat, nelem * nd * nd, on_device) );
    DevBuf< int > missing;
    OB_CHECK( missing.alloc(1) );
    OB_CUDA( cudaMemsetAsync(missing.p, 0, sizeof( int ), ctx->stream) );
    int grid = ctx->shape.grid(nelem * nd * nd, 256, 8);
    OB_LAUNCH(ctx, csr_assemble_kernel, grid, 256, 0, L.d, M.d, nd, nelem, A->rowptr.p, A->colind.p, A->val.p, missing.p);
    int h = 0;
    OB_CUDA( cudaMemcpyAsync(&h, missing.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    A->version++;
    OB_REQUIRE(h == 0, OB200_ESTRUCT, "csr_assemble: couldn't find %d entries in the sparse structure", h);
    return OB200_OK;
}

int ob200_csr_assemble_rect(ob200_csr *A, int32_t nr, int32_t nc, const int32_t *rloc, const int32_t *cloc, const double *mat, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && rloc && cloc && mat, OB200_EINVAL, "csr_assemble_rect: null argument");
    OB_REQUIRE(nr >= 0 && nc >= 0, OB200_EINVAL, "csr_assemble_rect: negative size");
    if ( nr == 0 || nc == 0 ) return OB200_OK;
    ob200_context *ctx = A->ctx;
    OB_CHECK( ob200_csr_materialize(A) );
    Staged< int32_t > R, Cc;
    Staged< double > M;
    OB_CHECK( R.stage(ctx, rloc, nr, on_device) );
    OB_CHECK( Cc.stage(ctx, cloc, nc, on_device) );
    OB_CHECK( M.stage(ctx, mat, (int64_t) nr * nc, on_device) );
    DevBuf< int > missing;
    OB_CHECK( missing.alloc(1) );
    OB_CUDA( cudaMemsetAsync(missing.p, 0, sizeof( int ), ctx->stream) );
    OB_LAUNCH(ctx, csr_assemble_rect_kernel, ctx->shape.grid((int64_t) nr * nc, 256, 8), 256, 0, R.d, Cc.d, M.d, nr, nc, A->rowptr.p,
              A->colind.p, A->val.p, missing.p);
    int h = 0;
    OB_CUDA( cudaMemcpyAsync(&h, missing.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    A->version++;
    OB_REQUIRE(h == 0, OB200_ESTRUCT, "csr_assemble_rect: couldn't find %d entries in the sparse structure", h);
    return OB200_OK;
}

int ob200_csr_times(ob200_csr *A, const double *x, double *y, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && ( A->neq == 0 || ( x && y ) ), OB200_EINVAL, "csr_times: null argument");
    Staged< double > X;
    StagedOut< double > Y;
    OB_CHECK( X.stage(A->ctx, x, A->neq, on_device) );
    OB_CHECK( Y.stage(A->ctx, y, A->neq, on_device) );
    OB_CHECK( spmv(A, X.d, Y.d) );
    return Y.finish(A->ctx);
}

int ob200_csr_times_t(ob200_csr *A, const double *x, double *y, int on_device)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && ( A->neq == 0 || ( x && y ) ), OB200_EINVAL, "csr_times_t: null argument");
    Staged< double > X;
    StagedOut< double > Y;
    OB_CHECK( X.stage(A->ctx, x, A->neq, on_device) );
    OB_CHECK( Y.stage(A->ctx, y, A->neq, on_device) );
    OB_CHECK( ob200_csr_materialize(A) );
    if ( A->neq ) {
        OB_CUDA( cudaMemsetAsync(Y.d, 0, sizeof( double ) * (size_t) A->neq, A->ctx->stream) );
        OB_LAUNCH(A->ctx, ob200::csr_times_t_kernel, A->ctx->shape.grid((int64_t) A->neq * 8, 256, 8), 256, 0, A->neq, A->rowptr.p,
                  A->colind.p, A->val.p, X.d, Y.d);
    }
    return Y.finish(A->ctx);
}

int ob200_csr_at(ob200_csr *A, int32_t i, int32_t j, double *value)
{
    if ( A ) ob200::bind_stream(A->ctx);
    OB_REQUIRE(A && value, OB200_EINVAL, "csr_at: null argument");
    // CompCol::at (compcol.C:376-390): "Array accessing exception -- out of bounds"
    OB_REQUIRE(i >= 1 && j >= 1 && i <= A->neq && j <= A->neq, OB200_EINVAL, "csr_at: (%d,%d) out of bounds", i, j);
    OB_CHECK( ob200_csr_materialize(A) );
    int32_t rp[2];
    OB_CUDA( cudaMemcpyAsync(rp, A->rowptr.p + ( i - 1 ), sizeof( int32_t ) * 2, cudaMemcpyDeviceToHost, A->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
    int n = rp[1] - rp[0];
    *value = 0.0;
    if ( n <= 0 ) return 1;
    std::vector< int32_t > cols(n);
    OB_CUDA( cudaMemcpyAsync(cols.data(), A->colind.p + rp[0], sizeof( int32_t ) * n, cudaMemcpyDeviceToHost, A->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
    for ( int k = 0; k < n; k++ )
        if ( cols[k] == j - 1 ) {
            OB_CUDA( cudaMemcpyAsync(value, A->val.p + rp[0] + k, sizeof( double ), cudaMemcpyDeviceToHost, A->ctx->stream) );
            OB_CUDA( cudaStreamSynchronize(A->ctx->stream) );
            return OB200_OK;
        }
    return 1;          // in bounds but not in the sparse structure: value 0 (CompCol::at const, compcol.C:392-399)
}

} // extern "C"

// scratch: time the three modes of the streamed SpMV on one GPU (route = -1 everywhere in mode 2)
extern "C" int ob200_debug_spmv_modes(ob200_csr *A, const double *x_dev, double *y_dev, int reps, float *ms3)
{
    ob200::bind_stream(A->ctx);
    ob200_context *ctx = A->ctx;
    ob200::DevBuf< int32_t > route;
    ob200::DevBuf< double > part;
    OB_CHECK( route.alloc(A->neq) );
    OB_CHECK( part.alloc(4096) );
    OB_CUDA( cudaMemsetAsync(route.p, 0xFF, sizeof( int32_t ) * (size_t) A->neq, ctx->stream) );
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for ( int mode = 0; mode < 3; mode++ ) {
        for ( int k = 0; k < reps + 2; k++ ) {
            if ( k == 2 ) cudaEventRecord(e0, ctx->stream);
            int nb = 0;
            if ( mode == 0 ) OB_CHECK( ob200::spmv(A, x_dev, y_dev) );
            if ( mode == 1 ) OB_CHECK( ob200::spmv_fused_dot(A, x_dev, y_dev, part.p, &nb, nullptr) );
            if ( mode == 2 ) {
                const ob200::SpmvHalo hv{ route.p, nullptr, nullptr, 0, 1 };
                OB_CHECK( ob200::spmv_fused_halo(A, x_dev, y_dev, part.p, &nb, nullptr, hv) );
            }
        }
        cudaEventRecord(e1, ctx->stream);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(ms3 + mode, e0, e1);
        ms3[mode] /= reps;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return OB200_OK;
}
