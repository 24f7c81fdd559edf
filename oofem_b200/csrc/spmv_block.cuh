// Blocked-index SpMV for finite element matrices (sm_100a).
//
// Same product as spmv.cuh (CompCol::times, src/core/compcol.C:119-134) on the same val array, but the
// column indices are read in compressed form.  The rows of a node (3 dofs) share one pattern and the
// columns of a neighbouring node are consecutive, so the matrix is described by
//   * row blocks: up to 3 consecutive rows with identical column sets (desc[rb] = {first row, first
//     entry, first column block, rows | row length << 8 | all-blocks-3-wide << 30 | shared-row flag << 31}), and
//   * column blocks per row block: up to 3 consecutive columns (bw = {first column, offset in the row |
//     width << 16}),
// both derived from rowptr/colind once per structure (csr_build_blocks).  Per non-zero the kernel
// streams 8 B of value and 8/9 B of index instead of 8 + 4 B, and gathers every x entry once per row
// block instead of once per row.  rowptr/colind themselves stay untouched (bit-exact with CompCol).
// Matrices without that structure (fewer than kBlkMinRatio non-zeros per column block on average, or a
// chunk that exceeds the stage) keep the plain CSR kernel.
#pragma once
#include "spmv.cuh"

namespace ob200 {

#ifndef OB200_BLK_CHUNK
#define OB200_BLK_CHUNK 2048
#endif
constexpr int kBlkChunk = OB200_BLK_CHUNK;                // non-zeros per chunk
constexpr int kBlkValCap = kBlkChunk + 520;               // values per stage (chunk + one row block + alignment)
constexpr int kBlkMinRatio = 4;                           // average non-zeros per column block required to use the blocked index
constexpr int kBlkCap = kBlkValCap / kBlkMinRatio + 120;  // column blocks per stage
constexpr int kRbCap = kBlkChunk / 8 - 6;                 // row blocks per stage
#ifndef OB200_BLK_STAGES
#define OB200_BLK_STAGES 2
#endif
#ifndef OB200_BLK_CTAS
#define OB200_BLK_CTAS 3
#endif
constexpr int kBlkStages = OB200_BLK_STAGES;      // the gathers of x, not the streamed copies, need the parallelism: more CTAs, shallower rings
constexpr int kBlkCtas = OB200_BLK_CTAS;

struct SpmvBlkStage {
    double val[kBlkValCap];
    uint2 bw[kBlkCap + 8];
    int4 desc[kRbCap + 6];
};
struct SpmvBlkShared {
    SpmvBlkStage st[kBlkStages];
    unsigned long long full[kBlkStages], empty[kBlkStages];
    int4 meta0[kBlkStages], meta1[kBlkStages];
    double scratch[32];
};

// rows with the same column set as the row before them
__global__ void __launch_bounds__(256)
blk_row_same_kernel(int32_t neq, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, unsigned char *__restrict__ same)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5, nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    for ( int64_t r = warp0; r < neq; r += nwarps ) {
        int eq = 0;
        if ( r > 0 ) {
            const int a0 = rowptr[r - 1], a1 = rowptr[r], a2 = rowptr[r + 1];
            eq = ( a1 - a0 == a2 - a1 ) && a2 > a1;
            if ( eq )
                for ( int k = lane; k < a2 - a1 && eq; k += 32 ) eq = colind[a0 + k] == colind[a1 + k];
            eq = __all_sync(0xffffffffu, eq);
        }
        if ( lane == 0 ) same[r] = (unsigned char) eq;
    }
}

// Column blocks of the first row of every row block: runs of consecutive columns cut into pieces of 3.
// FILL = false: count; FILL = true: write the block words and the row-block descriptors.
template< bool FILL >
__global__ void __launch_bounds__(256)
blk_columns_kernel(int32_t nrb, const int32_t *__restrict__ rbrow, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                   int32_t *__restrict__ cnt, const int64_t *__restrict__ b0, uint2 *__restrict__ bw, int4 *__restrict__ desc)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5, nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    for ( int64_t rb = warp0; rb < nrb; rb += nwarps ) {
        const int r0 = rbrow[rb], nr = rbrow[rb + 1] - r0;
        const int s = rowptr[r0], e = rowptr[r0 + 1];
        int carry = s, nblk = 0;                    // start of the run the previous 32 entries ended in
        bool regular = true;                        // every column block is 3 wide
        for ( int base = s; base < e; base += 32 ) {
            const int k = base + lane;
            const int c = k < e ? colind[k] : -2;
            const int cprev = ( k > s && k < e ) ? colind[k - 1] : -4;
            const bool brk = k < e && ( k == s || c != cprev + 1 );
            const unsigned int m = __ballot_sync(0xffffffffu, brk);
            const unsigned int below = m & ( 0xffffffffu >> ( 31 - lane ) );
            const int runstart = below ? base + 31 - __clz(below) : carry;
            const bool start = k < e && ( ( k - runstart ) % 3 == 0 );
            const unsigned int ms = __ballot_sync(0xffffffffu, start);
            if ( FILL ) {
                int w = 3;
                if ( start ) {
                    w = 1;
                    if ( k + 1 < e && colind[k + 1] == c + 1 ) {
                        w = 2;
                        if ( k + 2 < e && colind[k + 2] == c + 2 ) w = 3;
                    }
                }
                regular = __all_sync(0xffffffffu, w == 3) && regular;
                if ( start ) bw[b0[rb] + nblk + __popc(ms & ( ( 1u << lane ) - 1u ))] = make_uint2((unsigned int) c, (unsigned int)( k - s ) | ( (unsigned int) w << 16 ));
            }
            nblk += __popc(ms);
            if ( m ) carry = base + 31 - __clz(m);
        }
        if ( lane == 0 ) {
            if ( FILL ) desc[rb] = make_int4(r0, s, (int) b0[rb], nr | ( ( e - s ) << 8 ) | ( regular ? ( 1 << 30 ) : 0 ));
            else cnt[rb] = nblk;
        }
    }
}

// chunk table over row blocks: table[c] = {first row block whose first entry is >= c*kBlkChunk, its first entry, its first column block}
__global__ void blk_chunk_table_kernel(int32_t nrb, const int4 *__restrict__ desc, int32_t nchunks, int4 *__restrict__ table)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r <= nrb; r += stride ) {
        const int4 d = desc[r];
        int hi = d.y / kBlkChunk;
        const int lo = r == 0 ? 0 : desc[r - 1].y / kBlkChunk + 1;
        if ( hi > nchunks - 1 ) hi = nchunks - 1;
        for ( int c = lo; c <= hi; c++ ) table[c] = make_int4((int) r, d.y, d.z, 0);
        if ( r == nrb )
            for ( int c = ( hi + 1 > lo ? hi + 1 : lo ); c <= nchunks; c++ ) table[c] = make_int4(nrb, d.y, d.z, 0);
    }
}

// descriptors with the shared-row flag (distributed product): bit 31 of .w set if any row of the block is shared
__global__ void blk_flag_desc_kernel(int32_t nrb, const int4 *__restrict__ desc, const int32_t *__restrict__ route, int4 *__restrict__ out)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t rb = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; rb <= nrb; rb += stride ) {
        int4 d = desc[rb];
        if ( rb < nrb ) {
            const int nr = d.w & 0xFF;
            bool sh = false;
            for ( int i = 0; i < nr; i++ ) sh |= route[d.x + i] >= 0;
            if ( sh ) d.w |= (int) 0x80000000;
        }
        out[rb] = d;
    }
}

// y = A x on the blocked index.  MODE as in spmv_stream_kernel.  One warp per row block, one lane per column block.
template< int MODE >
__global__ void __launch_bounds__(kSpmvThreads, kBlkCtas)
spmv_block_kernel(const double *__restrict__ val, const uint2 *__restrict__ bw, const int4 *__restrict__ desc,
                  const int4 *__restrict__ table, int32_t nchunks, const double *__restrict__ x, double *__restrict__ y,
                  double *__restrict__ partials, const int *__restrict__ done, SpmvHalo halo)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SpmvBlkShared &sh = *reinterpret_cast< SpmvBlkShared * >( smem_raw );
    if ( done && *done ) return;
    const int tid = threadIdx.x;
    if ( tid == 0 ) {
#pragma unroll
        for ( int s = 0; s < kBlkStages; s++ ) {
            mbar_init(&sh.full[s], 1);
            mbar_init(&sh.empty[s], kSpmvConsumers / 32);
        }
        asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    __syncthreads();
    const int nmine = ( nchunks > (int) blockIdx.x ) ? ( nchunks - 1 - (int) blockIdx.x ) / (int) gridDim.x + 1 : 0;

    if ( tid >= kSpmvConsumers ) {
        // ---- producer: three bulk copies per chunk (values, column-block words, row-block descriptors) ----
        if ( tid != kSpmvConsumers ) return;
        const uint64_t policy = l2_policy_evict_first();
        int4 t0 = make_int4(0, 0, 0, 0), t1 = t0;
        if ( nmine > 0 ) {
            t0 = table[blockIdx.x];
            t1 = table[blockIdx.x + 1];
        }
        for ( int k = 0; k < nmine; k++ ) {
            const int s = k % kBlkStages;
            int4 n0 = t0, n1 = t1;
            if ( k + 1 < nmine ) {
                const int c = blockIdx.x + ( k + 1 ) * gridDim.x;
                n0 = table[c];
                n1 = table[c + 1];
            }
            if ( k >= kBlkStages ) mbar_wait(&sh.empty[s], ( ( k / kBlkStages ) - 1 ) & 1);
            SpmvBlkStage &st = sh.st[s];
            sh.meta0[s] = t0;
            sh.meta1[s] = t1;
            const int a0 = t0.y & ~1, nv = ( ( t1.y - a0 ) + 1 ) & ~1;          // 16-byte granules
            const int b0 = t0.z & ~1, nb = ( ( t1.z - b0 ) + 1 ) & ~1;
            const int nd = t1.x - t0.x + 1;                                      // + the descriptor after the last row block
            if ( t1.x > t0.x && nv > 0 ) {
                mbar_expect_tx(&sh.full[s], (uint32_t) nv * 8u + (uint32_t) nb * 8u + (uint32_t) nd * 16u);
                tma_load_1d(st.val, val + a0, (uint32_t) nv * 8u, &sh.full[s], policy);
                tma_load_1d(st.bw, bw + b0, (uint32_t) nb * 8u, &sh.full[s], policy);
                tma_load_1d(st.desc, desc + t0.x, (uint32_t) nd * 16u, &sh.full[s], policy);
            } else {
                mbar_expect_tx(&sh.full[s], 0);
            }
            t0 = n0;
            t1 = n1;
        }
        return;
    }

    // ---- consumers ----
    double pq = 0.0;
    const int wid = tid >> 5, lane = tid & 31;
    constexpr int kWarps = kSpmvConsumers / 32;
    for ( int k = 0; k < nmine; k++ ) {
        const int s = k % kBlkStages;
        const SpmvBlkStage &st = sh.st[s];
        mbar_wait(&sh.full[s], ( k / kBlkStages ) & 1);
        const int4 m0 = sh.meta0[s], m1 = sh.meta1[s];
        const int a0 = m0.y & ~1, b0 = m0.z & ~1;
        // row block rb belongs to warp rb mod 8 whatever chunk it falls into: chunks with few row blocks still load all warps evenly
        for ( int rb = m0.x + ( ( wid - m0.x ) & ( kWarps - 1 ) ); rb < m1.x; rb += kWarps ) {
            const int4 d = st.desc[rb - m0.x], dn = st.desc[rb - m0.x + 1];
            const int nr = d.w & 0xFF, rowlen = ( d.w >> 8 ) & 0x3FFFFF, nb = dn.z - d.z;
            const bool regular = ( d.w & ( 1 << 30 ) ) != 0;
            // what the epilogue needs, requested before the block is summed
            double xr = 0.0;
            if ( MODE != 0 && lane < nr ) xr = __ldg(x + d.x + lane);
            const double *v0 = st.val + ( d.y - a0 );
            const uint2 *bwp = st.bw + ( d.z - b0 );
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            if ( regular ) {
                // every column block is 3 wide: lane = entry of the row, so that the gathers of x and the reads of
                // the values are coalesced (a lane per column block costs three times the L1 requests)
                for ( int kb = 0; kb < rowlen; kb += 96 ) {
                    // three entries per lane, all gathers in flight before the first product
                    const int k0 = kb + lane, k1 = k0 + 32, k2 = k0 + 64;
                    const bool p0 = k0 < rowlen, p1 = k1 < rowlen, p2 = k2 < rowlen;
                    const int j0 = k0 / 3, j1 = k1 / 3, j2 = k2 / 3;
                    const double x0 = p0 ? __ldg(x + bwp[j0].x + ( k0 - 3 * j0 )) : 0.0;
                    const double x1 = p1 ? __ldg(x + bwp[j1].x + ( k1 - 3 * j1 )) : 0.0;
                    const double x2 = p2 ? __ldg(x + bwp[j2].x + ( k2 - 3 * j2 )) : 0.0;
                    const double *va = v0, *vb = v0 + ( nr > 1 ? rowlen : 0 ), *vc = v0 + ( nr > 2 ? 2 * rowlen : 0 );
                    if ( p0 ) { s0 += va[k0] * x0; s1 += vb[k0] * x0; s2 += vc[k0] * x0; }
                    if ( p1 ) { s0 += va[k1] * x1; s1 += vb[k1] * x1; s2 += vc[k1] * x1; }
                    if ( p2 ) { s0 += va[k2] * x2; s1 += vb[k2] * x2; s2 += vc[k2] * x2; }
                }
            } else
            for ( int j = lane; j < nb; j += 32 ) {
                const uint2 b = bwp[j];
                const int w = (int)( b.y >> 16 );
                const double *p = v0 + ( b.y & 0xFFFFu );
                const double *xp = x + b.x;
                const double x0 = __ldg(xp), x1 = w > 1 ? __ldg(xp + 1) : 0.0, x2 = w > 2 ? __ldg(xp + 2) : 0.0;
                if ( w == 3 && nr == 3 ) {
                    s0 += p[0] * x0 + p[1] * x1 + p[2] * x2;
                    p += rowlen;
                    s1 += p[0] * x0 + p[1] * x1 + p[2] * x2;
                    p += rowlen;
                    s2 += p[0] * x0 + p[1] * x1 + p[2] * x2;
                } else {
                    double t[3] = { 0.0, 0.0, 0.0 };
                    for ( int i = 0; i < nr; i++ ) {
                        double a = p[0] * x0;
                        if ( w > 1 ) a += p[1] * x1;
                        if ( w > 2 ) a += p[2] * x2;
                        t[i] = a;
                        p += rowlen;
                    }
                    s0 += t[0];
                    s1 += t[1];
                    s2 += t[2];
                }
            }
#pragma unroll
            for ( int o = 16; o > 0; o >>= 1 ) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if ( lane < nr ) {
                const double sv = lane == 0 ? s0 : ( lane == 1 ? s1 : s2 );
                const int row = d.x + lane;
                y[row] = sv;
                if ( MODE == 1 ) pq += sv * xr;
                if ( MODE == 2 ) {
                    int u = -1;
                    if ( d.w < 0 ) u = __ldg(halo.route + row);          // only row blocks flagged as shared look the row up
                    if ( u < 0 ) pq += sv * xr;
                    else if ( halo.dst )
                        for ( int i = __ldg(halo.uptr + u), e2 = __ldg(halo.uptr + u + 1); i < e2; i++ )
                            ll_store(halo.dst[i] + halo.half_words, sv, halo.seq);
                }
            }
        }
        __syncwarp();
        if ( lane == 0 ) asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( smem_u32(&sh.empty[s]) ) : "memory" );
    }
    if ( MODE != 0 ) {
        // block sum over the consumer warps in a fixed order
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
