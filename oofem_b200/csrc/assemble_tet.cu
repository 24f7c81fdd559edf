// Owner-computes assembly of the LTRSpace tangent into the cudacsr matrix ("node rows").
//
// Replaces EngngModel::assemble (src/core/engngm.C:889-929) + CompCol::assemble (src/core/compcol.C:263-299) for a whole
// LTRSpace element set (IsotropicLinearElasticMaterial and / or MisesMat) without atomics: the up to three matrix rows of a
// node are produced by one warp and written once, and the contributions of the elements around the node are added in
// ascending element number -- bit-reproducible run to run.
//
//   * per element (once per mesh): the constant gradients of FEI3dTetLin::evaldNdx (src/core/fei3dtetlin.C:116-166), the
//     volume (1-point rule, weight 1/6: gaussintegrationrule.C:500-507) and the volume-weighted Lame constants
//     (isolinearelasticmaterial.C:80-84): 16 doubles.  With a MisesMat in the set also, per call, the 6x6 algorithmic tangent
//     of the element's Gauss point (MisesMat::give3dMaterialStiffnessMatrix, misesmat.C:493-545).
//   * per node A: the column blocks (neighbour nodes B) are read off the first free row of A in the CSR structure; one lane
//     per block walks the elements around A (node -> element incidence built at create), and where the element contains B
//     adds V B_a^T D B_b (Structural3DElement::computeBmatrixAt, structural3delement.C:63-86) to its 3x3 accumulator.
//
// tet_bind verifies what the kernel relies on (the rows of a node have one column pattern, the free equations of a node are
// consecutive, every block an element produces is in the pattern); otherwise the slot-map path stays in use.
#include "element_device.cuh"
#include "elemset.h"
#include "scan.cuh"
#include <limits.h>
#include <string.h>
#include <stdlib.h>

namespace ob200 {

constexpr int kTetWarps = 8;
constexpr int kTetMaxBlk = 128;          // column blocks of one node's rows
constexpr int kTetRec = 16;              // doubles per element: g[4][3], V lambda, V mu, V, pad

// eqnode[eq-1] = node*4 + component
__global__ void tet_eqnode_kernel(int64_t nnode, const int32_t *__restrict__ nodeeq, int32_t *__restrict__ eqnode)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < nnode * 3; t += stride ) {
        const int eq = nodeeq[t];
        if ( eq > 0 ) eqnode[eq - 1] = (int32_t)( ( t / 3 ) * 4 + t % 3 );
    }
}

// One warp per node.  flags[0]: the structure is not what the kernel assumes; flags[1]: more than kTetMaxBlk column blocks;
// covered: number of matrix entries the kernel writes.
__global__ void __launch_bounds__(256)
tet_rows_check_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                      const int32_t *__restrict__ conn, const int32_t *__restrict__ nodeeq, const int32_t *__restrict__ eqnode,
                      const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int32_t neq,
                      int *__restrict__ flags, unsigned long long *__restrict__ covered)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    for ( int64_t A = warp0; A < nnode; A += nwarps ) {
        int eq[3], nfree = 0, rfirst = 0, bad = 0;
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            eq[i] = nodeeq[A * 3 + i];
            if ( eq[i] < 0 || eq[i] > neq ) { bad = 1; eq[i] = 0; }
            if ( eq[i] ) {
                if ( eqnode[eq[i] - 1] != (int32_t)( A * 4 + i ) ) bad = 1;      // an equation shared by two dofs
                if ( rfirst && eq[i] != rfirst + nfree ) bad = 1;                // not consecutive
                if ( !rfirst ) rfirst = eq[i];
                nfree++;
            }
        }
        if ( !rfirst ) {
            if ( bad && lane == 0 ) atomicOr(flags, 1);
            continue;
        }
        const int p0 = rowptr[rfirst - 1], len = rowptr[rfirst] - p0;
        // all rows of the node have the same columns
        for ( int i = 0; i < 3; i++ ) {
            if ( !eq[i] || eq[i] == rfirst ) continue;
            const int p = rowptr[eq[i] - 1];
            if ( rowptr[eq[i]] - p != len ) { bad = 1; continue; }
            for ( int t = lane; t < len; t += 32 ) if ( colind[p + t] != colind[p0 + t] ) bad = 1;
        }
        // number of column blocks
        int nb = 0, carry = -1;
        for ( int t0 = 0; t0 < len; t0 += 32 ) {
            const int t = t0 + lane;
            const int node = t < len ? ( eqnode[colind[p0 + t]] >> 2 ) : -2;
            if ( node == -1 ) bad = 1;                   // a column that is no dof of this element set
            int prev = __shfl_up_sync(0xffffffffu, node, 1);
            if ( lane == 0 ) prev = carry;
            nb += __popc(__ballot_sync(0xffffffffu, t < len && node != prev));
            carry = __shfl_sync(0xffffffffu, node, 31);
        }
        // every block the elements around the node produce is in the row
        const int v0 = ninc_start[A], nv = ninc_start[A + 1] - v0;
        for ( int it = lane; it < nv * 4; it += 32 ) {
            const int ea = ninc[v0 + ( it >> 2 )];
            const int B = conn[(int64_t)( ea >> 3 ) * 4 + ( it & 3 )] - 1;
            int cB = 0;
#pragma unroll
            for ( int j = 2; j >= 0; j-- ) if ( nodeeq[(int64_t) B * 3 + j] > 0 ) cB = nodeeq[(int64_t) B * 3 + j];
            if ( cB ) {
                int lo = p0, hi = p0 + len - 1, found = 0;
                while ( lo <= hi ) {
                    const int mid = ( lo + hi ) >> 1, v = colind[mid];
                    if ( v == cB - 1 ) { found = 1; break; }
                    if ( v < cB - 1 ) lo = mid + 1; else hi = mid - 1;
                }
                if ( !found ) bad = 1;
            }
        }
        if ( __any_sync(0xffffffffu, bad) && lane == 0 ) atomicOr(flags, 1);
        if ( lane == 0 ) {
            if ( nb > kTetMaxBlk ) atomicOr(flags + 1, 1);
            atomicAdd(covered, (unsigned long long) nfree * (unsigned long long) len);
        }
    }
}

// gradients, volume and volume-weighted Lame constants of every element
__global__ void __launch_bounds__(128)
tet_geometry_kernel(ElemSetView S, int64_t nelem, double *__restrict__ rec)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride ) {
        double g[4][3], c[12];
#pragma unroll
        for ( int a = 0; a < 4; a++ ) {
            const int node = S.conn[e * 4 + a] - 1;
#pragma unroll
            for ( int j = 0; j < 3; j++ ) c[3 * a + j] = S.coords[(int64_t) node * 3 + j];
        }
        const double dV = fabs(tet_dNdx(c, g)) * ( 1.0 / 6.0 );
        const MatParams mp = S.mat[S.matid[e]];
        double lam, mu;
        isole_lame(mp.E, mp.nu, lam, mu);
        double2 *o = reinterpret_cast< double2 * >( rec + e * kTetRec );
#pragma unroll
        for ( int a = 0; a < 2; a++ ) {
            o[3 * a] = make_double2(g[2 * a][0], g[2 * a][1]);
            o[3 * a + 1] = make_double2(g[2 * a][2], g[2 * a + 1][0]);
            o[3 * a + 2] = make_double2(g[2 * a + 1][1], g[2 * a + 1][2]);
        }
        o[6] = make_double2(dV * lam, dV * mu);
        o[7] = make_double2(dV, 0.0);
    }
}

// material tangent of every element's Gauss point (sets with a MisesMat)
__global__ void __launch_bounds__(128)
tet_tangent_kernel(ElemSetView S, int64_t nelem, double *__restrict__ Dout)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride ) {
        const MatParams mp = S.mat[S.matid[e]];
        double D[36];
        if ( mp.type == (double) OB200_MAT_MISES ) mises_tangent(mp, mises_ref(S, e), D);
        else isole_D(mp.E, mp.nu, D);
#pragma unroll
        for ( int i = 0; i < 36; i++ ) Dout[e * 36 + i] = D[i];
    }
}

struct TetRowsView {
    int64_t nnode;
    const int32_t *ninc_start, *ninc, *conn, *nodeeq, *eqnode, *rowptr, *colind;
    const double *rec, *D;
};

template< int MODE >        // bit 0: add to val (else overwrite), bit 1: general 6x6 tangent per element (else isotropic)
__global__ void __launch_bounds__(kTetWarps * 32)
ltrspace_rows_kernel(TetRowsView V, double *__restrict__ val)
{
    constexpr bool ACCUM = ( MODE & 1 ) != 0, GEN = ( MODE & 2 ) != 0;
    __shared__ int4 s_conn[kTetWarps][32];
    __shared__ int s_ea[kTetWarps][32];
    __shared__ int s_blk[kTetWarps][kTetMaxBlk];
    __shared__ unsigned short s_off[kTetWarps][kTetMaxBlk];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t) gridDim.x * kTetWarps;
    for ( int64_t A = (int64_t) blockIdx.x * kTetWarps + w; A < V.nnode; A += nwarps ) {
        int eqA[3];
#pragma unroll
        for ( int i = 0; i < 3; i++ ) eqA[i] = V.nodeeq[A * 3 + i];
        const int rfirst = eqA[0] ? eqA[0] : eqA[1] ? eqA[1] : eqA[2];
        if ( !rfirst ) continue;
        const int p0 = V.rowptr[rfirst - 1], len = V.rowptr[rfirst] - p0;
        // the column blocks of the node: runs of entries that belong to one column node
        int nb = 0, carry = -1;
        for ( int t0 = 0; t0 < len; t0 += 32 ) {
            const int t = t0 + lane;
            const int node = t < len ? ( V.eqnode[V.colind[p0 + t]] >> 2 ) : -2;
            int prev = __shfl_up_sync(0xffffffffu, node, 1);
            if ( lane == 0 ) prev = carry;
            const bool start = t < len && node != prev;
            const unsigned m = __ballot_sync(0xffffffffu, start);
            if ( start ) {
                const int idx = nb + __popc(m & ( ( 1u << lane ) - 1u ));
                if ( idx < kTetMaxBlk ) {
                    s_blk[w][idx] = node;
                    s_off[w][idx] = (unsigned short) t;
                }
            }
            nb += __popc(m);
            carry = __shfl_sync(0xffffffffu, node, 31);
        }
        if ( nb > kTetMaxBlk ) nb = kTetMaxBlk;          // tet_bind declines such a matrix
        __syncwarp();
        const int v0 = V.ninc_start[A], nv = V.ninc_start[A + 1] - v0;
        for ( int bb = 0; bb < nb; bb += 32 ) {
            const bool active = bb + lane < nb;
            const int B = active ? s_blk[w][bb + lane] : -1;
            double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            for ( int vb = 0; vb < nv; vb += 32 ) {
                if ( vb + lane < nv ) {
                    const int ea = V.ninc[v0 + vb + lane];
                    s_ea[w][lane] = ea;
                    s_conn[w][lane] = reinterpret_cast< const int4 * >( V.conn )[ea >> 3];
                }
                __syncwarp();
                const int n = min(32, nv - vb);
                for ( int k = 0; k < n; k++ ) {          // ascending element number
                    const int4 cn = s_conn[w][k];
                    const int b = cn.x - 1 == B ? 0 : cn.y - 1 == B ? 1 : cn.z - 1 == B ? 2 : cn.w - 1 == B ? 3 : -1;
                    if ( b >= 0 ) {
                        const int ea = s_ea[w][k];
                        const int64_t e = ea >> 3;
                        const double *rec = V.rec + e * kTetRec;
                        const int a = ea & 7;
                        const double ga[3] = { rec[3 * a], rec[3 * a + 1], rec[3 * a + 2] };
                        const double gb[3] = { rec[3 * b], rec[3 * b + 1], rec[3 * b + 2] };
                        if ( GEN ) block_general(acc, ga, gb, V.D + e * 36, rec[14]);
                        else block_iso(acc, ga, gb, rec[12], rec[13]);
                    }
                }
                __syncwarp();
            }
            if ( active ) {
                const int off = s_off[w][bb + lane];
                const bool fb[3] = { V.nodeeq[(int64_t) B * 3] > 0, V.nodeeq[(int64_t) B * 3 + 1] > 0, V.nodeeq[(int64_t) B * 3 + 2] > 0 };
#pragma unroll
                for ( int i = 0; i < 3; i++ ) {
                    if ( !eqA[i] ) continue;
                    double *o = val + V.rowptr[eqA[i] - 1] + off;
                    int jj = 0;
#pragma unroll
                    for ( int j = 0; j < 3; j++ )
                        if ( fb[j] ) {
                            if ( ACCUM ) o[jj] += acc[3 * i + j];
                            else o[jj] = acc[3 * i + j];
                            jj++;
                        }
                }
            }
        }
    }
}

// ---- the fast path: per-node tables built once per bound matrix ------------------------------------------------
//
// Per node (two int4): d0 = { first incidence, elements around the node, byte offset of its table, column blocks },
// d1 = { start of row 0 / 1 / 2 in val or -1 (prescribed), 0 }.  The table: per column block its first column << 3 | free-dof mask
// (u16), the units -- a column block's items in runs of at most kTet2Unit: start of every unit's item list (u8 [nu + 1]) and its block |
// first-unit-of-its-block << 6 | block-has-several-units << 7 (u8 [nu]) --, per item (incidence v, local node b) its column block or 0xFF
// (u8 [4 nv]), and the items ordered by (block, incidence) (u8 [4 nv]).  With it the kernel has no search and no divergence: one lane
// per (element, local node) item forms its 3x3 block and parks it in shared memory, then one lane per unit adds the parked blocks of
// its list in ascending element number (with kTet2Unit = all items a unit is a whole column block).  The order of every sum is
// fixed by the tables: bit-reproducible run to run.
constexpr int kTet2Warps = 4;
constexpr int kTet2MaxVal = 32;          // elements around a node
constexpr int kTet2MaxBlk = 64;          // column blocks of a node's rows
constexpr int kTet2Items = 4 * kTet2MaxVal;
constexpr int kTet2Unit = kTet2Items;   // items one lane adds.  Measured (scripts/sweep_tet.sh): cutting the long lists (the diagonal block has one item per
                                        // element around the node) into units of 8 shared by several lanes is slower, 1.96 vs 1.72 ms; so is staging
                                        // the element records in shared memory (1.97 ms).  The tables keep the unit format.
constexpr int kTet2MaxUnits = kTet2MaxBlk + kTet2Items / kTet2Unit;
constexpr int kTet2TabMax = 2 * kTet2MaxBlk + ( 2 * kTet2MaxUnits + 1 ) + 2 * kTet2Items + 8;

// most units of a node: one per block + one per kTet2Unit items
__device__ __forceinline__ int tet2_max_units(int nb, int nv) { return nb + ( 4 * nv ) / kTet2Unit; }
__device__ __forceinline__ int tet2_table_bytes(int nb, int nv)
{
    return ( ( 2 * nb + 3 ) & ~3 ) + ( ( 2 * tet2_max_units(nb, nv) + 1 + 8 * nv + 3 ) & ~3 );
}

template< bool FILL >
__global__ void __launch_bounds__(256)
tet_tables_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                  const int32_t *__restrict__ conn, const int32_t *__restrict__ nodeeq, const int32_t *__restrict__ eqnode,
                  const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                  int32_t *__restrict__ tbytes, const int32_t *__restrict__ toff, int4 *__restrict__ rdesc, unsigned char *__restrict__ tab,
                  int *__restrict__ flags)
{
    __shared__ int s_blk[8][kTet2MaxBlk];
    __shared__ unsigned short s_pk[8][kTet2MaxBlk];
    __shared__ unsigned char s_bidx[8][kTet2Items];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    for ( int64_t A = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5; A < nnode; A += nwarps ) {
        int eqA[3];
#pragma unroll
        for ( int i = 0; i < 3; i++ ) eqA[i] = nodeeq[A * 3 + i];
        const int rfirst = eqA[0] > 0 ? eqA[0] : eqA[1] > 0 ? eqA[1] : eqA[2] > 0 ? eqA[2] : 0;
        const int v0 = ninc_start[A], nv = ninc_start[A + 1] - v0;
        int nb = 0;
        if ( rfirst ) {
            const int p0 = rowptr[rfirst - 1], len = rowptr[rfirst] - p0;
            int carry = -1;
            for ( int t0 = 0; t0 < len; t0 += 32 ) {
                const int t = t0 + lane;
                const int node = t < len ? ( eqnode[colind[p0 + t]] >> 2 ) : -2;
                int prev = __shfl_up_sync(0xffffffffu, node, 1);
                if ( lane == 0 ) prev = carry;
                const bool start = t < len && node != prev;
                const unsigned m = __ballot_sync(0xffffffffu, start);
                if ( start ) {
                    const int idx = nb + __popc(m & ( ( 1u << lane ) - 1u ));
                    if ( idx < kTet2MaxBlk ) {
                        s_blk[w][idx] = node;
                        const int cm = ( nodeeq[(int64_t) node * 3] > 0 ? 1 : 0 ) | ( nodeeq[(int64_t) node * 3 + 1] > 0 ? 2 : 0 ) |
                                       ( nodeeq[(int64_t) node * 3 + 2] > 0 ? 4 : 0 );
                        s_pk[w][idx] = (unsigned short)( ( t << 3 ) | cm );
                    }
                }
                nb += __popc(m);
                carry = __shfl_sync(0xffffffffu, node, 31);
            }
            if ( lane == 0 && ( nb > kTet2MaxBlk || nv > kTet2MaxVal || len >= 8192 ) ) atomicOr(flags, 1);      // beyond the fast path
            if ( nb > kTet2MaxBlk ) nb = kTet2MaxBlk;
        }
        const int nvu = rfirst ? min(nv, kTet2MaxVal) : 0;
        if ( !FILL ) {
            if ( lane == 0 ) tbytes[A] = rfirst ? tet2_table_bytes(nb, nvu) : 0;
            continue;
        }
        if ( !rfirst ) {
            if ( lane == 0 ) {
                rdesc[2 * A] = make_int4(v0, 0, 0, 0);
                rdesc[2 * A + 1] = make_int4(-1, -1, -1, 0);
            }
            continue;
        }
        __syncwarp();
        unsigned char *T = tab + toff[A];
        unsigned short *colpk = reinterpret_cast< unsigned short * >( T );
        const int numax = tet2_max_units(nb, nvu);
        unsigned char *ustart = T + ( ( 2 * nb + 3 ) & ~3 );
        unsigned char *ublk = ustart + numax + 1, *bidx = ublk + numax, *items = bidx + 4 * nvu;
        for ( int B = lane; B < nb; B += 32 ) colpk[B] = s_pk[w][B];
        // block of every item
        for ( int it = lane; it < 4 * nvu; it += 32 ) {
            const int ea = ninc[v0 + ( it >> 2 )];
            const int Bn = conn[(int64_t)( ea >> 3 ) * 4 + ( it & 3 )] - 1;
            int idx = 0xFF;
            for ( int B = 0; B < nb; B++ )
                if ( s_blk[w][B] == Bn ) idx = B;
            s_bidx[w][it] = (unsigned char) idx;
            bidx[it] = (unsigned char) idx;
        }
        __syncwarp();
        // items ordered by (block, incidence), cut into units: lane = block counts its items, scans give the starts, then it lists them
        int run = 0, urun = 0;
        for ( int b0 = 0; b0 < nb; b0 += 32 ) {
            const int B = b0 + lane;
            int cnt = 0;
            if ( B < nb )
                for ( int it = 0; it < 4 * nvu; it++ ) cnt += s_bidx[w][it] == B;
            const int nun = B < nb ? max(1, ( cnt + kTet2Unit - 1 ) / kTet2Unit) : 0;
            int incl = cnt, uincl = nun;
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o), tu = __shfl_up_sync(0xffffffffu, uincl, o);
                if ( lane >= o ) { incl += t; uincl += tu; }
            }
            int at = run + incl - cnt;
            const int ut = urun + uincl - nun;
            if ( B < nb ) {
                for ( int k = 0; k < nun; k++ ) {
                    ustart[ut + k] = (unsigned char)( at + k * kTet2Unit );
                    ublk[ut + k] = (unsigned char)( B | ( k == 0 ? 0x40 : 0 ) | ( nun > 1 ? 0x80 : 0 ) );
                }
                for ( int it = 0; it < 4 * nvu; it++ )
                    if ( s_bidx[w][it] == B ) items[at++] = (unsigned char) it;
            }
            run += __shfl_sync(0xffffffffu, incl, 31);
            urun += __shfl_sync(0xffffffffu, uincl, 31);
        }
        if ( lane == 0 ) {
            ustart[urun] = (unsigned char) run;
            rdesc[2 * A] = make_int4(v0, nvu, toff[A], nb);
            rdesc[2 * A + 1] = make_int4(eqA[0] > 0 ? rowptr[eqA[0] - 1] : -1, eqA[1] > 0 ? rowptr[eqA[1] - 1] : -1,
                                         eqA[2] > 0 ? rowptr[eqA[2] - 1] : -1, urun);
        }
        __syncwarp();
    }
}

struct TetRows2View {
    int64_t nnode;
    const int4 *rdesc;
    const unsigned char *tab;
    const int32_t *ninc;
    const double *rec, *D;
};

template< int MODE >        // bit 0: add to val (else overwrite), bit 1: general 6x6 tangent per element (else isotropic)
__global__ void __launch_bounds__(kTet2Warps * 32)
ltrspace_rows2_kernel(const __grid_constant__ TetRows2View V, double *__restrict__ val)
{
    constexpr bool ACCUM = ( MODE & 1 ) != 0, GEN = ( MODE & 2 ) != 0;
    __shared__ double s_park[kTet2Warps][kTet2Items * 9];
    __shared__ __align__(16) unsigned char s_tab[kTet2Warps][( kTet2TabMax + 15 ) & ~15];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t) gridDim.x * kTet2Warps;
    int64_t A = (int64_t) blockIdx.x * kTet2Warps + w;
    if ( A >= V.nnode ) return;
    // the descriptor of the next node is requested one node ahead
    int4 d0 = V.rdesc[2 * A], d1 = V.rdesc[2 * A + 1];
    for ( ; A < V.nnode; A += nwarps ) {
        const int4 c0 = d0, c1 = d1;
        if ( A + nwarps < V.nnode ) {
            d0 = V.rdesc[2 * ( A + nwarps )];
            d1 = V.rdesc[2 * ( A + nwarps ) + 1];
        }
        const int v0 = c0.x, nv = c0.y, nb = c0.w;
        if ( nb == 0 ) continue;
        const int tbytes = tet2_table_bytes(nb, nv);
        {   // the node's table into shared memory, its element list into registers
            const uint32_t *src = reinterpret_cast< const uint32_t * >( V.tab + c0.z );
            uint32_t *dst = reinterpret_cast< uint32_t * >( s_tab[w] );
            for ( int t = lane; t < tbytes / 4; t += 32 ) dst[t] = src[t];
        }
        const int ea_l = lane < nv ? V.ninc[v0 + lane] : 0;
        __syncwarp();
        const unsigned short *colpk = reinterpret_cast< const unsigned short * >( s_tab[w] );
        const int nu = c1.w, numax = tet2_max_units(nb, nv);
        const unsigned char *ustart = s_tab[w] + ( ( 2 * nb + 3 ) & ~3 );
        const unsigned char *ublk = ustart + numax + 1, *bidx = ublk + numax, *items = bidx + 4 * nv;
        // one lane per (element, local node): K_ab = V B_a^T D B_b (Structural3DElement::computeBmatrixAt, structural3delement.C:63-86)
        for ( int it = lane; it - lane < 4 * nv; it += 32 ) {
            const int ea = __shfl_sync(0xffffffffu, ea_l, ( it >> 2 ) & 31);
            if ( it >= 4 * nv || bidx[it] == 0xFF ) continue;
            const int64_t e = ea >> 3;
            const int a = ea & 7, b = it & 3;
            const double *rec = V.rec + e * kTetRec;
            const double ga[3] = { rec[3 * a], rec[3 * a + 1], rec[3 * a + 2] };
            const double gb[3] = { rec[3 * b], rec[3 * b + 1], rec[3 * b + 2] };
            double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            if ( GEN ) block_general(acc, ga, gb, V.D + e * 36, rec[14]);
            else block_iso(acc, ga, gb, rec[12], rec[13]);
            double *pk = s_park[w] + it * 9;
#pragma unroll
            for ( int k = 0; k < 9; k++ ) pk[k] = acc[k];
        }
        __syncwarp();
        auto write_block = [&](int B, const double ( &acc )[9]) {
            const int pkc = colpk[B], cm = pkc & 7, off = pkc >> 3;
            const int rb[3] = { c1.x, c1.y, c1.z };
#pragma unroll
            for ( int i = 0; i < 3; i++ ) {
                if ( rb[i] < 0 ) continue;
                double *o = val + rb[i] + off;
                int jj = 0;
#pragma unroll
                for ( int j = 0; j < 3; j++ )
                    if ( cm & ( 1 << j ) ) {
                        if ( ACCUM ) o[jj] += acc[3 * i + j];
                        else o[jj] = acc[3 * i + j];
                        jj++;
                    }
            }
        };
        // one lane per unit (= column block, see kTet2Unit): its items in ascending element number
        for ( int u = lane; u < nu; u += 32 ) {
            double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            const int k0 = ustart[u], k1 = ustart[u + 1];
            for ( int k = k0; k < k1; k++ ) {
                const double *pk = s_park[w] + (int) items[k] * 9;
#pragma unroll
                for ( int q = 0; q < 9; q++ ) acc[q] += pk[q];
            }
            write_block(ublk[u] & 0x3F, acc);
        }
        __syncwarp();
    }
}

// ---- host side -----------------------------------------------------------------------------------

int tet_bind(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    S->rows_ok = false;
    if ( getenv("OB200_ASSEMBLY") && !strcmp(getenv("OB200_ASSEMBLY"), "slotmap") ) return OB200_OK;      // generic path (cross-checks)
    if ( S->etype != OB200_LTRSPACE || S->nelem == 0 || !S->ninc.p || A->neq == 0 || A->neq != S->neq ) return OB200_OK;
    if ( A->maxrow > 0xFFFF ) return OB200_OK;
    OB_CHECK( S->eqnode.alloc(A->neq) );
    OB_CUDA( cudaMemsetAsync(S->eqnode.p, 0xFF, sizeof( int32_t ) * (size_t) A->neq, ctx->stream) );
    OB_LAUNCH(ctx, tet_eqnode_kernel, ctx->shape.grid(S->nnode * 3, 256, 8), 256, 0, S->nnode, S->nodeeq.p, S->eqnode.p);
    DevBuf< int > flags;
    OB_CHECK( flags.alloc(4) );            // [0] mismatch, [1] capacity, [2..3] 64-bit count of covered entries
    OB_CUDA( cudaMemsetAsync(flags.p, 0, sizeof( int ) * 4, ctx->stream) );
    OB_LAUNCH(ctx, tet_rows_check_kernel, ctx->shape.grid(S->nnode * 32, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p,
              S->conn.p, S->nodeeq.p, S->eqnode.p, A->rowptr.p, A->colind.p, A->neq, flags.p,
              reinterpret_cast< unsigned long long * >( flags.p + 2 ));
    int h[4] = { 0, 0, 0, 0 };
    OB_CUDA( cudaMemcpyAsync(h, flags.p, sizeof( int ) * 4, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    S->rows_ok = ( h[0] == 0 && h[1] == 0 );
    unsigned long long cov;
    memcpy(&cov, h + 2, sizeof( cov ));
    S->covers_all = ( (int64_t) cov == A->nnz );
    if ( S->rows_ok && !S->trec.p ) {
        OB_CHECK( S->trec.alloc(S->nelem * kTetRec) );
        OB_LAUNCH(ctx, tet_geometry_kernel, ctx->shape.grid(S->nelem, 128, 8), 128, 0, S->view(), S->nelem, S->trec.p);
    }
    // the per-node tables of the fast kernel; meshes beyond its capacity (valence, blocks per row) keep ltrspace_rows_kernel
    S->tet_fast = false;
    const char *tk = getenv("OB200_TET_ROWS");
    if ( S->rows_ok && !( tk && !strcmp(tk, "search") ) ) {
        DevBuf< int32_t > tbytes;
        DevBuf< int64_t > s64;
        OB_CHECK( tbytes.alloc(S->nnode + 1) );
        OB_CHECK( s64.alloc(S->nnode + 1) );
        OB_CHECK( S->row_tstart.alloc(S->nnode + 1) );
        OB_CUDA( cudaMemsetAsync(tbytes.p, 0, sizeof( int32_t ) * ( S->nnode + 1 ), ctx->stream) );
        OB_CUDA( cudaMemsetAsync(flags.p, 0, sizeof( int ) * 4, ctx->stream) );
        const int grid = ctx->shape.grid(S->nnode * 32, 256, 8);
        OB_LAUNCH(ctx, tet_tables_kernel< false >, grid, 256, 0, S->nnode, S->ninc_start.p, S->ninc.p, S->conn.p, S->nodeeq.p, S->eqnode.p,
                  A->rowptr.p, A->colind.p, tbytes.p, (const int32_t *) nullptr, (int4 *) nullptr, (unsigned char *) nullptr, flags.p);
        int64_t total = 0;
        OB_CHECK( exclusive_scan(ctx, tbytes.p, s64.p, S->nnode + 1, &total) );
        OB_CUDA( cudaMemcpyAsync(h, flags.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        if ( h[0] == 0 && total < (int64_t) INT_MAX ) {
            OB_CHECK( narrow_i64_to_i32(ctx, s64.p, S->row_tstart.p, S->nnode + 1) );
            OB_CHECK( S->row_desc.alloc(S->nnode * 2 * 4) );
            OB_CHECK( S->row_vtab.alloc(total > 0 ? total : 4) );
            OB_LAUNCH(ctx, tet_tables_kernel< true >, grid, 256, 0, S->nnode, S->ninc_start.p, S->ninc.p, S->conn.p, S->nodeeq.p, S->eqnode.p,
                      A->rowptr.p, A->colind.p, tbytes.p, S->row_tstart.p, reinterpret_cast< int4 * >( S->row_desc.p ), S->row_vtab.p, flags.p);
            S->tet_fast = true;
        }
    }
    return OB200_OK;
}

int tet_assemble_ltrspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    const bool gen = !S->all_isole;
    if ( gen ) {
        if ( !S->tangent.p ) OB_CHECK( S->tangent.alloc(S->nelem * 36) );
        OB_LAUNCH(ctx, tet_tangent_kernel, ctx->shape.grid(S->nelem, 128, 8), 128, 0, S->view(), S->nelem, S->tangent.p);
    }
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    if ( S->tet_fast ) {
        TetRows2View V2{ S->nnode, reinterpret_cast< const int4 * >( S->row_desc.p ), S->row_vtab.p, S->ninc.p, S->trec.p, S->tangent.p };
        const int grid2 = ctx->shape.grid(S->nnode * 32, kTet2Warps * 32, 16);
        const int mode = ( A->zero_pending ? 0 : 1 ) | ( gen ? 2 : 0 );
        if ( mode == 0 ) OB_LAUNCH(ctx, ltrspace_rows2_kernel< 0 >, grid2, kTet2Warps * 32, 0, V2, A->val.p);
        else if ( mode == 1 ) OB_LAUNCH(ctx, ltrspace_rows2_kernel< 1 >, grid2, kTet2Warps * 32, 0, V2, A->val.p);
        else if ( mode == 2 ) OB_LAUNCH(ctx, ltrspace_rows2_kernel< 2 >, grid2, kTet2Warps * 32, 0, V2, A->val.p);
        else OB_LAUNCH(ctx, ltrspace_rows2_kernel< 3 >, grid2, kTet2Warps * 32, 0, V2, A->val.p);
        A->zero_pending = false;
        return OB200_OK;
    }
    TetRowsView V{ S->nnode, S->ninc_start.p, S->ninc.p, S->conn.p, S->nodeeq.p, S->eqnode.p, A->rowptr.p, A->colind.p, S->trec.p,
                   S->tangent.p };
    const int grid = ctx->shape.grid(S->nnode * 32, kTetWarps * 32, 8);
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    const bool accum = !A->zero_pending;           // else every entry of the pattern is written once: the pending zero() is absorbed
    if ( gen ) {
        if ( accum ) OB_LAUNCH(ctx, ltrspace_rows_kernel< 3 >, grid, kTetWarps * 32, 0, V, A->val.p);
        else OB_LAUNCH(ctx, ltrspace_rows_kernel< 2 >, grid, kTetWarps * 32, 0, V, A->val.p);
    } else {
        if ( accum ) OB_LAUNCH(ctx, ltrspace_rows_kernel< 1 >, grid, kTetWarps * 32, 0, V, A->val.p);
        else OB_LAUNCH(ctx, ltrspace_rows_kernel< 0 >, grid, kTetWarps * 32, 0, V, A->val.p);
    }
    A->zero_pending = false;
    return OB200_OK;
}

} // namespace ob200
