// Owner-computes ("gather") assembly of the LSpace tangent into the cudacsr matrix.
//
// Replaces EngngModel::assemble (src/core/engngm.C:889-929) + CompCol::assemble
// (src/core/compcol.C:263-299) for a whole element set without a single atomic: every matrix row
// is produced by exactly one warp and written once, coalesced -- bit-reproducible run to run.
//
//   * A "visit" is one (node, adjacent element) incidence.  A CTA first evaluates the geometry of
//     every DISTINCT element its visits touch, one thread per (element, Gauss point):
//         g_b = sqrt(dV) grad N_b,  b = 0..7   (FEI3dHexaLin::evaldNdx, fei3dhexalin.C:186-204)
//     into shared memory, then four threads per visit contract the eight 3x3 blocks K_ab (a = the
//     visited node, b = 0..7) of that element out of it:
//         G_ab = sum_gp g_a g_b^T,   K_ab = lambda G + mu G^T + mu tr(G) I
//     (IsotropicLinearElasticMaterial D, isolinearelasticmaterial.C:80-84, through the B matrix of
//     Structural3DElement::computeBmatrixAt, structural3delement.C:63-86) and park them in shared
//     memory at positions sorted by column.
//   * One warp per node then sums, for each neighbouring node, the parked blocks in a fixed order,
//     lays the node's (up to 3) rows out in shared memory and streams them to val.
//
// A CTA handles a "group": the nodes whose first visit index falls into one window of kGroupVisits
// visits.  Everything index-like (incidence lists, where each block is parked, the per-node list
// of column blocks) is computed once per mesh / per bound matrix by the kernels at the top.
#include "element_device.cuh"
#include "elemset.h"
#include "scan.cuh"
#include <limits.h>
#include <string.h>
#include <stdlib.h>

namespace ob200 {

#ifndef OB200_GROUP_VISITS
#define OB200_GROUP_VISITS 24
#endif
#ifndef OB200_ROUND_ELEMS
#define OB200_ROUND_ELEMS 16
#endif
#ifndef OB200_GATHER_CTAS
#define OB200_GATHER_CTAS 3
#endif
constexpr int kGroupVisits = OB200_GROUP_VISITS;     // visits per group window
constexpr int kMaxValence = 16;                      // elements around a node (fast path)
constexpr int kGatherVisits = kGroupVisits + kMaxValence;    // most visits a group can hold
constexpr int kGatherThreads = 128;                  // compute threads: one per (element, Gauss point), then four per visit
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kMaxRowLen = 128;                      // longest row the column-block schedule describes
constexpr int kBlkDoubles = 10;                      // a parked 3x3 block, padded to 80 B (16-byte aligned)
constexpr int kRoundElems = OB200_ROUND_ELEMS;                    // distinct elements whose gradients are resident at a time
constexpr int kGpStride = 26;                        // doubles per (element, Gauss point): 24 + pad (208 B: conflict-free 16-byte stores)
#ifndef OB200_EL_PAD
#define OB200_EL_PAD 8
#endif
constexpr int kElStride = 8 * kGpStride + OB200_EL_PAD;   // doubles per element: 1728 B = 64 mod 128, so the two visits a quarter-warp
                                                           // serves in phase A2 (elements 1 or 3 apart in the group's list) read disjoint banks (pad 2: 3.11 ms, pad 8: 2.93 ms)

// ---- mesh-only preprocessing (elemset create) ---------------------------------------------

__global__ void node_valence_kernel(const int32_t *__restrict__ conn, int64_t n, int64_t nnode, int32_t *__restrict__ cnt, int *__restrict__ bad)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        int node = conn[t] - 1;
        if ( node < 0 || node >= nnode ) atomicAdd(bad, 1);
        else atomicAdd(cnt + node, 1);
    }
}

__global__ void node_incidence_fill_kernel(const int32_t *__restrict__ conn, int64_t n, int nen, const int32_t *__restrict__ start,
                                           int32_t *__restrict__ fill, int32_t *__restrict__ ninc)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        int node = conn[t] - 1;
        int k = atomicAdd(fill + node, 1);
        ninc[start[node] + k] = (int32_t)( ( t / nen ) * 8 + ( t % nen ) );      // element*8 + local node index
    }
}

// sort each node's visits (ascending element number): the accumulation order must not depend on
// the order in which the atomics above happened to land
// vis[element * 8 + local node] = position of that incidence in the (sorted) node -> element lists
__global__ void visit_index_kernel(const int32_t *__restrict__ ninc, int64_t nvisit, int32_t *__restrict__ vis)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t p = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; p < nvisit; p += stride ) vis[ninc[p]] = (int32_t) p;
}

// eqcount[eq - 1] += 1 for every free nodal dof: an equation that belongs to two nodal dofs (master / slave) rules out the
// owner-computes vector assembly
__global__ void equation_owner_count_kernel(const int32_t *__restrict__ nodeeq, int64_t n, int32_t neq, int32_t *__restrict__ eqcount, int *__restrict__ bad)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        const int eq = nodeeq[t];
        if ( eq > neq ) atomicOr(bad, 1);
        else if ( eq > 0 && atomicAdd(eqcount + eq - 1, 1) > 0 ) atomicOr(bad, 1);
    }
}

__global__ void node_incidence_sort_kernel(int64_t nnode, const int32_t *__restrict__ start, int32_t *__restrict__ ninc)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t w = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; w < nnode; w += stride ) {
        const int b = start[w], e = start[w + 1];
        for ( int i = b + 1; i < e; i++ ) {
            int32_t v = ninc[i];
            int j = i - 1;
            while ( j >= b && ninc[j] > v ) { ninc[j + 1] = ninc[j]; j--; }
            ninc[j + 1] = v;
        }
    }
}

// parking base of every visit: 8 * (first visit of its node - first visit of the node's group)
__global__ void visit_base_kernel(int64_t nnode, const int32_t *__restrict__ start, const int2 *__restrict__ gtab, int32_t *__restrict__ vbase)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t w = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; w < nnode; w += stride ) {
        const int b = start[w], e = start[w + 1];
        if ( e == b ) continue;
        const int base = ( b - gtab[b / kGroupVisits].y ) * 8;
        for ( int i = b; i < e; i++ ) vbase[i] = base;
    }
}

// nodeeq[node][i] = equation number of dof i (0 = prescribed), read back from the location arrays
__global__ void node_equations_kernel(const int32_t *__restrict__ conn, const int32_t *__restrict__ loc, int64_t nelem, int nen,
                                      int32_t *__restrict__ nodeeq)
{
    const int64_t n = nelem * nen;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        int node = conn[t] - 1;
#pragma unroll
        for ( int i = 0; i < 3; i++ ) nodeeq[(int64_t) node * 3 + i] = loc[t * 3 + i];
    }
}

// vertex coordinates gathered per element: exyz[e][3*k + i] = coords[conn[e][k]][i]
__global__ void element_coords_kernel(const int32_t *__restrict__ conn, const double *__restrict__ coords, int64_t n, double *__restrict__ exyz)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        const double *c = coords + (int64_t)( conn[t] - 1 ) * 3;
        exyz[t * 3] = c[0];
        exyz[t * 3 + 1] = c[1];
        exyz[t * 3 + 2] = c[2];
    }
}

// group table (same construction as the SpMV chunk table): gtab[g] = {first node whose first visit
// index is >= g*kGroupVisits, that node's first visit index}
__global__ void group_table_kernel(int32_t nnode, const int32_t *__restrict__ start, int32_t ngroups, int2 *__restrict__ table)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r <= nnode; r += stride ) {
        int hi = start[r] / kGroupVisits;
        const int lo = r == 0 ? 0 : start[r - 1] / kGroupVisits + 1;
        if ( hi > ngroups - 1 ) hi = ngroups - 1;
        for ( int c = lo; c <= hi; c++ ) table[c] = make_int2((int) r, start[r]);
        if ( r == nnode )
            for ( int c = ( hi + 1 > lo ? hi + 1 : lo ); c <= ngroups; c++ ) table[c] = make_int2(nnode, start[nnode]);
    }
}

// Distinct elements of every group.  One warp per group: vu[visit] = rank of the visit's element
// among the distinct elements of its group (ascending element number), bit 7 set on the first
// visit of each distinct element (the producer warp builds the group's element list from those).
__global__ void __launch_bounds__(256)
visit_unique_kernel(int32_t ngroups, const int2 *__restrict__ gtab, const int32_t *__restrict__ ninc, unsigned char *__restrict__ vu)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    constexpr int Q = ( kGatherVisits + 31 ) / 32;
    for ( int64_t g = warp0; g < ngroups; g += nwarps ) {
        const int v0 = gtab[g].y, n = gtab[g + 1].y - v0;
        int el[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            const int it = q * 32 + lane;
            el[q] = it < n ? ( ninc[v0 + it] >> 3 ) : INT_MAX;
        }
        int first[Q], rank[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) { first[q] = 1; rank[q] = 0; }
        // pass 1: first occurrence?
        for ( int jt = 0; jt < n; jt++ ) {
            int sel = INT_MAX;
#pragma unroll
            for ( int qs = 0; qs < Q; qs++ ) if ( qs == ( jt >> 5 ) ) sel = el[qs];
            const int ej = __shfl_sync(0xffffffffu, sel, jt & 31);
#pragma unroll
            for ( int q = 0; q < Q; q++ ) if ( ej == el[q] && jt < q * 32 + lane ) first[q] = 0;
        }
        // pass 2: distinct smaller elements
        for ( int jt = 0; jt < n; jt++ ) {
            int sel = INT_MAX, self = 0;
#pragma unroll
            for ( int qs = 0; qs < Q; qs++ ) if ( qs == ( jt >> 5 ) ) { sel = el[qs]; self = first[qs]; }
            const int ej = __shfl_sync(0xffffffffu, sel, jt & 31), fj = __shfl_sync(0xffffffffu, self, jt & 31);
#pragma unroll
            for ( int q = 0; q < Q; q++ ) if ( fj && ej < el[q] ) rank[q]++;
        }
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            const int it = q * 32 + lane;
            if ( it < n ) vu[v0 + it] = (unsigned char)( rank[q] | ( first[q] ? 0x80 : 0 ) );
        }
    }
}

// ---- per bound matrix: where every block is parked, the column blocks of every node ---------
//
// One warp per node.  Items = (visit k, local node b) of the node's elements; key = first free
// equation of the neighbouring node conn[e_k][b].  Items sorted by (key, item index) give the
// parking position; each distinct key is one column block of the node's rows.
// flags[0]: a node's column blocks do not line up with the CSR row (free equations of a node not
// consecutive, or the pattern is not the one of this element set); flags[1]: capacity exceeded.
__global__ void __launch_bounds__(256)
node_blocks_allpairs_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                   const int32_t *__restrict__ conn, const int32_t *__restrict__ nodeeq,
                   const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int maxblk,
                   unsigned char *__restrict__ pos, unsigned char *__restrict__ nblk, unsigned short *__restrict__ blk,
                   int *__restrict__ flags, unsigned long long *__restrict__ covered)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    constexpr int Q = kMaxValence * 8 / 32;          // items per lane
    __shared__ int sm_key[8][Q * 32], sm_dkey[8][Q * 32], sm_dinfo[8][Q * 32], sm_dblk[8][Q * 32];
    for ( int64_t w = warp0; w < nnode; w += nwarps ) {
        const int v0 = ninc_start[w], nv = ninc_start[w + 1] - v0;
        const int nitems = nv * 8;
        if ( nv > kMaxValence ) {
            if ( lane == 0 ) atomicAdd(flags + 1, 1);
            continue;
        }
        int key[Q], cm[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            const int it = q * 32 + lane;
            key[q] = INT_MAX;
            cm[q] = 0;
            if ( it < nitems ) {
                const int e = ninc[v0 + ( it >> 3 )] >> 3;
                const int nb = conn[(int64_t) e * 8 + ( it & 7 )] - 1;
                const int e0 = nodeeq[(int64_t) nb * 3], e1 = nodeeq[(int64_t) nb * 3 + 1], e2 = nodeeq[(int64_t) nb * 3 + 2];
                cm[q] = ( e0 > 0 ? 1 : 0 ) | ( e1 > 0 ? 2 : 0 ) | ( e2 > 0 ? 4 : 0 );
                if ( cm[q] ) key[q] = e0 > 0 ? e0 : ( e1 > 0 ? e1 : e2 );
                // free equations of a node must be consecutive for its columns to form one block
                int prev = 0, ok = 1;
                if ( e0 > 0 ) prev = e0;
                if ( e1 > 0 ) { if ( prev && e1 != prev + 1 ) ok = 0; prev = e1; }
                if ( e2 > 0 ) { if ( prev && e2 != prev + 1 ) ok = 0; }
                if ( !ok ) atomicAdd(flags, 1);
            }
        }
        // the per-item data the passes below broadcast go through this warp's slice of shared memory
        int *skey = sm_key[threadIdx.x >> 5], *dkey = sm_dkey[threadIdx.x >> 5], *dinfo = sm_dinfo[threadIdx.x >> 5],
            *dblk = sm_dblk[threadIdx.x >> 5];
        __syncwarp();
#pragma unroll
        for ( int q = 0; q < Q; q++ ) skey[q * 32 + lane] = key[q];
        __syncwarp();
        // pass 1: rank inside my key (same key, earlier item) and size of my key
        int same_before[Q], same_total[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) same_before[q] = same_total[q] = 0;
        for ( int jt = 0; jt < nitems; jt++ ) {
            const int kj = skey[jt];
#pragma unroll
            for ( int q = 0; q < Q; q++ ) {
                same_total[q] += ( kj == key[q] );
                same_before[q] += ( kj == key[q] && jt < q * 32 + lane );
            }
        }
        // the distinct keys (= column blocks), compacted in item order
        int first[Q], width[Q], ndist = 0;
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            first[q] = ( same_before[q] == 0 && cm[q] != 0 ) ? 1 : 0;
            width[q] = __popc(cm[q]);
            const unsigned int m = __ballot_sync(0xffffffffu, first[q]);
            if ( first[q] ) {
                const int c = ndist + __popc(m & ( ( 1u << lane ) - 1u ));
                dkey[c] = key[q];
                dinfo[c] = width[q] | ( same_total[q] << 8 );
                first[q] = c + 1;                    // remember where this block sits in the compact list
            }
            ndist += __popc(m);
        }
        __syncwarp();
        // pass 2: distinct smaller keys (= my column block) and their total width (= my first column)
        int blkidx[Q], cstart[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) blkidx[q] = cstart[q] = 0;
        for ( int jd = 0; jd < ndist; jd++ ) {
            const int kj = dkey[jd], wj = dinfo[jd] & 0xFF;
#pragma unroll
            for ( int q = 0; q < Q; q++ )
                if ( kj < key[q] ) { blkidx[q]++; cstart[q] += wj; }
        }
#pragma unroll
        for ( int q = 0; q < Q; q++ )
            if ( first[q] ) dblk[first[q] - 1] = blkidx[q];
        __syncwarp();
        // pass 3: parking position.  Blocks are handled 32 at a time (one per lane) by the assembly
        // kernel, which walks "the i-th item of every block that has one" for i = 0, 1, ...; items are
        // parked in exactly that order (jagged-diagonal order) so that each step reads contiguous slots.
        int park[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) park[q] = 0;
        for ( int jd = 0; jd < ndist; jd++ ) {
            const int bj = dblk[jd], cj = dinfo[jd] >> 8;
#pragma unroll
            for ( int q = 0; q < Q; q++ ) {
                const int i = same_before[q];
                if ( ( bj >> 5 ) < ( blkidx[q] >> 5 ) ) park[q] += cj;
                else if ( ( bj >> 5 ) == ( blkidx[q] >> 5 ) ) park[q] += min(cj, i) + ( ( bj < blkidx[q] && cj > i ) ? 1 : 0 );
            }
        }
        // the node's own rows (any free dof; all of them share the pattern)
        const int r0 = nodeeq[w * 3], r1 = nodeeq[w * 3 + 1], r2 = nodeeq[w * 3 + 2];
        const int row = r0 > 0 ? r0 - 1 : ( r1 > 0 ? r1 - 1 : ( r2 > 0 ? r2 - 1 : -1 ) );
        int nb_max = 0, width_total = 0;
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            const int it = q * 32 + lane;
            if ( it < nitems ) {
                const bool live = cm[q] != 0 && row >= 0;
                pos[( (int64_t) v0 + ( it >> 3 ) ) * 8 + ( it & 7 )] = (unsigned char)( live ? park[q] : 0xFF );
                if ( live ) {
                    if ( park[q] >= 0xFF || blkidx[q] >= maxblk || same_total[q] > 0xFF ) atomicAdd(flags + 1, 1);
                    nb_max = max(nb_max, blkidx[q] + 1);
                    if ( first[q] ) {
                        width_total += width[q];
                        if ( blkidx[q] < maxblk ) {
                            // one entry per column block: number of parked items, free-dof mask of the column node
                            blk[w * maxblk + blkidx[q]] = (unsigned short)( same_total[q] | ( cm[q] << 8 ) );
                            if ( colind[rowptr[row] + cstart[q]] != key[q] - 1 ) atomicAdd(flags, 1);
                        }
                    }
                }
            }
        }
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) {
            nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, o));
            width_total += __shfl_xor_sync(0xffffffffu, width_total, o);
        }
        if ( row >= 0 ) {
            if ( width_total != rowptr[row + 1] - rowptr[row] || width_total > kMaxRowLen ) {
                if ( lane == 0 ) atomicAdd(flags + ( width_total > kMaxRowLen ? 1 : 0 ), 1);
            }
        }
        if ( lane == 0 ) {
            nblk[w] = (unsigned char)( row >= 0 ? nb_max : 0 );
            // matrix entries this node's rows account for (to know whether the set covers the whole pattern)
            if ( row >= 0 ) atomicAdd(covered, (unsigned long long) width_total * ( ( r0 > 0 ) + ( r1 > 0 ) + ( r2 > 0 ) ));
        }
    }
}

// The same schedule, sort-based: the warp sorts the node's items by (key, item index) with a bitonic network
// in registers (four items per lane), which turns the ranks of the all-pairs version above into scans:
// same_before = position - start of the run, column block = number of run starts before.  Produces exactly
// the arrays of node_blocks_allpairs_kernel (kept as the cross-check, OB200_NODE_BLOCKS=allpairs) in ~40 % of
// the instructions.
__global__ void __launch_bounds__(256)
node_blocks_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                   const int32_t *__restrict__ conn, const int32_t *__restrict__ nodeeq,
                   const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int maxblk,
                   unsigned char *__restrict__ pos, unsigned char *__restrict__ nblk, unsigned short *__restrict__ blk,
                   int *__restrict__ flags, unsigned long long *__restrict__ covered, unsigned char *__restrict__ ebidx)
{
    const int lane = threadIdx.x & 31, ws = threadIdx.x >> 5;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    constexpr int Q = 4, N = 128, NCH = N / 32;
    static_assert( kMaxValence * 8 == N, "four items per lane" );
    __shared__ int s_cnt[8][N], s_info[8][N], s_cstart[8][N], s_nacc[8][NCH][kMaxValence], s_cbase[8][NCH + 1];
    __shared__ unsigned int s_ball[8][NCH][kMaxValence];
    typedef unsigned long long u64;
    for ( int64_t w = warp0; w < nnode; w += nwarps ) {
        const int v0 = ninc_start[w], nv = ninc_start[w + 1] - v0;
        const int nitems = nv * 8;
        if ( nv > kMaxValence ) {
            if ( lane == 0 ) atomicAdd(flags + 1, 1);
            continue;
        }
        // composite sort key: equation number of the column node's first free dof | item index | free-dof mask
        u64 v[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            const int it = q * 32 + lane;
            int key = INT_MAX, cm = 0;
            if ( it < nitems ) {
                const int e = ninc[v0 + ( it >> 3 )] >> 3;
                const int nb = conn[(int64_t) e * 8 + ( it & 7 )] - 1;
                const int e0 = nodeeq[(int64_t) nb * 3], e1 = nodeeq[(int64_t) nb * 3 + 1], e2 = nodeeq[(int64_t) nb * 3 + 2];
                cm = ( e0 > 0 ? 1 : 0 ) | ( e1 > 0 ? 2 : 0 ) | ( e2 > 0 ? 4 : 0 );
                if ( cm ) key = e0 > 0 ? e0 : ( e1 > 0 ? e1 : e2 );
                // free equations of a node must be consecutive for its columns to form one block
                int prev = 0, ok = 1;
                if ( e0 > 0 ) prev = e0;
                if ( e1 > 0 ) { if ( prev && e1 != prev + 1 ) ok = 0; prev = e1; }
                if ( e2 > 0 ) { if ( prev && e2 != prev + 1 ) ok = 0; }
                if ( !ok ) atomicAdd(flags, 1);
            }
            v[q] = ( (u64) (unsigned int) key << 16 ) | ( (u64) it << 8 ) | (u64) cm;
        }
        // bitonic sort, ascending; element e = 4 * lane + q
#define OB_CE(a, b, up)                           \
    {                                             \
        const u64 lo__ = a < b ? a : b, hi__ = a < b ? b : a; \
        a = ( up ) ? lo__ : hi__;                 \
        b = ( up ) ? hi__ : lo__;                 \
    }
#pragma unroll
        for ( int k = 2; k <= N; k <<= 1 ) {
#pragma unroll
            for ( int j = k >> 1; j > 0; j >>= 1 ) {
                if ( j >= 4 ) {
                    const int lj = j >> 2;
                    const bool lower = ( lane & lj ) == 0;
                    const bool up = ( ( lane * 4 ) & k ) == 0;
#pragma unroll
                    for ( int q = 0; q < Q; q++ ) {
                        const u64 o = __shfl_xor_sync(0xffffffffu, v[q], lj);
                        const u64 mn = v[q] < o ? v[q] : o, mx = v[q] < o ? o : v[q];
                        v[q] = ( lower == up ) ? mn : mx;
                    }
                } else if ( j == 2 ) {
                    const bool up = ( ( lane * 4 ) & k ) == 0;
                    OB_CE(v[0], v[2], up);
                    OB_CE(v[1], v[3], up);
                } else {
                    const bool up01 = ( ( lane * 4 ) & k ) == 0, up23 = ( ( lane * 4 + 2 ) & k ) == 0;
                    OB_CE(v[0], v[1], up01);
                    OB_CE(v[2], v[3], up23);
                }
            }
        }
#undef OB_CE
        int key[Q], item[Q], cm[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            key[q] = (int)( v[q] >> 16 );
            item[q] = (int)( ( v[q] >> 8 ) & 0xFF );
            cm[q] = (int)( v[q] & 0xFF );
        }
        // runs of equal keys
        int prevk = __shfl_up_sync(0xffffffffu, key[Q - 1], 1);
        if ( lane == 0 ) prevk = -1;
        bool start[Q], valid[Q];
        int rs[Q], bi[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            start[q] = key[q] != ( q == 0 ? prevk : key[q - 1] );
            valid[q] = key[q] != INT_MAX;
        }
        {   // start of my run (inclusive max-scan of the start positions), my column block (inclusive count of starts - 1)
            int loc = -1, cntl = 0;
#pragma unroll
            for ( int q = 0; q < Q; q++ ) {
                if ( start[q] ) loc = lane * 4 + q;
                if ( start[q] && valid[q] ) cntl++;
                rs[q] = loc;
                bi[q] = cntl;
            }
            int incl = loc, csum = cntl;
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o), u = __shfl_up_sync(0xffffffffu, csum, o);
                if ( lane >= o ) { incl = max(incl, t); csum += u; }
            }
            int cin = __shfl_up_sync(0xffffffffu, incl, 1), bin = __shfl_up_sync(0xffffffffu, csum, 1);
            if ( lane == 0 ) { cin = -1; bin = 0; }
#pragma unroll
            for ( int q = 0; q < Q; q++ ) {
                rs[q] = max(rs[q], cin);
                bi[q] = bin + bi[q] - 1;
            }
        }
        const int ndist = __shfl_sync(0xffffffffu, bi[Q - 1] + 1, 31);        // valid run starts in the whole warp
        int *cnt = s_cnt[ws], *info = s_info[ws], *cst = s_cstart[ws];
        __syncwarp();
#pragma unroll
        for ( int q = 0; q < Q; q++ ) cnt[q * 32 + lane] = 0;
        __syncwarp();
        int sb[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            sb[q] = lane * 4 + q - rs[q];
            if ( valid[q] ) {
                atomicMax(&cnt[bi[q]], sb[q] + 1);
                if ( start[q] ) info[bi[q]] = __popc(cm[q]) | ( cm[q] << 8 );
            }
        }
        __syncwarp();
        // per column block (lane = block, 32 at a time): first column, and the jagged-diagonal bookkeeping
        int wcarry = 0, ccarry = 0;
        for ( int c = 0; c * 32 < ndist; c++ ) {
            const int b = c * 32 + lane;
            const int cb = b < ndist ? cnt[b] : 0, wb = b < ndist ? ( info[b] & 0xFF ) : 0;
            int incl = wb;
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ( lane >= o ) incl += t;
            }
            if ( b < ndist ) cst[b] = wcarry + incl - wb;
            wcarry += __shfl_sync(0xffffffffu, incl, 31);
            int run = 0;
            for ( int i = 0; i < kMaxValence; i++ ) {
                const unsigned int m = __ballot_sync(0xffffffffu, cb > i);
                if ( lane == 0 ) {
                    s_ball[ws][c][i] = m;
                    s_nacc[ws][c][i] = run;
                }
                run += __popc(m);
                if ( m == 0 ) break;
            }
            if ( lane == 0 ) s_cbase[ws][c] = ccarry;
            ccarry += run;
        }
        __syncwarp();
        // the node's own rows (any free dof; all of them share the pattern)
        const int r0 = nodeeq[w * 3], r1 = nodeeq[w * 3 + 1], r2 = nodeeq[w * 3 + 2];
        const int row = r0 > 0 ? r0 - 1 : ( r1 > 0 ? r1 - 1 : ( r2 > 0 ? r2 - 1 : -1 ) );
#pragma unroll
        for ( int q = 0; q < Q; q++ ) {
            if ( item[q] >= nitems ) continue;
            const bool live = valid[q] && row >= 0;
            int park = 0xFF;
            if ( live ) {
                const int B = bi[q], c = B >> 5, i = sb[q];
                park = s_cbase[ws][c] + s_nacc[ws][c][i] + __popc(s_ball[ws][c][i] & ( ( 1u << ( B & 31 ) ) - 1u ));
                const int cb = cnt[B];
                if ( park >= 0xFF || B >= maxblk || cb > 0xFF ) atomicAdd(flags + 1, 1);
                if ( start[q] && B < maxblk ) {
                    // one entry per column block: number of parked items, free-dof mask of the column node
                    blk[w * maxblk + B] = (unsigned short)( cb | ( cm[q] << 8 ) );
                    if ( colind[rowptr[row] + cst[B]] != key[q] - 1 ) atomicAdd(flags, 1);
                }
            }
            if ( pos ) pos[( (int64_t) v0 + ( item[q] >> 3 ) ) * 8 + ( item[q] & 7 )] = (unsigned char) park;
            if ( ebidx ) {
                // element-major copy for the cluster assembly (assemble_cluster.cu): index of the column block of local
                // node b in the block list of local node a's rows, 0xFF if there is no such block
                const int vis = ninc[v0 + ( item[q] >> 3 )];
                ebidx[(int64_t) vis * 8 + ( item[q] & 7 )] = (unsigned char)( live && bi[q] < 0xFF ? bi[q] : 0xFF );
            }
        }
        const int nb_max = row >= 0 ? ndist : 0, width_total = row >= 0 ? wcarry : 0;
        if ( row >= 0 ) {
            if ( width_total != rowptr[row + 1] - rowptr[row] || width_total > kMaxRowLen ) {
                if ( lane == 0 ) atomicAdd(flags + ( width_total > kMaxRowLen ? 1 : 0 ), 1);
            }
        }
        if ( lane == 0 ) {
            nblk[w] = (unsigned char) nb_max;
            // matrix entries this node's rows account for (to know whether the set covers the whole pattern)
            if ( row >= 0 ) atomicAdd(covered, (unsigned long long) width_total * ( ( r0 > 0 ) + ( r1 > 0 ) + ( r2 > 0 ) ));
        }
    }
}

// ---- the assembly kernel ------------------------------------------------------------------------

struct GatherView {
    const int32_t *ninc_start, *ninc, *ninc_node, *nodeeq;
    const unsigned char *pos, *nblk, *vu;
    const unsigned short *blk;
    const int2 *gtab;
    int maxblk;
};

// one pipeline stage: everything phase A needs for one group, written by the producer warp
struct GatherStage {
    double xyz[kGatherVisits][26];        // vertex coordinates of the group's distinct elements (row padded to 208 B)
    double lam[kGatherVisits], mu[kGatherVisits];     // per visit
    uint2 pos[kGatherVisits];             // parking positions of the visit's 8 blocks
    int ent[kGatherVisits];               // local node index | parking base << 3
    int uel[kGatherVisits];               // the distinct elements
    unsigned char vu[kGatherVisits];      // visit -> index into uel
    int4 meta;                            // first node, first visit, end node, end visit
    int nuniq;
};
struct GatherShared {
    GatherStage st[2];
    double park[kGatherVisits * 8 * kBlkDoubles];
    double grad[kRoundElems * kElStride];                      // sqrt(dV) grad N_b per (element, Gauss point)
    unsigned long long full[2], empty[2];
};

__device__ __forceinline__ void mbar_init_(unsigned long long *bar, int count)
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( (uint32_t) __cvta_generic_to_shared(bar) ), "r"( count ) );
}
__device__ __forceinline__ void mbar_arrive_(unsigned long long *bar)
{
    asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( (uint32_t) __cvta_generic_to_shared(bar) ) : "memory" );
}
// the producer's wait: it runs a whole group ahead, so it polls with a back-off instead of taking issue
// slots from the compute warps
__device__ __forceinline__ void mbar_wait_relaxed_(unsigned long long *bar, uint32_t parity)
{
    uint32_t done = 0;
    for ( ;; ) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"( done ) : "r"( (uint32_t) __cvta_generic_to_shared(bar) ), "r"( parity ) : "memory" );
        if ( done ) break;
        __nanosleep(256);
    }
}
__device__ __forceinline__ void mbar_wait_(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"( (uint32_t) __cvta_generic_to_shared(bar) ), "r"( parity ) : "memory" );
}

// Geometry of one element at one Gauss point: out[3 b + j] = sqrt(dV) dN_b/dx_j, b = 0..7.
// FEI3dHexaLin::evaldNdxi (fei3dhexalin.C:129-166) at the 2x2x2 rule (gaussintegrationrule.C:190-214,
// weights 1): dN_k/dxi = s_k^xi (1 + s_k^eta eta)(1 + s_k^zeta zeta) / 8, ...  Every factor takes one of
// two values at a Gauss point, so the twelve products are formed once and each node picks its own at
// compile time.  J = x dN/dxi, dN/dx = dN/dxi J^-1 (evaldNdx, fei3dhexalin.C:186-204); dV = |det J|
// (Structural3DElement::computeVolumeAround, structural3delement.C:328-338).
__device__ __forceinline__ void hexa_gradients(const double *__restrict__ xv, int gp, double *__restrict__ out)
{
    constexpr double kA = 0.577350269189626;
    const bool gu = ( gp & 4 ) != 0, gv = ( gp & 2 ) != 0, gw = ( gp & 1 ) != 0;
    // f?[1] = 1 + coordinate, f?[0] = 1 - coordinate
    double fx[2], fy[2], fz[2];
    fx[1] = gu ? 1.0 + kA : 1.0 - kA; fx[0] = gu ? 1.0 - kA : 1.0 + kA;
    fy[1] = gv ? 1.0 + kA : 1.0 - kA; fy[0] = gv ? 1.0 - kA : 1.0 + kA;
    fz[1] = gw ? 1.0 + kA : 1.0 - kA; fz[0] = gw ? 1.0 - kA : 1.0 + kA;
    double Pyz[2][2], Pxz[2][2], Pxy[2][2];
#pragma unroll
    for ( int s = 0; s < 2; s++ )
#pragma unroll
        for ( int t = 0; t < 2; t++ ) {
            Pyz[s][t] = 0.125 * fy[s] * fz[t];
            Pxz[s][t] = 0.125 * fx[s] * fz[t];
            Pxy[s][t] = 0.125 * fx[s] * fy[t];
        }
    double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    const double2 *xv2 = reinterpret_cast< const double2 * >( xv );
    double c[24];
#pragma unroll
    for ( int i = 0; i < 12; i++ ) {
        const double2 v = xv2[i];
        c[2 * i] = v.x;
        c[2 * i + 1] = v.y;
    }
#pragma unroll
    for ( int kk = 0; kk < 8; kk++ ) {
        // signs of node kk in (xi, eta, zeta) -- the node order of FEI3dHexaLin
        const int px = ( kk & 3 ) >= 2, py = ( ( kk & 3 ) == 1 || ( kk & 3 ) == 2 ), pz = kk < 4;
        const double d0 = px ? Pyz[py][pz] : -Pyz[py][pz];
        const double d1 = py ? Pxz[px][pz] : -Pxz[px][pz];
        const double d2 = pz ? Pxy[px][py] : -Pxy[px][py];
        const double x = c[3 * kk], y = c[3 * kk + 1], z = c[3 * kk + 2];
        J[0][0] += x * d0; J[0][1] += x * d1; J[0][2] += x * d2;
        J[1][0] += y * d0; J[1][1] += y * d1; J[1][2] += y * d2;
        J[2][0] += z * d0; J[2][1] += z * d1; J[2][2] += z * d2;
    }
    double Ji[3][3];
    const double sq = sqrt(fabs(inv3(J, Ji)));
#pragma unroll
    for ( int i = 0; i < 3; i++ )
#pragma unroll
        for ( int j = 0; j < 3; j++ ) Ji[i][j] *= sq;
    double g[24];
#pragma unroll
    for ( int kk = 0; kk < 8; kk++ ) {
        const int px = ( kk & 3 ) >= 2, py = ( ( kk & 3 ) == 1 || ( kk & 3 ) == 2 ), pz = kk < 4;
        const double d0 = px ? Pyz[py][pz] : -Pyz[py][pz];
        const double d1 = py ? Pxz[px][pz] : -Pxz[px][pz];
        const double d2 = pz ? Pxy[px][py] : -Pxy[px][py];
#pragma unroll
        for ( int j = 0; j < 3; j++ ) g[3 * kk + j] = d0 * Ji[0][j] + d1 * Ji[1][j] + d2 * Ji[2][j];
    }
    double2 *o = reinterpret_cast< double2 * >( out );
#pragma unroll
    for ( int i = 0; i < 12; i++ ) o[i] = make_double2(g[2 * i], g[2 * i + 1]);
}

// Persistent, warp-specialised: warp 4 is the producer -- it walks this CTA's groups one ahead of
// the compute warps and resolves the dependent index chain (group table -> visit -> element list ->
// coordinates, material) into a shared-memory stage; warps 0-3 never wait on that chain.
template< bool ACCUM >
__global__ void __launch_bounds__(kGatherThreads + 32, OB200_GATHER_CTAS)
lspace_gather_kernel(ElemSetView S, GatherView G, int32_t ngroups, const int32_t *__restrict__ rowptr, double *__restrict__ val)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GatherShared &sh = *reinterpret_cast< GatherShared * >( smem_raw );
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if ( tid == 0 ) {
#pragma unroll
        for ( int s = 0; s < 2; s++ ) {
            mbar_init_(&sh.full[s], 1);
            mbar_init_(&sh.empty[s], kGatherWarps);
        }
        asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    __syncthreads();
    const int nmine = ( ngroups > (int) blockIdx.x ) ? ( ngroups - 1 - (int) blockIdx.x ) / (int) gridDim.x + 1 : 0;

    if ( wid == kGatherWarps ) {
        // ---- producer ----
        // Software-pipelined over groups: the visit records of group k+1 (and the table entries of
        // group k+2) are requested while the element -> coordinate chain of group k resolves.
        auto stage_visit = [&](GatherStage &st, int t, int ent, int pbase, uint2 pw, int vub) {
            const int64_t e = ent >> 3;
            const MatParams *mp = S.mat + S.matid[e];
            double lam, mu;
            isole_lame(mp->E, mp->nu, lam, mu);
            st.lam[t] = lam;
            st.mu[t] = mu;
            st.pos[t] = pw;
            st.ent[t] = ( ent & 7 ) | ( pbase << 3 );
            st.vu[t] = (unsigned char)( vub & 0x7F );
            if ( vub & 0x80 ) st.uel[vub & 0x7F] = (int) e;
        };
        int2 g0 = make_int2(0, 0), g1 = g0, h0 = g0, h1 = g0;      // table entries of group k, k+1
        int ent = 0, pbase = 0, vub = 0;
        uint2 pw = make_uint2(0, 0);
        if ( nmine > 0 ) {
            g0 = G.gtab[blockIdx.x];
            g1 = G.gtab[blockIdx.x + 1];
            if ( lane < g1.y - g0.y ) {
                ent = G.ninc[g0.y + lane];
                pbase = G.ninc_node[g0.y + lane];
                vub = G.vu[g0.y + lane];
                pw = *reinterpret_cast< const uint2 * >( G.pos + (int64_t)( g0.y + lane ) * 8 );
            }
        }
        if ( nmine > 1 ) {
            h0 = G.gtab[blockIdx.x + gridDim.x];
            h1 = G.gtab[blockIdx.x + gridDim.x + 1];
        }
        for ( int k = 0; k < nmine; k++ ) {
            const int s = k & 1;
            // requests for the groups behind this one
            int2 n0 = h0, n1 = h1;
            int nent = 0, npbase = 0, nvub = 0;
            uint2 npw = make_uint2(0, 0);
            if ( k + 1 < nmine && lane < h1.y - h0.y ) {
                nent = G.ninc[h0.y + lane];
                npbase = G.ninc_node[h0.y + lane];
                nvub = G.vu[h0.y + lane];
                npw = *reinterpret_cast< const uint2 * >( G.pos + (int64_t)( h0.y + lane ) * 8 );
            }
            if ( k + 2 < nmine ) {
                const int g = blockIdx.x + ( k + 2 ) * gridDim.x;
                n0 = G.gtab[g];
                n1 = G.gtab[g + 1];
            }
            if ( k >= 2 ) mbar_wait_relaxed_(&sh.empty[s], ( ( k >> 1 ) - 1 ) & 1);
            GatherStage &st = sh.st[s];
            const int nvis = g1.y - g0.y;
            int nfirst = 0;
            if ( lane < nvis ) {
                stage_visit(st, lane, ent, pbase, pw, vub);
                nfirst = ( vub >> 7 ) & 1;
            }
            for ( int t = lane + 32; t < nvis; t += 32 ) {           // the few visits beyond the first 32
                const int64_t v = (int64_t) g0.y + t;
                const int vb = G.vu[v];
                stage_visit(st, t, G.ninc[v], G.ninc_node[v], *reinterpret_cast< const uint2 * >( G.pos + v * 8 ), vb);
                nfirst += ( vb >> 7 ) & 1;
            }
#pragma unroll
            for ( int o = 16; o > 0; o >>= 1 ) nfirst += __shfl_xor_sync(0xffffffffu, nfirst, o);
            __syncwarp();
            // coordinates of the distinct elements: 12 chunks of 16 B per element, six coalesced requests in flight per lane
            const int nchunk = nfirst * 12;
            for ( int c0 = 0; c0 < nchunk; c0 += 6 * 32 ) {
                double2 buf[6];
#pragma unroll
                for ( int r = 0; r < 6; r++ ) {
                    const int c = c0 + r * 32 + lane, u = c / 12, part = c - u * 12;
                    if ( c < nchunk ) buf[r] = reinterpret_cast< const double2 * >( S.exyz + (int64_t) st.uel[u] * 24 )[part];
                }
#pragma unroll
                for ( int r = 0; r < 6; r++ ) {
                    const int c = c0 + r * 32 + lane, u = c / 12, part = c - u * 12;
                    if ( c < nchunk ) reinterpret_cast< double2 * >( st.xyz[u] )[part] = buf[r];
                }
            }
            if ( lane == 0 ) {
                st.meta = make_int4(g0.x, g0.y, g1.x, g1.y);
                st.nuniq = nfirst;
            }
            __syncwarp();
            if ( lane == 0 ) mbar_arrive_(&sh.full[s]);
            g0 = h0; g1 = h1; h0 = n0; h1 = n1;
            ent = nent; pbase = npbase; pw = npw; vub = nvub;
        }
        return;
    }

    // ---- compute warps ----
    for ( int k = 0; k < nmine; k++ ) {
        const int s = k & 1;
        const GatherStage &st = sh.st[s];
        mbar_wait_(&sh.full[s], ( k >> 1 ) & 1);
        const int4 meta = st.meta;
        const int n0 = meta.x, v0 = meta.y, n1 = meta.z, nvis = meta.w - meta.y;
        const int nuniq = st.nuniq;

        // metadata of the node this warp will reduce first: requested now, consumed after the barrier
        int b_nb = 0, b_info = 0, b_start = 0, b_eq[3] = { 0, 0, 0 }, b_row[3] = { 0, 0, 0 };
        if ( n0 + wid < n1 ) {
            const int w = n0 + wid;
            b_nb = G.nblk[w];
            b_start = G.ninc_start[w];
            if ( lane < b_nb ) b_info = G.blk[(int64_t) w * G.maxblk + lane];
#pragma unroll
            for ( int i = 0; i < 3; i++ ) b_eq[i] = G.nodeeq[(int64_t) w * 3 + i];
        }

        // ---- phase A, kRoundElems distinct elements at a time ----
        for ( int r0 = 0; r0 < nuniq; r0 += kRoundElems ) {
            const int nu = min(kRoundElems, nuniq - r0);
            // A1: one thread per (element, Gauss point): sqrt(dV) grad N_b, b = 0..7, into shared memory
            for ( int t = tid; t < nu * 8; t += kGatherThreads )
                hexa_gradients(st.xyz[r0 + ( t >> 3 )], t & 7, sh.grad + ( t >> 3 ) * kElStride + ( t & 7 ) * kGpStride);
            asm volatile( "bar.sync 1, %0;" ::"n"( kGatherThreads ) : "memory" );
            // A2: four threads per visit; thread q contracts the blocks K_ab, b = 2q, 2q+1, over the Gauss points
            for ( int t = tid >> 2; t < nvis; t += kGatherThreads / 4 ) {
                const int u = (int) st.vu[t] - r0;
                if ( (unsigned) u >= (unsigned) nu ) continue;
                const int q = tid & 3;
                const uint2 pw = st.pos[t];
                const unsigned int pq = ( ( q < 2 ? pw.x : pw.y ) >> ( 16 * ( q & 1 ) ) ) & 0xFFFFu;
                if ( pq == 0xFFFFu ) continue;              // both blocks belong to prescribed columns (or the row is prescribed)
                const int entw = st.ent[t];
                const int la = entw & 7;
                const double *ge = sh.grad + u * kElStride;
                double acc[2][9];
#pragma unroll
                for ( int bb = 0; bb < 2; bb++ )
#pragma unroll
                    for ( int kk = 0; kk < 9; kk++ ) acc[bb][kk] = 0.0;
#pragma unroll
                for ( int gp = 0; gp < 8; gp++ ) {
                    const double *gg = ge + gp * kGpStride;
                    const double a0 = gg[3 * la], a1 = gg[3 * la + 1], a2 = gg[3 * la + 2];
                    const double2 *gb = reinterpret_cast< const double2 * >( gg + 6 * q );
                    const double2 b0 = gb[0], b1 = gb[1], b2 = gb[2];
                    acc[0][0] += a0 * b0.x; acc[0][1] += a0 * b0.y; acc[0][2] += a0 * b1.x;
                    acc[0][3] += a1 * b0.x; acc[0][4] += a1 * b0.y; acc[0][5] += a1 * b1.x;
                    acc[0][6] += a2 * b0.x; acc[0][7] += a2 * b0.y; acc[0][8] += a2 * b1.x;
                    acc[1][0] += a0 * b1.y; acc[1][1] += a0 * b2.x; acc[1][2] += a0 * b2.y;
                    acc[1][3] += a1 * b1.y; acc[1][4] += a1 * b2.x; acc[1][5] += a1 * b2.y;
                    acc[1][6] += a2 * b1.y; acc[1][7] += a2 * b2.x; acc[1][8] += a2 * b2.y;
                }
                const double lam = st.lam[t], mu = st.mu[t];
                const int base = entw >> 3;
#pragma unroll
                for ( int bb = 0; bb < 2; bb++ ) {
                    const unsigned int p = ( pq >> ( 8 * bb ) ) & 0xFFu;
                    if ( p == 0xFFu ) continue;
                    const double *g = acc[bb];
                    const double tr = mu * ( g[0] + g[4] + g[8] );
                    double2 *o = reinterpret_cast< double2 * >( sh.park + (size_t)( base + p ) * kBlkDoubles );
                    o[0] = make_double2(lam * g[0] + mu * g[0] + tr, lam * g[1] + mu * g[3]);
                    o[1] = make_double2(lam * g[2] + mu * g[6], lam * g[3] + mu * g[1]);
                    o[2] = make_double2(lam * g[4] + mu * g[4] + tr, lam * g[5] + mu * g[7]);
                    o[3] = make_double2(lam * g[6] + mu * g[2], lam * g[7] + mu * g[5]);
                    o[4].x = lam * g[8] + mu * g[8] + tr;
                }
            }
            if ( r0 + kRoundElems < nuniq ) asm volatile( "bar.sync 1, %0;" ::"n"( kGatherThreads ) : "memory" );
        }
        // second level of the phase-B metadata (its address arrived during phase A)
#pragma unroll
        for ( int i = 0; i < 3; i++ )
            if ( b_eq[i] > 0 ) b_row[i] = rowptr[b_eq[i] - 1];
        // the stage has been consumed: hand it back to the producer
        __syncwarp();
        if ( lane == 0 ) mbar_arrive_(&sh.empty[s]);
        asm volatile( "bar.sync 1, %0;" ::"n"( kGatherThreads ) : "memory" );

        // ---- phase B: one warp per node: sum the parked blocks per column block, write the rows ----
        for ( int w = n0 + wid; w < n1; w += kGatherWarps ) {
            const bool first = ( w == n0 + wid );
            const int nb = first ? b_nb : G.nblk[w];
            if ( nb == 0 ) continue;
            int eqs[3], rowbase[3];
#pragma unroll
            for ( int i = 0; i < 3; i++ ) {
                eqs[i] = first ? b_eq[i] : G.nodeeq[(int64_t) w * 3 + i];
                rowbase[i] = first ? b_row[i] : ( eqs[i] > 0 ? rowptr[eqs[i] - 1] : 0 );
            }
            const int vbase = ( ( first ? b_start : G.ninc_start[w] ) - v0 ) * 8;
            int ccarry = 0, slot0 = vbase;
            for ( int nn0 = 0; nn0 < nb; nn0 += 32 ) {
                const int n = nn0 + lane;
                int info = 0;
                if ( first && nn0 == 0 ) info = b_info;
                else if ( n < nb ) info = G.blk[(int64_t) w * G.maxblk + n];
                const int cnt = info & 0xFF, cm = info >> 8;
                const int width = __popc(cm);
                int incl = width;                                   // inclusive scan of the widths
#pragma unroll
                for ( int o = 1; o < 32; o <<= 1 ) {
                    int tsh = __shfl_up_sync(0xffffffffu, incl, o);
                    if ( lane >= o ) incl += tsh;
                }
                const int cstart = ccarry + incl - width;
                ccarry += __shfl_sync(0xffffffffu, incl, 31);
                // step i sums the i-th parked item of every block that has one; the items of a step sit in
                // consecutive slots (jagged-diagonal order fixed by node_blocks_kernel)
                double kb[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
                for ( int i = 0;; i++ ) {
                    const bool active = cnt > i;
                    const unsigned int mask = __ballot_sync(0xffffffffu, active);
                    if ( mask == 0 ) break;
                    if ( active ) {
                        const int slot = slot0 + __popc(mask & ( ( 1u << lane ) - 1u ));
                        const double2 *sp = reinterpret_cast< const double2 * >( sh.park + (size_t) slot * kBlkDoubles );
                        const double2 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3];
                        const double s4 = sp[4].x;
                        kb[0] += s0.x; kb[1] += s0.y; kb[2] += s1.x; kb[3] += s1.y;
                        kb[4] += s2.x; kb[5] += s2.y; kb[6] += s3.x; kb[7] += s3.y;
                        kb[8] += s4;
                    }
                    slot0 += __popc(mask);
                }
                if ( n < nb ) {
                    // the lanes of the warp cover the row contiguously (block after block): three 8-byte
                    // stores per row and lane, merged into full sectors in L2
#pragma unroll
                    for ( int i = 0; i < 3; i++ ) {
                        if ( eqs[i] <= 0 ) continue;
                        double *dst = val + rowbase[i] + cstart;
                        int c = 0;
#pragma unroll
                        for ( int j = 0; j < 3; j++ )
                            if ( cm & ( 1 << j ) ) {
                                dst[c] = ACCUM ? dst[c] + kb[3 * i + j] : kb[3 * i + j];
                                c++;
                            }
                    }
                }
            }
        }
        // no barrier here: the parking lot is next written in phase A2 of the following group, behind that group's
        // A1 -> A2 barrier, which every warp reaches only after its phase B of this group; warps without a node to
        // reduce start on the next group's geometry right away
    }
}

// ---- host side -----------------------------------------------------------------------------------

int gather_prepare_mesh(ob200_elemset *S)
{
    ob200_context *ctx = S->ctx;
    if ( S->nelem == 0 || S->nnode == 0 ) return OB200_OK;
    const int64_t n = S->nelem * S->nen;
    OB_REQUIRE(n < (int64_t) INT_MAX / 8, OB200_ECAPACITY, "elemset: %lld element nodes exceed the 32-bit visit index", (long long) n);
    S->nvisit = n;
    DevBuf< int32_t > cnt, fill;
    DevBuf< int64_t > start64;
    DevBuf< int > bad;
    OB_CHECK( cnt.alloc(S->nnode + 1) );
    OB_CHECK( fill.alloc(S->nnode + 1) );
    OB_CHECK( start64.alloc(S->nnode + 1) );
    OB_CHECK( bad.alloc(1) );
    OB_CUDA( cudaMemsetAsync(cnt.p, 0, sizeof( int32_t ) * ( S->nnode + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(fill.p, 0, sizeof( int32_t ) * ( S->nnode + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(bad.p, 0, sizeof( int ), ctx->stream) );
    const int grid = ctx->shape.grid(n, 256, 8);
    OB_LAUNCH(ctx, node_valence_kernel, grid, 256, 0, S->conn.p, n, S->nnode, cnt.p, bad.p);
    int hbad = 0;
    OB_CUDA( cudaMemcpyAsync(&hbad, bad.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    OB_REQUIRE(hbad == 0, OB200_EINVAL, "elemset_create: %d connectivity entries outside [1, nnode=%lld]", hbad, (long long) S->nnode);
    int64_t total = 0;
    OB_CHECK( exclusive_scan(ctx, cnt.p, start64.p, S->nnode + 1, &total) );
    OB_CHECK( S->ninc_start.alloc(S->nnode + 1) );
    OB_CHECK( narrow_i64_to_i32(ctx, start64.p, S->ninc_start.p, S->nnode + 1) );
    OB_CHECK( S->ninc.alloc(n) );
    OB_LAUNCH(ctx, node_incidence_fill_kernel, grid, 256, 0, S->conn.p, n, S->nen, S->ninc_start.p, fill.p, S->ninc.p);
    OB_LAUNCH(ctx, node_incidence_sort_kernel, ctx->shape.grid(S->nnode, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p);
    OB_CHECK( max_reduce(ctx, cnt.p, S->nnode, &S->maxval) );
    OB_CHECK( S->nodeeq.alloc(S->nnode * 3) );
    OB_CUDA( cudaMemsetAsync(S->nodeeq.p, 0, sizeof( int32_t ) * (size_t) S->nnode * 3, ctx->stream) );
    OB_CHECK( elemset_await_loc(S) );
    OB_LAUNCH(ctx, node_equations_kernel, grid, 256, 0, S->conn.p, S->loc.p, S->nelem, S->nen, S->nodeeq.p);
    // position of every (element, local node) incidence in the node lists (strip assembly, owner-computes vector assembly)
    OB_CHECK( S->row_vis.alloc(S->nelem * 8) );
    OB_LAUNCH(ctx, visit_index_kernel, grid, 256, 0, S->ninc.p, n, S->row_vis.p);
    {
        DevBuf< int32_t > eqcount;
        OB_CHECK( eqcount.alloc(S->neq > 0 ? S->neq : 1) );
        OB_CUDA( cudaMemsetAsync(eqcount.p, 0, sizeof( int32_t ) * (size_t)( S->neq > 0 ? S->neq : 1 ), ctx->stream) );
        OB_CUDA( cudaMemsetAsync(bad.p, 0, sizeof( int ), ctx->stream) );
        OB_LAUNCH(ctx, equation_owner_count_kernel, ctx->shape.grid(S->nnode * 3, 256, 8), 256, 0, S->nodeeq.p, S->nnode * 3, S->neq, eqcount.p, bad.p);
        OB_CUDA( cudaMemcpyAsync(&hbad, bad.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(ctx->stream) );
        S->eq_unique = ( hbad == 0 );
    }
    if ( S->etype != OB200_LSPACE ) return OB200_OK;          // the rest is the visit-group schedule of the LSpace kernels
    OB_CHECK( S->ninc_node.alloc(n) );
    S->ngroups = (int32_t)( ( S->nvisit - 1 ) / kGroupVisits + 1 );
    OB_CHECK( S->gtab.alloc(S->ngroups + 1) );
    OB_LAUNCH(ctx, group_table_kernel, ctx->shape.grid(S->nnode + 1, 256, 8), 256, 0, (int32_t) S->nnode, S->ninc_start.p, S->ngroups, S->gtab.p);
    OB_LAUNCH(ctx, visit_base_kernel, ctx->shape.grid(S->nnode, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->gtab.p, S->ninc_node.p);
    OB_CHECK( S->vu.alloc(n) );
    if ( S->maxval <= kMaxValence )
        OB_LAUNCH(ctx, visit_unique_kernel, ctx->shape.grid((int64_t) S->ngroups * 32, 256, 8), 256, 0, S->ngroups, S->gtab.p, S->ninc.p, S->vu.p);
    OB_CHECK( S->exyz.alloc(n * 3) );
    OB_LAUNCH(ctx, element_coords_kernel, grid, 256, 0, S->conn.p, S->coords.p, n, S->exyz.p);
    return OB200_OK;
}

int gather_bind(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    S->gather_ok = false;
    S->cluster_ok = false;
    S->strips_ok = false;
    const char *mode = getenv("OB200_ASSEMBLY");
    if ( mode && !strcmp(mode, "slotmap") ) return OB200_OK;      // generic path (cross-checks)
    if ( S->etype != OB200_LSPACE || S->nelem == 0 || !S->ninc.p ) return OB200_OK;
    // sets with a MisesMat (or OB200_ASSEMBLY=strips) take the element-strip assembly, which shares the node-block schedule
    const bool strips = !S->all_isole || ( mode && !strcmp(mode, "strips") );
    if ( S->maxval > kMaxValence || A->maxrow > kMaxRowLen || A->neq == 0 ) return OB200_OK;
    S->maxblk = ( ( A->maxrow + 3 ) & ~3 );          // a column block is at least one column wide
    if ( S->maxblk < 4 ) S->maxblk = 4;
    if ( !strips ) OB_CHECK( S->pos.alloc(S->nvisit * 8) );
    OB_CHECK( S->nblk.alloc(S->nnode) );
    OB_CHECK( S->blk.alloc(S->nnode * S->maxblk) );
    DevBuf< int > flags;
    OB_CHECK( flags.alloc(4) );            // [0] mismatch, [1] capacity, [2..3] 64-bit count of covered entries
    OB_CUDA( cudaMemsetAsync(flags.p, 0, sizeof( int ) * 4, ctx->stream) );
    OB_CUDA( cudaMemsetAsync(S->blk.p, 0, sizeof( unsigned short ) * (size_t) S->nnode * S->maxblk, ctx->stream) );
    OB_CHECK( S->ebidx.alloc(S->nvisit * 8) );
    const bool allpairs = !strips && getenv("OB200_NODE_BLOCKS") && !strcmp(getenv("OB200_NODE_BLOCKS"), "allpairs");
    if ( allpairs ) {
        OB_LAUNCH(ctx, node_blocks_allpairs_kernel, ctx->shape.grid(S->nnode * 32, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p,
                  S->conn.p, S->nodeeq.p, A->rowptr.p, A->colind.p, S->maxblk, S->pos.p, S->nblk.p, S->blk.p, flags.p,
                  reinterpret_cast< unsigned long long * >( flags.p + 2 ));
    } else {
        OB_LAUNCH(ctx, node_blocks_kernel, ctx->shape.grid(S->nnode * 32, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p,
                  S->conn.p, S->nodeeq.p, A->rowptr.p, A->colind.p, S->maxblk, strips ? nullptr : S->pos.p, S->nblk.p, S->blk.p, flags.p,
                  reinterpret_cast< unsigned long long * >( flags.p + 2 ), S->ebidx.p);
    }
    int h[4] = { 0, 0, 0, 0 };
    OB_CUDA( cudaMemcpyAsync(h, flags.p, sizeof( int ) * 4, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    // h[0]: the matrix pattern is not the pattern of this element set (or equations of a node are not
    // consecutive); h[1]: capacity -- both keep the generic slot-map path
    S->gather_ok = ( h[0] == 0 && h[1] == 0 && !strips );
    S->strips_ok = ( h[0] == 0 && h[1] == 0 && strips );
    unsigned long long cov;
    memcpy(&cov, h + 2, sizeof( cov ));
    S->covers_all = ( (int64_t) cov == A->nnz );
    if ( S->gather_ok && !allpairs ) OB_CHECK( cluster_bind(S, A) );
    if ( S->strips_ok ) OB_CHECK( strips_bind(S, A) );
    return OB200_OK;
}

int gather_assemble_lspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    static unsigned long long attr_set = 0;
    const int smem = (int) sizeof( GatherShared );
    if ( attr_needed(attr_set, ctx->device) ) {
        OB_CUDA( cudaFuncSetAttribute(lspace_gather_kernel< false >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
        OB_CUDA( cudaFuncSetAttribute(lspace_gather_kernel< true >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
    }
    GatherView G{ S->ninc_start.p, S->ninc.p, S->ninc_node.p, S->nodeeq.p, S->pos.p, S->nblk.p, S->vu.p, S->blk.p, S->gtab.p, S->maxblk };
    ElemSetView v = S->view();
    int grid = ctx->shape.sms * OB200_GATHER_CTAS;     // persistent: as many CTAs as fit an SM
    if ( grid > S->ngroups ) grid = S->ngroups;
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    if ( A->zero_pending ) {
        // every entry of the pattern is written by exactly one warp: the pending zero() is absorbed
        OB_LAUNCH(ctx, lspace_gather_kernel< false >, grid, kGatherThreads + 32, smem, v, G, S->ngroups, A->rowptr.p, A->val.p);
        A->zero_pending = false;
    } else {
        OB_LAUNCH(ctx, lspace_gather_kernel< true >, grid, kGatherThreads + 32, smem, v, G, S->ngroups, A->rowptr.p, A->val.p);
    }
    return OB200_OK;
}

} // namespace ob200
