// Owner-computes ("gather") assembly of the LSpace tangent into the cudacsr matrix.
//
// Replaces EngngModel::assemble (src/core/engngm.C:889-929) + CompCol::assemble
// (src/core/compcol.C:263-299) for a whole element set without a single atomic: every matrix row
// is produced by exactly one warp and written once, coalesced -- bit-reproducible run to run.
//
//   * A "visit" is one (node, adjacent element) incidence.  One thread per visit integrates the
//     eight 3x3 blocks K_ab (a = the visited node, b = 0..7) of that element:
//         G_ab = sum_gp dV (grad N_a)(grad N_b)^T,   K_ab = lambda G + mu G^T + mu tr(G) I
//     (IsotropicLinearElasticMaterial D, isolinearelasticmaterial.C:80-84, through the B matrix of
//     Structural3DElement::computeBmatrixAt, structural3delement.C:63-86) and parks them in shared
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

namespace ob200 {

constexpr int kGroupVisits = 32;                     // visits per group window
constexpr int kMaxValence = 16;                      // elements around a node (fast path)
constexpr int kGatherVisits = kGroupVisits + kMaxValence;    // most visits a group can hold
constexpr int kGatherThreads = 4 * kGroupVisits;     // 128: four lanes per visit; a group's few extra visits take a second pass
constexpr int kGatherWarps = kGatherThreads / 32;
constexpr int kMaxRowLen = 128;                      // longest row laid out in shared memory
constexpr int kBlkDoubles = 9;

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
__global__ void node_incidence_sort_kernel(int64_t nnode, const int32_t *__restrict__ start, int32_t *__restrict__ ninc,
                                           int32_t *__restrict__ ninc_node)
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
        for ( int i = b; i < e; i++ ) ninc_node[i] = (int32_t) w;
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

// ---- per bound matrix: where every block is parked, the column blocks of every node ---------
//
// One warp per node.  Items = (visit k, local node b) of the node's elements; key = first free
// equation of the neighbouring node conn[e_k][b].  Items sorted by (key, item index) give the
// parking position; each distinct key is one column block of the node's rows.
// flags[0]: a node's column blocks do not line up with the CSR row (free equations of a node not
// consecutive, or the pattern is not the one of this element set); flags[1]: capacity exceeded.
__global__ void __launch_bounds__(256)
node_blocks_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                   const int32_t *__restrict__ conn, const int32_t *__restrict__ nodeeq,
                   const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind, int maxblk,
                   unsigned char *__restrict__ pos, unsigned char *__restrict__ nblk, unsigned short *__restrict__ blk,
                   int *__restrict__ flags, unsigned long long *__restrict__ covered)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    constexpr int Q = kMaxValence * 8 / 32;          // items per lane
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
        // pass 1: items with a smaller key, same key before me, same key in total
        int less[Q], same_before[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) less[q] = same_before[q] = 0;
        const int nq = ( nitems + 31 ) >> 5;
        for ( int q2 = 0; q2 < nq; q2++ ) {
            int kq = INT_MAX;
#pragma unroll
            for ( int q = 0; q < Q; q++ ) if ( q == q2 ) kq = key[q];
            for ( int src = 0; src < 32; src++ ) {
                const int jt = q2 * 32 + src;
                if ( jt >= nitems ) break;
                const int kj = __shfl_sync(0xffffffffu, kq, src);
#pragma unroll
                for ( int q = 0; q < Q; q++ ) {
                    const int it = q * 32 + lane;
                    less[q] += ( kj < key[q] );
                    same_before[q] += ( kj == key[q] && jt < it );
                }
            }
        }
        // pass 2: distinct smaller keys (= my column block) and their total width (= my first column)
        int blkidx[Q], cstart[Q];
#pragma unroll
        for ( int q = 0; q < Q; q++ ) blkidx[q] = cstart[q] = 0;
        for ( int q2 = 0; q2 < nq; q2++ ) {
            int kq = INT_MAX, fq = 0, wq = 0;
#pragma unroll
            for ( int q = 0; q < Q; q++ ) if ( q == q2 ) { kq = key[q]; fq = ( same_before[q] == 0 ); wq = __popc(cm[q]); }
            for ( int src = 0; src < 32; src++ ) {
                const int jt = q2 * 32 + src;
                if ( jt >= nitems ) break;
                const int kj = __shfl_sync(0xffffffffu, kq, src);
                const int fj = __shfl_sync(0xffffffffu, fq, src);
                const int wj = __shfl_sync(0xffffffffu, wq, src);
#pragma unroll
                for ( int q = 0; q < Q; q++ )
                    if ( fj && kj < key[q] ) { blkidx[q]++; cstart[q] += wj; }
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
                const int p = less[q] + same_before[q];
                pos[( (int64_t) v0 + ( it >> 3 ) ) * 8 + ( it & 7 )] = (unsigned char)( live ? p : 0xFF );
                if ( live ) {
                    if ( p >= 0xFF || blkidx[q] >= maxblk ) atomicAdd(flags + 1, 1);
                    // the last item of a block (largest position) records where the block ends
                    nb_max = max(nb_max, blkidx[q] + 1);
                    if ( same_before[q] == 0 ) {
                        width_total += __popc(cm[q]);
                        if ( blkidx[q] < maxblk ) {
                            // count of this key: filled in below through seg_end (positions are contiguous)
                            if ( colind[rowptr[row] + cstart[q]] != key[q] - 1 ) atomicAdd(flags, 1);
                        }
                    }
                }
            }
        }
        // seg_end of block n = number of live items with block index <= n: every item bumps its block
        // (small shared-memory-free approach: lanes publish, block owners count)
        nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, 16));
        nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, 8));
        nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, 4));
        nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, 2));
        nb_max = max(nb_max, __shfl_xor_sync(0xffffffffu, nb_max, 1));
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) width_total += __shfl_xor_sync(0xffffffffu, width_total, o);
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
        // seg_end: the item with the largest position inside its block writes position + 1
        // (it is the one whose same-key count after it is zero); computed with one more sweep
        if ( row >= 0 ) {
            int same_total[Q];
#pragma unroll
            for ( int q = 0; q < Q; q++ ) same_total[q] = 0;
            for ( int q2 = 0; q2 < nq; q2++ ) {
                int kq = INT_MAX;
#pragma unroll
                for ( int q = 0; q < Q; q++ ) if ( q == q2 ) kq = key[q];
                for ( int src = 0; src < 32; src++ ) {
                    const int jt = q2 * 32 + src;
                    if ( jt >= nitems ) break;
                    const int kj = __shfl_sync(0xffffffffu, kq, src);
#pragma unroll
                    for ( int q = 0; q < Q; q++ ) same_total[q] += ( kj == key[q] );
                }
            }
#pragma unroll
            for ( int q = 0; q < Q; q++ ) {
                const int it = q * 32 + lane;
                if ( it < nitems && cm[q] != 0 && same_before[q] == 0 && blkidx[q] < maxblk )
                    blk[w * maxblk + blkidx[q]] = (unsigned short)( ( less[q] + same_total[q] ) | ( cm[q] << 8 ) );
            }
        }
    }
}

// ---- the assembly kernel ------------------------------------------------------------------------

struct GatherView {
    const int32_t *ninc_start, *ninc, *ninc_node, *nodeeq;
    const unsigned char *pos, *nblk;
    const unsigned short *blk;
    const int2 *gtab;
    int maxblk;
};

struct GatherShared {
    double park[kGatherVisits * 8 * kBlkDoubles];             // 36,864 B
    double rows[kGatherWarps][3][kMaxRowLen];                 // 24,576 B
};

template< bool ACCUM >
__global__ void __launch_bounds__(kGatherThreads, 4)
lspace_gather_kernel(ElemSetView S, GatherView G, const int32_t *__restrict__ rowptr, double *__restrict__ val)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    GatherShared &sh = *reinterpret_cast< GatherShared * >( smem_raw );
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int2 g0 = G.gtab[blockIdx.x], g1 = G.gtab[blockIdx.x + 1];
    const int nvis = g1.y - g0.y;

    // ---- phase A: four lanes per visit ----
    // Lane q of the quad inverts the Jacobian at Gauss points 2q, 2q+1 and forms dV * grad N_a there;
    // the quad then walks the eight Gauss points, the owner broadcasts (J^-1, dV grad N_a) with
    // shuffles, and every lane accumulates the two blocks K_ab, b = 2q, 2q+1 (18 accumulators).
    // metadata of the node this warp will reduce first: requested now, consumed after the barrier
    int b_nb = 0, b_info = 0, b_seg0 = 0, b_eq[3] = { 0, 0, 0 }, b_row[3] = { 0, 0, 0 };
    if ( g0.x + wid < g1.x ) {
        const int w = g0.x + wid;
        b_nb = G.nblk[w];
        if ( lane < b_nb ) b_info = G.blk[(int64_t) w * G.maxblk + lane];
        if ( lane > 0 && lane < b_nb ) b_seg0 = G.blk[(int64_t) w * G.maxblk + lane - 1] & 0xFF;
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            b_eq[i] = G.nodeeq[(int64_t) w * 3 + i];
            if ( b_eq[i] > 0 ) b_row[i] = rowptr[b_eq[i] - 1];
        }
    }
    for ( int t = tid >> 2; t < ( ( nvis + 7 ) & ~7 ); t += kGroupVisits ) {
        const int q = tid & 3;
        const bool live_visit = t < nvis;
        const int64_t v = (int64_t) g0.y + ( live_visit ? t : 0 );
        uint2 pw = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        if ( live_visit ) pw = *reinterpret_cast< const uint2 * >( G.pos + v * 8 );
        // whole quads are live or dead together, whole warps stay converged for the shuffles
        const bool work = !( pw.x == 0xFFFFFFFFu && pw.y == 0xFFFFFFFFu );
        if ( __any_sync(0xffffffffu, work) ) {
            const int ent = work ? G.ninc[v] : 0;
            const int64_t e = ent >> 3;
            const int la = ent & 7;
            double own[2][12];       // per owned Gauss point: J^-1 (9), dV * grad N_a (3)
            {
                double xyz[24];
                const int4 c0 = *reinterpret_cast< const int4 * >( S.conn + e * 8 );
                const int4 c1 = *reinterpret_cast< const int4 * >( S.conn + e * 8 + 4 );
                const int nd[8] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w };
#pragma unroll
                for ( int k = 0; k < 8; k++ ) {
                    const double *c = S.coords + (int64_t)( nd[k] - 1 ) * 3;
                    xyz[3 * k] = c[0]; xyz[3 * k + 1] = c[1]; xyz[3 * k + 2] = c[2];
                }
                double sxa, sya, sza;
                hexa_signs(la, sxa, sya, sza);
#pragma unroll
                for ( int h = 0; h < 2; h++ ) {
                    double u, vv, ww, Ji[3][3];
                    hexa_gp(2 * q + h, u, vv, ww);
                    const double dV = fabs(hexa_jacobian(xyz, u, vv, ww, Ji));
                    const double fu = 1.0 + sxa * u, fv = 1.0 + sya * vv, fw = 1.0 + sza * ww;
                    const double da0 = dV * sxa * 0.125 * fv * fw, da1 = dV * sya * 0.125 * fu * fw, da2 = dV * sza * 0.125 * fu * fv;
#pragma unroll
                    for ( int i = 0; i < 3; i++ )
#pragma unroll
                        for ( int j = 0; j < 3; j++ ) own[h][3 * i + j] = Ji[i][j];
#pragma unroll
                    for ( int j = 0; j < 3; j++ ) own[h][9 + j] = da0 * Ji[0][j] + da1 * Ji[1][j] + da2 * Ji[2][j];
                }
            }
            // the two shape functions of this lane: b = 2q, 2q + 1.  dN_b at a Gauss point is
            // +-0.125 * {(1+a)^2, (1-a^2), (1-a)^2} (a = 1/sqrt 3), picked by sign agreement between
            // the node and the Gauss point: selects, no FP64 work
            bool px[2], py[2], pz[2];
#pragma unroll
            for ( int bb = 0; bb < 2; bb++ ) {
                const int b = 2 * q + bb;
                px[bb] = ( b & 3 ) >= 2;
                py[bb] = ( b & 3 ) == 1 || ( b & 3 ) == 2;
                pz[bb] = b < 4;
            }
            constexpr double kA = 0.577350269189626;
            constexpr double cPP = 0.125 * ( 1.0 + kA ) * ( 1.0 + kA ), cPM = 0.125 * ( 1.0 + kA ) * ( 1.0 - kA ),
                             cMM = 0.125 * ( 1.0 - kA ) * ( 1.0 - kA );
            double acc[2][9];
#pragma unroll
            for ( int bb = 0; bb < 2; bb++ )
#pragma unroll
                for ( int k = 0; k < 9; k++ ) acc[bb][k] = 0.0;
#pragma unroll
            for ( int gp = 0; gp < 8; gp++ ) {
                double m[12];
#pragma unroll
                for ( int k = 0; k < 12; k++ ) m[k] = __shfl_sync(0xffffffffu, own[gp & 1][k], gp >> 1, 4);
                const bool gu = ( gp & 4 ) != 0, gv = ( gp & 2 ) != 0, gw = ( gp & 1 ) != 0;   // hexa_gp: sign of (u, v, w)
#pragma unroll
                for ( int bb = 0; bb < 2; bb++ ) {
                    const bool au = px[bb] == gu, av = py[bb] == gv, aw = pz[bb] == gw;
                    const double t0 = av ? ( aw ? cPP : cPM ) : ( aw ? cPM : cMM );
                    const double t1 = au ? ( aw ? cPP : cPM ) : ( aw ? cPM : cMM );
                    const double t2 = au ? ( av ? cPP : cPM ) : ( av ? cPM : cMM );
                    const double d0 = px[bb] ? t0 : -t0, d1 = py[bb] ? t1 : -t1, d2 = pz[bb] ? t2 : -t2;
                    double gb[3];
#pragma unroll
                    for ( int j = 0; j < 3; j++ ) gb[j] = d0 * m[j] + d1 * m[3 + j] + d2 * m[6 + j];
#pragma unroll
                    for ( int i = 0; i < 3; i++ )
#pragma unroll
                        for ( int j = 0; j < 3; j++ ) acc[bb][3 * i + j] += m[9 + i] * gb[j];
                }
            }
            if ( work ) {
                double lam, mu;
                {
                    const MatParams *mp = S.mat + S.matid[e];
                    isole_lame(mp->E, mp->nu, lam, mu);
                }
                const int base = ( G.ninc_start[G.ninc_node[v]] - g0.y ) * 8;
                const unsigned int pq = q < 2 ? pw.x : pw.y;
#pragma unroll
                for ( int bb = 0; bb < 2; bb++ ) {
                    const unsigned int p = ( pq >> ( 16 * ( q & 1 ) + 8 * bb ) ) & 0xFFu;
                    if ( p == 0xFFu ) continue;
                    const double *g = acc[bb];
                    const double tr = mu * ( g[0] + g[4] + g[8] );
                    double *o = sh.park + (size_t)( base + p ) * kBlkDoubles;
                    o[0] = lam * g[0] + mu * g[0] + tr;
                    o[1] = lam * g[1] + mu * g[3];
                    o[2] = lam * g[2] + mu * g[6];
                    o[3] = lam * g[3] + mu * g[1];
                    o[4] = lam * g[4] + mu * g[4] + tr;
                    o[5] = lam * g[5] + mu * g[7];
                    o[6] = lam * g[6] + mu * g[2];
                    o[7] = lam * g[7] + mu * g[5];
                    o[8] = lam * g[8] + mu * g[8] + tr;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase B: one warp per node: sum the parked blocks per column block, write the rows ----
    for ( int w = g0.x + wid; w < g1.x; w += kGatherWarps ) {
        const bool first = ( w == g0.x + wid );
        const int nb = first ? b_nb : G.nblk[w];
        if ( nb == 0 ) continue;
        int eqs[3], rowbase[3];
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            eqs[i] = first ? b_eq[i] : G.nodeeq[(int64_t) w * 3 + i];
            rowbase[i] = first ? b_row[i] : ( eqs[i] > 0 ? rowptr[eqs[i] - 1] : 0 );
        }
        const int vbase = ( G.ninc_start[w] - g0.y ) * 8;
        double( *rows )[kMaxRowLen] = sh.rows[wid];
        int ccarry = 0, rowlen = 0;
        for ( int n0 = 0; n0 < nb; n0 += 32 ) {
            const int n = n0 + lane;
            int info = 0, seg0 = 0;
            if ( first && n0 == 0 ) {
                info = b_info;
                seg0 = b_seg0;
            } else if ( n < nb ) {
                info = G.blk[(int64_t) w * G.maxblk + n];
                seg0 = n > 0 ? ( G.blk[(int64_t) w * G.maxblk + n - 1] & 0xFF ) : 0;
            }
            const int seg1 = info & 0xFF, cm = info >> 8;
            const int width = __popc(cm);
            int incl = width;                                   // inclusive scan of the widths
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ( lane >= o ) incl += t;
            }
            const int cstart = ccarry + incl - width;
            ccarry += __shfl_sync(0xffffffffu, incl, 31);
            if ( n < nb ) {
                double k[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
                for ( int p = seg0; p < seg1; p++ ) {
                    const double *s = sh.park + (size_t)( vbase + p ) * kBlkDoubles;
#pragma unroll
                    for ( int q = 0; q < 9; q++ ) k[q] += s[q];
                }
                int c = cstart;
#pragma unroll
                for ( int j = 0; j < 3; j++ )
                    if ( cm & ( 1 << j ) ) {
                        rows[0][c] = k[j];
                        rows[1][c] = k[3 + j];
                        rows[2][c] = k[6 + j];
                        c++;
                    }
            }
            rowlen = ccarry;
        }
        __syncwarp();
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            if ( eqs[i] <= 0 ) continue;
            double *dst = val + rowbase[i];
            for ( int c = lane; c < rowlen; c += 32 ) dst[c] = ACCUM ? dst[c] + rows[i][c] : rows[i][c];
        }
        __syncwarp();
    }
}

// ---- host side -----------------------------------------------------------------------------------

int gather_prepare_mesh(ob200_elemset *S)
{
    ob200_context *ctx = S->ctx;
    if ( S->etype != OB200_LSPACE || S->nelem == 0 || S->nnode == 0 ) return OB200_OK;
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
    OB_CHECK( S->ninc_node.alloc(n) );
    OB_LAUNCH(ctx, node_incidence_fill_kernel, grid, 256, 0, S->conn.p, n, S->nen, S->ninc_start.p, fill.p, S->ninc.p);
    OB_LAUNCH(ctx, node_incidence_sort_kernel, ctx->shape.grid(S->nnode, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p, S->ninc_node.p);
    OB_CHECK( max_reduce(ctx, cnt.p, S->nnode, &S->maxval) );
    OB_CHECK( S->nodeeq.alloc(S->nnode * 3) );
    OB_CUDA( cudaMemsetAsync(S->nodeeq.p, 0, sizeof( int32_t ) * (size_t) S->nnode * 3, ctx->stream) );
    OB_LAUNCH(ctx, node_equations_kernel, grid, 256, 0, S->conn.p, S->loc.p, S->nelem, S->nen, S->nodeeq.p);
    return OB200_OK;
}

int gather_bind(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    S->gather_ok = false;
    if ( S->etype != OB200_LSPACE || S->nelem == 0 || !S->all_isole || !S->ninc.p ) return OB200_OK;
    if ( S->maxval > kMaxValence || A->maxrow > kMaxRowLen || A->neq == 0 ) return OB200_OK;
    S->maxblk = ( ( A->maxrow + 3 ) & ~3 );          // a column block is at least one column wide
    if ( S->maxblk < 4 ) S->maxblk = 4;
    OB_CHECK( S->pos.alloc(S->nvisit * 8) );
    OB_CHECK( S->nblk.alloc(S->nnode) );
    OB_CHECK( S->blk.alloc(S->nnode * S->maxblk) );
    DevBuf< int > flags;
    OB_CHECK( flags.alloc(4) );            // [0] mismatch, [1] capacity, [2..3] 64-bit count of covered entries
    OB_CUDA( cudaMemsetAsync(flags.p, 0, sizeof( int ) * 4, ctx->stream) );
    OB_CUDA( cudaMemsetAsync(S->blk.p, 0, sizeof( unsigned short ) * (size_t) S->nnode * S->maxblk, ctx->stream) );
    OB_LAUNCH(ctx, node_blocks_kernel, ctx->shape.grid(S->nnode * 32, 256, 8), 256, 0, S->nnode, S->ninc_start.p, S->ninc.p,
              S->conn.p, S->nodeeq.p, A->rowptr.p, A->colind.p, S->maxblk, S->pos.p, S->nblk.p, S->blk.p, flags.p,
              reinterpret_cast< unsigned long long * >( flags.p + 2 ));
    S->ngroups = (int32_t)( ( S->nvisit - 1 ) / kGroupVisits + 1 );
    OB_CHECK( S->gtab.alloc(S->ngroups + 1) );
    OB_LAUNCH(ctx, group_table_kernel, ctx->shape.grid(S->nnode + 1, 256, 8), 256, 0, (int32_t) S->nnode, S->ninc_start.p, S->ngroups, S->gtab.p);
    int h[4] = { 0, 0, 0, 0 };
    OB_CUDA( cudaMemcpyAsync(h, flags.p, sizeof( int ) * 4, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    // h[0]: the matrix pattern is not the pattern of this element set (or equations of a node are not
    // consecutive); h[1]: capacity -- both keep the generic slot-map path
    S->gather_ok = ( h[0] == 0 && h[1] == 0 );
    unsigned long long cov;
    memcpy(&cov, h + 2, sizeof( cov ));
    S->covers_all = ( (int64_t) cov == A->nnz );
    return OB200_OK;
}

int gather_assemble_lspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    static bool attr_set = false;
    const int smem = (int) sizeof( GatherShared );
    if ( !attr_set ) {
        OB_CUDA( cudaFuncSetAttribute(lspace_gather_kernel< false >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
        OB_CUDA( cudaFuncSetAttribute(lspace_gather_kernel< true >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
        attr_set = true;
    }
    GatherView G{ S->ninc_start.p, S->ninc.p, S->ninc_node.p, S->nodeeq.p, S->pos.p, S->nblk.p, S->blk.p, S->gtab.p, S->maxblk };
    ElemSetView v = S->view();
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    if ( A->zero_pending ) {
        // every entry of the pattern is written by exactly one warp: the pending zero() is absorbed
        OB_LAUNCH(ctx, lspace_gather_kernel< false >, S->ngroups, kGatherThreads, smem, v, G, A->rowptr.p, A->val.p);
        A->zero_pending = false;
    } else {
        OB_LAUNCH(ctx, lspace_gather_kernel< true >, S->ngroups, kGatherThreads, smem, v, G, A->rowptr.p, A->val.p);
    }
    return OB200_OK;
}

} // namespace ob200
