// Cluster assembly of the LSpace / IsotropicLinearElasticMaterial tangent into the cudacsr matrix: the fast path of
// ob200_elemset_assemble_stiffness.
//
// Replaces EngngModel::assemble (src/core/engngm.C:889-929) over StructuralElement::computeStiffnessMatrix
// (src/sm/Elements/structuralelement.C:575-643) + CompCol::assemble (src/core/compcol.C:263-299) for a whole element
// set, owner-computes, without atomics, bit-reproducible run to run.
//
//   * The nodes are binned into spatial cells of ~4x4x4 nodes; a CLUSTER is up to 64 nodes of one cell.  One CTA owns
//     a cluster at a time: the 3x3 blocks of the rows of its nodes live in shared memory as nine planes
//     acc[3i+j][position], position = (node's first block) + (index of the column node in the node's block list).
//   * Every element touching a cluster node is evaluated for that cluster (1.95 evaluations per element on a
//     structured mesh; 5.3 in the round-1 kernel).  Geometry warps: one thread per (element, Gauss point) writes
//     H[i][a][gp] = sqrt(|det J|) dN_a/dx_i (FEI3dHexaLin::evaldNdx, fei3dhexalin.C:186-204; 2x2x2 rule,
//     gaussintegrationrule.C:190-214).  Contraction warps: G = H H^T on the FP64 tensor path -- with the rows of H
//     ordered component-major, tile (i,j) of the 24x24 product is the 8x8 matrix G_ab[i][j] over the node pairs (a,b),
//     exactly one mma.sync.m8n8k4.f64 accumulator (DMMA in SASS; B200 runs it at the DFMA pipe's rate -- profiles/
//     probe_fp64_r02.txt -- but one instruction carries 256 FMAs instead of 32, which is what frees the issue slots).
//     Only the six tiles i <= j are computed; the value of tile (i,j) at (a,b) is also entry (j,i) of block (b,a).
//     The lanes add their entries into the planes (each lane owns distinct positions within an element).
//   * Elements of a cluster are ordered by a greedy colouring (elements of one colour share no cluster node), so
//     conflicting elements are far apart in the order; correctness does not rest on that: every record carries, per
//     contraction warp, the number of elements that warp must have completed before this one may touch the planes
//     (the last earlier element sharing a cluster node), and warps publish their progress in shared memory.  The
//     order of the additions into any entry is therefore fixed by the schedule.
//   * When the cluster's elements are done the contraction warps apply K = lambda G + mu G^T + mu tr(G) I per block
//     (IsotropicLinearElasticMaterial, isolinearelasticmaterial.C:80-84 through the B matrix of
//     Structural3DElement::computeBmatrixAt, structural3delement.C:63-86) and stream the rows to val.  Elements of
//     different materials in one cluster form separate steps, the later ones adding to val.
//
// Everything index-like is computed once per (element set, matrix) by the kernels in the first half of this file and
// stored as a stream of self-contained 304-byte records (vertex coordinates, block indices, accumulator bases,
// first-touch bits, dependencies) that one thread per CTA moves into a shared-memory ring with bulk async copies.
#include "element_device.cuh"
#include "elemset.h"
#include "scan.cuh"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace ob200 {

#ifndef OB200_CL_NODES
#define OB200_CL_NODES 64
#endif
#ifndef OB200_CL_CTAS
#define OB200_CL_CTAS 1
#endif
constexpr int kClNodes = OB200_CL_NODES;     // most nodes a cluster owns
constexpr int kClBlocks = 27 * kClNodes;     // most 3x3 blocks (positions per plane) a cluster owns
constexpr int kClCtas = OB200_CL_CTAS;       // CTAs per SM (two need clusters of 32 nodes and the small rings to fit shared memory)
static_assert( kClNodes <= 64, "six bits of cluster-local node index in the flush table" );
constexpr int kClMaxValence = 16;            // elements around a node
constexpr int kClVisits = kClNodes * kClMaxValence;      // most (node, element) incidences of a cluster
constexpr int kClMaxCell = 2048;             // most nodes in one spatial cell (beyond: the path declines)
#ifndef OB200_CL_CWARPS
#define OB200_CL_CWARPS 8
#endif
#ifndef OB200_CL_GWARPS
#define OB200_CL_GWARPS 4
#endif
#ifndef OB200_CL_RECSLOTS
#define OB200_CL_RECSLOTS 16
#endif
#ifndef OB200_CL_HSLOTS
#define OB200_CL_HSLOTS 8
#endif
#ifndef OB200_CL_FLUSH_UNROLL
#define OB200_CL_FLUSH_UNROLL 2
#endif
#ifndef OB200_CL_SPIN_NS
#define OB200_CL_SPIN_NS 20
#endif
#ifndef OB200_CL_ABLATE
#define OB200_CL_ABLATE 0
#endif
constexpr int kFlushUnroll = OB200_CL_FLUSH_UNROLL;
constexpr int kAblate = OB200_CL_ABLATE;     // scratch (timing experiments only): 1 no accumulate, 2 no products, 4 no flush, 8 no geometry, 16 no dependency wait
constexpr int kCWarps = OB200_CL_CWARPS;     // contraction warps
constexpr int kGWarps = OB200_CL_GWARPS;     // geometry warps
constexpr int kClThreads = ( kCWarps + 1 + kGWarps ) * 32;
constexpr int kRecSlots = OB200_CL_RECSLOTS;                // ring of record packets (4 elements each): deep enough to cover the HBM latency
constexpr int kHSlots = OB200_CL_HSLOTS;                   // ring of gradient packets between the geometry and the contraction warps
constexpr int kHStride = 200;                // doubles per element in the gradient ring: H[kstep][3a'+i][gp & 3], kstep stride 100
#ifndef OB200_CL_BANKCAP
#define OB200_CL_BANKCAP ( OB200_CL_NODES == 64 ? 120 : ( 30 * OB200_CL_NODES ) / 16 )
#endif
constexpr int kBankCap = OB200_CL_BANKCAP;   // positions per shared-memory bank residue (16 residues of 8-byte words)
static_assert( 16 * kBankCap <= 2048 && 16 * kBankCap > kClBlocks, "plane positions: 11 bits in the flush table" );
constexpr int kPlane = 16 * kBankCap;        // plane stride: positions are handed out per bank residue (see cl_records_kernel)
static_assert( kCWarps <= 16, "dependency bytes per record" );
constexpr int kBuildThreads = 128;

struct ClRecord {                            // one (cluster, element) incidence
    double xyz[24];                          // vertex coordinates (the element's own node order)
    unsigned short pos[64];                  // [a'][b']: position of block (node at row slot a', node at slot b') in the planes,
                                             // bit 15: this element is the first of the step to touch it; 0xFFFF: no such block here
    unsigned char need[16];                  // elements contraction warp w must have completed in this step before this one
    int32_t elem;                            // element number, -1 = padding
    uint32_t slotof;                         // nibble a: row slot a' of the element's node a -- cluster nodes first, so that the
                                             // lanes with something to add are the first ones of a warp
    int32_t pad[2];                          // pad[0]: number of cluster nodes of the element (the first row slots)
};
static_assert( sizeof( ClRecord ) == 352, "record layout" );
constexpr int kPacketBytes = 4 * (int) sizeof( ClRecord );
constexpr int kChunkPk = 4;                  // packets per bulk copy (one producer thread: fewer, larger copies)
constexpr int kChunks = kRecSlots / kChunkPk;
static_assert( kRecSlots % kChunkPk == 0, "record ring holds whole chunks" );

struct ClStep {                              // elements of one material of one cluster
    int32_t rec_begin, npk;                  // first record, packets (4 records each)
    int32_t node_begin, nnodes;              // the cluster's nodes in cnodes[]
    int32_t matid, flags;                    // flags bit 0: add to val (not the first step of its cluster); bit 1: clear the planes
                                             // first (several materials in the cluster); bit 2: last step of its cluster
    int32_t nblocks, nsteps;                 // positions per plane in use; steps of this cluster (valid in its first step)
};
// What the flush of a step needs, contiguous in HBM so that one bulk copy brings it into shared memory:
// the header, per cluster node the start of its (up to 3) rows in val (-1: prescribed dof) and its first block (sign bit: some
// column node of the row has prescribed dofs), per block (A, B) (in row order: node after node, column block after column
// block) its position in the planes | position of the transposed block (B, A) << 11 | free-dof mask of the column node << 22 |
// (B is a cluster node) << 25 | cluster-local index of A << 26.
struct ClBlob {
    ClStep hdr;
    double lam, mu;                          // Lame constants of the step's material
    int32_t rowbase[kClNodes][4];
    uint32_t postab[kClBlocks];
};
static_assert( sizeof( ClBlob ) % 16 == 0, "blob layout" );
constexpr int kBlobHead = (int)( sizeof( ClStep ) + 2 * sizeof( double ) + sizeof( int32_t ) * kClNodes * 4 );

// ---- spatial cells -> clusters ---------------------------------------------------------------------

__device__ __forceinline__ unsigned long long order_key(double d)
{
    const unsigned long long b = (unsigned long long) __double_as_longlong(d);
    return ( b & 0x8000000000000000ull ) ? ~b : ( b | 0x8000000000000000ull );
}
static inline double order_key_inv(unsigned long long k)
{
    const unsigned long long b = ( k & 0x8000000000000000ull ) ? ( k & 0x7FFFFFFFFFFFFFFFull ) : ~k;
    double d;
    memcpy(&d, &b, sizeof( d ));
    return d;
}

// box[0..2] = min, box[3..5] = max over the nodes that own rows, as order-preserving integer keys
__global__ void cl_bbox_kernel(const double *__restrict__ coords, const unsigned char *__restrict__ nblk, int64_t nnode,
                               unsigned long long *__restrict__ box)
{
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t n = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; n < nnode; n += stride ) {
        if ( nblk[n] == 0 ) continue;
#pragma unroll
        for ( int d = 0; d < 3; d++ ) {
            const double x = coords[n * 3 + d];
            lo[d] = fmin(lo[d], x);
            hi[d] = fmax(hi[d], x);
        }
    }
#pragma unroll
    for ( int d = 0; d < 3; d++ ) {
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ( ( threadIdx.x & 31 ) == 0 ) {
            atomicMin(box + d, order_key(lo[d]));
            atomicMax(box + 3 + d, order_key(hi[d]));
        }
    }
}

// sum over the elements of their largest axis extent, in fixed point (integer additions: order independent)
__global__ void cl_extent_kernel(const double *__restrict__ coords, const int32_t *__restrict__ conn, int64_t nelem, double scale,
                                 unsigned long long *__restrict__ sum)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride ) {
        double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
#pragma unroll
        for ( int k = 0; k < 8; k++ ) {
            const double *c = coords + (int64_t)( conn[e * 8 + k] - 1 ) * 3;
#pragma unroll
            for ( int d = 0; d < 3; d++ ) {
                lo[d] = fmin(lo[d], c[d]);
                hi[d] = fmax(hi[d], c[d]);
            }
        }
        const double ext = fmax(hi[0] - lo[0], fmax(hi[1] - lo[1], hi[2] - lo[2]));
        acc += (unsigned long long)( ext * scale );
    }
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ( ( threadIdx.x & 31 ) == 0 && acc ) atomicAdd(sum, acc);
}

struct ClGrid {
    double x0[3], inv;         // cell index = floor((x - x0) * inv)
    int dim[3];
};

__global__ void cl_cell_kernel(const double *__restrict__ coords, const unsigned char *__restrict__ nblk, int64_t nnode, ClGrid g,
                               int32_t *__restrict__ cell, int32_t *__restrict__ count, unsigned char *__restrict__ npar)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t n = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; n < nnode; n += stride ) {
        int c = -1;
        {   // the node's position on the grid of element size, two bits per axis: bit 2 x, bit 1 y, bit 0 z parities, bits 5..3
            // the next bit of x, y, z (bank labels, see cl_records_kernel)
            int pr = 0;
#pragma unroll
            for ( int d = 0; d < 3; d++ ) {
                const int k = (int) floor(( coords[n * 3 + d] - g.x0[d] ) * ( 4.0 * g.inv ));
                pr |= ( k & 1 ) << ( 2 - d );
                pr |= ( ( k >> 1 ) & 1 ) << ( 5 - d );
            }
            npar[n] = (unsigned char) pr;
        }
        if ( nblk[n] ) {
            int k[3];
#pragma unroll
            for ( int d = 0; d < 3; d++ ) {
                k[d] = (int) floor(( coords[n * 3 + d] - g.x0[d] ) * g.inv);
                k[d] = min(max(k[d], 0), g.dim[d] - 1);
            }
            c = ( k[0] * g.dim[1] + k[1] ) * g.dim[2] + k[2];
            atomicAdd(count + c, 1);
        }
        cell[n] = c;
    }
}

__global__ void cl_cell_fill_kernel(const int32_t *__restrict__ cell, int64_t nnode, const int32_t *__restrict__ start,
                                    int32_t *__restrict__ fill, int32_t *__restrict__ cnodes)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t n = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; n < nnode; n += stride ) {
        const int c = cell[n];
        if ( c >= 0 ) cnodes[start[c] + atomicAdd(fill + c, 1)] = (int32_t) n;
    }
}

// One thread per cell: sort the cell's nodes (ascending -- the fill order of the atomics must not matter), then cut them
// into clusters of at most kClNodes nodes / kClBlocks blocks.  WRITE = false: count the clusters of the cell;
// WRITE = true: fill the cluster table and the per-node maps.
template< bool WRITE >
__global__ void cl_split_kernel(int32_t ncell, const int32_t *__restrict__ start, int32_t *__restrict__ cnodes,
                                const unsigned char *__restrict__ nblk, int32_t *__restrict__ nclus,
                                const int32_t *__restrict__ cl_off, int32_t *__restrict__ cl_begin, int32_t *__restrict__ ncl,
                                unsigned short *__restrict__ nbase, unsigned char *__restrict__ nloc)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += stride ) {
        const int b = start[c], e = start[c + 1];
        if ( !WRITE ) {
            for ( int i = b + 1; i < e; i++ ) {
                const int32_t v = cnodes[i];
                int j = i - 1;
                while ( j >= b && cnodes[j] > v ) { cnodes[j + 1] = cnodes[j]; j--; }
                cnodes[j + 1] = v;
            }
        }
        int k = 0, nn = 0, nb = 0;
        for ( int i = b; i < e; i++ ) {
            const int n = cnodes[i], w = nblk[n];
            if ( nn == kClNodes || nb + w > kClBlocks ) { k++; nn = 0; nb = 0; }
            if ( WRITE ) {
                if ( nn == 0 ) cl_begin[cl_off[c] + k] = i;
                ncl[n] = cl_off[c] + k;
                nbase[n] = (unsigned short) nb;
                nloc[n] = (unsigned char) nn;
            }
            nn++;
            nb += w;
        }
        if ( !WRITE ) nclus[c] = ( e > b ) ? k + 1 : 0;
    }
}

// ---- per cluster: element list, order, records ---------------------------------------------------------

struct ClBuildShared {
    unsigned long long key[kClVisits];       // sort keys
    int32_t elem[kClVisits];                 // distinct elements, ascending
    unsigned char oloc[kClVisits][8];        // cluster-local index of the element's node a, 0xFF if not a cluster node
    unsigned short outidx[kClVisits];        // position of sorted entry i in the cluster's record range
    short slot2ent[kClVisits + 4 * 64];      // record slot -> sorted entry, -1 = padding
    int minord[kClBlocks];                   // per block: first record of the step touching it
    unsigned short bpos[kClBlocks];          // per block (node's first block + index in the node's block list): its position
    unsigned short tbid[kClBlocks];          // per block (A, B): the block (B, A) if B is a cluster node, else 0xFFFF
    int bankcnt[16];
    unsigned char last[kClNodes][kCWarps];
    unsigned int colormask[kClNodes];
    int32_t nodes[kClNodes];
    int stepinfo[64][3];                     // per step: first record slot (cluster relative), record slots, matid
    int scan[kBuildThreads / 32];
    int ne, nslots, nsteps, nblocks;
};

__device__ __forceinline__ void cl_bitonic(unsigned long long *key, int n2, int tid, int nthreads)
{
    for ( int k = 2; k <= n2; k <<= 1 ) {
        for ( int j = k >> 1; j > 0; j >>= 1 ) {
            for ( int t = tid; t < n2 / 2; t += nthreads ) {
                const int i = ( ( t & ~( j - 1 ) ) << 1 ) | ( t & ( j - 1 ) ), p = i | j;
                const unsigned long long a = key[i], b = key[p];
                const bool up = ( i & k ) == 0;
                if ( ( a > b ) == up ) { key[i] = b; key[p] = a; }
            }
            __syncthreads();
        }
    }
}

// FILL = false: records and steps of every cluster are counted (rec_count, step_count);
// FILL = true: the records and step headers are written at the scanned offsets.
template< bool FILL >
__global__ void __launch_bounds__(kBuildThreads)
cl_records_kernel(int32_t nclusters, const int32_t *__restrict__ cl_begin, const int32_t *__restrict__ cnodes,
                  const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc, const int32_t *__restrict__ conn,
                  const double *__restrict__ coords, const int32_t *__restrict__ matid, const int32_t *__restrict__ ncl,
                  const unsigned short *__restrict__ nbase, const unsigned char *__restrict__ nloc,
                  const unsigned char *__restrict__ nblk, const unsigned char *__restrict__ ebidx,
                  int32_t *__restrict__ rec_count, int32_t *__restrict__ step_count,
                  const int32_t *__restrict__ rec_off, const int32_t *__restrict__ step_off,
                  ClRecord *__restrict__ recs, ClBlob *__restrict__ blobs, int *__restrict__ flags,
                  const int32_t *__restrict__ nodeeq, const int32_t *__restrict__ rowptr, const unsigned short *__restrict__ blk, int maxblk,
                  const MatParams *__restrict__ mat, const unsigned char *__restrict__ npar)
{
    __shared__ ClBuildShared sh;
    const int tid = threadIdx.x;
    for ( int cl = blockIdx.x; cl < nclusters; cl += gridDim.x ) {
        __syncthreads();
        const int nb0 = cl_begin[cl], nn = cl_begin[cl + 1] - nb0;
        // the cluster's nodes, their visits -> keys = element numbers
        if ( tid < kClNodes ) {
            sh.nodes[tid] = tid < nn ? cnodes[nb0 + tid] : -1;
            sh.colormask[tid] = 0;
        }
        if ( tid == 0 ) {
            const int n = cnodes[nb0 + nn - 1];
            sh.nblocks = nbase[n] + nblk[n];
        }
        __syncthreads();
        // visit counts per node, prefix by thread 0 (<= 64 entries)
        __shared__ int vstart[kClNodes + 1];
        if ( tid == 0 ) {
            int s = 0;
            for ( int k = 0; k < nn; k++ ) {
                vstart[k] = s;
                s += ninc_start[sh.nodes[k] + 1] - ninc_start[sh.nodes[k]];
            }
            vstart[nn] = s;
        }
        __syncthreads();
        const int nvis = vstart[nn];
        if ( nvis > kClVisits ) {          // cannot happen when the valence limit holds; decline instead of overflowing
            if ( tid == 0 ) atomicAdd(flags, 1);
            if ( !FILL && tid == 0 ) { rec_count[cl] = 0; step_count[cl] = 0; }
            continue;
        }
        int n2 = 32;
        while ( n2 < nvis ) n2 <<= 1;
        for ( int t = tid; t < n2; t += kBuildThreads ) sh.key[t] = ~0ull;
        __syncthreads();
        for ( int k = tid >> 4; k < nn; k += kBuildThreads / 16 ) {      // 16 lanes per node
            const int s0 = ninc_start[sh.nodes[k]], cnt = vstart[k + 1] - vstart[k];
            for ( int v = tid & 15; v < cnt; v += 16 ) sh.key[vstart[k] + v] = (unsigned long long)( ninc[s0 + v] >> 3 );
        }
        __syncthreads();
        cl_bitonic(sh.key, n2, tid, kBuildThreads);
        // distinct elements, compacted in ascending order (thread 0: <= 1024 entries)
        if ( tid == 0 ) {
            int ne = 0;
            unsigned long long prev = ~0ull;
            for ( int t = 0; t < nvis; t++ ) {
                const unsigned long long k = sh.key[t];
                if ( k != prev ) sh.elem[ne++] = (int32_t) k;
                prev = k;
            }
            sh.ne = ne;
        }
        __syncthreads();
        const int ne = sh.ne;
        // cluster-local indices of every element's nodes
        for ( int t = tid; t < ne * 8; t += kBuildThreads ) {
            const int node = conn[(int64_t) sh.elem[t >> 3] * 8 + ( t & 7 )] - 1;
            sh.oloc[t >> 3][t & 7] = ( ncl[node] == cl ) ? nloc[node] : (unsigned char) 0xFF;
        }
        __syncthreads();
        // greedy colouring in ascending element order: elements of one colour share no cluster node
        if ( tid == 0 ) {
            for ( int t = 0; t < ne; t++ ) {
                unsigned int m = 0;
#pragma unroll
                for ( int a = 0; a < 8; a++ ) {
                    const int l = sh.oloc[t][a];
                    if ( l != 0xFF ) m |= sh.colormask[l];
                }
                int c = __ffs(~m) - 1;
                if ( c < 0 ) c = 31;
#pragma unroll
                for ( int a = 0; a < 8; a++ ) {
                    const int l = sh.oloc[t][a];
                    if ( l != 0xFF ) sh.colormask[l] |= 1u << c;
                }
                // order: material, colour, element
                sh.key[t] = ( (unsigned long long) (unsigned int) matid[sh.elem[t]] << 32 ) | ( (unsigned long long) c << 16 ) | (unsigned long long) t;
            }
        }
        __syncthreads();
        int m2 = 32;
        while ( m2 < ne ) m2 <<= 1;
        for ( int t = ne + tid; t < m2; t += kBuildThreads ) sh.key[t] = ~0ull;
        __syncthreads();
        cl_bitonic(sh.key, m2, tid, kBuildThreads);
        // record slots: the elements of one material form a step, padded to whole packets
        if ( tid == 0 ) {
            int slot = 0, ns = 0;
            unsigned int prevmat = 0xFFFFFFFFu;
            for ( int t = 0; t < ne; t++ ) {
                const unsigned int mat = (unsigned int)( sh.key[t] >> 32 );
                if ( mat != prevmat ) {
                    if ( ns > 0 ) {
                        slot = ( slot + 3 ) & ~3;
                        sh.stepinfo[ns - 1][1] = slot - sh.stepinfo[ns - 1][0];
                    }
                    if ( ns < 64 ) {
                        sh.stepinfo[ns][0] = slot;
                        sh.stepinfo[ns][2] = (int) mat;
                    }
                    ns++;
                    prevmat = mat;
                }
                sh.outidx[t] = (unsigned short) slot;
                slot++;
            }
            slot = ( slot + 3 ) & ~3;
            if ( ns > 0 && ns <= 64 ) sh.stepinfo[ns - 1][1] = slot - sh.stepinfo[ns - 1][0];
            sh.nslots = slot;
            sh.nsteps = ns;
        }
        __syncthreads();
        const int nslots = sh.nslots, nsteps = sh.nsteps;
        if ( nsteps > 64 ) {               // more than 64 materials in one cluster: decline
            if ( tid == 0 ) atomicAdd(flags, 1);
            if ( !FILL && tid == 0 ) { rec_count[cl] = 0; step_count[cl] = 0; }
            continue;
        }
        if ( !FILL ) {
            if ( tid == 0 ) { rec_count[cl] = nslots; step_count[cl] = nsteps; }
            continue;
        }
        for ( int t = tid; t < nslots; t += kBuildThreads ) sh.slot2ent[t] = -1;
        __syncthreads();
        for ( int t = tid; t < ne; t += kBuildThreads ) sh.slot2ent[sh.outidx[t]] = (short) t;
        __syncthreads();
        // Positions in the planes.  The contraction lanes of a half-warp add to the blocks (row node A, column node B) of four
        // row nodes x four column nodes of one element at a time; the position decides the shared-memory bank.  With
        // position = bank + 16 k and bank = 4 (alpha(A) + kappa(B)) + beta(B), alpha = the antipodal class of A's grid parities (distinct on
        // every face of a brick), beta = (z, x) parities of B (distinct on the column sets the slot order below produces), the
        // sixteen lanes hit sixteen banks on a structured mesh (simulated: 1.08 wavefronts per access instead of 2.1 with
        // node-contiguous positions); on an unstructured mesh the labels are merely a hash.  First toucher allocates.
        for ( int t = tid; t < sh.nblocks; t += kBuildThreads ) { sh.bpos[t] = 0xFFFF; sh.tbid[t] = 0xFFFF; }
        if ( tid < 16 ) sh.bankcnt[tid] = 0;
        __syncthreads();
        for ( int t = tid; t < ne * 8; t += kBuildThreads ) {
            const int ei = t >> 3, a = t & 7, la = sh.oloc[ei][a];
            if ( la == 0xFF ) continue;
            const int e = sh.elem[ei], nodeA = sh.nodes[la], base = nbase[nodeA];
            const int pa = npar[nodeA], xa = ( pa >> 2 ) & 1;
            const int alpha = 2 * ( ( ( pa >> 1 ) & 1 ) ^ xa ) + ( ( pa & 1 ) ^ xa );
            const unsigned char *bx = ebidx + ( (int64_t) e * 8 + a ) * 8;
            for ( int b = 0; b < 8; b++ ) {
                if ( bx[b] == 0xFF ) continue;
                const int bid = base + bx[b];
                if ( atomicCAS(&sh.bpos[bid], (unsigned short) 0xFFFF, (unsigned short) 0xFFFE) != 0xFFFF ) continue;
                const int lb = sh.oloc[ei][b];
                if ( lb != 0xFF ) {           // the transposed block, also accumulated in this cluster
                    const unsigned char tb = ebidx[( (int64_t) e * 8 + b ) * 8 + a];
                    if ( tb != 0xFF ) sh.tbid[bid] = (unsigned short)( nbase[sh.nodes[lb]] + tb );
                }
                const int pb = npar[conn[(int64_t) e * 8 + b] - 1];
                // kappa(B) (any function of the column node keeps the sixteen lanes apart) spreads the blocks of one row node
                // over all banks for the flush: y parity and the next bit of y (searched: at most 2 blocks of 16 share a bank)
                const int kappa = 2 * ( ( pb >> 1 ) & 1 ) + ( ( pb >> 4 ) & 1 );
                int bank = ( 4 * ( ( alpha + kappa ) & 3 ) + 2 * ( pb & 1 ) + ( ( pb >> 2 ) & 1 ) ) & 15, k = 0;
                for ( int tries = 0; tries < 16; tries++ ) {
                    k = atomicAdd(&sh.bankcnt[bank], 1);
                    if ( k < kBankCap ) break;
                    atomicSub(&sh.bankcnt[bank], 1);
                    bank = ( bank + 1 ) & 15;
                }
                sh.bpos[bid] = (unsigned short)( bank + 16 * k );          // nblocks <= kClBlocks < 16 kBankCap: some bank has room
            }
        }
        __syncthreads();
        for ( int st = 0; st < nsteps; st++ ) {
            const int s0 = sh.stepinfo[st][0], sn = sh.stepinfo[st][1];
            // first touch of every block in this step
            for ( int t = tid; t < sh.nblocks; t += kBuildThreads ) sh.minord[t] = INT_MAX;
            if ( tid < kClNodes * kCWarps ) ( &sh.last[0][0] )[tid] = 0;
            for ( int t = tid + kBuildThreads; t < kClNodes * kCWarps; t += kBuildThreads ) ( &sh.last[0][0] )[t] = 0;
            __syncthreads();
            for ( int t = tid; t < sn * 8; t += kBuildThreads ) {
                const int r = t >> 3, a = t & 7, ent = sh.slot2ent[s0 + r];
                if ( ent < 0 ) continue;
                const int ei = (int)( sh.key[ent] & 0xFFFF );
                const int la = sh.oloc[ei][a];
                if ( la == 0xFF ) continue;
                const int base = nbase[sh.nodes[la]];
                const unsigned char *bx = ebidx + ( (int64_t) sh.elem[ei] * 8 + a ) * 8;
#pragma unroll
                for ( int b = 0; b < 8; b++ )
                    if ( bx[b] != 0xFF ) atomicMin(&sh.minord[base + bx[b]], r);
            }
            // dependencies: for record r (warp r % kCWarps, its element number r / kCWarps + 1), per warp w the number of
            // elements w must have completed = the last earlier record of w sharing a cluster node
            if ( tid < kCWarps ) {
                for ( int r = 0; r < sn; r++ ) {
                    const int ent = sh.slot2ent[s0 + r];
                    int nd = 0;
                    if ( ent >= 0 ) {
                        const int ei = (int)( sh.key[ent] & 0xFFFF );
#pragma unroll
                        for ( int a = 0; a < 8; a++ ) {
                            const int la = sh.oloc[ei][a];
                            if ( la != 0xFF ) nd = max(nd, (int) sh.last[la][tid]);
                        }
                        if ( tid == r % kCWarps ) nd = 0;          // program order
                    }
                    recs[rec_off[cl] + s0 + r].need[tid] = (unsigned char) nd;
                    __syncwarp(( 1u << kCWarps ) - 1u);
                    if ( ent >= 0 && tid == r % kCWarps ) {
                        const int ei = (int)( sh.key[ent] & 0xFFFF );
#pragma unroll
                        for ( int a = 0; a < 8; a++ ) {
                            const int la = sh.oloc[ei][a];
                            if ( la != 0xFF ) sh.last[la][tid] = (unsigned char)( r / kCWarps + 1 );
                        }
                    }
                    __syncwarp(( 1u << kCWarps ) - 1u);
                }
            }
            __syncthreads();
            // the records of the step: four threads per record
            ClRecord *out = recs + rec_off[cl] + s0;
            for ( int t = tid; t < sn * 4; t += kBuildThreads ) {
                const int r = t >> 2, part = t & 3, ent = sh.slot2ent[s0 + r];
                ClRecord &R = out[r];
                if ( ent < 0 ) {
                    if ( part == 0 ) {
                        R.elem = -1;
                        R.slotof = 0x76543210u;
#pragma unroll
                        for ( int w = kCWarps; w < 16; w++ ) R.need[w] = 0;
                    }
#pragma unroll
                    for ( int q = 0; q < 16; q++ ) R.pos[16 * part + q] = 0xFFFF;
                    continue;
                }
                const int ei = (int)( sh.key[ent] & 0xFFFF );
                const int e = sh.elem[ei];
                // row slots: the element's cluster nodes first, the others behind; inside each group nodes whose x and y
                // grid parities agree take the even places (as far as there are any): the column sets {0,2,4,6} and
                // {1,3,5,7} of the contraction lanes then hold nodes with distinct (z, x) parities
                int slotof[8], nodeat[8];
                {
                    unsigned int own = 0, dl = 0;
#pragma unroll
                    for ( int a = 0; a < 8; a++ ) {
                        if ( sh.oloc[ei][a] != 0xFF ) own |= 1u << a;
                        const int pr = npar[conn[(int64_t) e * 8 + a] - 1];
                        if ( ( ( pr >> 2 ) ^ ( pr >> 1 ) ) & 1 ) dl |= 1u << a;
                    }
                    int k = 0;
#pragma unroll
                    for ( int grp = 0; grp < 2; grp++ ) {
                        const unsigned int in = grp == 0 ? own : ( ~own & 0xFFu );
                        unsigned int me = in & ~dl, mo = in & dl;
                        while ( me | mo ) {
                            if ( me ) { const int a = __ffs(me) - 1; slotof[a] = k; nodeat[k] = a; k++; me &= me - 1; }
                            if ( mo ) { const int a = __ffs(mo) - 1; slotof[a] = k; nodeat[k] = a; k++; mo &= mo - 1; }
                        }
                    }
                }
                // part p: slots 2p, 2p+1 (and the coordinates of nodes 2p, 2p+1, which stay in the element's own order)
#pragma unroll
                for ( int aa = 0; aa < 2; aa++ ) {
                    const int a0 = 2 * part + aa;
                    const int node0 = conn[(int64_t) e * 8 + a0] - 1;
#pragma unroll
                    for ( int d = 0; d < 3; d++ ) R.xyz[3 * a0 + d] = coords[(int64_t) node0 * 3 + d];
                    const int sl = 2 * part + aa;
                    const int a = nodeat[sl];
                    const int node = conn[(int64_t) e * 8 + a] - 1;
                    const int la = sh.oloc[ei][a];
                    const int base = la != 0xFF ? (int) nbase[node] : 0;
                    const unsigned char *bx = ebidx + ( (int64_t) e * 8 + a ) * 8;
                    for ( int b = 0; b < 8; b++ ) {
                        const unsigned char bi = bx[b];
                        unsigned short pv = 0xFFFF;
                        if ( la != 0xFF && bi != 0xFF ) pv = (unsigned short)( sh.bpos[base + bi] | ( sh.minord[base + bi] == r ? 0x8000 : 0 ) );
                        R.pos[sl * 8 + slotof[b]] = pv;
                    }
                }
                if ( part == 0 ) {
                    R.elem = e;
                    int nown = 0;
#pragma unroll
                    for ( int a = 0; a < 8; a++ ) nown += sh.oloc[ei][a] != 0xFF;
                    R.pad[0] = nown;
                    uint32_t w = 0;
#pragma unroll
                    for ( int a = 0; a < 8; a++ ) w |= (uint32_t) slotof[a] << ( 4 * a );
                    R.slotof = w;
#pragma unroll
                    for ( int w2 = kCWarps; w2 < 16; w2++ ) R.need[w2] = 0;
                }
            }
            // the flush tables of the step
            ClBlob &B = blobs[step_off[cl] + st];
            if ( tid == 0 ) {
                ClStep &S = B.hdr;
                S.rec_begin = rec_off[cl] + s0;
                S.npk = sn / 4;
                S.node_begin = nb0;
                S.nnodes = nn;
                S.matid = sh.stepinfo[st][2];
                S.flags = ( st > 0 ? 1 : 0 ) | ( nsteps > 1 ? 2 : 0 ) | ( st == nsteps - 1 ? 4 : 0 );
                S.nblocks = sh.nblocks;
                S.nsteps = nsteps;
                const MatParams &mp = mat[sh.stepinfo[st][2]];
                isole_lame(mp.E, mp.nu, B.lam, B.mu);
            }
            for ( int k = tid; k < nn; k += kBuildThreads ) {
                const int w = sh.nodes[k], nb = nblk[w], base = nbase[w];
#pragma unroll
                for ( int i = 0; i < 3; i++ ) {
                    const int eq = nodeeq[(int64_t) w * 3 + i];
                    B.rowbase[k][i] = eq > 0 ? rowptr[eq - 1] : -1;
                }
                bool regular = true;
                for ( int n = 0; n < nb; n++ ) {
                    const int cm = blk[(int64_t) w * maxblk + n] >> 8;
                    regular = regular && cm == 7;
                    const int tb = sh.tbid[base + n];
                    B.postab[base + n] = (uint32_t) sh.bpos[base + n] | ( (uint32_t)( tb != 0xFFFF ? sh.bpos[tb] : 0 ) << 11 ) |
                                         ( (uint32_t) cm << 22 ) | ( tb != 0xFFFF ? 1u << 25 : 0u ) | ( (uint32_t) k << 26 );
                }
                B.rowbase[k][3] = base | ( regular ? 0 : (int) 0x80000000 );
            }
            __syncthreads();
        }
    }
}

// ---- the assembly kernel --------------------------------------------------------------------------------

struct ClShared {
    double acc[9 * kPlane];
    double H[kHSlots][4][kHStride];
    ClRecord rec[kChunks][kChunkPk * 4];
    ClBlob blob[2];
    unsigned long long full_p[kChunks], empty_r[kChunks], full_g[kHSlots], empty_h[kHSlots], blob_full[2], blob_empty[2];
    volatile unsigned int done[kCWarps];
    volatile unsigned int failed;            // a bounded wait ran out
};

__device__ __forceinline__ uint32_t cl_smem(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void cl_mbar_init(unsigned long long *bar, int count)
{
    asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( cl_smem(bar) ), "r"( count ) );
}
__device__ __forceinline__ void cl_mbar_arrive(unsigned long long *bar)
{
    asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( cl_smem(bar) ) : "memory" );
}
__device__ __forceinline__ void cl_mbar_arrive_n(unsigned long long *bar, uint32_t n)
{
    asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"( cl_smem(bar) ), "r"( n ) : "memory" );
}
__device__ __forceinline__ void cl_mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( cl_smem(bar) ), "r"( bytes ) : "memory" );
}
// Every wait in the kernel is bounded (about 2 s): a schedule or hardware fault then ends the launch with the error word set
// instead of hanging the GPU; cluster_assemble_lspace reports it as OB200_ECUDA.
__device__ int *cl_error_word = nullptr;
__device__ __forceinline__ void cl_fail(int code)
{
    if ( cl_error_word ) atomicExch(cl_error_word, code);
}
#ifndef OB200_CL_RELAX_NS
#define OB200_CL_RELAX_NS 200
#endif
#ifndef OB200_CL_WAIT_HINT_NS
#define OB200_CL_WAIT_HINT_NS 20000
#endif
#ifndef OB200_CL_BOUNDED
#define OB200_CL_BOUNDED 1
#endif
constexpr unsigned int kClSpinLimit = OB200_CL_BOUNDED ? 1u << 20 : 0xFFFFFFFFu;      // try_wait rounds (each suspends for a hardware-defined time) / polls
// slow path of cl_mbar_wait: bounded polling (kept out of line: the contraction warps' code stays as small as with a bare loop)
__device__ __noinline__ void cl_mbar_wait_slow(uint32_t bar, uint32_t parity, volatile unsigned int *failed)
{
    for ( unsigned int n = 0; n < kClSpinLimit; n++ ) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"( ok ) : "r"( bar ), "r"( parity ) : "memory" );
        if ( ok ) return;
    }
    *failed = 1u;
}
// the contraction warps' wait: almost always the phase is complete already (the other roles run ahead)
__device__ __forceinline__ void cl_mbar_wait(unsigned long long *bar, uint32_t parity, volatile unsigned int *failed)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"( ok ) : "r"( cl_smem(bar) ), "r"( parity ) : "memory" );
    if ( !ok ) cl_mbar_wait_slow(cl_smem(bar), parity, failed);
}
// The wait of the warps that run ahead of the contraction warps (producer thread, geometry warps): try_wait returns quickly when
// the phase is not complete, so a bare loop spins at full issue rate on a scheduler it shares with two contraction warps (the
// five-instruction bounded loop above cost 12 % of the kernel when every role used it).  Here the poll is followed by a sleep.
__device__ __forceinline__ void cl_mbar_wait_relaxed(unsigned long long *bar, uint32_t parity, volatile unsigned int *failed)
{
    for ( unsigned int n = 0; n < kClSpinLimit; n++ ) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"( ok ) : "r"( cl_smem(bar) ), "r"( parity ) : "memory" );
        if ( ok ) return;
        __nanosleep(OB200_CL_RELAX_NS);
    }
    *failed = 1u;
}
__device__ __forceinline__ void cl_bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  ::"r"( cl_smem(dst) ), "l"( src ), "r"( bytes ), "r"( cl_smem(bar) ) : "memory" );
}
__device__ __forceinline__ void cl_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"( c0 ), "+d"( c1 ) : "d"( a ), "d"( b ) );
}

// One thread per (element, Gauss point): H[ks][3a+i][gp & 3] = sqrt(|det J|) dN_a/dx_i, ks = gp >> 2.
// FEI3dHexaLin::evaldNdxi (fei3dhexalin.C:129-166) at the 2x2x2 rule (gaussintegrationrule.C:190-214, weights 1);
// J = x dN/dxi, dN/dx = dN/dxi J^-1 (evaldNdx, fei3dhexalin.C:186-204); dV = |det J|
// (Structural3DElement::computeVolumeAround, structural3delement.C:328-338).  J^-1 sqrt|det J| = adj(J) sign(det) / sqrt|det|:
// one reciprocal square root instead of a division and a square root.
__device__ __forceinline__ void cl_geometry(const double *__restrict__ xv, int gp, uint32_t slotof, double *__restrict__ Hel)
{
    constexpr double kA = 0.577350269189626;
    constexpr double pp = 0.125 * ( 1.0 + kA ) * ( 1.0 + kA ), pm = 0.125 * ( 1.0 + kA ) * ( 1.0 - kA ), mm = 0.125 * ( 1.0 - kA ) * ( 1.0 - kA );
    const bool gu = ( gp & 4 ) != 0, gv = ( gp & 2 ) != 0, gw = ( gp & 1 ) != 0;
    // P..[s][t] = (1 + s' u)(1 + t' v)/8 with s' = +1 for index 1: the factor is 1 + kA when the sign agrees with the Gauss point's
    double Pyz[2][2], Pxz[2][2], Pxy[2][2];
#pragma unroll
    for ( int s = 0; s < 2; s++ )
#pragma unroll
        for ( int t = 0; t < 2; t++ ) {
            const bool ay = ( s == 1 ) == gv, az = ( t == 1 ) == gw, ax = ( s == 1 ) == gu, ay2 = ( t == 1 ) == gv;
            Pyz[s][t] = ay ? ( az ? pp : pm ) : ( az ? pm : mm );
            Pxz[s][t] = ax ? ( az ? pp : pm ) : ( az ? pm : mm );
            Pxy[s][t] = ax ? ( ay2 ? pp : pm ) : ( ay2 ? pm : mm );
        }
    double c[24];
    const double2 *xv2 = reinterpret_cast< const double2 * >( xv );
#pragma unroll
    for ( int i = 0; i < 12; i++ ) {
        const double2 v = xv2[i];
        c[2 * i] = v.x;
        c[2 * i + 1] = v.y;
    }
    double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
    double dN[8][3];
#pragma unroll
    for ( int kk = 0; kk < 8; kk++ ) {
        // signs of node kk in (xi, eta, zeta) -- the node order of FEI3dHexaLin (fei3dhexalin.C:46-63)
        const int px = ( kk & 3 ) >= 2, py = ( ( kk & 3 ) == 1 || ( kk & 3 ) == 2 ), pz = kk < 4;
        dN[kk][0] = px ? Pyz[py][pz] : -Pyz[py][pz];
        dN[kk][1] = py ? Pxz[px][pz] : -Pxz[px][pz];
        dN[kk][2] = pz ? Pxy[px][py] : -Pxy[px][py];
        const double x = c[3 * kk], y = c[3 * kk + 1], z = c[3 * kk + 2];
        J[0][0] += x * dN[kk][0]; J[0][1] += x * dN[kk][1]; J[0][2] += x * dN[kk][2];
        J[1][0] += y * dN[kk][0]; J[1][1] += y * dN[kk][1]; J[1][2] += y * dN[kk][2];
        J[2][0] += z * dN[kk][0]; J[2][1] += z * dN[kk][1]; J[2][2] += z * dN[kk][2];
    }
    // adjugate (FloatMatrix::beInverseOf 3x3, floatmatrix.C:790-808, without the division)
    double A[3][3];
    A[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    A[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    A[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    A[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    A[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    A[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    A[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    A[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    A[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * A[0][0] + J[0][1] * A[1][0] + J[0][2] * A[2][0];
    double s = rsqrt(fabs(det));
    if ( det < 0.0 ) s = -s;
#pragma unroll
    for ( int i = 0; i < 3; i++ )
#pragma unroll
        for ( int j = 0; j < 3; j++ ) A[i][j] *= s;
    double *o = Hel + ( gp >> 2 ) * 100 + ( gp & 3 );
#pragma unroll
    for ( int kk = 0; kk < 8; kk++ ) {
        double *ok = o + 12 * ( ( slotof >> ( 4 * kk ) ) & 7u );          // row slot of node kk
#pragma unroll
        for ( int j = 0; j < 3; j++ ) ok[4 * j] = dN[kk][0] * A[0][j] + dN[kk][1] * A[1][j] + dN[kk][2] * A[2][j];
    }
}

struct ClView {
    const ClRecord *recs;
    const ClBlob *blobs;          // one per step
    const int32_t *cl_step;       // [nclusters + 1]: first step of every cluster
    int32_t nclusters;
};

// Walks the steps of this CTA's clusters (cluster c goes to CTA c mod gridDim) one cluster ahead of its use: the index of
// the next cluster's first step is requested while the current cluster is being worked on.
struct ClWalk {
    const ClView &V;
    int cl, st, st_end, nst, nst_end;
    __device__ ClWalk(const ClView &v) : V(v)
    {
        cl = blockIdx.x;
        st = st_end = nst = nst_end = 0;
        if ( cl < V.nclusters ) { st = V.cl_step[cl]; st_end = V.cl_step[cl + 1]; }
        prefetch();
    }
    __device__ void prefetch()
    {
        const int n = cl + gridDim.x;
        nst = nst_end = 0;
        if ( n < V.nclusters ) { nst = V.cl_step[n]; nst_end = V.cl_step[n + 1]; }
    }
    __device__ bool valid() const { return cl < V.nclusters; }
    __device__ void next()
    {
        if ( ++st < st_end ) return;
        cl += gridDim.x;
        st = nst;
        st_end = nst_end;
        if ( cl < V.nclusters ) prefetch();
    }
};

template< bool ACCUM >
__global__ void __launch_bounds__(kClThreads, kClCtas)
lspace_cluster_kernel(ElemSetView S, ClView V, double *__restrict__ val)
{
    extern __shared__ __align__(128) unsigned char cl_smem_raw[];
    ClShared &sh = *reinterpret_cast< ClShared * >( cl_smem_raw );
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if ( tid == 0 ) {
        for ( int s = 0; s < kChunks; s++ ) {
            cl_mbar_init(&sh.full_p[s], 1);
            cl_mbar_init(&sh.empty_r[s], kChunkPk * 4);
        }
        for ( int s = 0; s < kHSlots; s++ ) {
            cl_mbar_init(&sh.full_g[s], 1);
            cl_mbar_init(&sh.empty_h[s], 4);
        }
        for ( int b = 0; b < 2; b++ ) {
            cl_mbar_init(&sh.blob_full[b], 1);
            cl_mbar_init(&sh.blob_empty[b], 1);
        }
        asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
    }
    if ( tid < kCWarps ) sh.done[tid] = 0;
    if ( tid == 0 ) sh.failed = 0;
    __syncthreads();

    if ( wid == kCWarps ) {
        // ---- producer: one thread streams the flush tables of every step and its record packets into shared memory ----
        if ( lane != 0 ) return;
        unsigned int cseq = 0, bseq = 0;
        for ( ClWalk w(V); w.valid(); w.next(), bseq++ ) {
            const ClBlob *blob = V.blobs + w.st;
            const int rec_begin = blob->hdr.rec_begin, npk = blob->hdr.npk, nblocks = blob->hdr.nblocks;
            const int b = bseq & 1;
            if ( bseq >= 2 ) cl_mbar_wait_relaxed(&sh.blob_empty[b], ( ( bseq >> 1 ) - 1 ) & 1, &sh.failed);
            const uint32_t bytes = (uint32_t)( kBlobHead + ( ( nblocks * 4 + 15 ) & ~15 ) );
            cl_mbar_expect_tx(&sh.blob_full[b], bytes);
            cl_bulk_load(&sh.blob[b], blob, bytes, &sh.blob_full[b]);
            for ( int k = 0; k < npk; k += kChunkPk, cseq++ ) {
                const int s = cseq % kChunks, n = min(kChunkPk, npk - k);
                const unsigned int round = cseq / kChunks;
                if ( round > 0 ) cl_mbar_wait_relaxed(&sh.empty_r[s], ( round - 1 ) & 1, &sh.failed);
                if ( n < kChunkPk ) cl_mbar_arrive_n(&sh.empty_r[s], ( kChunkPk - n ) * 4);      // the records a short chunk lacks
                cl_mbar_expect_tx(&sh.full_p[s], n * kPacketBytes);
                cl_bulk_load(sh.rec[s], V.recs + rec_begin + 4 * k, n * kPacketBytes, &sh.full_p[s]);
            }
        }
        if ( sh.failed ) cl_fail((int) sh.failed);
        return;
    }

    if ( wid > kCWarps ) {
        // ---- geometry warps: lane = 8 * (element of the packet) + Gauss point ----
        const int g = wid - kCWarps - 1;
        unsigned int pseq = 0, cseq0 = 0;
        for ( ClWalk w(V); w.valid(); w.next() ) {
            const int npk = V.blobs[w.st].hdr.npk;
            for ( int k = 0; k < npk; k++, pseq++ ) {
                if ( (int)( pseq % kGWarps ) != g ) continue;
                const unsigned int cseq = cseq0 + k / kChunkPk;
                const int cs = cseq % kChunks, hs = pseq % kHSlots;
                cl_mbar_wait_relaxed(&sh.full_p[cs], ( cseq / kChunks ) & 1, &sh.failed);
                if ( pseq >= kHSlots ) cl_mbar_wait_relaxed(&sh.empty_h[hs], ( pseq / kHSlots - 1 ) & 1, &sh.failed);
                const ClRecord &R = sh.rec[cs][( k % kChunkPk ) * 4 + ( lane >> 3 )];
                if ( R.elem >= 0 && !( kAblate & 8 ) ) cl_geometry(R.xyz, lane & 7, R.slotof, sh.H[hs][lane >> 3]);
                __syncwarp();
                if ( lane == 0 ) cl_mbar_arrive(&sh.full_g[hs]);
            }
            cseq0 += ( npk + kChunkPk - 1 ) / kChunkPk;
        }
        if ( sh.failed ) cl_fail((int) sh.failed);
        return;
    }

    // ---- contraction warps ----
    const int a = lane >> 2, bp = lane & 3, b0 = 2 * bp, b1 = b0 + 1;
    unsigned int pseq0 = 0, cseq0 = 0;      // packet / chunk sequence numbers of the current step's first packet
    const int nmine = V.nclusters > (int) blockIdx.x ? ( V.nclusters - 1 - (int) blockIdx.x ) / (int) gridDim.x + 1 : 0;
    unsigned int bseq = 0;
    for ( int kcl = 0; kcl < nmine; kcl++ ) {
        bool last = false;
        while ( !last ) {
            const int bb = bseq & 1;
            cl_mbar_wait(&sh.blob_full[bb], ( bseq >> 1 ) & 1, &sh.failed);
            const ClBlob &blob = sh.blob[bb];
            const ClStep step = blob.hdr;
            last = ( step.flags & 4 ) != 0;
            if ( step.flags & 2 ) {         // several materials in this cluster: positions a step does not touch must read as zero
                for ( int t = tid; t < 9 * kPlane; t += kCWarps * 32 ) sh.acc[t] = 0.0;
                asm volatile( "bar.sync 1, %0;" ::"n"( kCWarps * 32 ) : "memory" );
            }
            const int nrec = step.npk * 4;
            unsigned int mydone = 0;
            for ( int r = wid; r < nrec; r += kCWarps ) {
                const unsigned int pseq = pseq0 + ( r >> 2 ), cseq = cseq0 + ( r >> 2 ) / kChunkPk;
                const int rs = cseq % kChunks, hs = pseq % kHSlots;
                cl_mbar_wait(&sh.full_p[rs], ( cseq / kChunks ) & 1, &sh.failed);        // the records (bulk copy)
                cl_mbar_wait(&sh.full_g[hs], ( pseq / kHSlots ) & 1, &sh.failed);        // the gradients (geometry warp)
                const ClRecord &R = sh.rec[rs][r % ( kChunkPk * 4 )];
                if ( R.elem >= 0 ) {
                    const double *H = sh.H[hs][r & 3] + 12 * a + bp;
                    // fragments: h[i][ks] = H[ks][3a+i][bp] -- both the A fragment of component i and the B fragment of component i
                    double h[3][2];
#pragma unroll
                    for ( int i = 0; i < 3; i++ ) {
                        h[i][0] = H[4 * i];
                        h[i][1] = H[100 + 4 * i];
                    }
                    // row slot a of this lane, column slots b0, b1
                    // lanes whose row node (slot a) belongs to the cluster add to block (a, b); lanes whose row node does not,
                    // but whose column node does, add the off-diagonal products to the transposed planes of block (b, a) -- for
                    // pairs inside the cluster the flush reads entry (j,i) of block (a,b) as entry (i,j) of block (b,a) instead
                    const bool rown = a < R.pad[0];
                    const unsigned int pp = rown ? *reinterpret_cast< const unsigned int * >( &R.pos[a * 8 + b0] )
                                                 : (unsigned int) R.pos[b0 * 8 + a] | ( (unsigned int) R.pos[b1 * 8 + a] << 16 );
                    const unsigned int nd = lane < kCWarps ? R.need[lane] : 0u;
                    // the six products (accumulators start at zero; two k-steps of four Gauss points)
                    double d[6][2];
#pragma unroll
                    for ( int t = 0; t < 6; t++ ) d[t][0] = d[t][1] = 0.0;
#pragma unroll
                    for ( int ks = 0; ks < ( ( kAblate & 2 ) ? 0 : 2 ); ks++ ) {
                        cl_dmma(d[0][0], d[0][1], h[0][ks], h[0][ks]);     // (0,0)
                        cl_dmma(d[1][0], d[1][1], h[1][ks], h[1][ks]);     // (1,1)
                        cl_dmma(d[2][0], d[2][1], h[2][ks], h[2][ks]);     // (2,2)
                        cl_dmma(d[3][0], d[3][1], h[0][ks], h[1][ks]);     // (0,1)
                        cl_dmma(d[4][0], d[4][1], h[0][ks], h[2][ks]);     // (0,2)
                        cl_dmma(d[5][0], d[5][1], h[1][ks], h[2][ks]);     // (1,2)
                    }
                    const bool v0 = ( pp & 0xFFFFu ) != 0xFFFFu && !( kAblate & 1 ), v1 = ( pp >> 16 ) != 0xFFFFu && !( kAblate & 1 );
                    // a first touch overwrites: the old value is then not read
                    const bool l0 = v0 && !( pp & 0x8000u ), l1 = v1 && !( pp & 0x80000000u );
                    // planes 0..5: entries (0,0) (1,1) (2,2) (0,1) (0,2) (1,2); planes 6..8: (1,0) (2,0) (2,1) of blocks whose column
                    // node is outside the cluster.  A row-owning lane starts at plane 0, the others at plane 6 with products 3..5.
                    double *const p0 = sh.acc + ( rown ? 0 : 6 * kPlane ) + ( pp & 0x7FFFu );
                    double *const p1 = sh.acc + ( rown ? 0 : 6 * kPlane ) + ( ( pp >> 16 ) & 0x7FFFu );
                    // wait until every earlier element sharing a cluster node with this one has been added
                    if ( nd && !( kAblate & 16 ) ) {
                        unsigned int spins = 0;
                        while ( sh.done[lane] < nd ) {
                            __nanosleep(OB200_CL_SPIN_NS);
                            if ( ++spins > kClSpinLimit ) {
                                sh.failed = 2u;
                                break;
                            }
                        }
                    }
                    __syncwarp();
                    __threadfence_block();
                    // All loads first, then the additions, then the stores: the positions of one lane are distinct.
                    double o0[6], o1[6];
#pragma unroll
                    for ( int t = 0; t < 6; t++ ) {
                        o0[t] = 0.0;
                        o1[t] = 0.0;
                        if ( t < 3 || rown ) {
                            if ( l0 ) o0[t] = p0[t * kPlane];
                            if ( l1 ) o1[t] = p1[t * kPlane];
                        }
                    }
#pragma unroll
                    for ( int t = 0; t < 3; t++ ) {
                        const double x0 = rown ? d[t][0] : d[3 + t][0], x1 = rown ? d[t][1] : d[3 + t][1];
                        if ( v0 ) p0[t * kPlane] = o0[t] + x0;
                        if ( v1 ) p1[t * kPlane] = o1[t] + x1;
                    }
#pragma unroll
                    for ( int t = 3; t < 6; t++ ) {
                        if ( rown && v0 ) p0[t * kPlane] = o0[t] + d[t][0];
                        if ( rown && v1 ) p1[t * kPlane] = o1[t] + d[t][1];
                    }
                }
                mydone++;
                __syncwarp();
                if ( lane == 0 ) {
                    __threadfence_block();
                    sh.done[wid] = mydone;
                    cl_mbar_arrive(&sh.empty_r[rs]);
                    cl_mbar_arrive(&sh.empty_h[hs]);
                }
            }
            pseq0 += step.npk;
            cseq0 += ( step.npk + kChunkPk - 1 ) / kChunkPk;
            asm volatile( "bar.sync 1, %0;" ::"n"( kCWarps * 32 ) : "memory" );

            // ---- flush: one thread per 3x3 block, in row order (consecutive threads write consecutive column blocks of a node) ----
            const double lam = blob.lam, mu = blob.mu;
            const bool add = ACCUM || ( step.flags & 1 );
#pragma unroll ( kFlushUnroll )
            for ( int q = tid; q < ( ( kAblate & 4 ) ? 0 : step.nblocks ); q += kCWarps * 32 ) {
                const uint32_t info = blob.postab[q];
                const int cm = ( info >> 22 ) & 7;
                const int4 rb = *reinterpret_cast< const int4 * >( blob.rowbase[info >> 26] );
                const int noff = rb.w & 0xFFFF;
                // column offset of the block in its row: three per block unless a neighbouring node has prescribed dofs
                int cstart = 3 * ( q - noff );
                if ( rb.w < 0 ) {
                    cstart = 0;
                    for ( int n = noff; n < q; n++ ) cstart += __popc(( blob.postab[n] >> 22 ) & 7u);
                }
                const double *pa = sh.acc + ( info & 0x7FF );
                // entries below the diagonal: from the transposed block if its row node is in the cluster, else from planes 6..8
                const double *pl = ( info >> 25 ) & 1 ? sh.acc + 3 * kPlane + ( ( info >> 11 ) & 0x7FF ) : pa + 6 * kPlane;
                double g[9];
                g[0] = pa[0];
                g[4] = pa[kPlane];
                g[8] = pa[2 * kPlane];
                g[1] = pa[3 * kPlane];
                g[2] = pa[4 * kPlane];
                g[5] = pa[5 * kPlane];
                g[3] = pl[0];
                g[6] = pl[kPlane];
                g[7] = pl[2 * kPlane];
                const double tr = mu * ( g[0] + g[4] + g[8] );
                double kb[9];
                kb[0] = lam * g[0] + mu * g[0] + tr; kb[1] = lam * g[1] + mu * g[3];      kb[2] = lam * g[2] + mu * g[6];
                kb[3] = lam * g[3] + mu * g[1];      kb[4] = lam * g[4] + mu * g[4] + tr; kb[5] = lam * g[5] + mu * g[7];
                kb[6] = lam * g[6] + mu * g[2];      kb[7] = lam * g[7] + mu * g[5];      kb[8] = lam * g[8] + mu * g[8] + tr;
                const int rowb[3] = { rb.x, rb.y, rb.z };
#pragma unroll
                for ( int i = 0; i < 3; i++ ) {
                    if ( rowb[i] < 0 ) continue;
                    double *dst = val + rowb[i] + cstart;
                    if ( cm == 7 ) {
                        if ( add ) { dst[0] += kb[3 * i]; dst[1] += kb[3 * i + 1]; dst[2] += kb[3 * i + 2]; }
                        else { dst[0] = kb[3 * i]; dst[1] = kb[3 * i + 1]; dst[2] = kb[3 * i + 2]; }
                    } else {
                        int c = 0;
#pragma unroll
                        for ( int j = 0; j < 3; j++ )
                            if ( cm & ( 1 << j ) ) {
                                dst[c] = add ? dst[c] + kb[3 * i + j] : kb[3 * i + j];
                                c++;
                            }
                    }
                }
            }
            if ( tid < kCWarps ) sh.done[tid] = 0;
            asm volatile( "bar.sync 1, %0;" ::"n"( kCWarps * 32 ) : "memory" );
            if ( tid == 0 ) cl_mbar_arrive(&sh.blob_empty[bb]);
            bseq++;
        }
    }
    if ( sh.failed ) cl_fail((int) sh.failed);
}

// ---- host side ---------------------------------------------------------------------------------------------

// Build the cluster schedule for the bound matrix.  Needs nblk / blk / ebidx from node_blocks_kernel (gather_bind).
int cluster_bind(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    S->cluster_ok = false;
    const bool disabled = getenv("OB200_ASSEMBLY") && !strcmp(getenv("OB200_ASSEMBLY"), "gather");
    if ( disabled || S->etype != OB200_LSPACE || S->nelem == 0 || !S->all_isole || !S->ebidx.p ) return OB200_OK;
    if ( S->maxval > kClMaxValence || S->maxblk > 255 ) return OB200_OK;
    const int64_t nnode = S->nnode;
    // bounding box of the nodes that own rows, mean element size
    DevBuf< unsigned long long > box;
    OB_CHECK( box.alloc(8) );
    unsigned long long hbox[8] = { ~0ull, ~0ull, ~0ull, 0, 0, 0, 0, 0 };
    OB_CUDA( cudaMemcpyAsync(box.p, hbox, sizeof( hbox ), cudaMemcpyHostToDevice, ctx->stream) );
    OB_LAUNCH(ctx, cl_bbox_kernel, ctx->shape.grid(nnode, 256, 4), 256, 0, S->coords.p, S->nblk.p, nnode, box.p);
    OB_CUDA( cudaMemcpyAsync(hbox, box.p, sizeof( hbox ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    if ( hbox[0] == ~0ull ) return OB200_OK;             // no node owns a row
    double lo[3], hi[3], diag = 0.0;
    for ( int d = 0; d < 3; d++ ) {
        lo[d] = order_key_inv(hbox[d]);
        hi[d] = order_key_inv(hbox[3 + d]);
        diag += ( hi[d] - lo[d] ) * ( hi[d] - lo[d] );
    }
    diag = sqrt(diag);
    if ( !( diag > 0.0 ) || !isfinite(diag) ) return OB200_OK;
    const double scale = 4294967296.0 / diag;
    OB_LAUNCH(ctx, cl_extent_kernel, ctx->shape.grid(S->nelem, 256, 4), 256, 0, S->coords.p, S->conn.p, S->nelem, scale, box.p + 6);
    OB_CUDA( cudaMemcpyAsync(hbox + 6, box.p + 6, sizeof( unsigned long long ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    double h = (double) hbox[6] / 4294967296.0 * diag / (double) S->nelem;
    if ( !( h > 0.0 ) ) return OB200_OK;
    // cells of 4h, the first one starting half an element before the box; at most 2^24 cells
    ClGrid g;
    double cs = 4.0 * h;
    for ( ;; ) {
        double total = 1.0;
        for ( int d = 0; d < 3; d++ ) {
            g.x0[d] = lo[d] - 0.125 * cs;
            const double n = floor(( hi[d] - g.x0[d] ) / cs) + 1.0;
            g.dim[d] = n < 1.0 ? 1 : ( n > 1e9 ? 1000000000 : (int) n );
            total *= (double) g.dim[d];
        }
        if ( total <= 16777216.0 ) break;
        cs *= 2.0;
    }
    g.inv = 1.0 / cs;
    const int32_t ncell = g.dim[0] * g.dim[1] * g.dim[2];
    DevBuf< int32_t > cell, ccount, cfill, cstart, nclus, cl_off;
    DevBuf< int64_t > s64;
    OB_CHECK( cell.alloc(nnode) );
    OB_CHECK( ccount.alloc(ncell + 1) );
    OB_CHECK( cfill.alloc(ncell + 1) );
    OB_CHECK( cstart.alloc(ncell + 1) );
    OB_CHECK( s64.alloc(ncell + 1) );
    OB_CUDA( cudaMemsetAsync(ccount.p, 0, sizeof( int32_t ) * ( ncell + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(cfill.p, 0, sizeof( int32_t ) * ( ncell + 1 ), ctx->stream) );
    OB_CHECK( S->npar.alloc(nnode) );
    OB_LAUNCH(ctx, cl_cell_kernel, ctx->shape.grid(nnode, 256, 4), 256, 0, S->coords.p, S->nblk.p, nnode, g, cell.p, ccount.p, S->npar.p);
    int32_t maxcell = 0;
    OB_CHECK( max_reduce(ctx, ccount.p, ncell, &maxcell) );
    if ( maxcell > kClMaxCell ) return OB200_OK;
    int64_t nowned = 0;
    OB_CHECK( exclusive_scan(ctx, ccount.p, s64.p, (int64_t) ncell + 1, &nowned) );
    OB_CHECK( narrow_i64_to_i32(ctx, s64.p, cstart.p, (int64_t) ncell + 1) );
    OB_CHECK( S->cnodes.alloc(nowned > 0 ? nowned : 1) );
    OB_LAUNCH(ctx, cl_cell_fill_kernel, ctx->shape.grid(nnode, 256, 4), 256, 0, cell.p, nnode, cstart.p, cfill.p, S->cnodes.p);
    OB_CHECK( nclus.alloc(ncell + 1) );
    OB_CHECK( cl_off.alloc(ncell + 1) );
    OB_CUDA( cudaMemsetAsync(nclus.p, 0, sizeof( int32_t ) * ( ncell + 1 ), ctx->stream) );
    OB_CHECK( S->ncl.alloc(nnode) );
    OB_CHECK( S->nbase.alloc(nnode) );
    OB_CHECK( S->nloc.alloc(nnode) );
    OB_CUDA( cudaMemsetAsync(S->ncl.p, 0xFF, sizeof( int32_t ) * (size_t) nnode, ctx->stream) );
    OB_LAUNCH(ctx, cl_split_kernel< false >, ctx->shape.grid(ncell, 128, 8), 128, 0, ncell, cstart.p, S->cnodes.p, S->nblk.p, nclus.p,
              (const int32_t *) nullptr, (int32_t *) nullptr, (int32_t *) nullptr, (unsigned short *) nullptr, (unsigned char *) nullptr);
    int64_t nclusters = 0;
    OB_CHECK( exclusive_scan(ctx, nclus.p, s64.p, (int64_t) ncell + 1, &nclusters) );
    OB_CHECK( narrow_i64_to_i32(ctx, s64.p, cl_off.p, (int64_t) ncell + 1) );
    if ( nclusters == 0 ) return OB200_OK;
    OB_CHECK( S->cl_begin.alloc(nclusters + 1) );
    OB_LAUNCH(ctx, cl_split_kernel< true >, ctx->shape.grid(ncell, 128, 8), 128, 0, ncell, cstart.p, S->cnodes.p, S->nblk.p, nclus.p,
              cl_off.p, S->cl_begin.p, S->ncl.p, S->nbase.p, S->nloc.p);
    const int32_t nown32 = (int32_t) nowned;
    OB_CUDA( cudaMemcpyAsync(S->cl_begin.p + nclusters, &nown32, sizeof( int32_t ), cudaMemcpyHostToDevice, ctx->stream) );
    // records: count, scan, fill
    DevBuf< int32_t > rcount, scount, roff;
    DevBuf< int > flags;
    OB_CHECK( rcount.alloc(nclusters + 1) );
    OB_CHECK( scount.alloc(nclusters + 1) );
    OB_CHECK( roff.alloc(nclusters + 1) );
    OB_CHECK( S->cl_step.alloc(nclusters + 1) );
    OB_CHECK( flags.alloc(1) );
    OB_CUDA( cudaMemsetAsync(rcount.p, 0, sizeof( int32_t ) * ( nclusters + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(scount.p, 0, sizeof( int32_t ) * ( nclusters + 1 ), ctx->stream) );
    OB_CUDA( cudaMemsetAsync(flags.p, 0, sizeof( int ), ctx->stream) );
    const int bgrid = (int)( nclusters < (int64_t) ctx->shape.sms * 8 ? nclusters : (int64_t) ctx->shape.sms * 8 );
    OB_LAUNCH(ctx, cl_records_kernel< false >, bgrid, kBuildThreads, 0, (int32_t) nclusters, S->cl_begin.p, S->cnodes.p, S->ninc_start.p,
              S->ninc.p, S->conn.p, S->coords.p, S->matid.p, S->ncl.p, S->nbase.p, S->nloc.p, S->nblk.p, S->ebidx.p, rcount.p, scount.p,
              (const int32_t *) nullptr, (const int32_t *) nullptr, (ClRecord *) nullptr, (ClBlob *) nullptr, flags.p,
              S->nodeeq.p, A->rowptr.p, S->blk.p, S->maxblk, (const MatParams *) S->mat.p, S->npar.p);
    int64_t nrec = 0, nsteps = 0;
    DevBuf< int64_t > c64;
    OB_CHECK( c64.alloc(nclusters + 1) );
    OB_CHECK( exclusive_scan(ctx, rcount.p, c64.p, nclusters + 1, &nrec) );
    OB_CHECK( narrow_i64_to_i32(ctx, c64.p, roff.p, nclusters + 1) );
    OB_CHECK( exclusive_scan(ctx, scount.p, c64.p, nclusters + 1, &nsteps) );
    OB_CHECK( narrow_i64_to_i32(ctx, c64.p, S->cl_step.p, nclusters + 1) );
    int hflag = 0;
    OB_CUDA( cudaMemcpyAsync(&hflag, flags.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    if ( hflag || nrec == 0 || nrec >= (int64_t) INT_MAX / 2 ) return OB200_OK;
    OB_CHECK( S->cl_recs.alloc(nrec * (int64_t) sizeof( ClRecord )) );
    OB_CHECK( S->cl_steps.alloc(nsteps * (int64_t) sizeof( ClBlob )) );
    OB_LAUNCH(ctx, cl_records_kernel< true >, bgrid, kBuildThreads, 0, (int32_t) nclusters, S->cl_begin.p, S->cnodes.p, S->ninc_start.p,
              S->ninc.p, S->conn.p, S->coords.p, S->matid.p, S->ncl.p, S->nbase.p, S->nloc.p, S->nblk.p, S->ebidx.p, rcount.p, scount.p,
              roff.p, S->cl_step.p, (ClRecord *) S->cl_recs.p, (ClBlob *) S->cl_steps.p, flags.p,
              S->nodeeq.p, A->rowptr.p, S->blk.p, S->maxblk, (const MatParams *) S->mat.p, S->npar.p);
    OB_CUDA( cudaMemcpyAsync(&hflag, flags.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    if ( hflag ) return OB200_OK;
    S->nclusters = (int32_t) nclusters;
    S->cl_nrec = nrec;
    S->cluster_ok = true;
    return OB200_OK;
}

int cluster_assemble_lspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    const int smem = (int) sizeof( ClShared );
    OB_CUDA( cudaFuncSetAttribute(lspace_cluster_kernel< false >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
    OB_CUDA( cudaFuncSetAttribute(lspace_cluster_kernel< true >, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
    ClView V{ (const ClRecord *) S->cl_recs.p, (const ClBlob *) S->cl_steps.p, S->cl_step.p, S->nclusters };
    ElemSetView v = S->view();
    int grid = ctx->shape.sms * kClCtas;                // persistent: kClCtas CTAs per SM
    if ( grid > S->nclusters ) grid = S->nclusters;
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    // error word: a failure of the previous launch is reported now (and by ob200_context_sync); no host synchronisation here
    if ( !ctx->kerr_dev ) {
        OB_CUDA( cudaMalloc(&ctx->kerr_dev, sizeof( int )) );
        OB_CUDA( cudaMallocHost((void **) &ctx->kerr_host, sizeof( int )) );
        *ctx->kerr_host = 0;
        OB_CUDA( cudaMemsetAsync(ctx->kerr_dev, 0, sizeof( int ), ctx->stream) );
    }
    if ( *ctx->kerr_host ) {
        const int code = *ctx->kerr_host;
        *ctx->kerr_host = 0;
        OB_CUDA( cudaMemsetAsync(ctx->kerr_dev, 0, sizeof( int ), ctx->stream) );
        OB_REQUIRE(false, OB200_ECUDA, "cluster assembly: a wait inside the previous launch timed out (code %d); its matrix values are invalid", code);
    }
    int *errp = ctx->kerr_dev;
    OB_CUDA( cudaMemcpyToSymbolAsync(cl_error_word, &errp, sizeof( errp ), 0, cudaMemcpyHostToDevice, ctx->stream) );
    if ( A->zero_pending ) {
        // every entry of the pattern is written by exactly one lane: the pending zero() is absorbed
        OB_LAUNCH(ctx, lspace_cluster_kernel< false >, grid, kClThreads, smem, v, V, A->val.p);
        A->zero_pending = false;
    } else {
        OB_LAUNCH(ctx, lspace_cluster_kernel< true >, grid, kClThreads, smem, v, V, A->val.p);
    }
    OB_CUDA( cudaMemcpyAsync((void *) ctx->kerr_host, ctx->kerr_dev, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) );
    return OB200_OK;
}

} // namespace ob200
