// Owner-computes assembly of the LSpace tangent with a general material stiffness ("element strips").
//
// Replaces EngngModel::assemble (src/core/engngm.C:889-929) + CompCol::assemble (src/core/compcol.C:263-299) for an LSpace
// element set that contains a MisesMat (per-Gauss-point, possibly unsymmetric algorithmic tangent, misesmat.C:493-545)
// without atomics and without the element -> slot map of the generic path.  Two kernels:
//
//   1. lspace_ke_dmma_kernel -- StructuralElement::computeStiffnessMatrix (structuralelement.C:575-643), one warp per
//      element, on the FP64 tensor path.  With the B matrix of Structural3DElement::computeBmatrixAt
//      (structural3delement.C:63-86), eps_s = sum_{i,p} [v(i,p) = s] dN/dx_p u_i (v = Voigt index of the pair), so
//          K_ab[i][j] = sum_gp sum_p  g_a[p] * Q[gp][v(i,p)][j][b],      Q[gp][s][j][b] = dV (D_gp B_b)[s][j],
//      i.e. for every component pair (i,j) an 8x8 (a,b) product with inner dimension (gp, p) = 24: one m8n8k4 accumulator
//      per pair, 6 k-steps, 54 DMMAs per element; the A fragment (the gradients) is shared by the nine pairs.
//      Q (6 x 8 x 24 doubles) is formed once per element by the lanes (two (gp, b) pairs each, 54 FMAs per pair) and
//      staged in shared memory with strides that make the fragment loads conflict-free.  Ke goes to HBM row-major
//      ([nelem][24][24]; each lane stores 48 contiguous bytes per row, a quarter-warp a whole 192-byte row).
//   2. lspace_rows_kernel -- CompCol::assemble as a gather: one warp per node sums, in ascending element number, the
//      3 x 24 strips of the element matrices around the node (rows of local node a) into the node's (up to) three matrix
//      rows in shared memory and writes each row once -- bit-reproducible run to run.  The column of every strip entry
//      comes from the node-block schedule that gather_bind builds (block index per (element, a, b), free-dof masks).
//
// Per element: 4608 B of Ke written and read once + its share of the matrix rows.
#include "element_device.cuh"
#include "elemset.h"
#include <string.h>
#include <stdlib.h>

namespace ob200 {

constexpr int kKeWarps = 4;
constexpr int kQgs = 28;                 // doubles between Gauss points of Q (= 12 mod 16: the four Gauss points of a k-step fall in distinct bank groups)
constexpr int kQss = 8 * kQgs;           // doubles between strain rows of Q
constexpr int kGgs = 12;                 // doubles between Gauss points of the transposed gradients
constexpr int kDs = 37;                  // doubles between the tangents of the Gauss points

struct KeShared {
    double xyz[24];                      // vertex coordinates
    double gT[3][8 * kGgs];              // gT[p][gp][a] = dN_a/dx_p
    double dV[8];
    double D[8][kDs];                    // tangent of every Gauss point
    double Q[6 * kQss];                  // Q[s][gp][j * 8 + b]
};

__device__ __forceinline__ void ke_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile( "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"( c0 ), "+d"( c1 ) : "d"( a ), "d"( b ) );
}

// Voigt index of the (displacement component, gradient component) pair: 11 22 33 23 13 12
__device__ __forceinline__ constexpr int voigt(int i, int p)
{
    return i == p ? i : 6 - i - p;
}

__global__ void __launch_bounds__(kKeWarps * 32)
lspace_ke_dmma_kernel(ElemSetView S, int64_t nelem, double *__restrict__ Ke)
{
    extern __shared__ __align__(16) unsigned char ke_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    KeShared &s = reinterpret_cast< KeShared * >( ke_smem )[wid];
    const int64_t stride = (int64_t) gridDim.x * kKeWarps;
    for ( int64_t e = (int64_t) blockIdx.x * kKeWarps + wid; e < nelem; e += stride ) {
        if ( lane < 24 ) {
            const int node = S.conn[e * 8 + lane / 3] - 1;
            s.xyz[lane] = S.coords[(int64_t) node * 3 + lane % 3];
        }
        const MatParams mp = S.mat[S.matid[e]];
        __syncwarp();
        {   // geometry: lane = 4 gp + sub builds the Jacobian of its Gauss point and the gradients of nodes 2 sub, 2 sub + 1
            // (FEI3dHexaLin::evaldNdx, fei3dhexalin.C:186-204; 2x2x2 rule, gaussintegrationrule.C:190-214)
            const int gp = lane >> 2, sub = lane & 3;
            double u, v, w, Ji[3][3];
            hexa_gp(gp, u, v, w);
            const double det = hexa_jacobian(s.xyz, u, v, w, Ji);
#pragma unroll
            for ( int n = 0; n < 2; n++ ) {
                const int k = 2 * sub + n;
                double dN[3];
                hexa_dNdxi(k, u, v, w, dN);
#pragma unroll
                for ( int j = 0; j < 3; j++ ) s.gT[j][gp * kGgs + k] = dN[0] * Ji[0][j] + dN[1] * Ji[1][j] + dN[2] * Ji[2][j];
            }
            if ( sub == 0 ) s.dV[gp] = fabs(det);          // weights 1 (structural3delement.C:328-338)
        }
        if ( lane < 8 ) {
            if ( mp.type == (double) OB200_MAT_MISES ) mises_tangent(mp, &S.state[e * 8 + lane], s.D[lane]);
            else isole_D(mp.E, mp.nu, s.D[lane]);
        }
        __syncwarp();
        {   // Q[s][gp][j][b] = dV (D B_b)[s][j]: lane = (gp, node pair)
            const int gp = lane >> 2, b0 = 2 * ( lane & 3 );
            const double dV = s.dV[gp];
            double gb[2][3];
#pragma unroll
            for ( int h = 0; h < 2; h++ )
#pragma unroll
                for ( int q = 0; q < 3; q++ ) gb[h][q] = dV * s.gT[q][gp * kGgs + b0 + h];
            const double *D = s.D[gp];
#pragma unroll
            for ( int sr = 0; sr < 6; sr++ ) {
#pragma unroll
                for ( int j = 0; j < 3; j++ ) {
                    const double d0 = D[6 * sr + voigt(j, 0)], d1 = D[6 * sr + voigt(j, 1)], d2 = D[6 * sr + voigt(j, 2)];
                    const double t0 = d0 * gb[0][0] + d1 * gb[0][1] + d2 * gb[0][2];
                    const double t1 = d0 * gb[1][0] + d1 * gb[1][1] + d2 * gb[1][2];
                    *reinterpret_cast< double2 * >( &s.Q[sr * kQss + gp * kQgs + j * 8 + b0] ) = make_double2(t0, t1);
                }
            }
        }
        __syncwarp();
        // the nine 8x8 products: rows a = lane >> 2 (A) / columns b = lane >> 2 (B), inner index (p, gp) in steps of four Gauss points
        double acc[9][2];
#pragma unroll
        for ( int t = 0; t < 9; t++ ) acc[t][0] = acc[t][1] = 0.0;
        const int kk = lane & 3, ab = lane >> 2;
#pragma unroll
        for ( int p = 0; p < 3; p++ )
#pragma unroll
            for ( int h = 0; h < 2; h++ ) {
                const int gp = 4 * h + kk;
                const double A = s.gT[p][gp * kGgs + ab];
#pragma unroll
                for ( int i = 0; i < 3; i++ ) {
                    const double *q = &s.Q[voigt(i, p) * kQss + gp * kQgs + ab];
#pragma unroll
                    for ( int j = 0; j < 3; j++ ) ke_dmma(acc[3 * i + j][0], acc[3 * i + j][1], A, q[j * 8]);
                }
            }
        // the lane holds the 3x3 blocks (a, 2 kk) and (a, 2 kk + 1): six consecutive entries of each of the rows 3a .. 3a+2
        double *o = Ke + e * 576 + ( 3 * ab ) * 24 + 6 * kk;
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            double2 *o2 = reinterpret_cast< double2 * >( o + i * 24 );
            o2[0] = make_double2(acc[3 * i][0], acc[3 * i + 1][0]);
            o2[1] = make_double2(acc[3 * i + 2][0], acc[3 * i][1]);
            o2[2] = make_double2(acc[3 * i + 1][1], acc[3 * i + 2][1]);
        }
        __syncwarp();
    }
}

// ---- rows from strips ------------------------------------------------------------------------------------

constexpr int kRowWarps = 8;
constexpr int kRowCap = 128;             // longest row (kMaxRowLen of assemble_gather.cu)
constexpr int kRowVisits = 8;            // strips requested at a time

struct RowsView {
    int64_t nnode;
    const int32_t *ninc_start, *ninc, *nodeeq, *rowptr;
    const unsigned char *ebidx, *nblk;
    const unsigned short *blk;
    int maxblk;
    const double *Ke;
};

template< bool ACCUM >
__global__ void __launch_bounds__(kRowWarps * 32)
lspace_rows_kernel(RowsView V, double *__restrict__ val)
{
    __shared__ double s_acc[kRowWarps][3][kRowCap];
    __shared__ unsigned short s_col[kRowWarps][kRowCap];          // per column block: first column << 3 | free-dof mask of the column node
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t) gridDim.x * kRowWarps;
    // position of this lane's three strip entries: t = lane + 32 r -> row i = t / 24, local column c = t % 24 (node b = c / 3, component j = c % 3)
    int ti[3], tb[3], tj[3];
#pragma unroll
    for ( int r = 0; r < 3; r++ ) {
        const int t = lane + 32 * r;
        ti[r] = t / 24;
        tb[r] = ( t % 24 ) / 3;
        tj[r] = t % 3;                    // 24 is a multiple of 3
    }
    for ( int64_t A = (int64_t) blockIdx.x * kRowWarps + w; A < V.nnode; A += nwarps ) {
        const int nb = V.nblk[A];
        if ( nb == 0 ) continue;                                  // no free equation at this node
        int eq[3], rowbase[3];
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            eq[i] = V.nodeeq[A * 3 + i];
            rowbase[i] = eq[i] > 0 ? V.rowptr[eq[i] - 1] : 0;
        }
        bool rfree[3];                                            // is the row of this lane's r-th strip entry a free equation
#pragma unroll
        for ( int r = 0; r < 3; r++ ) rfree[r] = ( ti[r] == 0 ? eq[0] : ti[r] == 1 ? eq[1] : eq[2] ) > 0;
        // first column of every block
        int width = 0;
        for ( int b0 = 0; b0 < nb; b0 += 32 ) {
            const int B = b0 + lane;
            const int cm = B < nb ? ( V.blk[A * V.maxblk + B] >> 8 ) : 0;
            const int wdt = __popc(cm);
            int incl = wdt;
#pragma unroll
            for ( int o = 1; o < 32; o <<= 1 ) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if ( lane >= o ) incl += t;
            }
            if ( B < nb ) s_col[w][B] = (unsigned short)( ( ( width + incl - wdt ) << 3 ) | cm );
            width += __shfl_sync(0xffffffffu, incl, 31);
        }
        for ( int c = lane; c < width; c += 32 ) s_acc[w][0][c] = s_acc[w][1][c] = s_acc[w][2][c] = 0.0;
        __syncwarp();
        const int v0 = V.ninc_start[A], nv = V.ninc_start[A + 1] - v0;
        for ( int vb = 0; vb < nv; vb += kRowVisits ) {
            // request the strips of up to kRowVisits elements, then add them in ascending element number
            double kv[kRowVisits][3];
            int cl[kRowVisits][3];
#pragma unroll
            for ( int v = 0; v < kRowVisits; v++ ) {
                if ( vb + v >= nv ) {
#pragma unroll
                    for ( int r = 0; r < 3; r++ ) cl[v][r] = -1;
                    continue;
                }
                const int ea = V.ninc[v0 + vb + v];
                const uint2 bi = *reinterpret_cast< const uint2 * >( V.ebidx + (int64_t) ea * 8 );
                const double *strip = V.Ke + (int64_t)( ea >> 3 ) * 576 + ( 3 * ( ea & 7 ) ) * 24;
#pragma unroll
                for ( int r = 0; r < 3; r++ ) {
                    cl[v][r] = -1;
                    if ( r == 2 && lane >= 8 ) continue;           // 72 entries
                    const unsigned int word = tb[r] < 4 ? bi.x : bi.y;
                    const int B = ( word >> ( 8 * ( tb[r] & 3 ) ) ) & 0xFF;
                    if ( B == 0xFF || !rfree[r] ) continue;
                    const int pk = s_col[w][B], cm = pk & 7;
                    if ( !( cm & ( 1 << tj[r] ) ) ) continue;
                    cl[v][r] = ( pk >> 3 ) + __popc(cm & ( ( 1 << tj[r] ) - 1 ));
                    kv[v][r] = strip[lane + 32 * r];
                }
            }
#pragma unroll
            for ( int v = 0; v < kRowVisits; v++ ) {
#pragma unroll
                for ( int r = 0; r < 3; r++ )
                    if ( cl[v][r] >= 0 ) s_acc[w][ti[r]][cl[v][r]] += kv[v][r];
                __syncwarp();
            }
        }
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            if ( eq[i] <= 0 ) continue;
            double *dst = val + rowbase[i];
            for ( int c = lane; c < width; c += 32 ) dst[c] = ACCUM ? dst[c] + s_acc[w][i][c] : s_acc[w][i][c];
        }
        __syncwarp();
    }
}

// ---- host side -----------------------------------------------------------------------------------

// LSpace stiffness matrices of the whole set into Ke [nelem][24][24] (device pointer)
int strips_element_matrices(ob200_elemset *S, double *Ke)
{
    ob200_context *ctx = S->ctx;
    static bool attr_set = false;
    const int smem = (int) sizeof( KeShared ) * kKeWarps;
    if ( !attr_set ) {
        OB_CUDA( cudaFuncSetAttribute(lspace_ke_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
        attr_set = true;
    }
    if ( S->nelem == 0 ) return OB200_OK;
    const int per_sm = 3;                                    // 62.7 KB of shared memory per CTA
    int grid = ctx->shape.sms * per_sm;
    const int64_t need = ( S->nelem + kKeWarps - 1 ) / kKeWarps;
    if ( grid > need ) grid = (int) need;
    OB_LAUNCH(ctx, lspace_ke_dmma_kernel, grid, kKeWarps * 32, smem, S->view(), S->nelem, Ke);
    return OB200_OK;
}

int strips_assemble_lspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    if ( !S->kebuf.p ) OB_CHECK( S->kebuf.alloc(S->nelem * 576) );
    OB_CHECK( strips_element_matrices(S, S->kebuf.p) );
    RowsView V{ S->nnode, S->ninc_start.p, S->ninc.p, S->nodeeq.p, A->rowptr.p, S->ebidx.p, S->nblk.p, S->blk.p, S->maxblk, S->kebuf.p };
    const int grid = ctx->shape.grid(S->nnode * 32, kRowWarps * 32, 8);
    if ( A->zero_pending && !S->covers_all ) OB_CHECK( ob200_csr_materialize(A) );
    if ( A->zero_pending ) {
        // every entry of the pattern is written by exactly one warp: the pending zero() is absorbed
        OB_LAUNCH(ctx, lspace_rows_kernel< false >, grid, kRowWarps * 32, 0, V, A->val.p);
        A->zero_pending = false;
    } else {
        OB_LAUNCH(ctx, lspace_rows_kernel< true >, grid, kRowWarps * 32, 0, V, A->val.p);
    }
    return OB200_OK;
}

} // namespace ob200
