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
#include "scan.cuh"
#include <limits.h>
#include <string.h>
#include <stdlib.h>

namespace ob200 {

constexpr int kKeWarps = 4;
constexpr int kKeMats = 16;               // materials kept in shared memory
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
lspace_ke_dmma_kernel(ElemSetView S, int nmat, int64_t nelem, const int32_t *__restrict__ vis, double *__restrict__ Ke)
{
    extern __shared__ __align__(16) unsigned char ke_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    KeShared &s = reinterpret_cast< KeShared * >( ke_smem )[wid];
    const int64_t stride = (int64_t) gridDim.x * kKeWarps;
    // the material table in shared memory (the first kKeMats materials; others are read from global memory)
    MatParams *s_mat = reinterpret_cast< MatParams * >( ke_smem + sizeof( KeShared ) * kKeWarps );
    double *s_aux = reinterpret_cast< double * >( s_mat + kKeMats );       // per material: G, K, 1 / (H + 3G), -
    for ( int t = threadIdx.x; t < min(nmat, kKeMats) * (int)( sizeof( MatParams ) / sizeof( double ) ); t += blockDim.x )
        reinterpret_cast< double * >( s_mat )[t] = reinterpret_cast< const double * >( S.mat )[t];
    if ( (int) threadIdx.x < min(nmat, kKeMats) ) {
        const MatParams m = S.mat[threadIdx.x];
        const double G = m.E / ( 2.0 * ( 1.0 + m.nu ) );
        s_aux[4 * threadIdx.x] = G;
        s_aux[4 * threadIdx.x + 1] = m.E / ( 3.0 * ( 1.0 - 2.0 * m.nu ) );
        s_aux[4 * threadIdx.x + 2] = 1.0 / ( m.H + 3.0 * G );
    }
    __syncthreads();
    // vertex coordinates and the material number are requested one element ahead, the connectivity two (conn -> coords is a
    // dependent chain)
    const int64_t e0 = (int64_t) blockIdx.x * kKeWarps + wid;
    double xyz_next = 0.0;
    int node_next = 0, matid_next = 0;
    // where the strip of local node lane >> 2 goes: element-major [nelem][24][24] (vis == nullptr), or the position of the
    // (element, node) incidence in the node -> element lists, so that the strips a node gathers are contiguous
    int vis_next = 0;
    if ( e0 < nelem ) {
        matid_next = S.matid[e0];
        vis_next = vis ? vis[e0 * 8 + ( lane >> 2 )] : (int)( e0 * 8 ) + ( lane >> 2 );
    }
    if ( lane < 24 ) {
        if ( e0 < nelem ) xyz_next = S.coords[(int64_t)( S.conn[e0 * 8 + lane / 3] - 1 ) * 3 + lane % 3];
        if ( e0 + stride < nelem ) node_next = S.conn[( e0 + stride ) * 8 + lane / 3] - 1;
    }
    for ( int64_t e = e0; e < nelem; e += stride ) {
        const int matid = matid_next;
        const MatParams mp = matid < kKeMats ? s_mat[matid] : S.mat[matid];
        const bool mises = mp.type == (double) OB200_MAT_MISES;
        const int gp = lane >> 2, sub = lane & 3;
        // the state of the Gauss point is requested now and used behind the geometry
        MisesTangentIn tin;
        if ( mises && sub < 3 ) mises_tangent_load(mises_ref(S, e * 8 + gp), sub, tin);
        const int strip = vis_next;
        if ( e + stride < nelem ) {
            matid_next = S.matid[e + stride];
            vis_next = vis ? vis[( e + stride ) * 8 + ( lane >> 2 )] : (int)( ( e + stride ) * 8 ) + ( lane >> 2 );
        }
        if ( lane < 24 ) {
            s.xyz[lane] = xyz_next;
            if ( e + stride < nelem ) xyz_next = S.coords[(int64_t) node_next * 3 + lane % 3];
            if ( e + 2 * stride < nelem ) node_next = S.conn[( e + 2 * stride ) * 8 + lane / 3] - 1;
        }
        __syncwarp();
        {   // geometry: lane = 4 gp + sub builds the Jacobian of its Gauss point and the gradients of nodes 2 sub, 2 sub + 1
            // (FEI3dHexaLin::evaldNdx, fei3dhexalin.C:186-204; 2x2x2 rule, gaussintegrationrule.C:190-214)
            // every factor (1 +- xi_gp)/... of dN/dxi takes one of two values at a Gauss point: the twelve products are constants
            // picked by the signs of the Gauss point and of the node (FEI3dHexaLin::evaldNdxi, fei3dhexalin.C:129-166)
            constexpr double kA = 0.577350269189626;
            constexpr double pp = 0.125 * ( 1.0 + kA ) * ( 1.0 + kA ), pm = 0.125 * ( 1.0 + kA ) * ( 1.0 - kA ), mm = 0.125 * ( 1.0 - kA ) * ( 1.0 - kA );
            const bool gu = ( gp & 4 ) != 0, gv = ( gp & 2 ) != 0, gw = ( gp & 1 ) != 0;
            double Pyz[2][2], Pxz[2][2], Pxy[2][2];
#pragma unroll
            for ( int a = 0; a < 2; a++ )
#pragma unroll
                for ( int b = 0; b < 2; b++ ) {
                    const bool ay = ( a == 1 ) == gv, az = ( b == 1 ) == gw, ax = ( a == 1 ) == gu, ay2 = ( b == 1 ) == gv;
                    Pyz[a][b] = ay ? ( az ? pp : pm ) : ( az ? pm : mm );
                    Pxz[a][b] = ax ? ( az ? pp : pm ) : ( az ? pm : mm );
                    Pxy[a][b] = ax ? ( ay2 ? pp : pm ) : ( ay2 ? pm : mm );
                }
            double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
            for ( int kk = 0; kk < 8; kk++ ) {
                const int px = ( kk & 3 ) >= 2, py = ( ( kk & 3 ) == 1 || ( kk & 3 ) == 2 ), pz = kk < 4;
                const double d0 = px ? Pyz[py][pz] : -Pyz[py][pz];
                const double d1 = py ? Pxz[px][pz] : -Pxz[px][pz];
                const double d2 = pz ? Pxy[px][py] : -Pxy[px][py];
                const double x = s.xyz[3 * kk], y = s.xyz[3 * kk + 1], z = s.xyz[3 * kk + 2];
                J[0][0] += x * d0; J[0][1] += x * d1; J[0][2] += x * d2;
                J[1][0] += y * d0; J[1][1] += y * d1; J[1][2] += y * d2;
                J[2][0] += z * d0; J[2][1] += z * d1; J[2][2] += z * d2;
            }
            double Ji[3][3];
            const double det = inv3(J, Ji);               // FEI3dHexaLin::evaldNdx (fei3dhexalin.C:186-204)
#pragma unroll
            for ( int n = 0; n < 2; n++ ) {
                const int k = 2 * sub + n;
                const bool px = ( k & 3 ) >= 2, py = ( ( k & 3 ) == 1 || ( k & 3 ) == 2 ), pz = k < 4;
                const double ayz = py ? ( pz ? Pyz[1][1] : Pyz[1][0] ) : ( pz ? Pyz[0][1] : Pyz[0][0] );
                const double axz = px ? ( pz ? Pxz[1][1] : Pxz[1][0] ) : ( pz ? Pxz[0][1] : Pxz[0][0] );
                const double axy = px ? ( py ? Pxy[1][1] : Pxy[1][0] ) : ( py ? Pxy[0][1] : Pxy[0][0] );
                const double d0 = px ? ayz : -ayz, d1 = py ? axz : -axz, d2 = pz ? axy : -axy;
#pragma unroll
                for ( int j = 0; j < 3; j++ ) s.gT[j][gp * kGgs + k] = d0 * Ji[0][j] + d1 * Ji[1][j] + d2 * Ji[2][j];
            }
            if ( sub == 0 ) s.dV[gp] = fabs(det);          // weights 1 (structural3delement.C:328-338)
            // tangent of the Gauss point: three lanes, two rows each
            if ( sub < 3 ) {
                double G, K, rh;
                if ( matid < kKeMats ) {
                    G = s_aux[4 * matid]; K = s_aux[4 * matid + 1]; rh = s_aux[4 * matid + 2];
                } else {
                    G = mp.E / ( 2.0 * ( 1.0 + mp.nu ) ); K = mp.E / ( 3.0 * ( 1.0 - 2.0 * mp.nu ) ); rh = 1.0 / ( mp.H + 3.0 * G );
                }
                if ( !mises ) tin.kappa = tin.tempKappa = 0.0;          // no plastic increment: the elastic stiffness (isolinearelasticmaterial.C:80-84)
                mises_tangent_rows(mp, G, K, rh, tin, sub, s.D[gp]);
            }
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
        double *o = Ke + (int64_t) strip * 72 + 6 * kk;
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
constexpr int kRowVisits = 8;            // strips of one chunk: requested together, added one after the other
constexpr int kChunkBytes = 72 * kRowVisits;      // column table of one chunk: [r][lane][visit] bytes, r = 0, 1 for 32 lanes, r = 2 for 8

// Per node (built once per bound matrix, strips_bind): d0 = { first incidence, elements around the node, first chunk of its
// column table, row length }, d1 = { start of row 0 / 1 / 2 in val or -1 (prescribed), 0 }.  The column table holds, per
// chunk of eight incidences and per strip entry t = lane + 32 r (strip row i = t / 24, local column t % 24), the column of
// the entry in the node's matrix rows, 0xFF if it has none (prescribed dof of the row or of the column node, no such visit).
struct RowsView {
    int64_t nnode;
    const int4 *rdesc;
    const unsigned char *vtab;
    const double *Ke;               // strips in incidence order: [nvisit][3][24]
};

// the tables: one warp per node; FILL = false counts the chunks of the node only
template< bool FILL >
__global__ void __launch_bounds__(kRowWarps * 32)
rows_tables_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ ninc,
                   const int32_t *__restrict__ nodeeq, const int32_t *__restrict__ rowptr, const unsigned char *__restrict__ ebidx,
                   const unsigned char *__restrict__ nblk, const unsigned short *__restrict__ blk, int maxblk,
                   int32_t *__restrict__ nchunk, const int32_t *__restrict__ tstart, int4 *__restrict__ rdesc, unsigned char *__restrict__ vtab)
{
    __shared__ unsigned short s_col[kRowWarps][kRowCap];          // per column block: first column << 3 | free-dof mask of the column node
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t) gridDim.x * kRowWarps;
    for ( int64_t A = (int64_t) blockIdx.x * kRowWarps + w; A < nnode; A += nwarps ) {
        const int nb = nblk[A];
        const int v0 = ninc_start[A], nv = ninc_start[A + 1] - v0;
        const int nch = nb ? ( nv + kRowVisits - 1 ) / kRowVisits : 0;
        if ( !FILL ) {
            if ( lane == 0 ) nchunk[A] = nch;
            continue;
        }
        int eq[3], rowbase[3];
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            eq[i] = nodeeq[A * 3 + i];
            rowbase[i] = ( nb && eq[i] > 0 ) ? rowptr[eq[i] - 1] : -1;
        }
        int width = 0;
        for ( int b0 = 0; b0 < nb; b0 += 32 ) {
            const int B = b0 + lane;
            const int cm = B < nb ? ( blk[A * maxblk + B] >> 8 ) : 0;
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
        if ( lane == 0 ) {
            rdesc[2 * A] = make_int4(v0, nch ? nv : 0, tstart[A], width);
            rdesc[2 * A + 1] = make_int4(rowbase[0], rowbase[1], rowbase[2], 0);
        }
        __syncwarp();
        for ( int v = 0; v < nch * kRowVisits; v++ ) {
            unsigned char *chunk = vtab + (int64_t)( tstart[A] + v / kRowVisits ) * kChunkBytes;
            uint2 bi = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
            if ( v < nv ) bi = *reinterpret_cast< const uint2 * >( ebidx + (int64_t) ninc[v0 + v] * 8 );
#pragma unroll
            for ( int r = 0; r < 3; r++ ) {
                if ( r == 2 && lane >= 8 ) continue;              // 72 entries
                const int t = lane + 32 * r, i = t / 24, b = ( t % 24 ) / 3, j = t % 3;
                int col = 0xFF;
                const unsigned int word = b < 4 ? bi.x : bi.y;
                const int B = ( word >> ( 8 * ( b & 3 ) ) ) & 0xFF;
                if ( B != 0xFF && ( i == 0 ? eq[0] : i == 1 ? eq[1] : eq[2] ) > 0 ) {
                    const int pk = s_col[w][B], cm = pk & 7;
                    if ( cm & ( 1 << j ) ) col = ( pk >> 3 ) + __popc(cm & ( ( 1 << j ) - 1 ));
                }
                chunk[( r * 32 + lane ) * kRowVisits + ( v % kRowVisits )] = (unsigned char) col;
            }
        }
        __syncwarp();
    }
}

// One warp per node, software-pipelined over the warp's nodes: the descriptor of the next node is requested at the top, its
// element list and column table behind the strip requests of the current node -- a node costs one memory round trip.
template< bool ACCUM >
__global__ void __launch_bounds__(kRowWarps * 32, 3)
lspace_rows_kernel(const __grid_constant__ RowsView V, double *__restrict__ val)
{
    __shared__ double s_acc[kRowWarps][3][kRowCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t) gridDim.x * kRowWarps;
    const int ti0 = lane / 24, ti1 = ( lane + 32 ) / 24;          // strip row of the lane's entries t = lane, lane + 32 (lane + 64: row 2)
    struct Desc {
        int4 d0, d1;
        uint2 t0, t1, t2;      // the lane's columns in the first chunk, one byte per incidence
    };
    auto level0 = [&](int64_t A, Desc &d) {
        d.d0 = V.rdesc[2 * A];
        d.d1 = V.rdesc[2 * A + 1];
    };
    auto load_tab = [&](int chunk, uint2 &t0, uint2 &t1, uint2 &t2) {
        const uint2 *tab = reinterpret_cast< const uint2 * >( V.vtab + (int64_t) chunk * kChunkBytes );
        t0 = tab[lane];
        t1 = tab[32 + lane];
        t2 = lane < 8 ? tab[64 + lane] : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
    };
    auto level1 = [&](Desc &d) {
        if ( d.d0.y > 0 ) load_tab(d.d0.z, d.t0, d.t1, d.t2);
    };
    int64_t A = (int64_t) blockIdx.x * kRowWarps + w;
    if ( A >= V.nnode ) return;
    Desc cur, nxt;
    level0(A, cur);
    level1(cur);
    nxt = cur;
    for ( ; A < V.nnode; A += nwarps ) {
        const bool have_next = A + nwarps < V.nnode;
        if ( have_next ) level0(A + nwarps, nxt);
        const int v0 = cur.d0.x, nv = cur.d0.y, width = nv > 0 ? cur.d0.w : 0;
        for ( int c = lane; c < width; c += 32 ) s_acc[w][0][c] = s_acc[w][1][c] = s_acc[w][2][c] = 0.0;
        __syncwarp();
        uint2 t0 = cur.t0, t1 = cur.t1, t2 = cur.t2;
        for ( int vb = 0; vb < nv; vb += kRowVisits ) {
            if ( vb > 0 ) load_tab(cur.d0.z + vb / kRowVisits, t0, t1, t2);
            // request the strips of half a chunk, then add them in ascending element number
#pragma unroll 1
            for ( int half = 0; half < 2; half++ ) {
                if ( vb + 4 * half >= nv ) break;
                const unsigned int w0 = half ? t0.y : t0.x, w1 = half ? t1.y : t1.x, w2 = half ? t2.y : t2.x;
                double kv[4][3];
#pragma unroll
                for ( int v = 0; v < 4; v++ ) {
                    const double *strip = V.Ke + (int64_t)( v0 + vb + 4 * half + v ) * 72 + lane;
                    if ( ( ( w0 >> ( 8 * v ) ) & 0xFFu ) != 0xFFu ) kv[v][0] = strip[0];
                    if ( ( ( w1 >> ( 8 * v ) ) & 0xFFu ) != 0xFFu ) kv[v][1] = strip[32];
                    if ( ( ( w2 >> ( 8 * v ) ) & 0xFFu ) != 0xFFu ) kv[v][2] = strip[64];
                }
                if ( vb == 0 && half == 0 && have_next ) level1(nxt);              // behind the first strip requests of this node
#pragma unroll
                for ( int v = 0; v < 4; v++ ) {
                    const unsigned int c0 = ( w0 >> ( 8 * v ) ) & 0xFFu, c1 = ( w1 >> ( 8 * v ) ) & 0xFFu, c2 = ( w2 >> ( 8 * v ) ) & 0xFFu;
                    if ( c0 != 0xFFu ) s_acc[w][ti0][c0] += kv[v][0];
                    if ( c1 != 0xFFu ) s_acc[w][ti1][c1] += kv[v][1];
                    if ( c2 != 0xFFu ) s_acc[w][2][c2] += kv[v][2];
                    __syncwarp();
                }
            }
        }
        if ( nv == 0 && have_next ) level1(nxt);
        if ( width > 0 ) {
            const int rb[3] = { cur.d1.x, cur.d1.y, cur.d1.z };
#pragma unroll
            for ( int i = 0; i < 3; i++ ) {
                if ( rb[i] < 0 ) continue;
                double *dst = val + rb[i];
                for ( int c = lane; c < width; c += 32 ) dst[c] = ACCUM ? dst[c] + s_acc[w][i][c] : s_acc[w][i][c];
            }
        }
        __syncwarp();
        cur = nxt;
    }
}

// ---- host side -----------------------------------------------------------------------------------

// LSpace stiffness matrices of the whole set into Ke (device pointer): [nelem][24][24], or with vis the 3 x 24 strips in
// incidence order
int strips_element_matrices(ob200_elemset *S, double *Ke, const int32_t *vis)
{
    ob200_context *ctx = S->ctx;
    static unsigned long long attr_set = 0;
    const int smem = (int) sizeof( KeShared ) * kKeWarps + kKeMats * ( (int) sizeof( MatParams ) + 4 * (int) sizeof( double ) );
    if ( attr_needed(attr_set, ctx->device) ) {
        OB_CUDA( cudaFuncSetAttribute(lspace_ke_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) );
    }
    if ( S->nelem == 0 ) return OB200_OK;
    const int per_sm = 3;                                    // 63.7 KB of shared memory per CTA
    int grid = ctx->shape.sms * per_sm;
    const int64_t need = ( S->nelem + kKeWarps - 1 ) / kKeWarps;
    if ( grid > need ) grid = (int) need;
    OB_LAUNCH(ctx, lspace_ke_dmma_kernel, grid, kKeWarps * 32, smem, S->view(), (int) S->nmat, S->nelem, vis, Ke);
    return OB200_OK;
}

// the per-node descriptors and column tables of lspace_rows_kernel (once per bound matrix; part of the cached schedule)
int strips_bind(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    if ( !S->strips_ok ) return OB200_OK;
    const int grid = ctx->shape.grid(S->nnode * 32, kRowWarps * 32, 8);
    DevBuf< int32_t > nchunk;
    DevBuf< int64_t > s64;
    OB_CHECK( nchunk.alloc(S->nnode + 1) );
    OB_CHECK( s64.alloc(S->nnode + 1) );
    OB_CHECK( S->row_tstart.alloc(S->nnode + 1) );
    OB_CUDA( cudaMemsetAsync(nchunk.p, 0, sizeof( int32_t ) * ( S->nnode + 1 ), ctx->stream) );
    OB_LAUNCH(ctx, rows_tables_kernel< false >, grid, kRowWarps * 32, 0, S->nnode, S->ninc_start.p, S->ninc.p, S->nodeeq.p, A->rowptr.p,
              S->ebidx.p, S->nblk.p, S->blk.p, S->maxblk, nchunk.p, (const int32_t *) nullptr, (int4 *) nullptr, (unsigned char *) nullptr);
    int64_t total = 0;
    OB_CHECK( exclusive_scan(ctx, nchunk.p, s64.p, S->nnode + 1, &total) );
    OB_REQUIRE(total < (int64_t) INT_MAX, OB200_ECAPACITY, "strip assembly: %lld chunks exceed the 32-bit chunk index", (long long) total);
    OB_CHECK( narrow_i64_to_i32(ctx, s64.p, S->row_tstart.p, S->nnode + 1) );
    OB_CHECK( S->row_desc.alloc(S->nnode * 2 * 4) );
    OB_CHECK( S->row_vtab.alloc(( total > 0 ? total : 1 ) * kChunkBytes) );
    OB_CUDA( cudaMemsetAsync(S->row_vtab.p, 0xFF, (size_t)( total > 0 ? total : 1 ) * kChunkBytes, ctx->stream) );
    OB_LAUNCH(ctx, rows_tables_kernel< true >, grid, kRowWarps * 32, 0, S->nnode, S->ninc_start.p, S->ninc.p, S->nodeeq.p, A->rowptr.p,
              S->ebidx.p, S->nblk.p, S->blk.p, S->maxblk, nchunk.p, S->row_tstart.p, reinterpret_cast< int4 * >( S->row_desc.p ), S->row_vtab.p);
    return OB200_OK;
}

int strips_assemble_lspace(ob200_elemset *S, ob200_csr *A)
{
    ob200_context *ctx = S->ctx;
    if ( !S->kebuf.p ) OB_CHECK( S->kebuf.alloc(S->nelem * 576) );
    OB_CHECK( strips_element_matrices(S, S->kebuf.p, S->row_vis.p) );
    RowsView V{ S->nnode, reinterpret_cast< const int4 * >( S->row_desc.p ), S->row_vtab.p, S->kebuf.p };
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
