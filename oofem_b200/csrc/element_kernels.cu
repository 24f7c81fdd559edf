// Batched element evaluation: LSpace (warp per element) and LTRSpace (thread per element).
//
// Replaces the per-element loop of EngngModel::assemble / assembleVector
// (src/core/engngm.C:889-929) over StructuralElement::computeStiffnessMatrix and
// giveInternalForcesVector (src/sm/Elements/structuralelement.C:575-643, 724-802).
#include "element_device.cuh"
#include "element_forces.cuh"
#include "elemset.h"
#include <stdlib.h>
#include <string.h>
#include <utility>

namespace ob200 {

constexpr int kHexWarps = 8;        // warps (= elements in flight) per CTA for the LSpace kernels

// output modes of the stiffness kernels
enum { OUT_KE = 0, OUT_CSR = 1 };

struct HexShared {
    double xyz[24];          // vertex coordinates
    double g[8][8][3];       // dN/dx per Gauss point and node
    double dV[8];            // |det J| * weight
    double D[8][36];         // per-GP tangent (MisesMat only)
    double ue[24];           // element displacement vector (KDU / internal forces)
    double sig[8][6];        // stresses (internal forces)
};

// Geometry phase shared by all LSpace kernels: lane = 4*gp + sub; every lane builds the Jacobian
// of its Gauss point and the gradients of nodes 2*sub, 2*sub+1.
__device__ __forceinline__ void hex_geometry(HexShared &s, int lane)
{
    int gp = lane >> 2, sub = lane & 3;
    double u, v, w, Ji[3][3];
    hexa_gp(gp, u, v, w);
    double det = hexa_jacobian(s.xyz, u, v, w, Ji);
#pragma unroll
    for ( int n = 0; n < 2; n++ ) {
        int k = 2 * sub + n;
        double dN[3];
        hexa_dNdxi(k, u, v, w, dN);
#pragma unroll
        for ( int j = 0; j < 3; j++ ) s.g[gp][k][j] = dN[0] * Ji[0][j] + dN[1] * Ji[1][j] + dN[2] * Ji[2][j];
    }
    if ( sub == 0 ) s.dV[gp] = fabs(det);      // weight 1*1*1 (structural3delement.C:328-338)
}

template< int MODE >
__global__ void __launch_bounds__(kHexWarps * 32)
lspace_stiffness_kernel(ElemSetView S, double *__restrict__ out, const int32_t *__restrict__ slot, int64_t e_begin, int64_t e_end)
{
    __shared__ HexShared sh[kHexWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    HexShared &s = sh[wid];
    const int64_t stride = (int64_t) gridDim.x * kHexWarps;
    for ( int64_t e = e_begin + (int64_t) blockIdx.x * kHexWarps + wid; e < e_end; e += stride ) {
        // stage connectivity -> coordinates through shared memory
        if ( lane < 24 ) {
            int node = S.conn[e * 8 + lane / 3] - 1;
            s.xyz[lane] = S.coords[(int64_t) node * 3 + lane % 3];
        }
        __syncwarp();
        hex_geometry(s, lane);
        const MatParams mp = S.mat[S.matid[e]];
        const bool mises = ( mp.type == (double) OB200_MAT_MISES );
        if ( mises && lane < 8 ) mises_tangent(mp, mises_ref(S, e * 8 + lane), s.D[lane]);
        __syncwarp();
        double lam, mu;
        isole_lame(mp.E, mp.nu, lam, mu);
#pragma unroll 1
        for ( int t = lane; t < 64; t += 32 ) {
            const int a = t >> 3, b = t & 7;
            double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
            if ( !mises ) {
#pragma unroll
                for ( int gp = 0; gp < 8; gp++ ) {
                    double w = s.dV[gp];
                    block_iso(acc, s.g[gp][a], s.g[gp][b], w * lam, w * mu);
                }
            } else {
#pragma unroll 2
                for ( int gp = 0; gp < 8; gp++ ) block_general(acc, s.g[gp][a], s.g[gp][b], s.D[gp], s.dV[gp]);
            }
            if ( MODE == OUT_KE ) {
                double *o = out + e * 576 + ( 3 * a ) * 24 + 3 * b;
#pragma unroll
                for ( int i = 0; i < 3; i++ )
#pragma unroll
                    for ( int j = 0; j < 3; j++ ) o[i * 24 + j] = acc[3 * i + j];
            } else {
                const int32_t *sl = slot + e * 576 + ( 3 * a ) * 24 + 3 * b;
#pragma unroll
                for ( int i = 0; i < 3; i++ )
#pragma unroll
                    for ( int j = 0; j < 3; j++ ) {
                        int32_t p = sl[i * 24 + j];
                        if ( p >= 0 ) atomicAdd(out + p, acc[3 * i + j]);
                    }
            }
        }
        __syncwarp();
    }
}

// ---- LTRSpace: one thread per element (1 Gauss point, constant gradients) ---------------

__device__ __forceinline__ double tet_setup(const ElemSetView &S, int64_t e, double g[4][3], int node[4])
{
    double c[12];
#pragma unroll
    for ( int a = 0; a < 4; a++ ) {
        node[a] = S.conn[e * 4 + a] - 1;
#pragma unroll
        for ( int j = 0; j < 3; j++ ) c[3 * a + j] = S.coords[(int64_t) node[a] * 3 + j];
    }
    double det = tet_dNdx(c, g);
    return fabs(det) * ( 1.0 / 6.0 );     // gaussintegrationrule.C:500-507 weight 1/6
}

// FEI3dTetLin::evaldNdx raises OOFEM_ERROR("negative volume") for detJ <= 0 (fei3dtetlin.C:145-147): checked for the whole
// set when it is created; bad[0] = number of such elements, bad[1] = the first one (smallest number)
__global__ void tet_volume_check_kernel(ElemSetView S, int64_t nelem, int *__restrict__ bad)
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
        if ( !( tet_dNdx(c, g) > 0.0 ) ) {
            atomicAdd(bad, 1);
            atomicMin(bad + 1, (int) min(e, (int64_t) 0x7FFFFFFF));
        }
    }
}

template< int MODE >
__global__ void __launch_bounds__(128)
ltrspace_stiffness_kernel(ElemSetView S, double *__restrict__ out, const int32_t *__restrict__ slot, int64_t e_begin, int64_t e_end)
{
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t e = e_begin + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < e_end; e += stride ) {
        double g[4][3];
        int node[4];
        double dV = tet_setup(S, e, g, node);
        const MatParams mp = S.mat[S.matid[e]];
        const bool mises = ( mp.type == (double) OB200_MAT_MISES );
        double D[36];
        double lam, mu;
        isole_lame(mp.E, mp.nu, lam, mu);
        if ( mises ) mises_tangent(mp, mises_ref(S, e), D);
#pragma unroll 1
        for ( int a = 0; a < 4; a++ ) {
#pragma unroll 1
            for ( int b = 0; b < 4; b++ ) {
                double acc[9] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
                if ( mises ) block_general(acc, g[a], g[b], D, dV);
                else block_iso(acc, g[a], g[b], dV * lam, dV * mu);
                if ( MODE == OUT_KE ) {
                    double *o = out + e * 144 + ( 3 * a ) * 12 + 3 * b;
#pragma unroll
                    for ( int i = 0; i < 3; i++ )
#pragma unroll
                        for ( int j = 0; j < 3; j++ ) o[i * 12 + j] = acc[3 * i + j];
                } else {
                    const int32_t *sl = slot + e * 144 + ( 3 * a ) * 12 + 3 * b;
#pragma unroll
                    for ( int i = 0; i < 3; i++ )
#pragma unroll
                        for ( int j = 0; j < 3; j++ ) {
                            int32_t p = sl[i * 12 + j];
                            if ( p >= 0 ) atomicAdd(out + p, acc[3 * i + j]);
                        }
                }
            }
        }
    }
}

template< int MODE >       // FORCE_INTERNAL / FORCE_TANGENT_DU (element_forces.cuh)
__global__ void __launch_bounds__(128)
ltrspace_forces_kernel(ElemSetView S, const double *__restrict__ u, double *__restrict__ fe,
                                double *__restrict__ fglob, double *__restrict__ gp_strain,
                                double *__restrict__ gp_stress, double *__restrict__ ebe_norm2,
                                double *__restrict__ fvis, const int32_t *__restrict__ vis,
                                int64_t e_begin, int64_t e_end)
{
    double ebe[3] = { 0.0, 0.0, 0.0 };
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t e = e_begin + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < e_end; e += stride ) {
        double g[4][3];
        int node[4];
        double dV = tet_setup(S, e, g, node);
        const MatParams mp = S.mat[S.matid[e]];
        double eps[6] = { 0, 0, 0, 0, 0, 0 }, sig[6];
#pragma unroll
        for ( int a = 0; a < 4; a++ ) {
            double ua[3] = { u[(int64_t) node[a] * 3], u[(int64_t) node[a] * 3 + 1], u[(int64_t) node[a] * 3 + 2] };
            strain_add(eps, g[a], ua);
        }
        point_stress< MODE >(mp, S, e, eps, sig);
        if ( gp_strain )
#pragma unroll
            for ( int i = 0; i < 6; i++ ) gp_strain[e * 6 + i] = eps[i];
        if ( gp_stress )
#pragma unroll
            for ( int i = 0; i < 6; i++ ) gp_stress[e * 6 + i] = sig[i];
#pragma unroll
        for ( int a = 0; a < 4; a++ ) {
            double fk[3];
            force_node(fk, g[a], sig, dV);
#pragma unroll
            for ( int i = 0; i < 3; i++ ) {
                ebe[i] += fk[i] * fk[i];
                if ( fe ) fe[e * 12 + 3 * a + i] = fk[i];
                if ( fvis ) fvis[(int64_t) vis[e * 8 + a] * 3 + i] = fk[i];
                if ( fglob ) {
                    int32_t r = S.loc[e * 12 + 3 * a + i];
                    if ( r > 0 && ( MODE == FORCE_INTERNAL || fk[i] != 0.0 ) ) atomicAdd(fglob + r - 1, fk[i]);
                }
            }
        }
    }
    if ( ebe_norm2 ) {
#pragma unroll
        for ( int c = 0; c < 3; c++ ) {
            double v = ebe[c];
#pragma unroll
            for ( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ( ( threadIdx.x & 31 ) == 0 && v != 0.0 ) atomicAdd(ebe_norm2 + c, v);
        }
    }
}

// Owner-computes vector assembly (EngngModel::assembleVectorFromElements, engngm.C:1351-, without atomics): one thread per
// node adds the nodal forces of the elements around it in ascending element number -- fvis holds them in incidence order, so a
// node reads 24 nv contiguous bytes -- into the node's free equations, and the squares of all of them per dof id
// (the element-by-element norms of engngm.C:1108-1133).  The block sums are added in block order by norm_finish_kernel: the
// whole result is bit-reproducible run to run.
constexpr int kForceGatherThreads = 256;
__global__ void __launch_bounds__(kForceGatherThreads)
node_force_gather_kernel(int64_t nnode, const int32_t *__restrict__ ninc_start, const int32_t *__restrict__ nodeeq,
                         const double *__restrict__ fvis, double *__restrict__ f, double *__restrict__ partial)
{
    __shared__ double s_red[kForceGatherThreads / 32][3];
    const int64_t A = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    double sq[3] = { 0.0, 0.0, 0.0 };
    if ( A < nnode ) {
        const int v0 = ninc_start[A], v1 = ninc_start[A + 1];
        double acc[3] = { 0.0, 0.0, 0.0 };
        for ( int v = v0; v < v1; v++ ) {
#pragma unroll
            for ( int i = 0; i < 3; i++ ) {
                const double x = fvis[(int64_t) v * 3 + i];
                acc[i] += x;
                sq[i] += x * x;
            }
        }
#pragma unroll
        for ( int i = 0; i < 3; i++ ) {
            const int eq = nodeeq[A * 3 + i];
            if ( eq > 0 && v1 > v0 ) f[eq - 1] += acc[i];
        }
    }
    if ( !partial ) return;
#pragma unroll
    for ( int i = 0; i < 3; i++ ) {
#pragma unroll
        for ( int o = 16; o > 0; o >>= 1 ) sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
        if ( ( threadIdx.x & 31 ) == 0 ) s_red[threadIdx.x >> 5][i] = sq[i];
    }
    __syncthreads();
    if ( threadIdx.x < 3 ) {
        double t = 0.0;
        for ( int w = 0; w < kForceGatherThreads / 32; w++ ) t += s_red[w][threadIdx.x];
        partial[(int64_t) blockIdx.x * 3 + threadIdx.x] = t;
    }
}

// out[c] = sum over blocks of partial[block][c], one block, fixed order
__global__ void __launch_bounds__(256) norm_finish_kernel(const double *__restrict__ partial, int64_t nblocks, double *__restrict__ out)
{
    __shared__ double s_red[256][3];
    double t[3] = { 0.0, 0.0, 0.0 };
    for ( int64_t b = threadIdx.x; b < nblocks; b += 256 )
#pragma unroll
        for ( int i = 0; i < 3; i++ ) t[i] += partial[b * 3 + i];
#pragma unroll
    for ( int i = 0; i < 3; i++ ) s_red[threadIdx.x][i] = t[i];
    __syncthreads();
    for ( int o = 128; o > 0; o >>= 1 ) {
        if ( (int) threadIdx.x < o )
#pragma unroll
            for ( int i = 0; i < 3; i++ ) s_red[threadIdx.x][i] += s_red[threadIdx.x + o][i];
        __syncthreads();
    }
    if ( threadIdx.x < 3 ) out[threadIdx.x] = s_red[0][threadIdx.x];
}

// MaterialStatus::updateYourself: temp -> committed
__global__ void mises_commit_kernel(double *state, int64_t n)
{
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if ( i >= n ) return;
    const MisesStateRef st(state, i);
#pragma unroll
    for ( int k = 0; k < 6; k++ ) st.plStrain(k) = st.tempPlStrain(k);
    st.kappa() = st.tempKappa();
    st.damage() = st.tempDamage();
}

// the state as the C ABI exchanges it, point-major records [n][29], from / to the field-major device array [29][n]
template< bool TO_RECORDS >
__global__ void mises_state_transpose_kernel(double *__restrict__ fields, double *__restrict__ records, int64_t n)
{
    const int64_t total = n * OB200_MISES_STATE_DOUBLES, stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride ) {
        const int64_t g = t / OB200_MISES_STATE_DOUBLES;
        const int k = (int)( t - g * OB200_MISES_STATE_DOUBLES );
        double &fld = MisesStateRef(fields, g).f(k);
        if ( TO_RECORDS ) records[t] = fld;
        else fld = records[t];
    }
}

// element -> CSR slot map: slot[e][i*nd+j] = position of A(loc_i, loc_j) in val, -1 if prescribed
__global__ void slot_map_kernel(const int32_t *__restrict__ loc, int nd, int64_t nelem,
                                const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                int32_t *__restrict__ slot, int *__restrict__ missing)
{
    const int64_t total = nelem * nd * nd;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride ) {
        int64_t e = t / ( nd * nd );
        int ij = (int)( t - e * nd * nd );
        int i = ij / nd, j = ij - i * nd;
        int r = loc[e * nd + i], c = loc[e * nd + j];
        int32_t p = -1;
        if ( r > 0 && c > 0 ) {
            int lo = rowptr[r - 1], hi = rowptr[r] - 1, key = c - 1;
            while ( lo <= hi ) {
                int mid = ( lo + hi ) >> 1;
                int v = colind[mid];
                if ( v == key ) { p = mid; break; }
                if ( v < key ) lo = mid + 1; else hi = mid - 1;
            }
            if ( p < 0 ) atomicAdd(missing, 1);
        }
        slot[t] = p;
    }
}

// ---- host side of ob200_elemset ---------------------------------------------------------

static int launch_stiffness(ob200_elemset *S, int mode, double *out, const int32_t *slot)
{
    ob200_context *ctx = S->ctx;
    ElemSetView v = S->view();
    if ( S->nelem == 0 ) return OB200_OK;
    if ( S->etype == OB200_LSPACE ) {
        int grid = ctx->shape.grid(S->nelem * 32, kHexWarps * 32, 4);
        if ( mode == OUT_KE ) OB_LAUNCH(ctx, lspace_stiffness_kernel< OUT_KE >, grid, kHexWarps * 32, 0, v, out, slot, 0, S->nelem);
        else OB_LAUNCH(ctx, lspace_stiffness_kernel< OUT_CSR >, grid, kHexWarps * 32, 0, v, out, slot, 0, S->nelem);
    } else {
        int grid = ctx->shape.grid(S->nelem, 128, 8);
        if ( mode == OUT_KE ) OB_LAUNCH(ctx, ltrspace_stiffness_kernel< OUT_KE >, grid, 128, 0, v, out, slot, 0, S->nelem);
        else OB_LAUNCH(ctx, ltrspace_stiffness_kernel< OUT_CSR >, grid, 128, 0, v, out, slot, 0, S->nelem);
    }
    return OB200_OK;
}

static int launch_internal_forces(ob200_elemset *S, const double *u, double *fe, double *fglob, double *eps, double *sig,
                                  double *ebe = nullptr, int mode = FORCE_INTERNAL)
{
    ob200_context *ctx = S->ctx;
    ElemSetView v = S->view();
    if ( S->nelem == 0 ) return OB200_OK;
    // assembly into the global vector: owner-computes (no atomics) when the node incidence is there and no equation belongs to two
    // nodal dofs; OB200_VECTOR_ASSEMBLY=atomic keeps the scatter with atomicAdd (cross-check)
    const char *va = getenv("OB200_VECTOR_ASSEMBLY");
    const bool gather = fglob && S->eq_unique && S->row_vis.p && S->ninc_start.p && !( va && !strcmp(va, "atomic") );
    double *fvis = nullptr;
    if ( gather ) {
        if ( !S->fvis.p ) OB_CHECK( S->fvis.alloc(S->nvisit * 3) );
        fvis = S->fvis.p;
    }
    double *fg = gather ? nullptr : fglob, *eb = gather ? nullptr : ebe;
    if ( S->etype == OB200_LSPACE ) {
        int grid = ctx->shape.grid(( S->nelem + 3 ) / 4 * 32, kFwWarps * 32, 8);
        if ( mode == FORCE_INTERNAL )
            OB_LAUNCH(ctx, lspace_forces_kernel< FORCE_INTERNAL >, grid, kFwWarps * 32, 0, v, u, fe, fg, eps, sig, eb, fvis, S->row_vis.p, S->nelem);
        else
            OB_LAUNCH(ctx, lspace_forces_kernel< FORCE_TANGENT_DU >, grid, kFwWarps * 32, 0, v, u, fe, fg, eps, sig, eb, fvis, S->row_vis.p, S->nelem);
    } else {
        int grid = ctx->shape.grid(S->nelem, 128, 8);
        if ( mode == FORCE_INTERNAL )
            OB_LAUNCH(ctx, ltrspace_forces_kernel< FORCE_INTERNAL >, grid, 128, 0, v, u, fe, fg, eps, sig, eb, fvis, S->row_vis.p, 0, S->nelem);
        else
            OB_LAUNCH(ctx, ltrspace_forces_kernel< FORCE_TANGENT_DU >, grid, 128, 0, v, u, fe, fg, eps, sig, eb, fvis, S->row_vis.p, 0, S->nelem);
    }
    if ( gather ) {
        const int64_t nblocks = ( S->nnode + kForceGatherThreads - 1 ) / kForceGatherThreads;
        DevBuf< double > partial;
        if ( ebe ) OB_CHECK( partial.alloc(nblocks * 3) );
        OB_LAUNCH(ctx, node_force_gather_kernel, (int) nblocks, kForceGatherThreads, 0, S->nnode, S->ninc_start.p, S->nodeeq.p, fvis, fglob,
                  ebe ? partial.p : nullptr);
        if ( ebe ) OB_LAUNCH(ctx, norm_finish_kernel, 1, 256, 0, partial.p, nblocks, ebe);
    }
    return OB200_OK;
}

// ---- schedule cache -----------------------------------------------------------------------------------
// Content hash of an array of 32-bit words: sum over i of mix(i, a[i]) (integer additions: order independent, so the
// result does not depend on how the grid happens to be scheduled).
__global__ void content_hash_kernel(const uint32_t *__restrict__ a, int64_t n, unsigned long long salt, unsigned long long *__restrict__ out)
{
    unsigned long long acc = 0;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride ) {
        unsigned long long z = ( (unsigned long long) i * 0x9E3779B97F4A7C15ull ) ^ ( (unsigned long long) a[i] * 0xD6E8FEB86659FD93ull ) ^ salt;
        z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
        z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
        acc += z ^ ( z >> 31 );
    }
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 ) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ( ( threadIdx.x & 31 ) == 0 && acc ) atomicAdd(out, acc);
}

static void sched_cache_free(void *p) { delete static_cast< ob200_sched * >( p ); }

// key of the schedule: sizes + content hashes of connectivity, location arrays, coordinates, material numbers and
// material table (the records of the cluster assembly hold coordinates; the steps hold Lame constants)
static int sched_key(ob200_elemset *S, unsigned long long key[8])
{
    ob200_context *ctx = S->ctx;
    DevBuf< unsigned long long > h;
    OB_CHECK( h.alloc(5) );
    OB_CUDA( cudaMemsetAsync(h.p, 0, sizeof( unsigned long long ) * 5, ctx->stream) );
    const void *arr[5] = { S->conn.p, S->loc.p, S->coords.p, S->matid.p, S->mat.p };
    const int64_t words[5] = { S->nelem * S->nen, S->nelem * S->nd, S->nnode * 6, S->nelem, (int64_t) S->nmat * OB200_MATPARAM_STRIDE * 2 };
    for ( int i = 0; i < 5; i++ )
        if ( words[i] > 0 )
            OB_LAUNCH(ctx, content_hash_kernel, ctx->shape.grid(words[i], 256, 8), 256, 0, (const uint32_t *) arr[i], words[i],
                      (unsigned long long)( i + 1 ), h.p + i);
    OB_CUDA( cudaMemcpyAsync(key, h.p, sizeof( unsigned long long ) * 5, cudaMemcpyDeviceToHost, ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(ctx->stream) );
    const char *env = getenv("OB200_ASSEMBLY");
    key[5] = ( (unsigned long long) S->etype << 56 ) ^ ( (unsigned long long) S->nnode << 28 ) ^ (unsigned long long) S->nelem;
    key[6] = ( (unsigned long long) (unsigned int) S->neq << 32 ) ^ (unsigned long long) S->nmat;
    key[7] = 0;                                                                                     // the modes decide which schedule is built
    for ( const char *c = env; c && *c; c++ ) key[7] = key[7] * 131 + (unsigned char) *c;
    for ( const char *c = getenv("OB200_TET_ROWS"); c && *c; c++ ) key[7] = key[7] * 137 + (unsigned char) *c;
    return OB200_OK;
}

} // namespace ob200

using namespace ob200;

// Location arrays from nodal equation numbers: loc[e][3a+i] = nodeeq[conn[e][a]][i] -- what Element::giveLocationArray
// (src/core/element.C) collects from the dofs D_u, D_v, D_w of the element's nodes.  bad[0] counts connectivity entries
// outside 1..nnode (their location entries are set to 0).
__global__ void loc_from_nodeeq_kernel(int64_t nelem, int nen, int64_t nnode, const int32_t *__restrict__ conn,
                                       const int32_t *__restrict__ nodeeq, int32_t *__restrict__ loc, int *__restrict__ bad)
{
    const int64_t n = nelem * nen, stride = (int64_t) gridDim.x * blockDim.x;
    for ( int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride ) {
        const int64_t node = (int64_t) conn[t] - 1;
        int32_t e0 = 0, e1 = 0, e2 = 0;
        if ( node >= 0 && node < nnode ) {
            e0 = nodeeq[3 * node];
            e1 = nodeeq[3 * node + 1];
            e2 = nodeeq[3 * node + 2];
        } else {
            atomicAdd(bad, 1);
        }
        loc[3 * t] = e0;
        loc[3 * t + 1] = e1;
        loc[3 * t + 2] = e2;
    }
}

// loc [nelem][3*nen] or (loc == nullptr) nodeeq [nnode][3]
static int elemset_create_impl(ob200_context *ctx, int etype, int64_t nnode, const double *coords, int64_t nelem,
                               const int32_t *conn, const int32_t *matid, int32_t nmat, const double *matparams,
                               const int32_t *loc, const int32_t *nodeeq, int32_t neq, int on_device, ob200_elemset **out)
{
    if ( ctx ) ob200::bind_stream(ctx);
    OB_REQUIRE(ctx && out, OB200_EINVAL, "elemset_create: null context/out");
    OB_REQUIRE(etype == OB200_LSPACE || etype == OB200_LTRSPACE, OB200_EINVAL, "elemset_create: unknown element type %d", etype);
    OB_REQUIRE(nnode >= 0 && nelem >= 0 && nmat >= 1, OB200_EINVAL, "elemset_create: negative size or no material");
    ob200_elemset *S = new ob200_elemset();
    S->ctx = ctx;
    S->etype = etype;
    S->nen = etype == OB200_LSPACE ? 8 : 4;
    S->ngp = etype == OB200_LSPACE ? 8 : 1;
    S->nd = 3 * S->nen;
    S->nnode = nnode;
    S->nelem = nelem;
    S->nmat = nmat;
    S->neq = neq;
    int rc = OB200_OK;
    auto put = [&](auto &buf, const auto *src, int64_t n) -> int {
        OB_CHECK( buf.alloc(n) );
        if ( n == 0 ) return OB200_OK;
        OB_CUDA( cudaMemcpyAsync(buf.p, src, sizeof( *src ) * (size_t) n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream) );
        return OB200_OK;
    };
    // the small tables and what the node incidence needs go first on the main stream; the location arrays
    // (the largest upload) follow on the copy stream and are awaited where they are first read
    if ( ( rc = put(S->mat, matparams, (int64_t) nmat * OB200_MATPARAM_STRIDE) ) < 0 || ( rc = put(S->conn, conn, nelem * S->nen) ) < 0 ||
         ( rc = put(S->coords, coords, nnode * 3) ) < 0 || ( rc = put(S->matid, matid, nelem) ) < 0 ||
         ( rc = S->loc.alloc(nelem * S->nd) ) < 0 ) {
        delete S;
        return rc;
    }
    if ( nelem > 0 && !loc ) {
        // nodal equation numbers: 12 bytes per node travel instead of 4 * nd bytes per element; loc is formed on the device
        DevBuf< int32_t > ne;
        DevBuf< int > bad;
        int hb = 0;
        if ( ( rc = put(ne, nodeeq, nnode * 3) ) < 0 || ( rc = bad.alloc(1) ) < 0 ) { delete S; return rc; }
        cudaMemsetAsync(bad.p, 0, sizeof( int ), ctx->stream);
        loc_from_nodeeq_kernel<<< ctx->shape.grid(nelem * S->nen, 256, 8), 256, 0, ctx->stream >>>(nelem, S->nen, nnode, S->conn.p, ne.p, S->loc.p, bad.p);
        ctx->launches++;
        if ( cudaMemcpyAsync(&hb, bad.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
             cudaStreamSynchronize(ctx->stream) != cudaSuccess ) {
            set_error("elemset_create_nodal: forming the location arrays failed (%s)", cudaGetErrorString(cudaGetLastError()));
            delete S;
            return OB200_ECUDA;
        }
        if ( hb > 0 ) {
            set_error("elemset_create_nodal: %d connectivity entries outside 1..%lld", hb, (long long) nnode);
            delete S;
            return OB200_EINVAL;
        }
    } else if ( nelem > 0 ) {
        cudaError_t e = cudaEventRecord(ctx->copy_event, ctx->stream);                 // the allocation is ordered on the main stream
        if ( e == cudaSuccess ) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_event, 0);
        if ( e == cudaSuccess )
            e = cudaMemcpyAsync(S->loc.p, loc, sizeof( int32_t ) * (size_t)( nelem * S->nd ),
                                on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->copy_stream);
        if ( e == cudaSuccess ) e = cudaEventRecord(ctx->copy_event, ctx->copy_stream);
        if ( e != cudaSuccess ) {
            set_error("elemset_create: upload of the location arrays failed (%s)", cudaGetErrorString(e));
            delete S;
            return OB200_ECUDA;
        }
        S->loc_pending = true;
    }
    // material state only when a MisesMat is present (host copy of the small parameter table decides)
    std::vector< double > mp( (size_t) nmat * OB200_MATPARAM_STRIDE );
    if ( cudaMemcpyAsync(mp.data(), S->mat.p, sizeof( double ) * mp.size(), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
         cudaStreamSynchronize(ctx->stream) != cudaSuccess ) {
        set_error("elemset_create: reading material table failed");
        delete S;
        return OB200_ECUDA;
    }
    for ( int m = 0; m < nmat; m++ ) {
        int t = (int) mp[(size_t) m * OB200_MATPARAM_STRIDE];
        if ( t != OB200_MAT_ISOLE && t != OB200_MAT_MISES ) {
            set_error("elemset_create: unsupported material type %d", t);
            delete S;
            return OB200_EINVAL;
        }
        if ( t == OB200_MAT_MISES ) {
            S->has_state = true;
            S->all_isole = false;
        }
    }
    if ( etype == OB200_LTRSPACE && nelem > 0 ) {
        DevBuf< int > bad;
        int hb[2] = { 0, 0x7FFFFFFF };
        if ( ( rc = bad.alloc(2) ) < 0 ) { delete S; return rc; }
        cudaMemcpyAsync(bad.p, hb, sizeof( hb ), cudaMemcpyHostToDevice, ctx->stream);
        tet_volume_check_kernel<<< ctx->shape.grid(nelem, 128, 8), 128, 0, ctx->stream >>>(S->view(), nelem, bad.p);
        ctx->launches++;
        if ( cudaMemcpyAsync(hb, bad.p, sizeof( hb ), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
             cudaStreamSynchronize(ctx->stream) != cudaSuccess ) {
            set_error("elemset_create: volume check failed (%s)", cudaGetErrorString(cudaGetLastError()));
            delete S;
            return OB200_ECUDA;
        }
        if ( hb[0] > 0 ) {
            set_error("FEI3dTetLin::evaldNdx: negative volume (%d LTRSpace elements, first: element %d)", hb[0], hb[1] + 1);
            delete S;
            return OB200_EINVAL;
        }
    }
    // the schedule of the last destroyed set is adopted if it was built from identical arrays (OB200_SCHED_CACHE=0 disables)
    const bool use_cache = !( getenv("OB200_SCHED_CACHE") && !strcmp(getenv("OB200_SCHED_CACHE"), "0") );
    bool adopted = false;
    if ( use_cache && nelem > 0 ) {
        if ( ( rc = elemset_await_loc(S) ) < 0 ) { delete S; return rc; }
        unsigned long long key[8];
        if ( ( rc = sched_key(S, key) ) < 0 ) { delete S; return rc; }
        ob200_sched *C = static_cast< ob200_sched * >( ctx->sched_cache );
        if ( C && C->have_mesh && !memcmp(C->key, key, sizeof( key )) ) {
            static_cast< ob200_sched & >( *S ) = std::move(*C);
            delete C;
            ctx->sched_cache = nullptr;
            adopted = true;
        }
        memcpy(S->key, key, sizeof( key ));
    }
    if ( !adopted ) {
        if ( ( rc = gather_prepare_mesh(S) ) < 0 ) { delete S; return rc; }
        S->have_mesh = true;
        S->have_bind = false;
    }
    if ( ( rc = elemset_await_loc(S) ) < 0 ) { delete S; return rc; }
    if ( !on_device && cudaStreamSynchronize(ctx->copy_stream) != cudaSuccess ) {      // the caller's host buffer is free again on return
        set_error("elemset_create: upload of the location arrays failed");
        delete S;
        return OB200_ECUDA;
    }
    if ( S->has_state ) {
        int64_t n = ( ( nelem * S->ngp + 31 ) / 32 ) * 32 * OB200_MISES_STATE_DOUBLES;        // whole blocks of 32 Gauss points
        if ( ( rc = S->state.alloc(n) ) < 0 ) { delete S; return rc; }
        if ( n && cudaMemsetAsync(S->state.p, 0, sizeof( double ) * (size_t) n, ctx->stream) != cudaSuccess ) {
            set_error("elemset_create: memset failed");
            delete S;
            return OB200_ECUDA;
        }
    }
    *out = S;
    return OB200_OK;
}

extern "C" {

int ob200_elemset_create(ob200_context *ctx, int etype, int64_t nnode, const double *coords, int64_t nelem,
                         const int32_t *conn, const int32_t *matid, int32_t nmat, const double *matparams,
                         const int32_t *loc, int32_t neq, int on_device, ob200_elemset **out)
{
    OB_REQUIRE(loc || nelem <= 0, OB200_EINVAL, "elemset_create: null location arrays");
    return elemset_create_impl(ctx, etype, nnode, coords, nelem, conn, matid, nmat, matparams, loc, nullptr, neq, on_device, out);
}

int ob200_elemset_create_nodal(ob200_context *ctx, int etype, int64_t nnode, const double *coords, int64_t nelem,
                               const int32_t *conn, const int32_t *matid, int32_t nmat, const double *matparams,
                               const int32_t *nodeeq, int32_t neq, int on_device, ob200_elemset **out)
{
    OB_REQUIRE(nodeeq || nnode <= 0 || nelem <= 0, OB200_EINVAL, "elemset_create_nodal: null nodal equation numbers");
    return elemset_create_impl(ctx, etype, nnode, coords, nelem, conn, matid, nmat, matparams, nullptr, nodeeq, neq, on_device, out);
}

void ob200_elemset_destroy(ob200_elemset *S)
{
    if ( !S ) return;
    // keep the schedule for an element set that may be created from the same arrays (a re-upload of an unchanged mesh)
    if ( S->ctx && S->have_mesh && S->key[5] != 0 && stream_alive(S->ctx->stream) ) {
        ob200::bind_stream(S->ctx);
        if ( S->ctx->sched_cache ) sched_cache_free(S->ctx->sched_cache);
        S->ctx->sched_cache = new ob200_sched(std::move(static_cast< ob200_sched & >( *S )));
        S->ctx->sched_cache_free = sched_cache_free;
    }
    delete S;
}

int64_t ob200_elemset_size(const ob200_elemset *S) { return S ? S->nelem : 0; }

int ob200_elemset_stiffness(ob200_elemset *S, double *Ke, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && Ke, OB200_EINVAL, "elemset_stiffness: null argument");
    StagedOut< double > o;
    OB_CHECK( o.stage(S->ctx, Ke, S->nelem * S->nd * S->nd, on_device) );
    // LSpace: the FP64 tensor-path kernel of the strip assembly (assemble_strips.cu); OB200_KE=dfma keeps the lane-per-block kernel
    const char *ke = getenv("OB200_KE");
    if ( S->etype == OB200_LSPACE && !( ke && !strcmp(ke, "dfma") ) && ( reinterpret_cast< uintptr_t >( o.d ) & 15 ) == 0 )
        OB_CHECK( strips_element_matrices(S, o.d, nullptr) );
    else
        OB_CHECK( launch_stiffness(S, OUT_KE, o.d, nullptr) );
    return o.finish(S->ctx);
}

int ob200_elemset_internal_forces(ob200_elemset *S, const double *u, double *fe, double *gp_strain, double *gp_stress, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && u, OB200_EINVAL, "elemset_internal_forces: null argument");
    Staged< double > du;
    StagedOut< double > of, oe, os;
    OB_CHECK( du.stage(S->ctx, u, S->nnode * 3, on_device) );
    if ( fe ) OB_CHECK( of.stage(S->ctx, fe, S->nelem * S->nd, on_device) );
    if ( gp_strain ) OB_CHECK( oe.stage(S->ctx, gp_strain, S->nelem * S->ngp * 6, on_device) );
    if ( gp_stress ) OB_CHECK( os.stage(S->ctx, gp_stress, S->nelem * S->ngp * 6, on_device) );
    OB_CHECK( launch_internal_forces(S, du.d, fe ? of.d : nullptr, nullptr, gp_strain ? oe.d : nullptr, gp_stress ? os.d : nullptr) );
    OB_CHECK( of.finish(S->ctx) );
    OB_CHECK( oe.finish(S->ctx) );
    return os.finish(S->ctx);
}

// element -> CSR slot map of the generic (atomic scatter) path; built on first use
static int build_slot_map(ob200_elemset *S, ob200_csr *A)
{
    if ( S->slot_built ) return OB200_OK;
    int64_t total = S->nelem * S->nd * S->nd;
    OB_CHECK( S->slot.alloc(total) );
    DevBuf< int > missing;
    OB_CHECK( missing.alloc(1) );
    OB_CUDA( cudaMemsetAsync(missing.p, 0, sizeof( int ), S->ctx->stream) );
    if ( total ) {
        int grid = S->ctx->shape.grid(total, 256, 8);
        OB_LAUNCH(S->ctx, slot_map_kernel, grid, 256, 0, S->loc.p, S->nd, S->nelem, A->rowptr.p, A->colind.p, S->slot.p, missing.p);
    }
    int h = 0;
    OB_CUDA( cudaMemcpyAsync(&h, missing.p, sizeof( int ), cudaMemcpyDeviceToHost, S->ctx->stream) );
    OB_CUDA( cudaStreamSynchronize(S->ctx->stream) );
    // CompCol::assemble DEBUG branch: "Couldn't find row %d in the sparse structure" (compcol.C:288-290)
    OB_REQUIRE(h == 0, OB200_ESTRUCT, "elemset_bind: %d element entries are not in the sparse structure", h);
    S->slot_built = true;
    return OB200_OK;
}

int ob200_elemset_bind(ob200_elemset *S, ob200_csr *A)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && A, OB200_EINVAL, "elemset_bind: null argument");
    OB_REQUIRE(A->rowptr.p, OB200_EINVAL, "elemset_bind: matrix has no structure (call ob200_csr_build_structure first)");
    S->bound = nullptr;
    if ( !( S->have_bind && S->bind_structure == A->structure_version ) ) {       // else: adopted from the schedule cache
        S->have_bind = false;
        S->slot_built = false;
        // preferred: owner-computes assembly (no atomics); its preparation also verifies that the
        // matrix pattern is the one of this element set
        OB_CHECK( gather_bind(S, A) );
        if ( !S->gather_ok && !S->strips_ok ) OB_CHECK( tet_bind(S, A) );
        if ( !S->gather_ok && !S->strips_ok && !S->rows_ok ) OB_CHECK( build_slot_map(S, A) );
        S->have_bind = true;
        S->bind_structure = A->structure_version;
    }
    S->bound = A;
    S->bound_version = A->structure_version;
    return OB200_OK;
}

int ob200_elemset_assemble_stiffness(ob200_elemset *S, ob200_csr *A)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && A, OB200_EINVAL, "elemset_assemble_stiffness: null argument");
    if ( S->bound != A || S->bound_version != A->structure_version ) OB_CHECK( ob200_elemset_bind(S, A) );
    if ( S->gather_ok ) {
        if ( S->nelem ) OB_CHECK( S->cluster_ok ? cluster_assemble_lspace(S, A) : gather_assemble_lspace(S, A) );
        ob200_csr_touch(A);
        return OB200_OK;
    }
    if ( S->strips_ok ) {
        if ( S->nelem ) OB_CHECK( strips_assemble_lspace(S, A) );
        ob200_csr_touch(A);
        return OB200_OK;
    }
    if ( S->rows_ok ) {
        if ( S->nelem ) OB_CHECK( tet_assemble_ltrspace(S, A) );
        ob200_csr_touch(A);
        return OB200_OK;
    }
    OB_CHECK( build_slot_map(S, A) );
    OB_CHECK( ob200_csr_materialize(A) );
    OB_CHECK( launch_stiffness(S, OUT_CSR, A->val.p, S->slot.p) );
    ob200_csr_touch(A);
    return OB200_OK;
}

int ob200_elemset_assemble_internal_forces(ob200_elemset *S, const double *u, double *f, double *ebe_norm2, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && u && f, OB200_EINVAL, "elemset_assemble_internal_forces: null argument");
    Staged< double > du;
    StagedOut< double > of;
    OB_CHECK( du.stage(S->ctx, u, S->nnode * 3, on_device) );
    OB_CHECK( of.stage(S->ctx, f, S->neq_hint(), on_device, true) );
    DevBuf< double > ebe;
    if ( ebe_norm2 ) {
        OB_CHECK( ebe.alloc(3) );
        OB_CUDA( cudaMemsetAsync(ebe.p, 0, sizeof( double ) * 3, S->ctx->stream) );
    }
    OB_CHECK( launch_internal_forces(S, du.d, nullptr, of.d, nullptr, nullptr, ebe.p) );
    if ( ebe_norm2 ) {      // always returned to the host: three scalars for the convergence test
        OB_CUDA( cudaMemcpyAsync(ebe_norm2, ebe.p, sizeof( double ) * 3, cudaMemcpyDeviceToHost, S->ctx->stream) );
        OB_CUDA( cudaStreamSynchronize(S->ctx->stream) );
    }
    return of.finish(S->ctx);
}

int ob200_elemset_assemble_extrapolated_forces(ob200_elemset *S, const double *du, double *f, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && du && f, OB200_EINVAL, "elemset_assemble_extrapolated_forces: null argument");
    Staged< double > d;
    StagedOut< double > of;
    OB_CHECK( d.stage(S->ctx, du, S->nnode * 3, on_device) );
    OB_CHECK( of.stage(S->ctx, f, S->neq_hint(), on_device, true) );
    OB_CHECK( launch_internal_forces(S, d.d, nullptr, of.d, nullptr, nullptr, nullptr, FORCE_TANGENT_DU) );
    return of.finish(S->ctx);
}

int ob200_elemset_commit(ob200_elemset *S)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S, OB200_EINVAL, "elemset_commit: null argument");
    if ( !S->has_state ) return OB200_OK;
    int64_t n = S->nelem * S->ngp;
    if ( n ) OB_LAUNCH(S->ctx, mises_commit_kernel, (int) ceil_div(n, 256), 256, 0, S->state.p, n);
    return OB200_OK;
}

int ob200_elemset_get_state(ob200_elemset *S, double *state, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && state, OB200_EINVAL, "elemset_get_state: null argument");
    OB_REQUIRE(S->has_state, OB200_EINVAL, "elemset_get_state: element set has no MisesMat state");
    // the device array is field-major (element_device.cuh); the caller gets the records [ngp][29]
    const int64_t n = S->nelem * S->ngp;
    StagedOut< double > o;
    OB_CHECK( o.stage(S->ctx, state, n * OB200_MISES_STATE_DOUBLES, on_device) );
    if ( n ) OB_LAUNCH(S->ctx, mises_state_transpose_kernel< true >, S->ctx->shape.grid(n * OB200_MISES_STATE_DOUBLES, 256, 8), 256, 0, S->state.p, o.d, n);
    OB_CHECK( o.finish(S->ctx) );
    OB_CUDA( cudaStreamSynchronize(S->ctx->stream) );
    return OB200_OK;
}

int ob200_elemset_set_state(ob200_elemset *S, const double *state, int on_device)
{
    if ( S ) ob200::bind_stream(S->ctx);
    OB_REQUIRE(S && state, OB200_EINVAL, "elemset_set_state: null argument");
    OB_REQUIRE(S->has_state, OB200_EINVAL, "elemset_set_state: element set has no MisesMat state");
    const int64_t n = S->nelem * S->ngp;
    Staged< double > in;
    OB_CHECK( in.stage(S->ctx, state, n * OB200_MISES_STATE_DOUBLES, on_device) );
    if ( n ) OB_LAUNCH(S->ctx, mises_state_transpose_kernel< false >, S->ctx->shape.grid(n * OB200_MISES_STATE_DOUBLES, 256, 8), 256, 0, S->state.p,
                       const_cast< double * >( in.d ), n);
    OB_CUDA( cudaStreamSynchronize(S->ctx->stream) );
    return OB200_OK;
}

} // extern "C"

// ---- scratch probes (not part of the ABI; used by scripts/probe_atomics.py) ----------------
namespace ob200 {
// mode 0: RED.F64 through the slot map, no math; mode 1: plain store through the slot map;
// mode 2: slot-map read only
__global__ void __launch_bounds__(256) probe_scatter_kernel(const int32_t *__restrict__ slot, double *__restrict__ val,
                                                             int64_t nelem, int mode, int *sink)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ( (int64_t) blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
    const int64_t nwarps = ( (int64_t) gridDim.x * blockDim.x ) >> 5;
    int acc = 0;
    for ( int64_t e = warp0; e < nelem; e += nwarps ) {
        const int32_t *sl = slot + e * 576;
#pragma unroll 6
        for ( int t = lane; t < 576; t += 32 ) {
            int32_t p = sl[t];
            if ( mode == 2 ) acc += p;
            else if ( p >= 0 ) {
                if ( mode == 0 ) atomicAdd(val + p, 1.0);
                else val[p] = 1.0;
            }
        }
    }
    if ( mode == 2 && acc == 123456789 ) *sink = acc;
}
}
extern "C" int ob200_debug_probe_scatter(ob200_elemset *S, ob200_csr *A, int mode, int blocks_per_sm)
{
    if ( S ) ob200::bind_stream(S->ctx);
    if ( S->bound != A || S->bound_version != A->structure_version ) OB_CHECK( ob200_elemset_bind(S, A) );
    OB_CHECK( build_slot_map(S, A) );
    const int32_t *rowptr, *colind;
    double *val;
    OB_CHECK( ob200_csr_device_arrays(A, &rowptr, &colind, &val) );
    int grid = S->ctx->shape.sms * blocks_per_sm;
    OB_LAUNCH(S->ctx, probe_scatter_kernel, grid, 256, 0, S->slot.p, val, S->nelem, mode, (int *) S->slot.p);
    return OB200_OK;
}
