// Nodal force vectors of LSpace / LTRSpace elements: StructuralElement::giveInternalForcesVector with useUpdatedGpRecord = 0
// (src/sm/Elements/structuralelement.C:724-802) and the tangent applied to a displacement increment, K_e du_e
// (StaticStructural::assembleExtrapolatedForces through TangentAssembler::vectorFromElement), evaluated as B^T (D B du) dV.
//
// LSpace: a warp handles four elements, one thread per (element, Gauss point): Jacobian, gradients of all eight nodes, strain,
// stress (IsotropicLinearElasticMaterial, or the MisesMat radial return with the state written to the temp record), and the
// point's 24 nodal force components, which go to shared memory; then one lane per (element, node) adds the eight Gauss points in
// order (answer.plusProduct(B, stress, dV) of the reference's loop) and stores / scatters the node's three components.
// LTRSpace: one thread per element.
#pragma once
#include "element_device.cuh"
#include "elemset.h"

namespace ob200 {

enum { FORCE_INTERNAL = 0, FORCE_TANGENT_DU = 1 };

// sigma = D_tangent eps with the tangent of MisesMat::give3dMaterialStiffnessMatrix (misesmat.C:493-545), without forming D:
// D = ( De + f1 t t^T + f2 Idev ) ( 1 - omega ) + scalar es t^T
__device__ __forceinline__ void mises_tangent_apply(const MatParams &mp, const MisesStateRef &st, const double eps[6], double sig[6])
{
    const double G = mp.E / ( 2.0 * ( 1.0 + mp.nu ) );
    double lam, mu;
    isole_lame(mp.E, mp.nu, lam, mu);
    iso_stress(lam, mu, eps, sig);
    const double kappa = st.kappa(), tempKappa = st.tempKappa();
    const double dKappa = tempKappa - kappa;
    if ( dKappa <= 0.0 ) return;
    double t[6], es[6];
#pragma unroll
    for ( int i = 0; i < 6; i++ ) {
        t[i] = st.trialStressDev(i);
        es[i] = st.effStress(i);
    }
    const double sigmaY = mp.sig0 + mp.H * kappa;
    const double trialS = dev_norm(t);
    const double factor = -2.0 * sqrt(6.0) * G * G / trialS;
    const double factor1 = factor * sigmaY / ( ( mp.H + 3.0 * G ) * trialS * trialS );
    const double factor2 = factor * dKappa;
    const double omega = st.tempDamage();
    const double omegaPrime = tempKappa >= 0.0 ? mp.omega_crit * mp.a * exp(-mp.a * tempKappa) : 0.0;
    const double scalar = -omegaPrime * sqrt(6.0) * G / ( 3.0 * G + mp.H ) / trialS;
    double te = 0.0;
#pragma unroll
    for ( int j = 0; j < 6; j++ ) te += t[j] * eps[j];
    const double tr3 = ( eps[0] + eps[1] + eps[2] ) * ( 1.0 / 3.0 );
#pragma unroll
    for ( int i = 0; i < 6; i++ ) {
        const double idev = i < 3 ? eps[i] - tr3 : 0.5 * eps[i];
        sig[i] = ( sig[i] + factor1 * t[i] * te + factor2 * idev ) * ( 1.0 - omega ) + scalar * es[i] * te;
    }
}

// stress of one Gauss point for the two modes
template< int MODE >
__device__ __forceinline__ void point_stress(const MatParams &mp, const ElemSetView &S, int64_t gpoint, const double eps[6], double sig[6])
{
    if ( mp.type == (double) OB200_MAT_MISES ) {
        const MisesStateRef st = mises_ref(S, gpoint);
        if ( MODE == FORCE_INTERNAL ) mises_stress(mp, eps, st, sig);
        else mises_tangent_apply(mp, st, eps, sig);
    } else {
        double lam, mu;
        isole_lame(mp.E, mp.nu, lam, mu);
        iso_stress(lam, mu, eps, sig);          // LinearElasticMaterial::giveRealStressVector_3d; the tangent is the same matrix
    }
}

constexpr int kFwWarps = 4;
constexpr int kFwXs = 26;                // doubles per element in the coordinate / displacement stage (24 + pad)
constexpr int kFwFs = 25;                // doubles per (element, Gauss point) in the force stage (24 + pad)

struct FwShared {
    double x[4][kFwXs], u[4][kFwXs];
    double f[4 * 8 * kFwFs];
};

// fe: element vectors [nelem][24]; fglob: scatter-add through loc (atomicAdd); fvis: the nodal forces in incidence order for the
// owner-computes assembly (node_force_gather_kernel); gp_strain / gp_stress [nelem * 8][6]; ebe_norm2[3]: element-by-element norms
template< int MODE >
__global__ void __launch_bounds__(kFwWarps * 32, 3)
lspace_forces_kernel(ElemSetView S, const double *__restrict__ u, double *__restrict__ fe, double *__restrict__ fglob,
                     double *__restrict__ gp_strain, double *__restrict__ gp_stress, double *__restrict__ ebe_norm2,
                     double *__restrict__ fvis, const int32_t *__restrict__ vis, int64_t nelem)
{
    __shared__ FwShared sh[kFwWarps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int el = lane >> 3, gp = lane & 7;
    FwShared &s = sh[wid];
    double ebe[3] = { 0.0, 0.0, 0.0 };   // sums of f^2 per dof id over this lane's (element, node) pairs
    const int64_t ngroups = ( nelem + 3 ) >> 2, stride = (int64_t) gridDim.x * kFwWarps;
    for ( int64_t grp = (int64_t) blockIdx.x * kFwWarps + wid; grp < ngroups; grp += stride ) {
        const int64_t e = grp * 4 + el;
        const bool valid = e < nelem;
        {   // stage: this lane's node of its element (connectivity -> coordinates, displacements)
            const int64_t node = valid ? S.conn[e * 8 + gp] - 1 : 0;
#pragma unroll
            for ( int j = 0; j < 3; j++ ) {
                s.x[el][3 * gp + j] = S.coords[node * 3 + j];
                s.u[el][3 * gp + j] = u[node * 3 + j];       // computeVectorOf(VM_Total) / the increment
            }
        }
        __syncwarp();
        if ( valid ) {
            // geometry at the Gauss point: every factor (1 +- xi_gp)/... of dN/dxi takes one of two values, the twelve products are
            // constants picked by the signs of the Gauss point and of the node (FEI3dHexaLin::evaldNdxi, fei3dhexalin.C:129-166)
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
            double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } }, dN[8][3];
            const double *xe = s.x[el];
#pragma unroll
            for ( int kk = 0; kk < 8; kk++ ) {
                const int px = ( kk & 3 ) >= 2, py = ( ( kk & 3 ) == 1 || ( kk & 3 ) == 2 ), pz = kk < 4;
                dN[kk][0] = px ? Pyz[py][pz] : -Pyz[py][pz];
                dN[kk][1] = py ? Pxz[px][pz] : -Pxz[px][pz];
                dN[kk][2] = pz ? Pxy[px][py] : -Pxy[px][py];
                const double x = xe[3 * kk], y = xe[3 * kk + 1], z = xe[3 * kk + 2];
                J[0][0] += x * dN[kk][0]; J[0][1] += x * dN[kk][1]; J[0][2] += x * dN[kk][2];
                J[1][0] += y * dN[kk][0]; J[1][1] += y * dN[kk][1]; J[1][2] += y * dN[kk][2];
                J[2][0] += z * dN[kk][0]; J[2][1] += z * dN[kk][1]; J[2][2] += z * dN[kk][2];
            }
            double Ji[3][3];
            const double dV = fabs(inv3(J, Ji));          // FEI3dHexaLin::evaldNdx (fei3dhexalin.C:186-204); weights 1
            double g[8][3], eps[6] = { 0, 0, 0, 0, 0, 0 }, sig[6];
            const double *ue = s.u[el];
#pragma unroll
            for ( int kk = 0; kk < 8; kk++ ) {
#pragma unroll
                for ( int j = 0; j < 3; j++ ) g[kk][j] = dN[kk][0] * Ji[0][j] + dN[kk][1] * Ji[1][j] + dN[kk][2] * Ji[2][j];
                const double uk[3] = { ue[3 * kk], ue[3 * kk + 1], ue[3 * kk + 2] };
                strain_add(eps, g[kk], uk);               // strain = B u
            }
            const MatParams mp = S.mat[S.matid[e]];
            point_stress< MODE >(mp, S, e * 8 + gp, eps, sig);
            if ( gp_strain )
#pragma unroll
                for ( int i = 0; i < 6; i++ ) gp_strain[( e * 8 + gp ) * 6 + i] = eps[i];
            if ( gp_stress )
#pragma unroll
                for ( int i = 0; i < 6; i++ ) gp_stress[( e * 8 + gp ) * 6 + i] = sig[i];
            double *fo = s.f + ( el * 8 + gp ) * kFwFs;
#pragma unroll
            for ( int kk = 0; kk < 8; kk++ ) {
                double fk[3];
                force_node(fk, g[kk], sig, dV);
                fo[3 * kk] = fk[0];
                fo[3 * kk + 1] = fk[1];
                fo[3 * kk + 2] = fk[2];
            }
        }
        __syncwarp();
        if ( valid ) {   // lane = (element, node gp): the eight Gauss points in order
            double f[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
            for ( int q = 0; q < 8; q++ ) {
                const double *fq = s.f + ( el * 8 + q ) * kFwFs + 3 * gp;
                f[0] += fq[0];
                f[1] += fq[1];
                f[2] += fq[2];
            }
#pragma unroll
            for ( int c = 0; c < 3; c++ ) {
                ebe[c] += f[c] * f[c];
                if ( fe ) fe[e * 24 + 3 * gp + c] = f[c];
                if ( fglob ) {
                    const int32_t r = S.loc[e * 24 + 3 * gp + c];
                    if ( r > 0 && ( MODE == FORCE_INTERNAL || f[c] != 0.0 ) ) atomicAdd(fglob + r - 1, f[c]);
                }
            }
            if ( fvis ) {
                double *o = fvis + (int64_t) vis[e * 8 + gp] * 3;
                o[0] = f[0];
                o[1] = f[1];
                o[2] = f[2];
            }
        }
        __syncwarp();
    }
    if ( ebe_norm2 ) {     // element-by-element norm per dof id (EngngModel::assembleVector eNorms, engngm.C:1108-1133)
#pragma unroll
        for ( int c = 0; c < 3; c++ ) {
            double v = ebe[c];
#pragma unroll
            for ( int o = 16; o > 0; o >>= 1 ) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ( lane == 0 && v != 0.0 ) atomicAdd(ebe_norm2 + c, v);
        }
    }
}

} // namespace ob200
