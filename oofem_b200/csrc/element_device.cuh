// Device-side element mathematics: FEI3dHexaLin / FEI3dTetLin geometry, B-matrix algebra,
// IsotropicLinearElasticMaterial and MisesMat.  Reference lines are cited per function.
#pragma once
#include "common.cuh"
#include "elemset.h"

namespace ob200 {

// Per-Gauss-point MisesMat state (the MisesMatStatus fields the 3D path uses, src/sm/Materials/misesmat.h): 29 doubles per
// Gauss point.  MisesState is the record as the C ABI exchanges it (ob200_elemset_get_state / set_state: [ngp][29]).  On the device
// the state is field-major within blocks of 32 Gauss points, state[ngp / 32][29][32]: the lanes of a warp work on consecutive
// Gauss points, so every access is a coalesced run of doubles (with point-major records the internal-force kernel spent 2.7 of its
// 3.3 ms at 1M hex on 29 scattered 8-byte accesses per point), and the 29 fields of a block stay within 7.4 KB (a kernel that
// reads the eight points of one element per warp keeps its DRAM locality).  MisesStateRef is one Gauss point of that array.
struct MisesState {
    double plStrain[6];
    double kappa;
    double damage;
    double tempPlStrain[6];
    double tempKappa;
    double tempDamage;
    double trialStressDev[6];
    double trialStressVol;
    double effStress[6];
};
static_assert( sizeof( MisesState ) == OB200_MISES_STATE_DOUBLES * sizeof( double ), "state layout" );

struct MisesStateRef {
    double *b;          // the Gauss point's slot in field 0 of its block
    __device__ __forceinline__ MisesStateRef(double *state, int64_t g) : b(state + ( g >> 5 ) * ( OB200_MISES_STATE_DOUBLES * 32 ) + ( g & 31 )) {}
    __device__ __forceinline__ double &f(int k) const { return b[k * 32]; }
    __device__ __forceinline__ double &plStrain(int i) const { return f(i); }
    __device__ __forceinline__ double &kappa() const { return f(6); }
    __device__ __forceinline__ double &damage() const { return f(7); }
    __device__ __forceinline__ double &tempPlStrain(int i) const { return f(8 + i); }
    __device__ __forceinline__ double &tempKappa() const { return f(14); }
    __device__ __forceinline__ double &tempDamage() const { return f(15); }
    __device__ __forceinline__ double &trialStressDev(int i) const { return f(16 + i); }
    __device__ __forceinline__ double &trialStressVol() const { return f(22); }
    __device__ __forceinline__ double &effStress(int i) const { return f(23 + i); }
};

// Gauss point g (element * ngp + point) of an element set's state
__device__ __forceinline__ MisesStateRef mises_ref(const ElemSetView &S, int64_t g) { return MisesStateRef(S.state, g); }

struct MatParams {   // [type, E, nu, sig0, H, omega_crit, a, pad]
    double type, E, nu, sig0, H, omega_crit, a, pad;
};

// ---- hexahedron ---------------------------------------------------------------------

// node signs of FEI3dHexaLin (src/core/fei3dhexalin.C:46-63): N_k = (1+sx u)(1+sy v)(1+sz w)/8
__device__ __forceinline__ void hexa_signs(int k, double &sx, double &sy, double &sz)
{
    int q = k & 3;
    sx = ( q >= 2 ) ? 1.0 : -1.0;
    sy = ( q == 1 || q == 2 ) ? 1.0 : -1.0;
    sz = ( k < 4 ) ? 1.0 : -1.0;
}

// Gauss point g of the 2x2x2 rule, loop order of GaussIntegrationRule::SetUpPointsOnCube
// (src/core/gaussintegrationrule.C:202-210): xi outermost, zeta innermost; weights are 1.
__device__ __forceinline__ void hexa_gp(int g, double &u, double &v, double &w)
{
    const double a = 0.577350269189626;     // gaussintegrationrule.C:1450
    u = ( g & 4 ) ? a : -a;
    v = ( g & 2 ) ? a : -a;
    w = ( g & 1 ) ? a : -a;
}

// dN_k/d(xi,eta,zeta) at (u,v,w): FEI3dHexaLin::evaldNdxi (fei3dhexalin.C:129-166)
__device__ __forceinline__ void hexa_dNdxi(int k, double u, double v, double w, double dN[3])
{
    double sx, sy, sz;
    hexa_signs(k, sx, sy, sz);
    double fu = 1.0 + sx * u, fv = 1.0 + sy * v, fw = 1.0 + sz * w;
    dN[0] = sx * 0.125 * fv * fw;
    dN[1] = sy * 0.125 * fu * fw;
    dN[2] = sz * 0.125 * fu * fv;
}

// FloatMatrix::beInverseOf 3x3 (src/core/floatmatrix.C:790-808) and giveDeterminant (1100-1104)
__device__ __forceinline__ double inv3(const double s[3][3], double a[3][3])
{
    double det = s[0][0] * s[1][1] * s[2][2] + s[0][1] * s[1][2] * s[2][0] +
                 s[0][2] * s[1][0] * s[2][1] - s[0][2] * s[1][1] * s[2][0] -
                 s[1][2] * s[2][1] * s[0][0] - s[2][2] * s[0][1] * s[1][0];
    double r = 1.0 / det;
    a[0][0] = ( s[1][1] * s[2][2] - s[1][2] * s[2][1] ) * r;
    a[1][0] = ( s[1][2] * s[2][0] - s[1][0] * s[2][2] ) * r;
    a[2][0] = ( s[1][0] * s[2][1] - s[1][1] * s[2][0] ) * r;
    a[0][1] = ( s[0][2] * s[2][1] - s[0][1] * s[2][2] ) * r;
    a[1][1] = ( s[0][0] * s[2][2] - s[0][2] * s[2][0] ) * r;
    a[2][1] = ( s[0][1] * s[2][0] - s[0][0] * s[2][1] ) * r;
    a[0][2] = ( s[0][1] * s[1][2] - s[0][2] * s[1][1] ) * r;
    a[1][2] = ( s[0][2] * s[1][0] - s[0][0] * s[1][2] ) * r;
    a[2][2] = ( s[0][0] * s[1][1] - s[0][1] * s[1][0] ) * r;
    return det;
}

// Jacobian of the hexahedron at Gauss point (u,v,w) from xyz[8][3] in shared memory:
// J = coords(3x8) * dNdxi(8x3) (FEI3dHexaLin::evaldNdx, fei3dhexalin.C:186-204); returns
// det J and the inverse.
__device__ __forceinline__ double hexa_jacobian(const double *xyz, double u, double v, double w, double Ji[3][3])
{
    double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for ( int k = 0; k < 8; k++ ) {
        double dN[3];
        hexa_dNdxi(k, u, v, w, dN);
#pragma unroll
        for ( int i = 0; i < 3; i++ )
#pragma unroll
            for ( int j = 0; j < 3; j++ ) J[i][j] += xyz[3 * k + i] * dN[j];
    }
    return inv3(J, Ji);
}

// ---- tetrahedron --------------------------------------------------------------------

// FEI3dTetLin::evaldNdx (src/core/fei3dtetlin.C:116-166): constant gradients g[4][3], returns detJ
__device__ __forceinline__ double tet_dNdx(const double c[12], double g[4][3])
{
    double x1 = c[0], y1 = c[1], z1 = c[2], x2 = c[3], y2 = c[4], z2 = c[5];
    double x3 = c[6], y3 = c[7], z3 = c[8], x4 = c[9], y4 = c[10], z4 = c[11];
    double detJ = ( ( x4 - x1 ) * ( y2 - y1 ) * ( z3 - z1 ) - ( x4 - x1 ) * ( y3 - y1 ) * ( z2 - z1 ) +
                    ( x3 - x1 ) * ( y4 - y1 ) * ( z2 - z1 ) - ( x2 - x1 ) * ( y4 - y1 ) * ( z3 - z1 ) +
                    ( x2 - x1 ) * ( y3 - y1 ) * ( z4 - z1 ) - ( x3 - x1 ) * ( y2 - y1 ) * ( z4 - z1 ) );
    g[0][0] = -( ( y3 - y2 ) * ( z4 - z2 ) - ( y4 - y2 ) * ( z3 - z2 ) );
    g[1][0] = ( y4 - y3 ) * ( z1 - z3 ) - ( y1 - y3 ) * ( z4 - z3 );
    g[2][0] = -( ( y1 - y4 ) * ( z2 - z4 ) - ( y2 - y4 ) * ( z1 - z4 ) );
    g[3][0] = ( y2 - y1 ) * ( z3 - z1 ) - ( y3 - y1 ) * ( z2 - z1 );
    g[0][1] = -( ( x4 - x2 ) * ( z3 - z2 ) - ( x3 - x2 ) * ( z4 - z2 ) );
    g[1][1] = ( x1 - x3 ) * ( z4 - z3 ) - ( x4 - x3 ) * ( z1 - z3 );
    g[2][1] = -( ( x2 - x4 ) * ( z1 - z4 ) - ( x1 - x4 ) * ( z2 - z4 ) );
    g[3][1] = ( x3 - x1 ) * ( z2 - z1 ) - ( x2 - x1 ) * ( z3 - z1 );
    g[0][2] = -( ( x3 - x2 ) * ( y4 - y2 ) - ( x4 - x2 ) * ( y3 - y2 ) );
    g[1][2] = ( x4 - x3 ) * ( y1 - y3 ) - ( x1 - x3 ) * ( y4 - y3 );
    g[2][2] = -( ( x1 - x4 ) * ( y2 - y4 ) - ( x2 - x4 ) * ( y1 - y4 ) );
    g[3][2] = ( x2 - x1 ) * ( y3 - y1 ) - ( x3 - x1 ) * ( y2 - y1 );
    double f = 1.0 / detJ;
#pragma unroll
    for ( int i = 0; i < 4; i++ )
#pragma unroll
        for ( int j = 0; j < 3; j++ ) g[i][j] *= f;
    return detJ;
}

// ---- materials ----------------------------------------------------------------------

// IsotropicLinearElasticMaterial::initTangents (isolinearelasticmaterial.C:80-84):
// D = 2G I_dev + K I(x)I  ==  lambda on the 3x3 normal block, +2mu on its diagonal, mu on shear.
__device__ __forceinline__ void isole_lame(double E, double nu, double &lam, double &mu)
{
    double G = E / ( 2.0 * ( 1.0 + nu ) );
    double K = E / ( 3.0 * ( 1.0 - 2.0 * nu ) );
    mu = G;
    lam = K - 2.0 * G * ( 1.0 / 3.0 );      // K*1 + 2G*(-1/3)
}

__device__ __forceinline__ void isole_D(double E, double nu, double D[36])
{
    double G = E / ( 2.0 * ( 1.0 + nu ) );
    double K = E / ( 3.0 * ( 1.0 - 2.0 * nu ) );
#pragma unroll
    for ( int i = 0; i < 36; i++ ) D[i] = 0.0;
#pragma unroll
    for ( int i = 0; i < 3; i++ ) {
#pragma unroll
        for ( int j = 0; j < 3; j++ ) D[6 * i + j] = 2.0 * G * ( i == j ? 2.0 / 3.0 : -1.0 / 3.0 ) + K;
        D[6 * ( i + 3 ) + i + 3] = 2.0 * G * 0.5;
    }
}

__device__ __forceinline__ double dev_norm(const double t[6])   // StructuralMaterial::computeStressNorm
{
    return sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + 2.0 * t[3] * t[3] + 2.0 * t[4] * t[4] + 2.0 * t[5] * t[5]);
}

// MisesMat::giveRealStressVector_3d + performPlasticityReturn (misesmat.C:161-176, 181-255), hType 0.
__device__ __forceinline__ void mises_stress(const MatParams &mp, const double strain[6], const MisesStateRef &st, double stress[6])
{
    double G = mp.E / ( 2.0 * ( 1.0 + mp.nu ) );
    double K = mp.E / ( 3.0 * ( 1.0 - 2.0 * mp.nu ) );
    double pl[6], dev[6], tdev[6];
    double kappa = st.kappa();
#pragma unroll
    for ( int i = 0; i < 6; i++ ) {
        pl[i] = st.plStrain(i);
        dev[i] = strain[i] - pl[i];
    }
    double mean = ( dev[0] + dev[1] + dev[2] ) / 3.0;
    dev[0] -= mean; dev[1] -= mean; dev[2] -= mean;
    tdev[0] = 2.0 * G * dev[0]; tdev[1] = 2.0 * G * dev[1]; tdev[2] = 2.0 * G * dev[2];
    tdev[3] = G * dev[3]; tdev[4] = G * dev[4]; tdev[5] = G * dev[5];
    double trialVol = 3.0 * K * mean;
#pragma unroll
    for ( int i = 0; i < 6; i++ ) st.trialStressDev(i) = tdev[i];
    st.trialStressVol() = trialVol;
    double trialS = dev_norm(tdev);
    double yieldValue = sqrt(3.0 / 2.0) * trialS - ( mp.sig0 + mp.H * kappa );
    if ( yieldValue > 0.0 ) {
        double dKappa = yieldValue / ( mp.H + 3.0 * G );
        kappa += dKappa;
        double f = sqrt(3.0 / 2.0) * dKappa / trialS;
        pl[0] += f * tdev[0]; pl[1] += f * tdev[1]; pl[2] += f * tdev[2];
        pl[3] += f * 2.0 * tdev[3]; pl[4] += f * 2.0 * tdev[4]; pl[5] += f * 2.0 * tdev[5];
        double sc = 1.0 - sqrt(6.0) * G * dKappa / trialS;
#pragma unroll
        for ( int i = 0; i < 6; i++ ) tdev[i] *= sc;
    }
    tdev[0] += trialVol; tdev[1] += trialVol; tdev[2] += trialVol;
    double dam = kappa > 0.0 ? mp.omega_crit * ( 1.0 - exp(-mp.a * kappa) ) : 0.0;   // computeDamageParam (449-456)
    if ( st.damage() > dam ) dam = st.damage();                                        // computeDamage (470-481)
#pragma unroll
    for ( int i = 0; i < 6; i++ ) {
        st.effStress(i) = tdev[i];
        st.tempPlStrain(i) = pl[i];
        stress[i] = tdev[i] * ( 1.0 - dam );
    }
    st.tempKappa() = kappa;
    st.tempDamage() = dam;
}

// MisesMat::give3dMaterialStiffnessMatrix, TangentStiffness (misesmat.C:493-545)
__device__ __forceinline__ void mises_tangent(const MatParams &mp, const MisesStateRef &st, double D[36])
{
    double G = mp.E / ( 2.0 * ( 1.0 + mp.nu ) );
    isole_D(mp.E, mp.nu, D);
    double kappa = st.kappa(), tempKappa = st.tempKappa();
    double dKappa = tempKappa - kappa;
    if ( dKappa <= 0.0 ) return;
    double sigmaY = mp.sig0 + mp.H * kappa;
    double t[6], es[6];
#pragma unroll
    for ( int i = 0; i < 6; i++ ) {
        t[i] = st.trialStressDev(i);
        es[i] = st.effStress(i);
    }
    double trialS = dev_norm(t);
    double factor = -2.0 * sqrt(6.0) * G * G / trialS;
    double factor1 = factor * sigmaY / ( ( mp.H + 3.0 * G ) * trialS * trialS );
    double factor2 = factor * dKappa;
    double omega = st.tempDamage();
    double omegaPrime = tempKappa >= 0.0 ? mp.omega_crit * mp.a * exp(-mp.a * tempKappa) : 0.0;
    double scalar = -omegaPrime * sqrt(6.0) * G / ( 3.0 * G + mp.H ) / trialS;
#pragma unroll
    for ( int i = 0; i < 6; i++ )
#pragma unroll
        for ( int j = 0; j < 6; j++ ) {
            double idev = ( i < 3 && j < 3 ) ? ( i == j ? 2.0 / 3.0 : -1.0 / 3.0 ) : ( i == j ? 0.5 : 0.0 );
            double d = D[6 * i + j] + factor1 * ( t[i] * t[j] ) + factor2 * idev;
            D[6 * i + j] = d * ( 1.0 - omega ) + scalar * ( es[i] * t[j] );
        }
}

// The same tangent for kernels that spread one Gauss point over several lanes: rows 2 sub, 2 sub + 1 (sub < 3).  The state
// is read first (mises_tangent_load, early in the kernel) and used later.
struct MisesTangentIn {
    double t[6], ti[2], esi[2], kappa, tempKappa, tempDamage;
};
__device__ __forceinline__ void mises_tangent_load(const MisesStateRef &st, int sub, MisesTangentIn &in)
{
#pragma unroll
    for ( int j = 0; j < 6; j++ ) in.t[j] = st.trialStressDev(j);
#pragma unroll
    for ( int n = 0; n < 2; n++ ) {
        in.ti[n] = st.trialStressDev(2 * sub + n);
        in.esi[n] = st.effStress(2 * sub + n);
    }
    in.kappa = st.kappa();
    in.tempKappa = st.tempKappa();
    in.tempDamage = st.tempDamage();
}
// G, K: shear and bulk modulus, rh = 1 / (H + 3G) (per material, computed once by the caller).  One division (1 / trialS)
// instead of the five of mises_tangent: the results agree to round-off.
__device__ __forceinline__ void mises_tangent_rows(const MatParams &mp, double G, double K, double rh, const MisesTangentIn &in, int sub, double *D)
{
    const double dKappa = in.tempKappa - in.kappa;
    const bool plastic = dKappa > 0.0;
    double factor1 = 0.0, factor2 = 0.0, omega = 0.0, scalar = 0.0;
    if ( plastic ) {
        const double sigmaY = mp.sig0 + mp.H * in.kappa;
        const double r = 1.0 / dev_norm(in.t);
        const double factor = -2.0 * sqrt(6.0) * G * G * r;
        factor1 = factor * sigmaY * rh * r * r;
        factor2 = factor * dKappa;
        omega = in.tempDamage;
        const double omegaPrime = ( in.tempKappa >= 0.0 && mp.omega_crit != 0.0 ) ? mp.omega_crit * mp.a * exp(-mp.a * in.tempKappa) : 0.0;
        scalar = -omegaPrime * sqrt(6.0) * G * rh * r;
    }
#pragma unroll
    for ( int n = 0; n < 2; n++ ) {
        const int i = 2 * sub + n;
#pragma unroll
        for ( int j = 0; j < 6; j++ ) {
            const bool nn = i < 3 && j < 3;
            const double De = nn ? 2.0 * G * ( i == j ? 2.0 / 3.0 : -1.0 / 3.0 ) + K : ( i == j ? 2.0 * G * 0.5 : 0.0 );
            const double idev = nn ? ( i == j ? 2.0 / 3.0 : -1.0 / 3.0 ) : ( i == j ? 0.5 : 0.0 );
            double d = De;
            if ( plastic ) {
                d = De + factor1 * ( in.ti[n] * in.t[j] ) + factor2 * idev;
                d = d * ( 1.0 - omega ) + scalar * ( in.esi[n] * in.t[j] );
            }
            D[6 * i + j] = d;
        }
    }
}

// ---- 3x3 block algebra ----------------------------------------------------------------
// Contribution of one Gauss point to the (a,b) node block of Ke: w * Ba^T D Bb with the
// B-matrix layout of Structural3DElement::computeBmatrixAt (structural3delement.C:63-86).

// isotropic D:  K_ij += w ( lam ga_i gb_j + mu ga_j gb_i + delta_ij mu ga.gb )
__device__ __forceinline__ void block_iso(double acc[9], const double ga[3], const double gb[3], double wl, double wm)
{
    double la0 = wl * ga[0], la1 = wl * ga[1], la2 = wl * ga[2];
    double ma0 = wm * ga[0], ma1 = wm * ga[1], ma2 = wm * ga[2];
    double d = ma0 * gb[0] + ma1 * gb[1] + ma2 * gb[2];
    acc[0] += la0 * gb[0] + ma0 * gb[0] + d;
    acc[1] += la0 * gb[1] + ma1 * gb[0];
    acc[2] += la0 * gb[2] + ma2 * gb[0];
    acc[3] += la1 * gb[0] + ma0 * gb[1];
    acc[4] += la1 * gb[1] + ma1 * gb[1] + d;
    acc[5] += la1 * gb[2] + ma2 * gb[1];
    acc[6] += la2 * gb[0] + ma0 * gb[2];
    acc[7] += la2 * gb[1] + ma1 * gb[2];
    acc[8] += la2 * gb[2] + ma2 * gb[2] + d;
}

// general (possibly unsymmetric) D[36]:  T = D Bb (6x3),  K += w Ba^T T
__device__ __forceinline__ void block_general(double acc[9], const double ga[3], const double gb[3], const double *D, double w)
{
    double T[6][3];
#pragma unroll
    for ( int q = 0; q < 6; q++ ) {
        const double *d = D + 6 * q;
        T[q][0] = d[0] * gb[0] + d[4] * gb[2] + d[5] * gb[1];
        T[q][1] = d[1] * gb[1] + d[3] * gb[2] + d[5] * gb[0];
        T[q][2] = d[2] * gb[2] + d[3] * gb[1] + d[4] * gb[0];
    }
    double ax = w * ga[0], ay = w * ga[1], az = w * ga[2];
#pragma unroll
    for ( int j = 0; j < 3; j++ ) {
        acc[0 + j] += ax * T[0][j] + az * T[4][j] + ay * T[5][j];
        acc[3 + j] += ay * T[1][j] + az * T[3][j] + ax * T[5][j];
        acc[6 + j] += az * T[2][j] + ay * T[3][j] + ax * T[4][j];
    }
}

// strain contribution of node k: eps += B_k u_k
__device__ __forceinline__ void strain_add(double e[6], const double g[3], const double u[3])
{
    e[0] += g[0] * u[0];
    e[1] += g[1] * u[1];
    e[2] += g[2] * u[2];
    e[3] += g[2] * u[1] + g[1] * u[2];
    e[4] += g[2] * u[0] + g[0] * u[2];
    e[5] += g[1] * u[0] + g[0] * u[1];
}

// nodal force of node k: f_k = w B_k^T sigma
__device__ __forceinline__ void force_node(double f[3], const double g[3], const double s[6], double w)
{
    f[0] = w * ( g[0] * s[0] + g[2] * s[4] + g[1] * s[5] );
    f[1] = w * ( g[1] * s[1] + g[2] * s[3] + g[0] * s[5] );
    f[2] = w * ( g[2] * s[2] + g[1] * s[3] + g[0] * s[4] );
}

__device__ __forceinline__ void iso_stress(double lam, double mu, const double e[6], double s[6])
{
    double tr = e[0] + e[1] + e[2];
    s[0] = lam * tr + 2.0 * mu * e[0];
    s[1] = lam * tr + 2.0 * mu * e[1];
    s[2] = lam * tr + 2.0 * mu * e[2];
    s[3] = mu * e[3];
    s[4] = mu * e[4];
    s[5] = mu * e[5];
}

} // namespace ob200
