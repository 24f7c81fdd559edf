/*
 * oofem_oracle.c -- CPU restatement of the reference's structural hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (oofem_b200/) never does.  Parity is PINNED: tests/test_oracle_pinning.py checks this
 * file against (a) the #%BEGIN_CHECK% vectors embedded in the reference's own
 * tests/sm/patch300.in, patch301.in and (b) full-precision dumps produced by the
 * unmodified reference built by oracle/build_ref.py (fixtures in tests/golden/).
 *
 * Every function cites the reference file:line it follows; loops keep the reference's
 * summation order so that results agree to the last few ulps, not just to 1e-12.
 *
 * Conventions: matrices are row-major C arrays; equation numbers in location arrays
 * are 1-based with 0 = prescribed dof, exactly like OOFEM's IntArray loc.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define ORC_LSPACE 1
#define ORC_LTRSPACE 2

#define ORC_MAT_ISOLE 1
#define ORC_MAT_MISES 2

/* ------------------------------------------------------------------------------------ */
/* Integration rules                                                                    */
/* ------------------------------------------------------------------------------------ */

/* GaussIntegrationRule::SetUpPointsOnCube (src/core/gaussintegrationrule.C:190-214) with
 * nPoints=8 -> 2x2x2, loop order i (xi1) outer, j, k (xi3) inner; line rule from
 * giveLineCoordsAndWeights case 2 (gaussintegrationrule.C:1449-1452). */
static const double G2[2] = { -0.577350269189626, 0.577350269189626 };

static void lspace_gp(int g, double lc[3], double *w)
{
    int i = g >> 2, j = ( g >> 1 ) & 1, k = g & 1;
    lc[0] = G2[i]; lc[1] = G2[j]; lc[2] = G2[k];
    *w = 1.0 * 1.0 * 1.0;
}

/* ------------------------------------------------------------------------------------ */
/* FEI3dHexaLin                                                                         */
/* ------------------------------------------------------------------------------------ */

/* FEI3dHexaLin::evaldNdxi (src/core/fei3dhexalin.C:129-166): dN(node, dir). */
static void hexa_dNdxi(const double lc[3], double dN[8][3])
{
    double u = lc[0], v = lc[1], w = lc[2];
    dN[0][0] = -0.125 * ( 1. - v ) * ( 1. + w );
    dN[1][0] = -0.125 * ( 1. + v ) * ( 1. + w );
    dN[2][0] =  0.125 * ( 1. + v ) * ( 1. + w );
    dN[3][0] =  0.125 * ( 1. - v ) * ( 1. + w );
    dN[4][0] = -0.125 * ( 1. - v ) * ( 1. - w );
    dN[5][0] = -0.125 * ( 1. + v ) * ( 1. - w );
    dN[6][0] =  0.125 * ( 1. + v ) * ( 1. - w );
    dN[7][0] =  0.125 * ( 1. - v ) * ( 1. - w );

    dN[0][1] = -0.125 * ( 1. - u ) * ( 1. + w );
    dN[1][1] =  0.125 * ( 1. - u ) * ( 1. + w );
    dN[2][1] =  0.125 * ( 1. + u ) * ( 1. + w );
    dN[3][1] = -0.125 * ( 1. + u ) * ( 1. + w );
    dN[4][1] = -0.125 * ( 1. - u ) * ( 1. - w );
    dN[5][1] =  0.125 * ( 1. - u ) * ( 1. - w );
    dN[6][1] =  0.125 * ( 1. + u ) * ( 1. - w );
    dN[7][1] = -0.125 * ( 1. + u ) * ( 1. - w );

    dN[0][2] =  0.125 * ( 1. - u ) * ( 1. - v );
    dN[1][2] =  0.125 * ( 1. - u ) * ( 1. + v );
    dN[2][2] =  0.125 * ( 1. + u ) * ( 1. + v );
    dN[3][2] =  0.125 * ( 1. + u ) * ( 1. - v );
    dN[4][2] = -0.125 * ( 1. - u ) * ( 1. - v );
    dN[5][2] = -0.125 * ( 1. - u ) * ( 1. + v );
    dN[6][2] = -0.125 * ( 1. + u ) * ( 1. + v );
    dN[7][2] = -0.125 * ( 1. + u ) * ( 1. - v );
}

/* FloatMatrix::giveDeterminant, 3x3 branch (src/core/floatmatrix.C:1100-1104) */
static double det3(const double m[3][3])
{
    return m[0][0] * m[1][1] * m[2][2] + m[1][0] * m[2][1] * m[0][2] + m[2][0] * m[0][1] * m[1][2]
           - m[0][2] * m[1][1] * m[2][0] - m[1][2] * m[2][1] * m[0][0] - m[2][2] * m[0][1] * m[1][0];
}

/* FloatMatrix::beInverseOf, 3x3 branch (src/core/floatmatrix.C:790-808) */
static int inv3(const double s[3][3], double a[3][3])
{
    double det = s[0][0] * s[1][1] * s[2][2] + s[0][1] * s[1][2] * s[2][0] +
                 s[0][2] * s[1][0] * s[2][1] - s[0][2] * s[1][1] * s[2][0] -
                 s[1][2] * s[2][1] * s[0][0] - s[2][2] * s[0][1] * s[1][0];
    if ( !( fabs(det) > 1.e-30 ) ) return 0;
    a[0][0] = ( s[1][1] * s[2][2] - s[1][2] * s[2][1] ) / det;
    a[1][0] = ( s[1][2] * s[2][0] - s[1][0] * s[2][2] ) / det;
    a[2][0] = ( s[1][0] * s[2][1] - s[1][1] * s[2][0] ) / det;
    a[0][1] = ( s[0][2] * s[2][1] - s[0][1] * s[2][2] ) / det;
    a[1][1] = ( s[0][0] * s[2][2] - s[0][2] * s[2][0] ) / det;
    a[2][1] = ( s[0][1] * s[2][0] - s[0][0] * s[2][1] ) / det;
    a[0][2] = ( s[0][1] * s[1][2] - s[0][2] * s[1][1] ) / det;
    a[1][2] = ( s[0][2] * s[1][0] - s[0][0] * s[1][2] ) / det;
    a[2][2] = ( s[0][0] * s[1][1] - s[0][1] * s[1][0] ) / det;
    return 1;
}

/* FEI3dHexaLin::evaldNdx (src/core/fei3dhexalin.C:186-205): J = coords(3x8)*dNduvw(8x3),
 * dNdx = dNduvw * inv(J); returns det J.  Sums run over k ascending as in
 * FloatMatrix::beProductOf (src/core/floatmatrix.C:384-393). */
static double hexa_dNdx(const double *xyz /* [8][3] */, const double lc[3], double dNdx[8][3])
{
    double dN[8][3], J[3][3], Ji[3][3];
    hexa_dNdxi(lc, dN);
    for ( int i = 0; i < 3; i++ )
        for ( int j = 0; j < 3; j++ ) {
            double c = 0.;
            for ( int k = 0; k < 8; k++ ) c += xyz[3 * k + i] * dN[k][j];
            J[i][j] = c;
        }
    inv3(J, Ji);
    for ( int k = 0; k < 8; k++ )
        for ( int j = 0; j < 3; j++ ) {
            double c = 0.;
            for ( int m = 0; m < 3; m++ ) c += dN[k][m] * Ji[m][j];
            dNdx[k][j] = c;
        }
    return det3(J);
}

/* ------------------------------------------------------------------------------------ */
/* FEI3dTetLin                                                                          */
/* ------------------------------------------------------------------------------------ */

/* FEI3dTetLin::evaldNdx (src/core/fei3dtetlin.C:116-166); returns detJ (== 6 V). */
static double tet_dNdx(const double *c /* [4][3] */, double a[4][3])
{
    double x1 = c[0], y1 = c[1], z1 = c[2], x2 = c[3], y2 = c[4], z2 = c[5];
    double x3 = c[6], y3 = c[7], z3 = c[8], x4 = c[9], y4 = c[10], z4 = c[11];
    double detJ = ( ( x4 - x1 ) * ( y2 - y1 ) * ( z3 - z1 ) - ( x4 - x1 ) * ( y3 - y1 ) * ( z2 - z1 ) +
                    ( x3 - x1 ) * ( y4 - y1 ) * ( z2 - z1 ) - ( x2 - x1 ) * ( y4 - y1 ) * ( z3 - z1 ) +
                    ( x2 - x1 ) * ( y3 - y1 ) * ( z4 - z1 ) - ( x3 - x1 ) * ( y2 - y1 ) * ( z4 - z1 ) );
    a[0][0] = -( ( y3 - y2 ) * ( z4 - z2 ) - ( y4 - y2 ) * ( z3 - z2 ) );
    a[1][0] = ( y4 - y3 ) * ( z1 - z3 ) - ( y1 - y3 ) * ( z4 - z3 );
    a[2][0] = -( ( y1 - y4 ) * ( z2 - z4 ) - ( y2 - y4 ) * ( z1 - z4 ) );
    a[3][0] = ( y2 - y1 ) * ( z3 - z1 ) - ( y3 - y1 ) * ( z2 - z1 );

    a[0][1] = -( ( x4 - x2 ) * ( z3 - z2 ) - ( x3 - x2 ) * ( z4 - z2 ) );
    a[1][1] = ( x1 - x3 ) * ( z4 - z3 ) - ( x4 - x3 ) * ( z1 - z3 );
    a[2][1] = -( ( x2 - x4 ) * ( z1 - z4 ) - ( x1 - x4 ) * ( z2 - z4 ) );
    a[3][1] = ( x3 - x1 ) * ( z2 - z1 ) - ( x2 - x1 ) * ( z3 - z1 );

    a[0][2] = -( ( x3 - x2 ) * ( y4 - y2 ) - ( x4 - x2 ) * ( y3 - y2 ) );
    a[1][2] = ( x4 - x3 ) * ( y1 - y3 ) - ( x1 - x3 ) * ( y4 - y3 );
    a[2][2] = -( ( x1 - x4 ) * ( y2 - y4 ) - ( x2 - x4 ) * ( y1 - y4 ) );
    a[3][2] = ( x2 - x1 ) * ( y3 - y1 ) - ( x3 - x1 ) * ( y2 - y1 );
    double f = 1. / detJ;              /* answer.times(1. / detJ), fei3dtetlin.C:164 */
    for ( int i = 0; i < 4; i++ ) for ( int j = 0; j < 3; j++ ) a[i][j] *= f;
    return detJ;
}

/* ------------------------------------------------------------------------------------ */
/* B matrix and volume                                                                  */
/* ------------------------------------------------------------------------------------ */

/* Structural3DElement::computeBmatrixAt (src/sm/Elements/structural3delement.C:63-86),
 * LSpace::computeBmatrixAt without reduced shear integration (src/sm/Elements/3D/lspace.C:105-130).
 * Rows: eps_x, eps_y, eps_z, gamma_yz, gamma_zx, gamma_xy.  B is [6][3*nen]. */
static void bmatrix(int nen, const double ( *dNdx )[3], double *B)
{
    int nd = 3 * nen;
    memset( B, 0, sizeof( double ) * 6 * nd );
    for ( int i = 0; i < nen; i++ ) {
        B[0 * nd + 3 * i + 0] = dNdx[i][0];
        B[1 * nd + 3 * i + 1] = dNdx[i][1];
        B[2 * nd + 3 * i + 2] = dNdx[i][2];
        B[4 * nd + 3 * i + 0] = B[3 * nd + 3 * i + 1] = dNdx[i][2];
        B[5 * nd + 3 * i + 0] = B[3 * nd + 3 * i + 2] = dNdx[i][1];
        B[5 * nd + 3 * i + 1] = B[4 * nd + 3 * i + 2] = dNdx[i][0];
    }
}

/* Number of element nodes / Gauss points: LSpace ctor (lspace.C:62-70), LTRSpace ctor (ltrspace.C:61-70) */
int orc_nen(int etype) { return etype == ORC_LSPACE ? 8 : 4; }
int orc_ngp(int etype) { return etype == ORC_LSPACE ? 8 : 1; }

/* B and dV = |det J| * weight (Structural3DElement::computeVolumeAround, structural3delement.C:328-338)
 * at Gauss point g of an element with vertex coordinates xyz[nen][3]. */
static double elem_B_dV(int etype, const double *xyz, int g, double *B)
{
    if ( etype == ORC_LSPACE ) {
        double lc[3], w, dNdx[8][3];
        lspace_gp(g, lc, &w);
        double det = hexa_dNdx(xyz, lc, dNdx);   /* same J as giveTransformationJacobian (fei3dhexalin.C:614-624) */
        bmatrix(8, dNdx, B);
        return fabs(det) * w;
    } else {
        double dNdx[4][3];
        /* SetUpPointsOnTetrahedra / giveTetCoordsAndWeights case 1 (gaussintegrationrule.C:500-507): w = 1/6 */
        double det = tet_dNdx(xyz, dNdx);        /* == giveTransformationJacobian (fei3dtetlin.C:246-270) */
        bmatrix(4, dNdx, B);
        return fabs(det) * ( 1. / 6. );
    }
}

/* ------------------------------------------------------------------------------------ */
/* Materials                                                                            */
/* ------------------------------------------------------------------------------------ */

/* IsotropicLinearElasticMaterial::initTangents (src/sm/Materials/isolinearelasticmaterial.C:80-84):
 * tangent = 2 G I_dev6 + K I6_I6 with I_dev6 / I6_I6 from src/core/floatmatrixf.h:954-971.
 * NOTE the reference writes -1/3. and 2/3. in some slots as "-1/3." and in others as "-1/3"
 * (integer division == 0!): floatmatrixf.h:955-957 has `-1/3.` (a double) so all are -0.333..,
 * and `2/3.` likewise.  G = E/(2(1+nu)) (isolinearelasticmaterial.C:74), K = E/(3(1-2nu)). */
void orc_isole_D(double E, double nu, double *D /* [36] */)
{
    double G = E / ( 2.0 * ( 1. + nu ) );
    double K = E / ( 3.0 * ( 1. - 2. * nu ) );
    static const double Idev[36] = {
        2. / 3., -1. / 3., -1 / 3., 0., 0., 0.,
        -1. / 3., 2. / 3., -1 / 3., 0., 0., 0.,
        -1. / 3., -1. / 3., 2 / 3., 0., 0., 0.,
        0., 0., 0., 0.5, 0., 0.,
        0., 0., 0., 0., 0.5, 0.,
        0., 0., 0., 0., 0., 0.5,
    };
    for ( int i = 0; i < 36; i++ ) {
        int r = i / 6, c = i % 6;
        double ii = ( r < 3 && c < 3 ) ? 1. : 0.;
        D[i] = 2 * G * Idev[i] + K * ii;
    }
}

/* Per-Gauss-point MisesMat state, the fields of MisesMatStatus that the 3D path touches
 * (src/sm/Materials/misesmat.h).  "temp" values are what the reference keeps until
 * updateYourself() commits them. */
typedef struct {
    double plStrain[6];       /* committed plastic strain */
    double kappa;             /* committed cumulative plastic strain */
    double damage;            /* committed damage */
    double tempPlStrain[6];
    double tempKappa;
    double tempDamage;
    double trialStressDev[6];
    double trialStressVol;
    double effStress[6];      /* temp effective stress */
} orc_mises_state;            /* 29 doubles */

int orc_mises_state_doubles(void) { return (int)( sizeof( orc_mises_state ) / sizeof( double ) ); }

/* material parameter block: [type, E, nu, sig0, H, omega_crit, a] */
#define MP_STRIDE 8

/* MisesMat::giveRealStressVector_3d + performPlasticityReturn, 3D branch, hType == 0
 * (src/sm/Materials/misesmat.h/.C:161-176, 181-255) with the helpers of
 * src/sm/Materials/structuralmaterial.C:1488-1597.  Thermal strain is zero here
 * (giveStressDependentPartOfStrainVector subtracts nothing without temperature loads). */
static void mises_stress(const double *mp, const double strain[6], orc_mises_state *st, double stress[6])
{
    double E = mp[1], nu = mp[2], sig0 = mp[3], H = mp[4], omega_crit = mp[5], a = mp[6];
    double G = E / ( 2.0 * ( 1. + nu ) );          /* linearElasticMaterial.giveShearModulus() */
    double K = E / ( 3.0 * ( 1. - 2. * nu ) );     /* giveBulkModulus() */
    double plStrain[6], elStrain[6], dev[6], tdev[6];
    double kappa = st->kappa;
    for ( int i = 0; i < 6; i++ ) { plStrain[i] = st->plStrain[i]; elStrain[i] = strain[i] - plStrain[i]; }
    /* computeDeviatoricVolumetricSplit */
    double vol = elStrain[0] + elStrain[1] + elStrain[2];
    double mean = vol / 3.0;
    for ( int i = 0; i < 6; i++ ) dev[i] = elStrain[i];
    dev[0] -= mean; dev[1] -= mean; dev[2] -= mean;
    /* applyDeviatoricElasticStiffness */
    tdev[0] = 2. * G * dev[0]; tdev[1] = 2. * G * dev[1]; tdev[2] = 2. * G * dev[2];
    tdev[3] = G * dev[3]; tdev[4] = G * dev[4]; tdev[5] = G * dev[5];
    double trialStressVol = 3 * K * mean;
    for ( int i = 0; i < 6; i++ ) st->trialStressDev[i] = tdev[i];
    st->trialStressVol = trialStressVol;
    /* computeStressNorm */
    double trialS = sqrt(tdev[0] * tdev[0] + tdev[1] * tdev[1] + tdev[2] * tdev[2] +
                         2. * tdev[3] * tdev[3] + 2. * tdev[4] * tdev[4] + 2. * tdev[5] * tdev[5]);
    double yieldValue = sqrt(3. / 2.) * trialS - ( sig0 + H * kappa );   /* computeYieldStress, hType 0 */
    if ( yieldValue > 0. ) {
        double dKappa = yieldValue / ( H + 3. * G );
        kappa += dKappa;
        /* applyDeviatoricElasticCompliance(trialStressDev, 0.5) */
        double dPl[6] = { 1. / ( 2. * 0.5 ) * tdev[0], 1. / ( 2. * 0.5 ) * tdev[1], 1. / ( 2. * 0.5 ) * tdev[2],
                          1. / 0.5 * tdev[3], 1. / 0.5 * tdev[4], 1. / 0.5 * tdev[5] };
        double f = sqrt(3. / 2.) * dKappa / trialS;
        for ( int i = 0; i < 6; i++ ) plStrain[i] += f * dPl[i];
        double sc = 1. - sqrt(6.) * G * dKappa / trialS;
        for ( int i = 0; i < 6; i++ ) tdev[i] *= sc;
    }
    double full[6];
    for ( int i = 0; i < 6; i++ ) full[i] = tdev[i];
    full[0] += trialStressVol; full[1] += trialStressVol; full[2] += trialStressVol;
    for ( int i = 0; i < 6; i++ ) { st->effStress[i] = full[i]; st->tempPlStrain[i] = plStrain[i]; }
    st->tempKappa = kappa;
    /* computeDamage (misesmat.C:470-481), computeDamageParam (449-456) */
    double tempDam = kappa > 0. ? omega_crit * ( 1.0 - exp(-a * kappa) ) : 0.;
    if ( st->damage > tempDam ) tempDam = st->damage;
    st->tempDamage = tempDam;
    for ( int i = 0; i < 6; i++ ) stress[i] = full[i] * ( 1 - tempDam );
}

/* MisesMat::give3dMaterialStiffnessMatrix, TangentStiffness (src/sm/Materials/misesmat.C:493-545) */
static void mises_tangent(const double *mp, const orc_mises_state *st, double D[36])
{
    double E = mp[1], nu = mp[2], sig0 = mp[3], H = mp[4], omega_crit = mp[5], a = mp[6];
    double G = E / ( 2.0 * ( 1. + nu ) );
    orc_isole_D(E, nu, D);
    double kappa = st->kappa, tempKappa = st->tempKappa;
    double dKappa = tempKappa - kappa;
    if ( dKappa <= 0.0 ) return;
    double sigmaY = sig0 + H * kappa;
    const double *t = st->trialStressDev;
    double trialS = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + 2. * t[3] * t[3] + 2. * t[4] * t[4] + 2. * t[5] * t[5]);
    double factor = -2. * sqrt(6.) * G * G / trialS;
    double factor1 = factor * sigmaY / ( ( H + 3. * G ) * trialS * trialS );
    for ( int i = 0; i < 6; i++ ) for ( int j = 0; j < 6; j++ ) D[6 * i + j] += factor1 * ( t[i] * t[j] );
    double factor2 = factor * dKappa;
    static const double Idev[36] = {
        2. / 3., -1. / 3., -1 / 3., 0., 0., 0.,
        -1. / 3., 2. / 3., -1 / 3., 0., 0., 0.,
        -1. / 3., -1. / 3., 2 / 3., 0., 0., 0.,
        0., 0., 0., 0.5, 0., 0.,
        0., 0., 0., 0., 0.5, 0.,
        0., 0., 0., 0., 0., 0.5,
    };
    for ( int i = 0; i < 36; i++ ) D[i] += factor2 * Idev[i];
    double omega = st->tempDamage;
    for ( int i = 0; i < 36; i++ ) D[i] *= 1. - omega;
    double omegaPrime = tempKappa >= 0. ? omega_crit * a * exp(-a * tempKappa) : 0.;   /* computeDamageParamPrime */
    double scalar = -omegaPrime * sqrt(6.) * G / ( 3. * G + H ) / trialS;
    for ( int i = 0; i < 6; i++ ) for ( int j = 0; j < 6; j++ ) D[6 * i + j] += scalar * ( st->effStress[i] * t[j] );
}

/* MisesMatStatus::updateYourself: temp -> committed (src/sm/Materials/misesmat.C, status class) */
void orc_mises_commit(int64_t ngp, double *state)
{
    orc_mises_state *s = (orc_mises_state *) state;
    for ( int64_t g = 0; g < ngp; g++ ) {
        memcpy( s[g].plStrain, s[g].tempPlStrain, sizeof( double ) * 6 );
        s[g].kappa = s[g].tempKappa;
        s[g].damage = s[g].tempDamage;
    }
}

/* MisesMatStatus::initTempStatus equivalent for freshly allocated state: temp = committed */
void orc_mises_init(int64_t ngp, double *state)
{
    orc_mises_state *s = (orc_mises_state *) state;
    for ( int64_t g = 0; g < ngp; g++ ) {
        memset( &s[g], 0, sizeof( orc_mises_state ) );
    }
}

/* ------------------------------------------------------------------------------------ */
/* Element level                                                                        */
/* ------------------------------------------------------------------------------------ */

/* StructuralElement::computeStiffnessMatrix / NLStructuralElement::computeStiffnessMatrix
 * (src/sm/Elements/structuralelement.C:627-642, nlstructuralelement.C:345-372, 444-446):
 *   for gp: B, D, dV; DB = D*B; symmetric material -> plusProductSymmUpper + symmetrized()
 *                                   else            -> plusProductUnsym.
 * D is [ngp][36] (dstride=36) or one shared matrix (dstride=0). */
void orc_element_stiffness(int etype, const double *xyz, const double *D, int dstride, int symm, double *Ke)
{
    int nen = orc_nen(etype), nd = 3 * nen, ngp = orc_ngp(etype);
    double B[6 * 24], DB[6 * 24];
    memset( Ke, 0, sizeof( double ) * nd * nd );
    for ( int g = 0; g < ngp; g++ ) {
        const double *Dg = D + (size_t) g * dstride;
        double dV = elem_B_dV(etype, xyz, g, B);
        /* DB.beProductOf(D, B)  (floatmatrix.C:384-393) */
        for ( int i = 0; i < 6; i++ )
            for ( int j = 0; j < nd; j++ ) {
                double c = 0.;
                for ( int k = 0; k < 6; k++ ) c += Dg[6 * i + k] * B[k * nd + j];
                DB[i * nd + j] = c;
            }
        if ( symm ) {
            /* plusProductSymmUpper (floatmatrix.C:1289-1299) */
            for ( int i = 0; i < nd; i++ )
                for ( int j = i; j < nd; j++ ) {
                    double s = 0.;
                    for ( int k = 0; k < 6; k++ ) s += B[k * nd + i] * DB[k * nd + j];
                    Ke[i * nd + j] += s * dV;
                }
        } else {
            /* plusProductUnsym (floatmatrix.C:708-718) */
            for ( int i = 0; i < nd; i++ )
                for ( int j = 0; j < nd; j++ ) {
                    double s = 0.;
                    for ( int k = 0; k < 6; k++ ) s += B[k * nd + i] * DB[k * nd + j];
                    Ke[i * nd + j] += s * dV;
                }
        }
    }
    if ( symm ) {
        /* symmetrized (floatmatrix.C:1147-1151) */
        for ( int i = 1; i < nd; i++ ) for ( int j = 0; j < i; j++ ) Ke[i * nd + j] = Ke[j * nd + i];
    }
}

/* Batched element evaluation: Ke for every element of a homogeneous batch.
 *   conn  [nelem][nen] 1-based node numbers, coords [nnode][3]
 *   matid [nelem] 0-based index into matparams [nmat][MP_STRIDE]
 *   state [nelem*ngp][29] MisesMat state (may be NULL when no Mises material)
 *   Ke    [nelem][nd*nd] row-major */
void orc_batch_stiffness(int etype, int64_t nelem, const int32_t *conn, const double *coords,
                         const int32_t *matid, const double *matparams, const double *state, double *Ke)
{
    int nen = orc_nen(etype), nd = 3 * nen, ngp = orc_ngp(etype);
    const orc_mises_state *st = (const orc_mises_state *) state;
    for ( int64_t e = 0; e < nelem; e++ ) {
        double xyz[24], D[8 * 36];
        for ( int a = 0; a < nen; a++ ) {
            int64_t n = conn[e * nen + a] - 1;
            xyz[3 * a] = coords[3 * n]; xyz[3 * a + 1] = coords[3 * n + 1]; xyz[3 * a + 2] = coords[3 * n + 2];
        }
        const double *mp = matparams + (size_t) matid[e] * MP_STRIDE;
        int type = (int) mp[0];
        if ( type == ORC_MAT_ISOLE ) {
            orc_isole_D(mp[1], mp[2], D);
            orc_element_stiffness(etype, xyz, D, 0, 1, Ke + (size_t) e * nd * nd);
        } else {
            for ( int g = 0; g < ngp; g++ ) mises_tangent(mp, &st[e * ngp + g], D + 36 * g);
            /* MisesMat::isCharacteristicMtrxSymmetric returns false (misesmat.h:125) */
            orc_element_stiffness(etype, xyz, D, 36, 0, Ke + (size_t) e * nd * nd);
        }
    }
}

/* StructuralElement::giveInternalForcesVector / NLStructuralElement (nlGeometry 0)
 * (src/sm/Elements/structuralelement.C:748-795, nlstructuralelement.C:166-236):
 *   strain = B u; stress = material(strain); answer.plusProduct(B, stress, dV)
 * ue [nelem][nd] are the element displacement vectors (computeVectorOf(VM_Total)).
 * Also returns per-GP strain and stress ([nelem*ngp][6]) when the pointers are non-NULL. */
void orc_batch_internal_forces(int etype, int64_t nelem, const int32_t *conn, const double *coords,
                               const int32_t *matid, const double *matparams, double *state,
                               const double *ue, double *fe, double *gp_strain, double *gp_stress)
{
    int nen = orc_nen(etype), nd = 3 * nen, ngp = orc_ngp(etype);
    orc_mises_state *st = (orc_mises_state *) state;
    for ( int64_t e = 0; e < nelem; e++ ) {
        double xyz[24], B[6 * 24], D[36];
        for ( int a = 0; a < nen; a++ ) {
            int64_t n = conn[e * nen + a] - 1;
            xyz[3 * a] = coords[3 * n]; xyz[3 * a + 1] = coords[3 * n + 1]; xyz[3 * a + 2] = coords[3 * n + 2];
        }
        const double *mp = matparams + (size_t) matid[e] * MP_STRIDE;
        int type = (int) mp[0];
        const double *u = ue + (size_t) e * nd;
        double *f = fe + (size_t) e * nd;
        memset( f, 0, sizeof( double ) * nd );
        if ( type == ORC_MAT_ISOLE ) orc_isole_D(mp[1], mp[2], D);
        for ( int g = 0; g < ngp; g++ ) {
            double strain[6], stress[6];
            double dV = elem_B_dV(etype, xyz, g, B);
            for ( int i = 0; i < 6; i++ ) {        /* strain.beProductOf(b, u) */
                double c = 0.;
                for ( int k = 0; k < nd; k++ ) c += B[i * nd + k] * u[k];
                strain[i] = c;
            }
            if ( type == ORC_MAT_ISOLE ) {
                /* LinearElasticMaterial::giveRealStressVector_3d (linearelasticmaterial.C:102-124): stress = dot(d, strain) */
                for ( int i = 0; i < 6; i++ ) {
                    double c = 0.;
                    for ( int k = 0; k < 6; k++ ) c += D[6 * i + k] * strain[k];
                    stress[i] = c;
                }
            } else {
                mises_stress(mp, strain, &st[e * ngp + g], stress);
            }
            if ( gp_strain ) memcpy( gp_strain + ( (size_t) e * ngp + g ) * 6, strain, sizeof( strain ) );
            if ( gp_stress ) memcpy( gp_stress + ( (size_t) e * ngp + g ) * 6, stress, sizeof( stress ) );
            /* FloatArray::plusProduct (src/core/floatarray.C:309-316): a_i += (sum_j b(j,i) s_j) * dV */
            for ( int i = 0; i < nd; i++ ) {
                double s = 0.;
                for ( int j = 0; j < 6; j++ ) s += B[j * nd + i] * stress[j];
                f[i] += s * dV;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* CompCol                                                                              */
/* ------------------------------------------------------------------------------------ */

static int cmp_i32(const void *a, const void *b)
{
    int32_t x = *(const int32_t *) a, y = *(const int32_t *) b;
    return ( x > y ) - ( x < y );
}

/* CompCol::buildInternalStructure (src/core/compcol.C:167-260): columns[jj-1].insert(ii-1)
 * for every pair of non-zero entries of every element location array; colptr / rowind
 * are the sorted, de-duplicated per-column row sets.  Implemented with the same result
 * by counting, bucketing and sort+unique per column.
 *   loc [nelem][nd] 1-based, 0 = prescribed.  Returns nnz; *rowind_out malloc'ed. */
int64_t orc_compcol_build(int64_t nelem, int nd, const int32_t *loc, int32_t neq,
                          int32_t *colptr /* [neq+1] */, int32_t **rowind_out)
{
    int64_t *cnt = (int64_t *) calloc( (size_t) neq + 1, sizeof( int64_t ) );
    for ( int64_t e = 0; e < nelem; e++ ) {
        const int32_t *l = loc + e * nd;
        int nz = 0;
        for ( int i = 0; i < nd; i++ ) nz += l[i] != 0;
        for ( int j = 0; j < nd; j++ ) if ( l[j] ) cnt[l[j]] += nz;
    }
    for ( int32_t j = 0; j < neq; j++ ) cnt[j + 1] += cnt[j];       /* cnt[j] = start of column j */
    int64_t total = cnt[neq];
    int32_t *buf = (int32_t *) malloc( sizeof( int32_t ) * ( total ? total : 1 ) );
    int64_t *fill = (int64_t *) malloc( sizeof( int64_t ) * ( (size_t) neq + 1 ) );
    memcpy( fill, cnt, sizeof( int64_t ) * ( (size_t) neq + 1 ) );
    for ( int64_t e = 0; e < nelem; e++ ) {
        const int32_t *l = loc + e * nd;
        for ( int j = 0; j < nd; j++ ) {
            if ( !l[j] ) continue;
            int64_t *p = &fill[l[j] - 1];
            for ( int i = 0; i < nd; i++ ) if ( l[i] ) buf[( *p )++] = l[i] - 1;
        }
    }
    int64_t nnz = 0;
    int32_t *rowind = (int32_t *) malloc( sizeof( int32_t ) * ( total ? total : 1 ) );
    for ( int32_t j = 0; j < neq; j++ ) {
        int64_t s = cnt[j], n = cnt[j + 1] - cnt[j];
        qsort( buf + s, (size_t) n, sizeof( int32_t ), cmp_i32 );
        colptr[j] = (int32_t) nnz;
        for ( int64_t t = 0; t < n; t++ )
            if ( t == 0 || buf[s + t] != buf[s + t - 1] ) rowind[nnz++] = buf[s + t];
    }
    colptr[neq] = (int32_t) nnz;
    free(buf); free(fill); free(cnt);
    *rowind_out = (int32_t *) realloc( rowind, sizeof( int32_t ) * ( nnz ? nnz : 1 ) );
    return nnz;
}

void orc_free(void *p) { free(p); }

/* CompCol::assemble(loc, mat) (src/core/compcol.C:263-299) for a batch of element matrices,
 * in element order like EngngModel::assemble (src/core/engngm.C:904-929). */
void orc_compcol_assemble(int64_t nelem, int nd, const int32_t *loc, const double *Ke,
                          const int32_t *colptr, const int32_t *rowind, double *val)
{
    for ( int64_t e = 0; e < nelem; e++ ) {
        const int32_t *l = loc + e * nd;
        const double *m = Ke + (size_t) e * nd * nd;
        for ( int j = 0; j < nd; j++ ) {
            int jj = l[j];
            if ( !jj ) continue;
            int cstart = colptr[jj - 1];
            for ( int i = 0; i < nd; i++ ) {
                int ii = l[i];
                if ( !ii ) continue;
                int t = cstart;
                while ( rowind[t] < ii - 1 ) t++;
                val[t] += m[i * nd + j];
            }
        }
    }
}

/* Assemble element vectors into a global vector (EngngModel::assembleVector -> FloatArray::assemble):
 * answer[loc_i - 1] += fe_i for loc_i != 0. */
void orc_assemble_vector(int64_t nelem, int nd, const int32_t *loc, const double *fe, double *answer)
{
    for ( int64_t e = 0; e < nelem; e++ )
        for ( int i = 0; i < nd; i++ ) {
            int ii = loc[e * nd + i];
            if ( ii ) answer[ii - 1] += fe[(size_t) e * nd + i];
        }
}

/* CompCol::times (src/core/compcol.C:119-134) */
void orc_compcol_times(int32_t n, const int32_t *colptr, const int32_t *rowind, const double *val,
                       const double *x, double *y)
{
    for ( int32_t i = 0; i < n; i++ ) y[i] = 0.;
    for ( int32_t j = 0; j < n; j++ ) {
        double rhs = x[j];
        for ( int32_t t = colptr[j]; t < colptr[j + 1]; t++ ) y[rowind[t]] += val[t] * rhs;
    }
}

/* ------------------------------------------------------------------------------------ */
/* IML++ CG                                                                             */
/* ------------------------------------------------------------------------------------ */

static double vdot(int32_t n, const double *a, const double *b)
{
    double s = 0.;
    for ( int32_t i = 0; i < n; i++ ) s += a[i] * b[i];
    return s;
}

/* CG template (iml/cg.h:23-72) as instantiated by IMLSolver::solve (src/core/iml/imlsolver.C:122-123)
 * with DiagPreconditioner (src/core/iml/diagpre.C:41-68; precond = 1) or VoidPreconditioner (precond = 0).
 * In: x initial guess, max_iter, tol.  Out: x, *iters, *resid.  Returns 0 converged / 1 not. */
int orc_cg(int32_t n, const int32_t *colptr, const int32_t *rowind, const double *val,
           const double *b, double *x, int precond, int max_iter, double tol, int *iters, double *resid_out)
{
    double *r = (double *) malloc( sizeof( double ) * 5 * ( n ? n : 1 ) );
    double *p = r + n, *z = p + n, *q = z + n, *diag = q + n;
    double rho = 0., rho_1 = 0., resid;
    if ( precond == 1 )
        for ( int32_t i = 0; i < n; i++ ) {           /* DiagPreconditioner::init: diag = 1/A(i,i) */
            double d = 0.;
            for ( int32_t t = colptr[i]; t < colptr[i + 1]; t++ ) if ( rowind[t] == i ) { d = val[t]; break; }
            diag[i] = 1. / d;
        }
    double normb = sqrt(vdot(n, b, b));
    orc_compcol_times(n, colptr, rowind, val, x, q);
    for ( int32_t i = 0; i < n; i++ ) r[i] = b[i] - q[i];
    if ( normb == 0.0 ) normb = 1;
    if ( ( resid = sqrt(vdot(n, r, r)) / normb ) <= tol ) {
        *resid_out = resid; *iters = 0; free(r); return 0;
    }
    for ( int i = 1; i <= max_iter; i++ ) {
        if ( precond == 1 ) for ( int32_t k = 0; k < n; k++ ) z[k] = r[k] * diag[k];
        else memcpy( z, r, sizeof( double ) * n );
        rho = vdot(n, r, z);
        if ( i == 1 ) memcpy( p, z, sizeof( double ) * n );
        else {
            double beta = rho / rho_1;
            for ( int32_t k = 0; k < n; k++ ) p[k] = z[k] + beta * p[k];
        }
        orc_compcol_times(n, colptr, rowind, val, p, q);
        double alpha = rho / vdot(n, p, q);
        for ( int32_t k = 0; k < n; k++ ) x[k] += alpha * p[k];
        for ( int32_t k = 0; k < n; k++ ) r[k] -= alpha * q[k];
        if ( ( resid = sqrt(vdot(n, r, r)) / normb ) <= tol ) {
            *resid_out = resid; *iters = i; free(r); return 0;
        }
        rho_1 = rho;
    }
    *resid_out = resid; *iters = max_iter; free(r);
    return 1;
}
