#include "cudacsr.h"
#include "cudacontext.h"

#include "activebc.h"
#include "assemblercallback.h"
#include "classfactory.h"
#include "dofmanager.h"
#include "domain.h"
#include "element.h"
#include "engngm.h"
#include "error.h"
#include "floatarray.h"
#include "floatmatrix.h"
#include "crosssection.h"
#include "gausspoint.h"
#include "integrationrule.h"
#include "material.h"
#include "timestep.h"
#include "unknownnumberingscheme.h"
#include "sm/Elements/3D/lspace.h"
#include "sm/Elements/3D/ltrspace.h"
#include "sm/Materials/isolinearelasticmaterial.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <typeinfo>

namespace oofem {
REGISTER_SparseMtrx(CudaCSR, SMT_CudaCSR);

// contributions staged on the host before one batched ob200_csr_assemble call
static const int64_t kPendingDoubles = 1 << 22;        // 32 MB of matrices

CudaCSR :: CudaCSR(int n) : SparseMtrx(n, n)
{
    CudaContext :: check(ob200_csr_create(CudaContext :: get(), & A), "CudaCSR");
}

CudaCSR :: ~CudaCSR()
{
    this->dropElementSet();
    ob200_csr_destroy(A);
}

void CudaCSR :: dropElementSet()
{
    if ( set ) {
        ob200_elemset_destroy(set);
        set = nullptr;
    }
    setDomain = nullptr;
    setTried = false;
}

int CudaCSR :: buildInternalStructure(EngngModel *eModel, int di, const UnknownNumberingScheme &s)
{
    // the pattern CompCol builds (compcol.C:174-267): union over elements of loc x loc, plus the
    // location arrays of the active boundary conditions
    Domain *domain = eModel->giveDomain(di);
    int neq = eModel->giveNumberOfDomainEquations(di, s);
    std :: vector< IntArray >locs;
    locs.reserve(domain->giveNumberOfElements() + 8);
    IntArray loc;
    int width = 0;
    for ( auto &elem : domain->giveElements() ) {
        elem->giveLocationArray(loc, s);
        width = std :: max(width, loc.giveSize());
        locs.push_back(loc);
    }
    std :: vector< IntArray >r_locs, c_locs;
    for ( auto &gbc : domain->giveBcs() ) {
        ActiveBoundaryCondition *bc = dynamic_cast< ActiveBoundaryCondition * >( gbc.get() );
        if ( bc ) {
            bc->giveLocationArrays(r_locs, c_locs, UnknownCharType, s, s);
            for ( std :: size_t k = 0; k < r_locs.size(); k++ ) {
                // the library takes square blocks loc x loc: rows and columns of an active bc couple
                // symmetrically in every bc of the reference (Lagrange multipliers), so their union is exact
                // whenever r x c and c x r are both present; otherwise it is a superset with explicit zeros
                IntArray u(r_locs [ k ]);
                u.followedBy(c_locs [ k ]);
                width = std :: max(width, u.giveSize());
                locs.push_back(u);
            }
        }
    }
    std :: vector< int32_t >flat(locs.size() * ( size_t ) std :: max(width, 1), 0);
    for ( std :: size_t e = 0; e < locs.size(); e++ ) {
        for ( int k = 0; k < locs [ e ].giveSize(); k++ ) {
            flat [ e * width + k ] = locs [ e ] [ k ];
        }
    }
    pendLoc.clear();
    pendMat.clear();
    pendCount = 0;
    cells.clear();
    this->dropElementSet();
    CudaContext :: check(ob200_csr_build_structure(A, neq, ( int64_t ) locs.size(), std :: max(width, 1), flat.data(), 0),
                         "CudaCSR::buildInternalStructure");
    nRows = nColumns = neq;
    this->version++;
    OOFEM_LOG_DEBUG("CudaCSR info: neq is %d, nwk is %ld\n", neq, ( long ) ob200_csr_nnz(A));
    return true;
}

void CudaCSR :: flush() const
{
    if ( pendCount ) {
        CudaContext :: check(ob200_csr_assemble(A, pendCount, pendDofs, pendLoc.data(), pendMat.data(), 0), "CudaCSR::assemble");
        pendLoc.clear();
        pendMat.clear();
        pendCount = 0;
    }
    for ( auto &c : cells ) {
        double d = c.second.cur - c.second.seen;
        if ( d != 0.0 ) {
            int32_t l [ 2 ] = { c.first.first, c.first.second };
            CudaContext :: check(ob200_csr_assemble_rect(A, 1, 1, l, l + 1, & d, 0), "CudaCSR::at");
        }
    }
    cells.clear();
}

int CudaCSR :: assemble(const IntArray &loc, const FloatMatrix &mat)
{
    int n = loc.giveSize();
    if ( n != mat.giveNumberOfRows() || n != mat.giveNumberOfColumns() ) {
        OOFEM_ERROR("dimension of 'k' and 'loc' mismatch");                   // compcol.C:268-270
    }
    if ( n == 0 ) {
        return 1;
    }
    if ( pendCount && ( pendDofs != n || ( int64_t ) pendMat.size() + ( int64_t ) n * n > kPendingDoubles ) ) {
        this->flush();
    }
    if ( !cells.empty() ) {
        this->flush();
    }
    pendDofs = n;
    pendLoc.insert( pendLoc.end(), loc.begin(), loc.end() );
    size_t base = pendMat.size();
    pendMat.resize(base + ( size_t ) n * n);
    for ( int i = 0; i < n; i++ ) {         // FloatMatrix is column-major, the C ABI row-major
        for ( int j = 0; j < n; j++ ) {
            pendMat [ base + ( size_t ) i * n + j ] = mat(i, j);
        }
    }
    pendCount++;
    this->version++;
    return 1;
}

int CudaCSR :: assemble(const IntArray &rloc, const IntArray &cloc, const FloatMatrix &mat)
{
    // rectangular contribution: only rloc x cloc is touched, like CompCol::assemble(rloc, cloc, mat) (compcol.C:301-336)
    int nr = rloc.giveSize(), nc = cloc.giveSize();
    if ( nr != mat.giveNumberOfRows() || nc != mat.giveNumberOfColumns() ) {
        OOFEM_ERROR("dimension of 'k' and 'loc' mismatch");
    }
    if ( nr == 0 || nc == 0 ) {
        return 1;
    }
    this->flush();
    std :: vector< double >m( ( size_t ) nr * nc );
    for ( int i = 0; i < nr; i++ ) {         // FloatMatrix is column-major, the C ABI row-major
        for ( int j = 0; j < nc; j++ ) {
            m [ ( size_t ) i * nc + j ] = mat(i, j);
        }
    }
    CudaContext :: check(ob200_csr_assemble_rect(A, nr, nc, rloc.givePointer(), cloc.givePointer(), m.data(), 0), "CudaCSR::assemble");
    this->version++;
    return 1;
}

void CudaCSR :: times(const FloatArray &x, FloatArray &answer) const
{
    if ( x.giveSize() != nColumns ) {
        OOFEM_ERROR("incompatible dimensions");                               // compcol.C:121-123
    }
    this->flush();
    answer.resize(nRows);
    if ( nRows ) {
        CudaContext :: check(ob200_csr_times(A, x.givePointer(), answer.givePointer(), 0), "CudaCSR::times");
    }
}

void CudaCSR :: timesT(const FloatArray &x, FloatArray &answer) const
{
    if ( x.giveSize() != nRows ) {
        OOFEM_ERROR("Error in CompCol -- incompatible dimensions");           // compcol.C:148-150
    }
    this->flush();
    answer.resize(nColumns);
    if ( nRows ) {
        CudaContext :: check(ob200_csr_times_t(A, x.givePointer(), answer.givePointer(), 0), "CudaCSR::timesT");
    }
}

void CudaCSR :: times(double x)
{
    this->flush();
    CudaContext :: check(ob200_csr_scale(A, x), "CudaCSR::times");
    this->version++;
}

void CudaCSR :: zero()
{
    pendLoc.clear();
    pendMat.clear();
    pendCount = 0;
    cells.clear();
    CudaContext :: check(ob200_csr_zero(A), "CudaCSR::zero");
    this->version++;
}

double CudaCSR :: at(int i, int j) const
{
    this->flush();
    double v = 0.;
    int rc = ob200_csr_at(A, i, j, & v);
    if ( rc == OB200_ESTRUCT ) {
        return 0.;                                                            // compcol.C:390-400: not stored = zero
    }
    CudaContext :: check(rc, "CudaCSR::at");
    return v;
}

double &CudaCSR :: at(int i, int j)
{
    // the caller may write through the reference: remember what it saw, write the difference back
    // at the next operation that reads the matrix
    auto key = std :: make_pair(i, j);
    auto it = cells.find(key);
    if ( it == cells.end() ) {
        if ( pendCount ) {
            this->flush();
        }
        double v = 0.;
        int rc = ob200_csr_at(A, i, j, & v);
        if ( rc != OB200_OK ) {               // out of bounds, or not in the sparse structure: CompCol::at raises for both
            OOFEM_ERROR("Array accessing exception -- (%d,%d) out of bounds", i, j);    // compcol.C:372
        }
        it = cells.insert({ key, Cell { v, v } }).first;
        this->version++;
    }
    return it->second.cur;
}

bool CudaCSR :: isAllocatedAt(int i, int j) const
{
    double v;
    return ob200_csr_at(A, i, j, & v) == OB200_OK;
}

void CudaCSR :: giveStructure(IntArray &rowptr, IntArray &colind) const
{
    rowptr.resize(nRows + 1);
    colind.resize( ( int ) ob200_csr_nnz(A) );
    CudaContext :: check(ob200_csr_get_structure(A, rowptr.givePointer(), colind.givePointer(), 0), "CudaCSR::giveStructure");
}

void CudaCSR :: toFloatMatrix(FloatMatrix &answer) const
{
    this->flush();
    IntArray rp, ci;
    this->giveStructure(rp, ci);
    std :: vector< double >v(std :: max< int64_t >(ob200_csr_nnz(A), 1));
    CudaContext :: check(ob200_csr_get_values(A, v.data(), 0), "CudaCSR::toFloatMatrix");
    answer.resize(nRows, nColumns);
    answer.zero();
    for ( int i = 0; i < nRows; i++ ) {
        for ( int k = rp [ i ]; k < rp [ i + 1 ]; k++ ) {
            answer(i, ci [ k ]) = v [ k ];
        }
    }
}

void CudaCSR :: printStatistics() const
{
    OOFEM_LOG_INFO("CudaCSR info: neq is %d, nwk is %ld, batched assembly %s\n", nRows, ( long ) ob200_csr_nnz(A),
                   set ? "on" : "off");
}

// ---- batched element-evaluation hook ------------------------------------------------------------

bool CudaCSR :: buildElementSet(EngngModel *eModel, const UnknownNumberingScheme &s, Domain *domain)
{
    // Accepts a domain made of plain LSpace or plain LTRSpace elements (small strain, default
    // integration rule) with IsotropicLinearElasticMaterial and nodes without local coordinate
    // systems.  Anything else keeps the reference's host loop.
    int nelem = domain->giveNumberOfElements();
    int nnode = domain->giveNumberOfDofManagers();
    if ( nelem == 0 || nnode == 0 ) {
        return false;
    }
    int etype = 0, nen = 0;
    const char *cn = domain->giveElement(1)->giveClassName();
    if ( !std :: strcmp(cn, "LSpace") ) {
        etype = OB200_LSPACE;
        nen = 8;
    } else if ( !std :: strcmp(cn, "LTRSpace") ) {
        etype = OB200_LTRSPACE;
        nen = 4;
    } else {
        return false;
    }
    const int ngp = etype == OB200_LSPACE ? 8 : 1;
    std :: vector< int32_t >conn( ( size_t ) nelem * nen), matid(nelem), loc( ( size_t ) nelem * nen * 3);
    std :: map< int, int >matIndex;                  // material number -> row of the parameter table
    std :: vector< double >matparams;
    IntArray l, ids;
    FloatMatrix R;
    for ( int e = 1; e <= nelem; e++ ) {
        Element *elem = domain->giveElement(e);
        if ( std :: strcmp(elem->giveClassName(), cn) ) {
            return false;
        }
        NLStructuralElement *se = dynamic_cast< NLStructuralElement * >( elem );
        if ( !se || se->giveGeometryMode() != 0 || elem->giveRotationMatrix(R) ) {
            return false;
        }
        IntegrationRule *iRule = elem->giveDefaultIntegrationRulePtr();
        if ( !iRule || iRule->giveNumberOfIntegrationPoints() != ngp || elem->giveNumberOfIntegrationRules() != 1 ) {
            return false;
        }
        const IntArray &dm = elem->giveDofManArray();
        if ( dm.giveSize() != nen ) {
            return false;
        }
        for ( int k = 0; k < nen; k++ ) {
            conn [ ( size_t ) ( e - 1 ) * nen + k ] = dm [ k ];
        }
        elem->giveLocationArray(l, s, & ids);
        if ( l.giveSize() != 3 * nen ) {
            return false;
        }
        for ( int k = 0; k < 3 * nen; k++ ) {
            if ( ids [ k ] != D_u + k % 3 ) {
                return false;
            }
            loc [ ( size_t ) ( e - 1 ) * nen * 3 + k ] = l [ k ];
        }
        // the material comes through the cross section (Structural3DElement::computeConstitutiveMatrixAt,
        // structural3delement.C:99-103); one material per element on this path
        Material *m = elem->giveCrossSection()->giveMaterial( iRule->getIntegrationPoint(0) );
        for ( int g = 1; g < ngp; g++ ) {
            if ( elem->giveCrossSection()->giveMaterial( iRule->getIntegrationPoint(g) ) != m ) {
                return false;
            }
        }
        auto found = matIndex.find( m->giveNumber() );
        if ( found == matIndex.end() ) {
            if ( std :: strcmp(m->giveClassName(), "IsotropicLinearElasticMaterial") ) {
                return false;
            }
            auto *iso = static_cast< IsotropicLinearElasticMaterial * >( m );
            found = matIndex.insert({ m->giveNumber(), ( int ) matIndex.size() }).first;
            double row [ OB200_MATPARAM_STRIDE ] = { ( double ) OB200_MAT_ISOLE, iso->giveYoungsModulus(), iso->givePoissonsRatio(), 0., 0., 0., 0., 0. };
            matparams.insert(matparams.end(), row, row + OB200_MATPARAM_STRIDE);
        }
        matid [ e - 1 ] = found->second;
    }
    std :: vector< double >coords( ( size_t ) nnode * 3, 0.);
    for ( int n = 1; n <= nnode; n++ ) {
        DofManager *dman = domain->giveDofManager(n);
        const FloatArray &c = dman->giveCoordinates();
        for ( int k = 0; k < std :: min(3, c.giveSize()); k++ ) {
            coords [ ( size_t ) ( n - 1 ) * 3 + k ] = c [ k ];
        }
    }
    int neq = eModel->giveNumberOfDomainEquations(domain->giveNumber(), s);
    CudaContext :: check(ob200_elemset_create(CudaContext :: get(), etype, nnode, coords.data(), nelem, conn.data(), matid.data(),
                                              ( int32_t ) matIndex.size(), matparams.data(), loc.data(), neq, 0, & set),
                         "CudaCSR: element set");
    CudaContext :: check(ob200_elemset_bind(set, A), "CudaCSR: element set bind");
    setDomain = domain;
    setDomainVersion = domain->giveSerialNumber();
    return true;
}

bool CudaCSR :: assembleBatched(EngngModel *eModel, TimeStep *tStep, const MatrixAssembler &ma,
                                const UnknownNumberingScheme &s, Domain *domain)
{
    // tangent stiffness only (for the linear elastic material every MatResponseMode gives the same D)
    if ( typeid( ma ) != typeid( TangentAssembler ) || std :: getenv("OOFEM_B200_NO_BATCH") ) {
        return false;
    }
    if ( set && ( setDomain != domain || setDomainVersion != domain->giveSerialNumber() ) ) {
        this->dropElementSet();
    }
    if ( !set ) {
        if ( setTried ) {
            return false;
        }
        setTried = true;
        if ( !this->buildElementSet(eModel, s, domain) ) {
            return false;
        }
    }
    // the conditions the host loop tests per element and step (engngm.C:909-911)
    for ( auto &elem : domain->giveElements() ) {
        if ( elem->giveParallelMode() == Element_remote || !elem->isActivated(tStep) || !eModel->isElementActivated( elem.get() ) ) {
            return false;
        }
    }
    this->flush();
    CudaContext :: check(ob200_elemset_assemble_stiffness(set, A), "CudaCSR::assembleBatched");
    this->version++;
    return true;
}
} // namespace oofem
