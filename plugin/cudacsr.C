#include "cudacsr.h"
#include "cudacontext.h"
#include "cudaparallel.h"

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
#include "fieldmanager.h"
#include "gausspoint.h"
#include "generalboundarycondition.h"
#include "integrationrule.h"
#include "material.h"
#include "timestep.h"
#include "unknownnumberingscheme.h"
#include "sm/Elements/3D/lspace.h"
#include "sm/Elements/3D/ltrspace.h"
#include "sm/Materials/isolinearelasticmaterial.h"
#include "sm/Materials/misesmat.h"
#include "sm/Materials/structuralms.h"
#include "sm/EngineeringModels/structengngmodel.h"
#include "dof.h"
#include "dofiditem.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <typeindex>
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
    ob200_csr_destroy(A);
}

int CudaCSR :: buildInternalStructure(EngngModel *eModel, int di, const UnknownNumberingScheme &s)
{
    CudaPhaseTimer timer("structure_s");
    // the pattern CompCol builds (compcol.C:174-267): union over elements of loc x loc, plus the
    // location arrays of the active boundary conditions
    Domain *domain = eModel->giveDomain(di);
    int neq = eModel->giveNumberOfDomainEquations(di, s);
    const int nelemAll = domain->giveNumberOfElements();
    std :: vector< IntArray >locs(nelemAll);
    locs.reserve(nelemAll + 8);
    std :: vector< int >widths(cudaPluginThreads(), 0);
    parallelFor(nelemAll, [ & ](long b, long e, int t) {
        for ( long i = b; i < e; i++ ) {
            domain->giveElement( ( int ) i + 1)->giveLocationArray(locs [ i ], s);
            widths [ t ] = std :: max(widths [ t ], locs [ i ].giveSize());
        }
    });
    int width = * std :: max_element(widths.begin(), widths.end());
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
    parallelFor( ( long ) locs.size(), [ & ](long b, long e2, int) {
        for ( long e = b; e < e2; e++ ) {
            for ( int k = 0; k < locs [ e ].giveSize(); k++ ) {
                flat [ ( size_t ) e * width + k ] = locs [ e ] [ k ];
            }
        }
    });
    pendLoc.clear();
    pendMat.clear();
    pendCount = 0;
    cells.clear();
    batchedUsed = false;
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
                   batchedUsed ? "on" : "off");
}

// ---- batched element-evaluation hook ------------------------------------------------------------
//
// One resident element set per domain (ob200_elemset: connectivity, coordinates, location arrays, material table, MisesMat
// state in HBM).  Accepted: a domain made of plain LSpace or plain LTRSpace elements (small strain, default integration
// rule, no local coordinate systems) whose materials are IsotropicLinearElasticMaterial or MisesMat (hType 0).  Anything
// else keeps the reference's host loops.

namespace {
// protected members the hook has to look at (the classes have no accessors for them)
struct TangentPeek : public TangentAssembler {
    using TangentAssembler :: rmode;
};
struct MisesPeek : public MisesMat {
    using MisesMat :: hType;
    using MisesMat :: linearElasticMaterial;
    using MisesMat :: omega_crit;
    using MisesMat :: a;
};

struct BatchedDomain {
    ob200_elemset *set = nullptr;
    Domain *domain = nullptr;
    int domainVersion = -1, neq = -1, etype = 0, nen = 0, ngp = 0, nelem = 0, nnode = 0;
    bool tried = false, hasMises = false;
    std :: type_index scheme = typeid( void );      // numbering the set was built with (loc, neq)
    // (step number, solution state counter) of the last status synchronisation / commit
    int syncStep = -1, commitStep = -1;
    long syncCounter = -1, commitCounter = -1;
    std :: vector< double >matparams;              // the table the set was created with
    std :: vector< Material * >mats;               // row -> material
    std :: vector< double >u;                      // scratch: nodal displacements

    void drop()
    {
        if ( set ) {
            ob200_elemset_destroy(set);
            set = nullptr;
        }
        tried = false;
        syncStep = commitStep = -1;
    }

    static bool materialRow(Material *m, GaussPoint *gp, TimeStep *tStep, double row [ OB200_MATPARAM_STRIDE ])
    {
        for ( int k = 0; k < OB200_MATPARAM_STRIDE; k++ ) {
            row [ k ] = 0.;
        }
        if ( !std :: strcmp(m->giveClassName(), "IsotropicLinearElasticMaterial") ) {
            auto *iso = static_cast< IsotropicLinearElasticMaterial * >( m );
            if ( m->giveCastingTime() >= 0. ) {
                return false;                       // reduced stiffness before casting, incremental stress update
            }                                       // (linearelasticmaterial.C:82-121): host loop
            row [ 0 ] = OB200_MAT_ISOLE;
            row [ 1 ] = iso->giveYoungsModulus();
            row [ 2 ] = iso->givePoissonsRatio();
            return true;
        }
        if ( !std :: strcmp(m->giveClassName(), "MisesMat") ) {
            auto *mm = static_cast< MisesMat * >( m );
            const MisesPeek *pk = static_cast< const MisesPeek * >( mm );
            if ( pk->hType != 0 ) {
                return false;                       // user-defined hardening curve: host loop
            }
            row [ 0 ] = OB200_MAT_MISES;
            row [ 1 ] = pk->linearElasticMaterial.giveYoungsModulus();
            row [ 2 ] = pk->linearElasticMaterial.givePoissonsRatio();
            row [ 3 ] = mm->computeYieldStress(0., gp, tStep);          // sig0 (a ScalarFunction: re-read every call)
            row [ 4 ] = mm->computeYieldStressPrime(0.);                // H
            row [ 5 ] = pk->omega_crit;
            row [ 6 ] = pk->a;
            return true;
        }
        return false;
    }

    bool build(EngngModel *eModel, TimeStep *tStep, const UnknownNumberingScheme &s, Domain *d)
    {
        CudaPhaseTimer timer("elemset_build_s");
        nelem = d->giveNumberOfElements();
        nnode = d->giveNumberOfDofManagers();
        if ( nelem == 0 || nnode == 0 ) {
            return false;
        }
        const char *cn = d->giveElement(1)->giveClassName();
        if ( !std :: strcmp(cn, "LSpace") ) {
            etype = OB200_LSPACE;
            nen = 8;
        } else if ( !std :: strcmp(cn, "LTRSpace") ) {
            etype = OB200_LTRSPACE;
            nen = 4;
        } else {
            return false;
        }
        ngp = etype == OB200_LSPACE ? 8 : 1;
        std :: vector< int32_t >conn( ( size_t ) nelem * nen), matid(nelem), loc( ( size_t ) nelem * nen * 3);
        std :: map< int, int >matIndex;                  // material number -> row of the parameter table
        matparams.clear();
        mats.clear();
        hasMises = false;
        // pass 1 (threads): what every element is asked -- class, geometry mode, integration rule, connectivity, location
        // array with its dof ids, the material behind the cross section
        std :: vector< Material * >emat(nelem, nullptr);
        std :: vector< char >reject(cudaPluginThreads(), 0);
        const int nen_ = nen, ngp_ = ngp;
        parallelFor(nelem, [ & ](long b, long e2, int t) {
            IntArray l, ids;
            FloatMatrix R;
            for ( long i = b; i < e2 && !reject [ t ]; i++ ) {
                Element *elem = d->giveElement( ( int ) i + 1);
                reject [ t ] = 1;                                   // cleared at the end of a clean pass
                if ( std :: strcmp(elem->giveClassName(), cn) ) {
                    break;
                }
                NLStructuralElement *se = dynamic_cast< NLStructuralElement * >( elem );
                if ( !se || se->giveGeometryMode() != 0 || elem->giveRotationMatrix(R) ) {
                    break;
                }
                IntegrationRule *iRule = elem->giveDefaultIntegrationRulePtr();
                if ( !iRule || iRule->giveNumberOfIntegrationPoints() != ngp_ || elem->giveNumberOfIntegrationRules() != 1 ) {
                    break;
                }
                const IntArray &dm = elem->giveDofManArray();
                if ( dm.giveSize() != nen_ ) {
                    break;
                }
                for ( int k = 0; k < nen_; k++ ) {
                    conn [ ( size_t ) i * nen_ + k ] = dm [ k ];
                }
                elem->giveLocationArray(l, s, & ids);
                if ( l.giveSize() != 3 * nen_ ) {
                    break;
                }
                bool ok = true;
                for ( int k = 0; k < 3 * nen_; k++ ) {
                    ok = ok && ids [ k ] == D_u + k % 3;
                    loc [ ( size_t ) i * nen_ * 3 + k ] = l [ k ];
                }
                if ( !ok ) {
                    break;
                }
                // the material comes through the cross section (Structural3DElement::computeConstitutiveMatrixAt,
                // structural3delement.C:99-103), a plain SimpleCrossSection forwards to it; one material per element on this path
                if ( std :: strcmp(elem->giveCrossSection()->giveClassName(), "SimpleCrossSection") ) {
                    break;
                }
                Material *m = elem->giveCrossSection()->giveMaterial( iRule->getIntegrationPoint(0) );
                for ( int g = 1; g < ngp_ && ok; g++ ) {
                    ok = elem->giveCrossSection()->giveMaterial( iRule->getIntegrationPoint(g) ) == m;
                }
                if ( !ok || !m ) {
                    break;
                }
                emat [ i ] = m;
                reject [ t ] = 0;
            }
        });
        if ( std :: any_of(reject.begin(), reject.end(), [](char c) { return c != 0; }) ) {
            return false;
        }
        // pass 2 (serial): the parameter table, one row per material in order of first use
        Material *lastMat = nullptr;
        int lastRow = -1;
        for ( int e = 1; e <= nelem; e++ ) {
            Material *m = emat [ e - 1 ];
            if ( m != lastMat ) {
                auto found = matIndex.find( m->giveNumber() );
                if ( found == matIndex.end() ) {
                    double row [ OB200_MATPARAM_STRIDE ];
                    GaussPoint *gp0 = d->giveElement(e)->giveDefaultIntegrationRulePtr()->getIntegrationPoint(0);
                    if ( !materialRow(m, gp0, tStep, row) ) {
                        return false;
                    }
                    hasMises = hasMises || row [ 0 ] == OB200_MAT_MISES;
                    found = matIndex.insert({ m->giveNumber(), ( int ) matIndex.size() }).first;
                    matparams.insert(matparams.end(), row, row + OB200_MATPARAM_STRIDE);
                    mats.push_back(m);
                }
                lastMat = m;
                lastRow = found->second;
            }
            matid [ e - 1 ] = lastRow;
        }
        std :: vector< double >coords( ( size_t ) nnode * 3, 0.);
        for ( int n = 1; n <= nnode; n++ ) {
            const FloatArray &c = d->giveDofManager(n)->giveCoordinates();
            for ( int k = 0; k < std :: min(3, c.giveSize()); k++ ) {
                coords [ ( size_t ) ( n - 1 ) * 3 + k ] = c [ k ];
            }
        }
        neq = eModel->giveNumberOfDomainEquations(d->giveNumber(), s);
        CudaContext :: check(ob200_elemset_create(CudaContext :: get(), etype, nnode, coords.data(), nelem, conn.data(), matid.data(),
                                                  ( int32_t ) matIndex.size(), matparams.data(), loc.data(), neq, 0, & set),
                             "CudaCSR: element set");
        domain = d;
        domainVersion = d->giveSerialNumber();
        scheme = typeid( s );
        return true;
    }

    /// The set for this domain, built on first use with the numbering of that call; nullptr if the domain is not of the
    /// accepted kind (or a condition the host loops test per element and step does not hold now).  A call with ANOTHER
    /// numbering scheme (reactions: EModelDefaultPrescribedEquationNumbering) gets the existing set -- whose location
    /// arrays are then not the caller's: it must scatter itself -- and never builds or rebuilds one (a rebuild would
    /// lose the resident material state).
    ob200_elemset *get(EngngModel *eModel, TimeStep *tStep, const UnknownNumberingScheme &s, Domain *d)
    {
        if ( std :: getenv("OOFEM_B200_NO_BATCH") ) {
            return nullptr;
        }
        const bool own = !set || scheme == std :: type_index( typeid( s ) );
        if ( set && ( domain != d || domainVersion != d->giveSerialNumber() ||
                      ( own && neq != eModel->giveNumberOfDomainEquations(d->giveNumber(), s) ) ) ) {
            if ( !own ) {
                return nullptr;
            }
            this->drop();
        }
        if ( !set ) {
            if ( tried ) {
                return nullptr;
            }
            tried = true;
            if ( !this->build(eModel, tStep, s, d) ) {
                return nullptr;
            }
        }
        // the conditions the host loops test per element and step (engngm.C:909-911, 1386-1392)
        std :: vector< char >inactive(cudaPluginThreads(), 0);
        parallelFor(d->giveNumberOfElements(), [ & ](long b, long e, int t) {
            for ( long i = b; i < e; i++ ) {
                Element *elem = d->giveElement( ( int ) i + 1);
                if ( elem->giveParallelMode() == Element_remote || !elem->isActivated(tStep) || !eModel->isElementActivated(elem) ) {
                    inactive [ t ] = 1;
                    return;
                }
            }
        }, 65536);
        if ( std :: any_of(inactive.begin(), inactive.end(), [](char c) { return c != 0; }) ) {
            return nullptr;
        }
        // stress-independent strains -- temperature or eigenstrain loads, temperature / eigenstrain fields
        // (StructuralMaterial::computeStressIndependentStrainVector_3d, structuralmaterial.C:2268-2340) -- enter the host's
        // stresses and internal forces; the resident kernels do not evaluate them: such a domain keeps the host loops
        for ( auto &bc : d->giveBcs() ) {
            const bcValType vt = bc->giveBCValType();
            if ( vt == TemperatureBVT || vt == EigenstrainBVT ) {
                return nullptr;
            }
        }
        if ( FieldManager *fm = eModel->giveContext()->giveFieldManager() ) {
            if ( fm->isFieldRegistered(FT_Temperature) || fm->isFieldRegistered(FT_EigenStrain) ) {
                return nullptr;
            }
        }
        // material parameters that may depend on time (MisesMat sig0 is a ScalarFunction): a change sends the job back to the host
        for ( std :: size_t m = 0; m < mats.size(); m++ ) {
            double row [ OB200_MATPARAM_STRIDE ];
            GaussPoint *gp = d->giveElement(1)->giveDefaultIntegrationRulePtr()->getIntegrationPoint(0);
            if ( !materialRow(mats [ m ], gp, tStep, row) || std :: memcmp(row, & matparams [ m * OB200_MATPARAM_STRIDE ], sizeof( row ) ) ) {
                return nullptr;
            }
        }
        return set;
    }

    /// nodal values of (D_u, D_v, D_w) in the given mode: what StructuralElement::computeVectorOf collects per element
    const double *displacements(ValueModeType mode, TimeStep *tStep)
    {
        u.assign( ( size_t ) nnode * 3, 0.);
        parallelFor(nnode, [ & ](long b, long e, int) {
            for ( long n = b; n < e; n++ ) {
                DofManager *dman = domain->giveDofManager( ( int ) n + 1);
                for ( int k = 0; k < 3; k++ ) {
                    auto it = dman->findDofWithDofId( ( DofIDItem ) ( D_u + k ) );
                    if ( it != dman->end() ) {
                        u [ ( size_t ) n * 3 + k ] = ( * it )->giveUnknown(mode, tStep);
                    }
                }
            }
        }, 16384);
        return u.data();
    }
};

std :: map< Domain *, BatchedDomain > &batchedDomains()
{
    static std :: map< Domain *, BatchedDomain >m;
    return m;
}
} // namespace

int CudaCSR :: batchedVectorCalls = 0;
int CudaCSR :: batchedUpdateCalls = 0;
int CudaCSR :: batchedStateCalls = 0;
int CudaCSR :: batchedReactionCalls = 0;

bool CudaCSR :: assembleBatched(EngngModel *eModel, TimeStep *tStep, const MatrixAssembler &ma,
                                const UnknownNumberingScheme &s, Domain *domain)
{
    if ( typeid( ma ) != typeid( TangentAssembler ) ) {
        return false;
    }
    CudaPhaseTimer timer("assemble_matrix_s");
    BatchedDomain &bd = batchedDomains() [ domain ];
    ob200_elemset *set = bd.get(eModel, tStep, s, domain);
    if ( !set || bd.scheme != std :: type_index( typeid( s ) ) ) {
        return false;           // (a set built for another numbering scheme scatters through other location arrays: host loop)
    }
    // MisesMat::give3dMaterialStiffnessMatrix answers the elastic matrix for every mode but TangentStiffness
    // (misesmat.C:496-503); the kernels evaluate the algorithmic tangent
    if ( bd.hasMises && static_cast< const TangentPeek & >( static_cast< const TangentAssembler & >( ma ) ).rmode != TangentStiffness ) {
        return false;
    }
    this->flush();
    CudaContext :: check(ob200_elemset_assemble_stiffness(set, A), "CudaCSR::assembleBatched");
    this->version++;
    if ( !batchedUsed ) {
        OOFEM_LOG_INFO("CudaCSR: batched tangent assembly on the GPU (%d elements)\n", bd.nelem);
    }
    batchedUsed = true;
    return true;
}

bool batchedAssembleVector(EngngModel *eModel, FloatArray &answer, TimeStep *tStep, const VectorAssembler &va, ValueModeType mode,
                           const UnknownNumberingScheme &s, Domain *domain, FloatArray *eNorms)
{
    // internal forces of the total state only: InternalForceAssembler::vectorFromElement asks the element for
    // InternalForcesVector (StructuralElement::giveInternalForcesVector, useUpdatedGpRecord = 0);
    // LastEquilibratedInternalForceAssembler (StructuralEngngModel::computeReaction, structengngmodel.C:183-201) for the same
    // integral over the stresses stored in the committed statuses (useUpdatedGpRecord = 1, structuralelement.C:926-930)
    const bool last = typeid( va ) == typeid( LastEquilibratedInternalForceAssembler );
    if ( !( last || typeid( va ) == typeid( InternalForceAssembler ) ) || mode != VM_Total ) {
        return false;
    }
    auto it = batchedDomains().find(domain);
    if ( it == batchedDomains().end() ) {
        return false;                               // no cudacsr matrix on this domain: not our job
    }
    CudaPhaseTimer timer(last ? "reaction_forces_s" : "assemble_vector_s");
    BatchedDomain &bd = it->second;
    ob200_elemset *set = bd.get(eModel, tStep, s, domain);
    if ( !set ) {
        return false;
    }
    // the stored stresses are those of the state the resident set committed for this very step: evaluating the
    // displacements of tStep reproduces them; any other moment (another step, before the update) stays on the host.
    // With a MisesMat the evaluation would also rewrite the temporary state behind the commit (tempKappa = kappa up to
    // round-off decides the branch of the next tangent, misesmat.C:507-512): those sets leave the reactions to the host.
    if ( last && ( bd.hasMises || !( bd.commitStep == tStep->giveNumber() && bd.commitCounter == ( long ) tStep->giveSolutionStateCounter() ) ) ) {
        return false;
    }
    const bool own = bd.scheme == std :: type_index( typeid( s ) );
    if ( own ) {
        if ( answer.giveSize() != bd.neq ) {
            return false;
        }
        double ebe [ 3 ] = { 0., 0., 0. };
        CudaContext :: check(ob200_elemset_assemble_internal_forces(set, bd.displacements(mode, tStep), answer.givePointer(),
                                                                    eNorms ? ebe : nullptr, 0), "batched internal forces");
        if ( eNorms ) {
            for ( int k = 0; k < 3; k++ ) {
                if ( D_u + k <= eNorms->giveSize() ) {
                    eNorms->at(D_u + k) += ebe [ k ];
                }
            }
        }
    } else {
        // another numbering than the set's (reactions: the prescribed equations): element vectors from the GPU, scattered
        // here through the caller's location arrays exactly as the host loop does (engngm.C:1396-1407)
        const int nd = 3 * bd.nen;
        std :: vector< double >fe( ( size_t ) bd.nelem * nd);
        CudaContext :: check(ob200_elemset_internal_forces(set, bd.displacements(mode, tStep), fe.data(), nullptr, nullptr, 0),
                             "batched internal forces (element vectors)");
        std :: vector< int32_t >locs( ( size_t ) bd.nelem * nd), ids( ( size_t ) bd.nelem * nd);
        std :: vector< int >badElem(cudaPluginThreads(), 0);
        parallelFor(bd.nelem, [ & ](long b, long e2, int t) {
            IntArray loc, dofids;
            for ( long e = b; e < e2; e++ ) {
                va.locationFromElement(loc, * domain->giveElement( ( int ) e + 1), s, & dofids);
                if ( loc.giveSize() != nd || dofids.giveSize() != nd ) {
                    badElem [ t ] = ( int ) e + 1;
                    return;
                }
                for ( int k = 0; k < nd; k++ ) {
                    locs [ ( size_t ) e * nd + k ] = loc [ k ];
                    ids [ ( size_t ) e * nd + k ] = dofids [ k ];
                }
            }
        });
        for ( int be : badElem ) {
            if ( be ) {
                OOFEM_ERROR("batched internal forces: element %d does not have %d location entries", be, nd);
            }
        }
        FloatArray charVec(nd);
        IntArray loc(nd), dofids(nd);
        for ( int e = 0; e < bd.nelem; e++ ) {
            bool any = eNorms != nullptr;
            for ( int k = 0; k < nd && !any; k++ ) {
                any = locs [ ( size_t ) e * nd + k ] != 0;
            }
            if ( !any ) {
                continue;
            }
            for ( int k = 0; k < nd; k++ ) {
                charVec [ k ] = fe [ ( size_t ) e * nd + k ];
                loc [ k ] = locs [ ( size_t ) e * nd + k ];
                dofids [ k ] = ids [ ( size_t ) e * nd + k ];
            }
            answer.assemble(charVec, loc);
            if ( eNorms ) {
                eNorms->assembleSquared(charVec, dofids);
            }
        }
    }
    if ( last ) {
        if ( CudaCSR :: batchedReactionCalls++ == 0 ) {
            OOFEM_LOG_INFO("CudaCSR: batched reaction forces on the GPU (%d elements)\n", bd.nelem);
        }
    } else if ( CudaCSR :: batchedVectorCalls++ == 0 ) {
        OOFEM_LOG_INFO("CudaCSR: batched internal forces on the GPU (%d elements)\n", bd.nelem);
    }
    return true;
}

namespace {
/// Strains, stresses (and the MisesMat variables) of the current solution into the temporary statuses of the host elements:
/// what StructuralElement::updateInternalState leaves there (structuralelement.C:960-972); Element::updateYourself commits
/// them (structuralelement.C:944, misesmat.C:672-690) and the output modules read them.
void syncStatuses(BatchedDomain &bd, TimeStep *tStep)
{
    Domain *domain = bd.domain;
    const size_t ngpt = ( size_t ) bd.nelem * bd.ngp;
    std :: vector< double >eps(ngpt * 6), sig(ngpt * 6), state;
    CudaContext :: check(ob200_elemset_internal_forces(bd.set, bd.displacements(VM_Total, tStep), nullptr, eps.data(), sig.data(), 0),
                         "batched update: strains and stresses");
    if ( bd.hasMises ) {
        state.resize(ngpt * OB200_MISES_STATE_DOUBLES);
        CudaContext :: check(ob200_elemset_get_state(bd.set, state.data(), 0), "batched update: material state");
    }
    // one contiguous range of elements per thread: the statuses of different Gauss points are independent objects
    // (created on first use by Material::giveStatus, as in the reference's own parallel element loops)
    parallelFor(bd.nelem, [ & ](long eb, long ee, int) {
    FloatArray v6(6);
    for ( long e = eb + 1; e <= ee; e++ ) {
        Element *elem = domain->giveElement( ( int ) e);
        IntegrationRule *iRule = elem->giveDefaultIntegrationRulePtr();
        for ( int g = 0; g < bd.ngp; g++ ) {
            GaussPoint *gp = iRule->getIntegrationPoint(g);
            Material *m = elem->giveCrossSection()->giveMaterial(gp);
            auto *st = static_cast< StructuralMaterialStatus * >( m->giveStatus(gp) );
            const size_t o = ( ( size_t ) ( e - 1 ) * bd.ngp + g );
            for ( int k = 0; k < 6; k++ ) v6 [ k ] = eps [ o * 6 + k ];
            st->letTempStrainVectorBe(v6);
            for ( int k = 0; k < 6; k++ ) v6 [ k ] = sig [ o * 6 + k ];
            st->letTempStressVectorBe(v6);
            if ( bd.hasMises && !std :: strcmp(m->giveClassName(), "MisesMat") ) {
                // layout of ob200 MisesState: plStrain[6] kappa damage tempPlStrain[6] tempKappa tempDamage
                // trialStressDev[6] trialStressVol effStress[6]
                const double *q = & state [ o * OB200_MISES_STATE_DOUBLES ];
                auto *ms = static_cast< MisesMatStatus * >( st );
                for ( int k = 0; k < 6; k++ ) v6 [ k ] = q [ 8 + k ];
                ms->letTempPlasticStrainBe(v6);
                ms->setTempCumulativePlasticStrain(q [ 14 ]);
                ms->setTempDamage(q [ 15 ]);
                for ( int k = 0; k < 6; k++ ) v6 [ k ] = q [ 16 + k ];
                ms->letTrialStressDevBe(v6);
                ms->setTrialStressVol(q [ 22 ]);
                for ( int k = 0; k < 6; k++ ) v6 [ k ] = q [ 23 + k ];
                ms->letTempEffectiveStressBe(v6);
            }
        }
    }
    }, 1024);
    bd.syncStep = tStep->giveNumber();
    bd.syncCounter = ( long ) tStep->giveSolutionStateCounter();
}
} // namespace

bool batchedInternalState(EngngModel *eModel, TimeStep *tStep, Domain *domain)
{
    auto it = batchedDomains().find(domain);
    if ( it == batchedDomains().end() || !it->second.set || CudaCSR :: batchedVectorCalls == 0 ||
         std :: getenv("OOFEM_B200_NO_STATUS_SYNC") ) {
        return false;           // the resident material state is only current if the internal forces went through the set
    }
    BatchedDomain &bd = it->second;
    if ( !bd.get(eModel, tStep, EModelDefaultEquationNumbering(), domain) ) {
        return false;
    }
    CudaPhaseTimer timer("status_update_s");
    syncStatuses(bd, tStep);
    if ( CudaCSR :: batchedStateCalls++ == 0 ) {
        OOFEM_LOG_INFO("CudaCSR: batched internal state update on the GPU (%d elements)\n", bd.nelem);
    }
    return true;
}

void batchedUpdate(EngngModel *eModel, TimeStep *tStep, Domain *domain)
{
    auto it = batchedDomains().find(domain);
    if ( it == batchedDomains().end() || !it->second.set || CudaCSR :: batchedVectorCalls == 0 ) {
        return;
    }
    CudaPhaseTimer timer("status_update_s");
    BatchedDomain &bd = it->second;
    // The converged state into the temporary statuses of the host elements, unless StructuralEngngModel::updateInternalState
    // has just done so for this very state (batchedInternalState).  OOFEM_B200_NO_STATUS_SYNC=1 skips the copy (no element
    // output then).
    if ( !std :: getenv("OOFEM_B200_NO_STATUS_SYNC") &&
         !( bd.syncStep == tStep->giveNumber() && bd.syncCounter == ( long ) tStep->giveSolutionStateCounter() ) ) {
        syncStatuses(bd, tStep);
    }
    // MaterialStatus::updateYourself for the resident state: temp -> committed
    CudaContext :: check(ob200_elemset_commit(bd.set), "batched update: commit");
    bd.commitStep = tStep->giveNumber();
    bd.commitCounter = ( long ) tStep->giveSolutionStateCounter();
    CudaCSR :: batchedUpdateCalls++;
}
} // namespace oofem
