// SparseMtrx "cudacsr": OOFEM-side binding of the CSR matrix of liboofem_b200.so.
//
// Same observable behaviour as CompCol (src/core/compcol.C) -- structure from the location arrays
// of the domain, assemble by location array, times, zero, at -- with storage and arithmetic on the
// GPU.  Selected from an unmodified input record with `smtype 11` (SparseMtrxType is read as an int
// and cast: linearstatic.C:100-101, staticstructural.C:110-111).
#ifndef oofem_b200_cudacsr_h
#define oofem_b200_cudacsr_h

#include "sparsemtrx.h"
#include "intarray.h"
#include "batchedassembly.h"
#include "oofem_b200.h"

#include <map>
#include <utility>
#include <vector>

namespace oofem {
/// The enumerator a maintainer appends to SparseMtrxType (src/core/sparsemtrxtype.h:55).
constexpr SparseMtrxType SMT_CudaCSR = static_cast< SparseMtrxType >( 11 );

class CudaCSR : public SparseMtrx, public BatchedAssemblyTarget
{
protected:
    ob200_csr *A = nullptr;
    // per-element contributions handed over by the host loop, staged and sent in batches
    mutable std :: vector< int32_t >pendLoc;
    mutable std :: vector< double >pendMat;
    mutable int pendDofs = 0;
    mutable int64_t pendCount = 0;
    // entries handed out by reference through at(i,j): value when handed out, current value
    struct Cell {
        double seen, cur;
    };
    mutable std :: map< std :: pair< int, int >, Cell >cells;
    // batched element-evaluation hook: the resident element set belongs to the domain (BatchedDomain in cudacsr.C), every
    // CudaCSR of that domain assembles from it
    bool batchedUsed = false;

    void flush() const;                 // send pending contributions and write back modified cells

public:
    CudaCSR(int n = 0);
    CudaCSR(const CudaCSR &) = delete;
    ~CudaCSR() override;

    int buildInternalStructure(EngngModel *eModel, int di, const UnknownNumberingScheme &s) override;
    int assemble(const IntArray &loc, const FloatMatrix &mat) override;
    int assemble(const IntArray &rloc, const IntArray &cloc, const FloatMatrix &mat) override;
    bool assembleBatched(EngngModel *eModel, TimeStep *tStep, const MatrixAssembler &ma,
                         const UnknownNumberingScheme &s, Domain *domain) override;
    void times(const FloatArray &x, FloatArray &answer) const override;
    void timesT(const FloatArray &x, FloatArray &answer) const override;
    void times(double x) override;
    void zero() override;
    double &at(int i, int j) override;
    double at(int i, int j) const override;
    bool isAllocatedAt(int i, int j) const override;
    bool canBeFactorized() const override { return false; }
    void toFloatMatrix(FloatMatrix &answer) const override;
    void printStatistics() const override;
    SparseMtrxType giveType() const override { return SMT_CudaCSR; }
    bool isAsymmetric() const override { return true; }
    const char *giveClassName() const override { return "CudaCSR"; }

    /// Library handle with every pending host-side contribution applied (for the solver).
    ob200_csr *giveHandle() { this->flush(); return A; }
    int64_t giveNumberOfNonzeros() const { return ob200_csr_nnz(A); }
    /// Structure as CompCol stores it (colptr / rowind of the symmetric pattern), for tests.
    void giveStructure(IntArray &rowptr, IntArray &colind) const;
    bool usesBatchedAssembly() const { return batchedUsed; }
    /// how often the vector hook (internal forces) and the status update ran on the GPU (tests)
    static int batchedVectorCalls, batchedUpdateCalls, batchedStateCalls, batchedReactionCalls;
};
} // namespace oofem
#endif
