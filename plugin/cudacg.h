// SparseLinearSystemNM "cudacg": the IML++ preconditioned CG of IMLSolver (src/core/iml/imlsolver.C:
// 102-146, iml/cg.h) on the GPU.  Selected from an unmodified input record with `lstype 9`; the
// keywords lstol / lsiter / lsprecond keep IMLSolver's meaning and defaults (imlsolver.h:48-52).
#ifndef oofem_b200_cudacg_h
#define oofem_b200_cudacg_h

#include "sparselinsystemnm.h"
#include "convergedreason.h"
#include "cudacsr.h"

namespace oofem {
/// The enumerator a maintainer appends to LinSystSolverType (src/core/linsystsolvertype.h:56).
constexpr LinSystSolverType ST_CudaCG = static_cast< LinSystSolverType >( 9 );

class CudaCGSolver : public SparseLinearSystemNM
{
protected:
    double tol = 1.e-5;
    int maxite = 200;
    int precondType = OB200_PRECOND_VOID;
    int lastIterations = 0;
    double lastResidual = 0.;

public:
    CudaCGSolver(Domain *d, EngngModel *m) : SparseLinearSystemNM(d, m) { }

    void initializeFrom(InputRecord &ir) override;
    ConvergedReason solve(SparseMtrx &A, FloatArray &b, FloatArray &x) override;
    const char *giveClassName() const override { return "CudaCGSolver"; }
    LinSystSolverType giveLinSystSolverType() const override { return ST_CudaCG; }
    SparseMtrxType giveRecommendedMatrix(bool symmetric) const override { return SMT_CudaCSR; }
    int giveLastIterations() const { return lastIterations; }
    double giveLastResidual() const { return lastResidual; }
};
} // namespace oofem
#endif
