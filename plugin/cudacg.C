#include "cudacg.h"
#include "cudacontext.h"

#include "classfactory.h"
#include "error.h"
#include "floatarray.h"
#include "inputrecord.h"
#include "iml/imlsolver.h"

namespace oofem {
REGISTER_SparseLinSolver(CudaCGSolver, ST_CudaCG);

void CudaCGSolver :: initializeFrom(InputRecord &ir)
{
    int val = 0;
    IR_GIVE_OPTIONAL_FIELD(ir, val, _IFT_IMLSolver_stype);
    if ( val != 0 ) {
        throw ValueInputException(ir, _IFT_IMLSolver_stype, "cudacg implements the CG solver (stype 0) only");
    }
    tol = 1.e-5;
    IR_GIVE_OPTIONAL_FIELD(ir, tol, _IFT_IMLSolver_lstol);
    maxite = 200;
    IR_GIVE_OPTIONAL_FIELD(ir, maxite, _IFT_IMLSolver_lsiter);
    val = 0;
    IR_GIVE_OPTIONAL_FIELD(ir, val, _IFT_IMLSolver_lsprecond);
    if ( val != OB200_PRECOND_VOID && val != OB200_PRECOND_DIAG ) {
        throw ValueInputException(ir, _IFT_IMLSolver_lsprecond, "unknown preconditioner type");   // imlsolver.C:93
    }
    precondType = val;
}

ConvergedReason CudaCGSolver :: solve(SparseMtrx &A, FloatArray &b, FloatArray &x)
{
    if ( x.giveSize() != b.giveSize() ) {
        OOFEM_ERROR("size mismatch");                                              // imlsolver.C:105-107
    }
    CudaCSR *M = dynamic_cast< CudaCSR * >( & A );
    if ( !M ) {
        OOFEM_ERROR("cudacg needs a cudacsr matrix (smtype 11), got %s", A.giveClassName());
    }
    CudaPhaseTimer timer("solve_s");
    int it = 0;
    double res = 0.;
    int flag = ob200_cg_solve(M->giveHandle(), b.givePointer(), x.givePointer(), precondType, maxite, tol, & it, & res, 0);
    if ( flag < 0 ) {
        OOFEM_ERROR("%s", ob200_last_error());
    }
    lastIterations = it;
    lastResidual = res;
    OOFEM_LOG_INFO("CudaCG(%s): flag=%d, nite %d, achieved tol. %g\n", precondType == OB200_PRECOND_DIAG ? "DiagPrec" : "VoidPrec", flag, it, res);
    return flag == 0 ? CR_CONVERGED : CR_DIVERGED_ITS;
}
} // namespace oofem
