// TEST / BENCH INFRASTRUCTURE ONLY -- times the UNMODIFIED reference's own implementation of
// the hot path on the host: CompCol::buildInternalStructure, EngngModel::assemble with
// TangentAssembler (LSpace::computeStiffnessMatrix per element + CompCol::assemble) and
// IMLSolver::solve (IML++ CG, DiagPreconditioner) for a fixed number of iterations.
// Links the reference objects built by oracle/build_ref.py; set-up follows src/main/main.C.
//
//   oofem_bench <input.in> <cg_iterations> <repeats>
// prints one JSON line.
#include "oofemenv.h"
#include "engngm.h"
#include "domain.h"
#include "element.h"
#include "timestep.h"
#include "oofemtxtdatareader.h"
#include "util.h"
#include "compcol.h"
#include "floatarray.h"
#include "unknownnumberingscheme.h"
#include "assemblercallback.h"
#include "dynamicinputrecord.h"
#include "iml/imlsolver.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oofem;
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    if ( argc < 4 ) { fprintf(stderr, "usage: %s input.in cg_iters repeats\n", argv[0]); return 2; }
    int iters = atoi(argv[2]), repeats = atoi(argv[3]);
    OOFEMTXTDataReader dr(argv[1]);
    auto problem = InstanciateProblem(dr, _processor, 0, NULL, false);
    dr.finish();
    if ( !problem ) return 1;
    problem->checkProblemConsistency();
    problem->init();
    problem->solveYourself();            // the input asks for a cheap solve; sets up numbering and the time step
    Domain *d = problem->giveDomain(1);
    TimeStep *tStep = problem->giveCurrentStep();
    EModelDefaultEquationNumbering en;
    int neq = problem->giveNumberOfDomainEquations(1, en);
    int nelem = d->giveNumberOfElements();

    CompCol A(0);
    double t0 = now();
    A.buildInternalStructure(problem.get(), 1, en);
    double t_struct = now() - t0;
    double t_asm = 1e300;
    for ( int r = 0; r < repeats; r++ ) {
        A.zero();
        t0 = now();
        problem->assemble(A, tStep, TangentAssembler(TangentStiffness), en, d);
        double t = now() - t0;
        if ( t < t_asm ) t_asm = t;
    }
    IMLSolver s(d, problem.get());
    DynamicInputRecord ir;
    ir.setField(1.e-300, "lstol");
    ir.setField(iters, "lsiter");
    ir.setField(1, "lsprecond");
    s.initializeFrom(ir);
    FloatArray b(neq), x(neq);
    for ( int i = 0; i < neq; i++ ) b[i] = 1.0 + 0.001 * ( ( i * 7919 ) % 1013 );
    double t_cg = 1e300;
    for ( int r = 0; r < repeats; r++ ) {
        x.zero();
        t0 = now();
        s.solve(A, b, x);
        double t = now() - t0;
        if ( t < t_cg ) t_cg = t;
    }
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    printf("{\"nelem\": %d, \"neq\": %d, \"nnz\": %d, \"threads\": %d, \"t_structure_s\": %.6f, \"t_assemble_s\": %.6f, "
           "\"cg_iters\": %d, \"t_cg_s\": %.6f, \"elements_per_s\": %.3f, \"cg_iters_per_s\": %.3f}\n",
           nelem, neq, A.giveRowIndex().giveSize(), threads, t_struct, t_asm, iters, t_cg, nelem / t_asm, iters / t_cg);
    fflush(stdout);
    return 0;
}
