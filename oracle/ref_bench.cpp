// TEST / BENCH INFRASTRUCTURE ONLY -- times the UNMODIFIED reference's own implementation of
// the hot path on the host: CompCol::buildInternalStructure, EngngModel::assemble with
// TangentAssembler (LSpace::computeStiffnessMatrix per element + CompCol::assemble) and
// IMLSolver::solve (IML++ CG, DiagPreconditioner) for a fixed number of iterations.
// Links the reference objects built by oracle/build_ref.py; set-up follows src/main/main.C.
//
//   oofem_bench <input.in> <cg_iterations> <repeats> [<warmup> <sample_elements>]
// prints one JSON line.  With sample_elements > 0 every repeat ("step") is a BOUNDED SAMPLE of the work on the
// full mesh: the loop body of EngngModel::assemble (engngm.C:902-925: TangentAssembler::matrixFromElement,
// locationFromElement, CompCol::assemble under the same OpenMP pragmas) over a window of sample_elements
// consecutive elements (the window moves every step), and cg_iterations IML CG iterations on the full matrix;
// the reference's own EngngModel::assemble over ALL elements is timed once beside it (t_assemble_full_s).
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
    int warmup = argc > 4 ? atoi(argv[4]) : 0, sample = argc > 5 ? atoi(argv[5]) : 0;
    double t_start = now();
    OOFEMTXTDataReader dr(argv[1]);
    auto problem = InstanciateProblem(dr, _processor, 0, NULL, false);
    dr.finish();
    if ( !problem ) return 1;
    problem->checkProblemConsistency();
    problem->init();
    double t_parse = now() - t_start;
    // what EngngModel::solveYourself does before solveYourselfAt (engngm.C:597-609), without the solve itself:
    // at the full 1M-element size the throw-away solve (structure + assembly + output) costs minutes
    problem->preInitializeNextStep();
    problem->giveNextStep();
    problem->forceEquationNumbering();
    problem->initializeYourself( problem->giveCurrentStep() );
    Domain *d = problem->giveDomain(1);
    TimeStep *tStep = problem->giveCurrentStep();
    EModelDefaultEquationNumbering en;
    int neq = problem->giveNumberOfDomainEquations(1, en);
    int nelem = d->giveNumberOfElements();

    CompCol A(0);
    double t0 = now();
    A.buildInternalStructure(problem.get(), 1, en);
    double t_struct = now() - t0;
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    // the reference's own assembly of the whole mesh, once (first touch of every element: Gauss points, cross sections)
    A.zero();
    t0 = now();
    problem->assemble(A, tStep, TangentAssembler(TangentStiffness), en, d);
    double t_full_first = now() - t0;
    IMLSolver s(d, problem.get());
    DynamicInputRecord ir;
    ir.setField(1.e-300, "lstol");
    ir.setField(iters, "lsiter");
    ir.setField(1, "lsprecond");
    s.initializeFrom(ir);
    FloatArray b(neq), x(neq);
    for ( int i = 0; i < neq; i++ ) b[i] = 1.0 + 0.001 * ( ( i * 7919 ) % 1013 );

    double t_asm = 1e300, t_cg = 1e300, t_asm_sum = 0, t_cg_sum = 0;
    int nsample = nelem;
    if ( sample <= 0 || sample >= nelem ) {
        // whole-mesh mode: every repeat is the full EngngModel::assemble + a full CG run; best of `repeats`
        for ( int r = 0; r < repeats; r++ ) {
            A.zero();
            t0 = now();
            problem->assemble(A, tStep, TangentAssembler(TangentStiffness), en, d);
            double t = now() - t0;
            if ( t < t_asm ) t_asm = t;
            x.zero();
            t0 = now();
            s.solve(A, b, x);
            t = now() - t0;
            if ( t < t_cg ) t_cg = t;
        }
        if ( repeats <= 0 ) { t_asm = t_full_first; t_cg = 0; }
        t_asm_sum = t_asm * repeats;
        t_cg_sum = t_cg * repeats;
    } else {
        nsample = sample;
        TangentAssembler ma(TangentStiffness);
        for ( int r = -warmup; r < repeats; r++ ) {
            const int e0 = (int)( ( (long long)( r + warmup ) * sample ) % ( nelem - sample + 1 ) );
            IntArray loc;
            FloatMatrix mat, R;
            t0 = now();
#ifdef _OPENMP
#pragma omp parallel for shared(A) private(mat, R, loc)
#endif
            for ( int ielem = e0 + 1; ielem <= e0 + sample; ielem++ ) {
                auto element = d->giveElement(ielem);
                ma.matrixFromElement(mat, *element, tStep);
                if ( mat.isNotEmpty() ) {
                    ma.locationFromElement(loc, *element, en);
                    if ( element->giveRotationMatrix(R) ) mat.rotatedWith(R);
#ifdef _OPENMP
#pragma omp critical
#endif
                    if ( A.assemble(loc, mat) == 0 ) OOFEM_ERROR("sparse matrix assemble error");
                }
            }
            double ta = now() - t0;
            x.zero();
            t0 = now();
            s.solve(A, b, x);
            double tc = now() - t0;
            if ( r >= 0 ) {
                t_asm_sum += ta;
                t_cg_sum += tc;
                if ( ta < t_asm ) t_asm = ta;
                if ( tc < t_cg ) t_cg = tc;
            }
        }
    }
    const double asm_mean = repeats > 0 ? t_asm_sum / repeats : t_asm, cg_mean = repeats > 0 ? t_cg_sum / repeats : t_cg;
    printf("{\"nelem\": %d, \"neq\": %d, \"nnz\": %d, \"threads\": %d, \"t_parse_s\": %.3f, \"t_structure_s\": %.6f, "
           "\"t_assemble_full_s\": %.6f, \"sample_elements\": %d, \"steps\": %d, \"warmup\": %d, \"t_assemble_s\": %.6f, "
           "\"t_assemble_best_s\": %.6f, \"cg_iters\": %d, \"t_cg_s\": %.6f, \"t_cg_best_s\": %.6f, \"elements_per_s\": %.3f, "
           "\"elements_per_s_full\": %.3f, \"cg_iters_per_s\": %.3f}\n",
           nelem, neq, A.giveRowIndex().giveSize(), threads, t_parse, t_struct, t_full_first, nsample, repeats, warmup, asm_mean, t_asm,
           iters, cg_mean, t_cg, nsample / asm_mean, nelem / t_full_first, cg_mean > 0 ? iters / cg_mean : 0.0);
    fflush(stdout);
    return 0;
}
