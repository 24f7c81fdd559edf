// Driver for the plugin parity tests: runs an input file through OOFEM (set-up as in src/main/main.C)
// with whatever `lstype` / `smtype` the input selects, then dumps at full precision
//   node_u            the solution the engineering model produced (through cudacg + cudacsr when selected)
//   rowptr/colind/val a CudaCSR built by buildInternalStructure + EngngModel::assemble (batched hook)
//   val_hostloop      the same matrix assembled by the host loop into CudaCSR::assemble(loc, mat)
//   spmv_x / spmv_y   CudaCSR::times on a fixed vector
//   meta              [neq, nnode, nelem, batched matrix hook used (0/1), batched internal-force calls, batched status updates]
// Record format as oracle/ref_dump.cpp: [int32 namelen][name][int32 dtype][int64 count][payload].
//
//   oofem_dump_cuda <input.in> <out.bin>
#include "oofemenv.h"
#include "engngm.h"
#include "domain.h"
#include "element.h"
#include "chartype.h"
#include "dofmanager.h"
#include "dof.h"
#include "timestep.h"
#include "oofemtxtdatareader.h"
#include "util.h"
#include "floatarray.h"
#include "intarray.h"
#include "unknownnumberingscheme.h"
#include "assemblercallback.h"
#include "valuemodetype.h"
#include "cudacsr.h"
#include <cstdio>
#include <csignal>
#include <execinfo.h>
#include <unistd.h>
#include <cstdlib>
#include <cstdint>
#include <string>
#include <vector>

using namespace oofem;

static FILE *fo;
static void rec(const std::string &name, int dtype, int64_t n, const void *p)
{
    int32_t l = (int32_t) name.size();
    fwrite(&l, 4, 1, fo);
    fwrite(name.data(), 1, l, fo);
    int32_t d = dtype;
    fwrite(&d, 4, 1, fo);
    fwrite(&n, 8, 1, fo);
    fwrite(p, dtype ? 8 : 4, n, fo);
}

static std::vector<double> values(CudaCSR &A)
{
    std::vector<double> v(A.giveNumberOfNonzeros() > 0 ? A.giveNumberOfNonzeros() : 1);
    ob200_csr_get_values(A.giveHandle(), v.data(), 0);
    v.resize(A.giveNumberOfNonzeros());
    return v;
}

static void on_segv(int sig)
{
    void *frames[64];
    int n = backtrace(frames, 64);
    fprintf(stderr, "signal %d\n", sig);
    backtrace_symbols_fd(frames, n, 2);
    _exit(128 + sig);
}

int main(int argc, char **argv)
{
    signal(SIGSEGV, on_segv);
    if ( argc < 3 ) {
        fprintf(stderr, "usage: %s input.in out.bin\n", argv[0]);
        return 2;
    }
    OOFEMTXTDataReader dr(argv[1]);
    auto problem = InstanciateProblem(dr, _processor, 0, NULL, false);
    dr.finish();
    if ( !problem ) return 1;
    problem->checkProblemConsistency();
    problem->init();
    problem->solveYourself();

    fo = fopen(argv[2], "wb");
    Domain *d = problem->giveDomain(1);
    TimeStep *tStep = problem->giveCurrentStep();
    EModelDefaultEquationNumbering en;
    int neq = problem->giveNumberOfDomainEquations(1, en);

    std::vector<double> u;
    for ( auto &dm : d->giveDofManagers() )
        for ( Dof *dof : *dm ) u.push_back(dof->giveUnknown(VM_Total, tStep));
    rec("node_u", 1, (int64_t) u.size(), u.data());

    // same sequence as oracle/ref_dump.cpp: the internal forces are evaluated first (for a history-dependent
    // material this refreshes the temporary status the tangent is then taken at)
    for ( auto &e : d->giveElements() ) {
        FloatArray f;
        e->giveCharacteristicVector(f, InternalForcesVector, VM_Total, tStep);
    }
    CudaCSR A(0);
    A.buildInternalStructure(problem.get(), 1, en);
    A.zero();
    problem->assemble(A, tStep, TangentAssembler(TangentStiffness), en, d);
    IntArray rp, ci;
    A.giveStructure(rp, ci);
    std::vector<int32_t> rpv(rp.begin(), rp.end()), civ(ci.begin(), ci.end());
    rec("rowptr", 0, (int64_t) rpv.size(), rpv.data());
    rec("colind", 0, (int64_t) civ.size(), civ.data());
    std::vector<double> v = values(A);
    rec("val", 1, (int64_t) v.size(), v.data());
    std::vector<int32_t> meta = { neq, d->giveNumberOfDofManagers(), d->giveNumberOfElements(), A.usesBatchedAssembly() ? 1 : 0,
                                  CudaCSR::batchedVectorCalls, CudaCSR::batchedUpdateCalls };
    rec("meta", 0, (int64_t) meta.size(), meta.data());

    FloatArray x(neq), y;
    for ( int i = 0; i < neq; i++ ) x[i] = 1.0 + 0.001 * ( ( i * 7919 ) % 1013 );
    A.times(x, y);
    rec("spmv_x", 1, neq, x.givePointer());
    rec("spmv_y", 1, neq, y.givePointer());

    // the unhooked route: host element loop -> CudaCSR::assemble(loc, mat), staged and sent in batches
    setenv("OOFEM_B200_NO_BATCH", "1", 1);
    CudaCSR B(0);
    B.buildInternalStructure(problem.get(), 1, en);
    B.zero();
    problem->assemble(B, tStep, TangentAssembler(TangentStiffness), en, d);
    std::vector<double> vb = values(B);
    rec("val_hostloop", 1, (int64_t) vb.size(), vb.data());
    // SparseMtrx::at, read and write through the reference
    double a11 = ( (const CudaCSR &) B ).at(1, 1);
    B.at(1, 1) += 1.0;
    B.at(1, 2) -= 0.5;
    double probe[4] = { a11, ( (const CudaCSR &) B ).at(1, 1), ( (const CudaCSR &) B ).at(1, 2), vb.size() > 1 ? vb[1] : 0.0 };
    rec("at_probe", 1, 4, probe);
    fclose(fo);
    problem->terminateAnalysis();
    return 0;
}
