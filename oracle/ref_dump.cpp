// TEST INFRASTRUCTURE ONLY -- a driver that links the UNMODIFIED reference objects
// (built by oracle/build_ref.py) and dumps, at full double precision, the quantities the
// parity tests need: equation numbering, element location arrays, the CompCol sparsity
// pattern produced by the reference's own buildInternalStructure, element characteristic
// matrices / internal force vectors from the reference's own element code, the assembled
// values and the solution vector.  It follows src/main/main.C:321-392 for problem set-up.
//
//   oofem_dump <input.in> <out.bin>
//
// Output (little endian): a sequence of named records
//   [int32 namelen][name][int32 dtype (0=int32,1=float64)][int64 count][payload]
#include "oofemenv.h"
#include "engngm.h"
#include "domain.h"
#include "element.h"
#include "dofmanager.h"
#include "dof.h"
#include "timestep.h"
#include "oofemtxtdatareader.h"
#include "util.h"
#include "compcol.h"
#include "floatmatrix.h"
#include "floatarray.h"
#include "intarray.h"
#include "unknownnumberingscheme.h"
#include "assemblercallback.h"
#include "chartype.h"
#include "valuemodetype.h"
#include "error.h"
#include "logger.h"
#include <cstdio>
#include <vector>
#include <string>
#include <cstdint>

using namespace oofem;

static FILE *fo;
static void rec(const std::string &name, int dtype, int64_t n, const void *p) {
    int32_t l = (int32_t) name.size();
    fwrite(&l, 4, 1, fo); fwrite(name.data(), 1, l, fo);
    int32_t d = dtype; fwrite(&d, 4, 1, fo); fwrite(&n, 8, 1, fo);
    fwrite(p, dtype ? 8 : 4, n, fo);
}
static void reci(const std::string &n, const std::vector<int32_t> &v) { rec(n, 0, (int64_t) v.size(), v.data()); }
static void recd(const std::string &n, const std::vector<double> &v) { rec(n, 1, (int64_t) v.size(), v.data()); }

int main(int argc, char **argv) {
    if ( argc < 3 ) { fprintf(stderr, "usage: %s input.in out.bin\n", argv[0]); return 2; }
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
    std::vector<int32_t> meta = { neq, d->giveNumberOfDofManagers(), d->giveNumberOfElements() };
    reci("meta", meta);

    // numbering + solution, node by node, dof by dof (src/core/dofmanager.h iterator)
    std::vector<int32_t> eqs, ndofs; std::vector<double> u, xyz;
    for ( auto &dm : d->giveDofManagers() ) {
        int c = 0;
        for ( Dof *dof : *dm ) { eqs.push_back(dof->giveEquationNumber(en)); u.push_back(dof->giveUnknown(VM_Total, tStep)); c++; }
        ndofs.push_back(c);
        for ( int k = 1; k <= 3; k++ ) xyz.push_back(dm->giveCoordinates().giveSize() >= k ? dm->giveCoordinate(k) : 0.);
    }
    reci("node_ndofs", ndofs); reci("node_eq", eqs); recd("node_u", u); recd("node_xyz", xyz);

    // element location arrays, stiffness matrices (row-major), internal forces
    std::vector<int32_t> locs, ndofel; std::vector<double> ke, fint;
    for ( auto &e : d->giveElements() ) {
        IntArray loc; e->giveLocationArray(loc, en);
        ndofel.push_back(loc.giveSize());
        for ( int i = 0; i < loc.giveSize(); i++ ) locs.push_back(loc[i]);
        FloatMatrix m; e->giveCharacteristicMatrix(m, TangentStiffnessMatrix, tStep);
        for ( int i = 1; i <= m.giveNumberOfRows(); i++ ) for ( int j = 1; j <= m.giveNumberOfColumns(); j++ ) ke.push_back(m.at(i, j));
        FloatArray f; e->giveCharacteristicVector(f, InternalForcesVector, VM_Total, tStep);
        for ( int i = 0; i < f.giveSize(); i++ ) fint.push_back(f[i]);
    }
    reci("elem_ndof", ndofel); reci("elem_loc", locs); recd("elem_ke", ke); recd("elem_fint", fint);

    // the reference's own CompCol: pattern from buildInternalStructure, values from EngngModel::assemble
    CompCol A(0);
    A.buildInternalStructure(problem.get(), 1, en);
    problem->assemble(A, tStep, TangentAssembler(TangentStiffness), en, d);
    std::vector<int32_t> cp, ri; std::vector<double> av;
    for ( int i = 0; i < A.giveColPtr().giveSize(); i++ ) cp.push_back(A.giveColPtr()[i]);
    for ( int i = 0; i < A.giveRowIndex().giveSize(); i++ ) ri.push_back(A.giveRowIndex()[i]);
    for ( int i = 0; i < A.giveValues().giveSize(); i++ ) av.push_back(A.giveValues()[i]);
    reci("colptr", cp); reci("rowind", ri); recd("val", av);

    // one product with a fixed, reproducible vector (CompCol::times) for SpMV parity
    FloatArray x(neq), y;
    for ( int i = 0; i < neq; i++ ) x[i] = 1.0 + 0.001 * ( ( i * 7919 ) % 1013 );
    A.times(x, y);
    std::vector<double> xv(neq), yv(neq);
    for ( int i = 0; i < neq; i++ ) { xv[i] = x[i]; yv[i] = y[i]; }
    recd("spmv_x", xv); recd("spmv_y", yv);
    fclose(fo);
    problem->terminateAnalysis();
    return 0;
}
