// Batched element-evaluation hook: the one interface OOFEM's assembly loops probe for.
//
// EngngModel::assemble (src/core/engngm.C:889-929) walks the elements on the host, asks each for
// its characteristic matrix and hands it to SparseMtrx::assemble.  A matrix that also implements
// this interface is offered the whole loop first; if it answers true the host loop is skipped
// (the boundary-condition part of EngngModel::assemble still runs).  The hook line itself is in
// plugin/engngm_hook.patch; nothing else of the reference changes.
#ifndef oofem_b200_batchedassembly_h
#define oofem_b200_batchedassembly_h

namespace oofem {
class EngngModel;
class TimeStep;
class MatrixAssembler;
class UnknownNumberingScheme;
class Domain;

class BatchedAssemblyTarget
{
public:
    virtual ~BatchedAssemblyTarget() { }
    /// Assemble the contributions of ALL elements of the domain; false = not handled, run the host loop.
    virtual bool assembleBatched(EngngModel *eModel, TimeStep *tStep, const MatrixAssembler &ma,
                                 const UnknownNumberingScheme &s, Domain *domain) = 0;
};
} // namespace oofem
#endif
