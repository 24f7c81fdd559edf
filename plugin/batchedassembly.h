// Batched element-evaluation hook: the interface OOFEM's assembly loops probe for.
//
// EngngModel::assemble (src/core/engngm.C:889-929) and EngngModel::assembleVectorFromElements (engngm.C:1351-1407) walk the
// elements on the host, ask each for its characteristic matrix / vector and add it to the global object;
// EngngModel::updateYourself (engngm.C:692-722) commits the material statuses at the end of a step;
// StructuralEngngModel::updateInternalState (src/sm/EngineeringModels/structengngmodel.C:304-312) asks every element to
// recompute the strains and stresses of its Gauss points.  With the three hook sites of plugin/engngm_hook.patch and the one
// of plugin/structengngmodel_hook.patch the loops first offer the whole job to this interface; if it answers true the host
// loop over the elements is skipped (the boundary-condition and load parts of the reference functions still run).  Nothing
// else of the reference changes.
#ifndef oofem_b200_batchedassembly_h
#define oofem_b200_batchedassembly_h

#include "valuemodetype.h"

namespace oofem {
class EngngModel;
class TimeStep;
class MatrixAssembler;
class VectorAssembler;
class UnknownNumberingScheme;
class Domain;
class FloatArray;

class BatchedAssemblyTarget
{
public:
    virtual ~BatchedAssemblyTarget() { }
    /// Assemble the contributions of ALL elements of the domain; false = not handled, run the host loop.
    virtual bool assembleBatched(EngngModel *eModel, TimeStep *tStep, const MatrixAssembler &ma,
                                 const UnknownNumberingScheme &s, Domain *domain) = 0;
};

/// EngngModel::assembleVectorFromElements: the element loop (va.vectorFromElement -> answer.assemble, eNorms->assembleSquared)
/// for ALL elements of the domain; false = not handled.  Defined in cudacsr.C.
bool batchedAssembleVector(EngngModel *eModel, FloatArray &answer, TimeStep *tStep, const VectorAssembler &va, ValueModeType mode,
                           const UnknownNumberingScheme &s, Domain *domain, FloatArray *eNorms);
/// EngngModel::updateYourself, before the elements update themselves: brings the temporary material statuses of the host
/// elements up to date with the GPU-resident ones and commits the latter.  Defined in cudacsr.C.
void batchedUpdate(EngngModel *eModel, TimeStep *tStep, Domain *domain);
/// StructuralEngngModel::updateInternalState: the loop elem->updateInternalState(tStep) (StructuralElement::updateInternalState,
/// structuralelement.C:960-972: strain and stress of every Gauss point for the current solution, left in the temporary
/// material statuses) for ALL elements of the domain; false = not handled.  Defined in cudacsr.C.
bool batchedInternalState(EngngModel *eModel, TimeStep *tStep, Domain *domain);
} // namespace oofem
#endif
