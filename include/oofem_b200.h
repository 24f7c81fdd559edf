/*
 * oofem_b200.h -- C ABI of the B200-native structural hot path for OOFEM.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Each entry
 * point names the reference interface it replaces.  All integer index arrays use OOFEM's
 * conventions: node numbers and equation numbers are 1-based, equation number 0 means
 * "prescribed dof" (IntArray loc as produced by Element::giveLocationArray).
 *
 * Memory spaces: every array argument is accompanied (per call) by `on_device`:
 *   0 -> host pointers; the library copies to / from HBM inside the call (what OOFEM's
 *        own C++ passes), 1 -> device pointers on the context's GPU (resident data).
 *
 * Error behaviour: every function returns OB200_OK (0) or a negative OB200_E* code and
 * records a message retrievable with ob200_last_error() -- the analogue of OOFEM_ERROR
 * at the same places the reference raises (dimension mismatch, missing sparsity entry,
 * zero diagonal in the preconditioner, ...).  There is NO CPU fallback: without a CUDA
 * device ob200_context_create fails with OB200_ENODEVICE.
 */
#ifndef OOFEM_B200_H
#define OOFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OB200_OK          0
#define OB200_ENODEVICE  -1   /* no CUDA device / driver */
#define OB200_ECUDA      -2   /* CUDA runtime error (message has details) */
#define OB200_EINVAL     -3   /* bad argument / dimension mismatch */
#define OB200_ESTRUCT    -4   /* entry not in sparse structure (CompCol::assemble DEBUG error) */
#define OB200_EZERODIAG  -5   /* DiagPreconditioner::init: zero diagonal */
#define OB200_ECAPACITY  -6   /* internal capacity exceeded (row with too many couplings) */
#define OB200_ENCCL      -7   /* NCCL error */

/* element types of the batched element-evaluation hook */
#define OB200_LSPACE   1      /* src/sm/Elements/3D/lspace.C   : 8-node brick, 8 Gauss points */
#define OB200_LTRSPACE 2      /* src/sm/Elements/3D/ltrspace.C : 4-node tetra, 1 Gauss point  */

/* materials; parameter block per material = 8 doubles [type, E, nu, sig0, H, omega_crit, a, 0] */
#define OB200_MAT_ISOLE 1     /* src/sm/Materials/isolinearelasticmaterial.C */
#define OB200_MAT_MISES 2     /* src/sm/Materials/misesmat.C (hType 0) */
#define OB200_MATPARAM_STRIDE 8
#define OB200_MISES_STATE_DOUBLES 29   /* per Gauss point, see oofem_b200/csrc/element_kernels.cuh */

/* IML preconditioner selector (IMLSolver::IMLPrecondType, src/core/iml/imlsolver.h:70) */
#define OB200_PRECOND_VOID 0
#define OB200_PRECOND_DIAG 1

typedef struct ob200_context ob200_context;
typedef struct ob200_csr     ob200_csr;
typedef struct ob200_elemset ob200_elemset;
typedef struct ob200_comm    ob200_comm;

const char *ob200_last_error(void);
const char *ob200_version(void);

/* ---- context: one per process / GPU ------------------------------------------------- */
int  ob200_context_create(int device, ob200_context **out);
void ob200_context_destroy(ob200_context *ctx);
int  ob200_context_sync(ob200_context *ctx);
/* cudaStream_t all kernels of this context are launched on (for CUDA-event timing) */
void *ob200_context_stream(ob200_context *ctx);
/* number of kernels launched through this context so far (bench.py's gpu_launches) */
int64_t ob200_context_launch_count(ob200_context *ctx);
/* per-kernel CUDA-event timing on the launching stream (bench.py's roofline): enable, run,
 * then read "kernel-name<TAB>total-ms<TAB>launches" lines */
int  ob200_context_set_profiling(ob200_context *ctx, int enable);
int  ob200_context_profile_reset(ob200_context *ctx);
int  ob200_context_profile_report(ob200_context *ctx, char *buf, int64_t buflen);
/* device memory helpers so that callers without a CUDA runtime binding can keep data resident */
int  ob200_malloc(ob200_context *ctx, int64_t bytes, void **dptr);
int  ob200_free(ob200_context *ctx, void *dptr);
int  ob200_memcpy_h2d(ob200_context *ctx, void *dst, const void *src, int64_t bytes);
int  ob200_memcpy_d2h(ob200_context *ctx, void *dst, const void *src, int64_t bytes);
int  ob200_memset(ob200_context *ctx, void *dst, int value, int64_t bytes);
/* write `bytes` of a scratch buffer (> L2) to evict L2 between timed iterations */
int  ob200_flush_l2(ob200_context *ctx);

/* ---- SparseMtrx "cudacsr" (src/core/sparsemtrx.h, modelled on CompCol, src/core/compcol.C) */
int  ob200_csr_create(ob200_context *ctx, ob200_csr **out);
void ob200_csr_destroy(ob200_csr *A);
/* CompCol::buildInternalStructure (compcol.C:167-260): pattern = union over elements of
 * loc x loc (non-zero entries), rows sorted ascending.  loc is [nelem][ndofel]. */
int  ob200_csr_build_structure(ob200_csr *A, int32_t neq, int64_t nelem, int32_t ndofel,
                               const int32_t *loc, int on_device);
int32_t ob200_csr_rows(const ob200_csr *A);
/* How SparseMtrx::times reads the matrix: info[0] = 1 if the blocked index is in use (rows of a node share
 * one pattern, columns of a node are consecutive: oofem_b200/csrc/spmv_block.cuh), info[1] = row blocks,
 * info[2] = column blocks; all 0 for plain CSR.  rowptr/colind are the same in both cases. */
int  ob200_csr_spmv_layout(ob200_csr *A, int64_t *info3);                 /* SparseMtrx::giveNumberOfRows */
int64_t ob200_csr_nnz(const ob200_csr *A);
/* copy out rowptr[neq+1], colind[nnz] (0-based, like CompCol's colptr/rowind) and val[nnz] */
int  ob200_csr_get_structure(const ob200_csr *A, int32_t *rowptr, int32_t *colind, int on_device);
int  ob200_csr_get_values(const ob200_csr *A, double *val, int on_device);
int  ob200_csr_set_values(ob200_csr *A, const double *val, int on_device);
/* device pointers of the resident arrays (rowptr, colind, val) */
int  ob200_csr_device_arrays(ob200_csr *A, const int32_t **rowptr, const int32_t **colind, double **val);
int  ob200_csr_zero(ob200_csr *A);                          /* SparseMtrx::zero (compcol.C:339) */
int  ob200_csr_scale(ob200_csr *A, double s);               /* SparseMtrx::times(double) (compcol.C:159) */
/* SparseMtrx::assemble(loc, mat) (compcol.C:263-299) for a batch: loc [nelem][ndofel],
 * mat [nelem][ndofel*ndofel] row-major; A(loc_i, loc_j) += mat(i,j).  nelem = 1 is the
 * plain per-element call the reference makes. */
int  ob200_csr_assemble(ob200_csr *A, int64_t nelem, int32_t ndofel, const int32_t *loc,
                        const double *mat, int on_device);
/* SparseMtrx::assemble(rloc, cloc, mat) (compcol.C:301-336): one rectangular block, mat [nr][nc] row-major; only the
 * rloc x cloc entries are looked up */
int  ob200_csr_assemble_rect(ob200_csr *A, int32_t nr, int32_t nc, const int32_t *rloc, const int32_t *cloc,
                             const double *mat, int on_device);
int  ob200_csr_times(ob200_csr *A, const double *x, double *y, int on_device);   /* SparseMtrx::times (compcol.C:119) */
/* SparseMtrx::timesT (compcol.C:146-163): y = A^T x */
int  ob200_csr_times_t(ob200_csr *A, const double *x, double *y, int on_device);
/* SparseMtrx::at, 1-based (compcol.C:376): returns 0 and the value if (i,j) is stored, 1 and value 0 if it is in bounds but not in
 * the sparse structure (CompCol::at const; isAllocatedAt is false), OB200_EINVAL out of bounds */
int  ob200_csr_at(ob200_csr *A, int32_t i, int32_t j, double *value);
int64_t ob200_csr_version(const ob200_csr *A);              /* SparseMtrx::giveVersion */

/* ---- batched element-evaluation hook ------------------------------------------------ */
/* A homogeneous set of elements resident in HBM: replaces the per-element loop of
 * EngngModel::assemble / assembleVector (src/core/engngm.C:889-929) over
 * StructuralElement::computeStiffnessMatrix / giveInternalForcesVector
 * (src/sm/Elements/structuralelement.C:575-643, 724-802).
 *   coords [nnode][3], conn [nelem][nen] (1-based), matid [nelem] (0-based),
 *   matparams [nmat][8], loc [nelem][3*nen], neq = number of equations of the numbering
 *   loc refers to (length of the global vectors the set scatters into). */
int  ob200_elemset_create(ob200_context *ctx, int etype, int64_t nnode, const double *coords,
                          int64_t nelem, const int32_t *conn, const int32_t *matid,
                          int32_t nmat, const double *matparams, const int32_t *loc, int32_t neq,
                          int on_device, ob200_elemset **out);
/* The same set from NODAL equation numbers nodeeq [nnode][3] (1-based, 0 = prescribed; dofs D_u, D_v, D_w of every
 * node -- DofManager::giveLocationArray, src/core/dofmanager.C) instead of per-element location arrays:
 * loc[e][3a+i] = nodeeq[conn[e][a]][i] is formed on the device, as Element::giveLocationArray (src/core/element.C)
 * forms it on the host.  12 bytes per node cross the bus instead of 12 bytes per element node. */
int  ob200_elemset_create_nodal(ob200_context *ctx, int etype, int64_t nnode, const double *coords,
                                int64_t nelem, const int32_t *conn, const int32_t *matid,
                                int32_t nmat, const double *matparams, const int32_t *nodeeq, int32_t neq,
                                int on_device, ob200_elemset **out);
void ob200_elemset_destroy(ob200_elemset *S);
int64_t ob200_elemset_size(const ob200_elemset *S);
/* element matrices Ke [nelem][nd*nd] (computeStiffnessMatrix, TangentStiffness) */
int  ob200_elemset_stiffness(ob200_elemset *S, double *Ke, int on_device);
/* element internal force vectors fe [nelem][nd] for nodal displacements u [nnode][3]
 * (giveInternalForcesVector, useUpdatedGpRecord = 0); updates the temp material state.
 * fe and gp_strain / gp_stress [nelem*ngp][6] are optional (NULL to skip). */
int  ob200_elemset_internal_forces(ob200_elemset *S, const double *u, double *fe,
                                   double *gp_strain, double *gp_stress, int on_device);
/* bind the set to a matrix: precomputes the element -> CSR slot map */
int  ob200_elemset_bind(ob200_elemset *S, ob200_csr *A);
/* fused: A += sum_e Ke scattered through loc (EngngModel::assemble with TangentAssembler) */
int  ob200_elemset_assemble_stiffness(ob200_elemset *S, ob200_csr *A);
/* fused: f[neq] += sum_e fe scattered through loc (assembleVector with InternalForceAssembler).
 * ebe_norm2 (HOST pointer, 3 doubles, may be NULL) receives the element-by-element squared
 * norms per dof id u,v,w -- the eNorms of EngngModel::assembleVector (engngm.C:1108-1133)
 * that NRSolver::checkConvergence scales the force error with (nrsolver.C:752-760). */
int  ob200_elemset_assemble_internal_forces(ob200_elemset *S, const double *u, double *f,
                                            double *ebe_norm2, int on_device);
/* f[neq] += sum_e Ke * du_e for nodal increments du [nnode][3]
 * (StaticStructural::assembleExtrapolatedForces, staticstructural.C:255) */
int  ob200_elemset_assemble_extrapolated_forces(ob200_elemset *S, const double *du, double *f, int on_device);
/* MaterialStatus::updateYourself for every Gauss point: temp -> committed */
int  ob200_elemset_commit(ob200_elemset *S);
/* raw MisesMat state [nelem*ngp][29] (tests / restart) */
int  ob200_elemset_get_state(ob200_elemset *S, double *state, int on_device);
int  ob200_elemset_set_state(ob200_elemset *S, const double *state, int on_device);

/* ---- SparseLinearSystemNM "cudacg" (src/core/iml/imlsolver.C:101-146, iml/cg.h) ------- */
/* Preconditioned CG exactly as the IML++ template: x is the initial guess on entry and
 * the solution on exit; *iters = iterations performed, *resid = ||r||/||b|| reached.
 * Returns 0 converged (CR_CONVERGED), 1 not converged within max_iter (CR_DIVERGED_ITS),
 * negative on error. */
int  ob200_cg_solve(ob200_csr *A, const double *b, double *x, int precond, int max_iter, double tol,
                    int *iters, double *resid, int on_device);

/* ---- multi-GPU: element partitions, shared-node halo exchange over NCCL --------------- */
/* nccl_unique_id is the 128-byte ncclUniqueId created by rank 0 (ob200_comm_unique_id)
 * and distributed by the caller (torch.distributed / MPI). */
int  ob200_comm_unique_id(void *id128);
int  ob200_comm_create(ob200_context *ctx, int nranks, int rank, const void *id128, ob200_comm **out);
void ob200_comm_destroy(ob200_comm *c);
/* Describe the shared dofs of this partition: for each of nneigh neighbour ranks the
 * local equation numbers (0-based) shared with it, in an order both sides agree on
 * (ascending global dof id); owned[neq] = 1 where this rank owns the dof (each shared dof
 * is owned by exactly one rank -- used to count it once in dot products). */
int  ob200_comm_set_halo(ob200_comm *c, int32_t neq, int nneigh, const int32_t *neigh_rank,
                         const int64_t *neigh_offset /* [nneigh+1] */, const int32_t *shared_eq,
                         const uint8_t *owned);
/* Peer-memory transport for the ranks of one NVLink/NVSwitch node (optional; NCCL is used without it):
 * every rank exports a mailbox in its HBM as a 64-byte CUDA IPC handle (cap = most dofs any two
 * ranks share), the caller gathers the handles of all ranks (torch.distributed / MPI_Allgather, as
 * OOFEM's ProblemCommunicator would) and hands the nranks x 64 bytes to _open.  From then on the
 * halo sum and the CG reductions are written straight into the peers' memory by the kernels. */
int  ob200_comm_p2p_export(ob200_comm *c, int64_t cap, void *handle64);
int  ob200_comm_p2p_open(ob200_comm *c, const void *handles);
int  ob200_comm_p2p_enabled(const ob200_comm *c);
int  ob200_comm_p2p_disable(ob200_comm *c);   /* all ranks, if _open failed on any of them: NCCL stays in use */
/* y <- y + contributions of the neighbours for shared dofs (OOFEM: updateSharedDofManagers) */
int  ob200_comm_exchange_add(ob200_comm *c, double *y_dev);
/* distributed PCG: A is the local sub-assembled matrix of this partition, b must already
 * be summed over partitions on shared dofs (fully assembled, consistent on all sharers). */
int  ob200_cg_solve_dist(ob200_csr *A, ob200_comm *c, const double *b, double *x, int precond,
                         int max_iter, double tol, int *iters, double *resid, int on_device);

#ifdef __cplusplus
}
#endif
#endif /* OOFEM_B200_H */
