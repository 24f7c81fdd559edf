// Host-side objects behind the opaque C handles.
#pragma once
#include "common.cuh"

namespace ob200 {
struct MatParams;

// what the element kernels see
struct ElemSetView {
    const double *coords;      // [nnode][3]
    const int32_t *conn;       // [nelem][nen], 1-based
    const int32_t *matid;      // [nelem], 0-based
    const MatParams *mat;      // [nmat]
    const int32_t *loc;        // [nelem][nd], 1-based, 0 = prescribed
    double *state;             // MisesMat state, [nstate / 32][29][32] (element_device.cuh: MisesStateRef), or nullptr
    int64_t nstate;            // nelem * ngp
    const double *exyz;        // [nelem][nen*3] vertex coordinates gathered per element (LSpace gather path) or nullptr
};
} // namespace ob200

constexpr int kCsrPad = 8;

struct ob200_csr {
    ob200_context *ctx = nullptr;
    int32_t neq = 0;
    int64_t nnz = 0;
    int64_t version = 0;
    // identity of the sparsity structure: a process-wide counter value taken by every build_structure (never reused,
    // so an element set bound to an earlier structure -- or to a destroyed matrix at the same address -- rebinds)
    int64_t structure_version = 0;
    ob200::DevBuf< int32_t > rowptr, colind;     // colind / val padded by kCsrPad entries (16-byte TMA granules)
    ob200::DevBuf< double > val;
    // streamed SpMV (spmv.cuh): per-chunk {first row, first entry}
    ob200::DevBuf< int2 > chunks;
    int32_t nchunks = 0;
    int32_t maxrow = 0;                           // longest row
    // SparseMtrx::zero() is lazy: the memset is skipped when the next writer overwrites every entry
    // (the gather assembly does); any other reader/writer materialises it first (csr_materialize)
    bool zero_pending = false;
    // blocked index of the SpMV (spmv_block.cuh), derived from rowptr/colind on first use
    bool blk_tried = false, blk_ok = false;
    int32_t nrb = 0, nbchunks = 0;
    int64_t nblk = 0;
    ob200::DevBuf< uint2 > bw;
    ob200::DevBuf< int4 > rbdesc, rbdesc_flag, bchunks;
    // distributed product (spmv.cuh MODE 2): copy of rowptr with the sign bit set on rows shared with other
    // partitions, valid for the halo description `flag_route` it was built from
    ob200::DevBuf< int32_t > rowptr_flag;
    const int32_t *flag_route = nullptr;
    // CG work vectors (allocated on first solve)
    ob200::DevBuf< double > work;
    ob200::DevBuf< double > diag;
    int64_t diag_version = -1;
};

// Everything the assembly kernels derive from the mesh and its equation numbers (built at create) and from the sparsity
// structure of the bound matrix (built at bind).  Kept apart from the element set so that it can outlive it: the context
// caches the schedule of the last destroyed set, and a set created from identical arrays adopts it (ob200_elemset_create).
struct ob200_sched {
    bool have_mesh = false, have_bind = false;
    int64_t bind_structure = -1;                   // structure_version of the matrix the bind part belongs to
    unsigned long long key[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };   // content hashes + sizes of the arrays it was built from
    ob200::DevBuf< double > exyz;
    ob200::DevBuf< int32_t > slot;
    bool slot_built = false;
    // owner-computes ("gather") assembly, assemble_gather.cu: node -> element incidence, built at create
    int32_t maxval = 0;                            // largest number of elements around a node
    int64_t nvisit = 0;                            // nelem * nen
    ob200::DevBuf< int32_t > ninc_start, ninc, ninc_node, nodeeq;
    ob200::DevBuf< int32_t > row_vis;              // [nelem][8]: position of the (element, local node) incidence in ninc
    bool eq_unique = false;                        // no equation belongs to two nodal dofs: vectors can be assembled owner-computes
    // ... and what depends on the bound matrix
    bool gather_ok = false, covers_all = false;
    int32_t maxblk = 0, ngroups = 0;
    ob200::DevBuf< unsigned char > pos, nblk, vu;
    ob200::DevBuf< unsigned short > blk;
    ob200::DevBuf< int2 > gtab;
    // cluster assembly (assemble_cluster.cu): schedule built per bound matrix
    bool cluster_ok = false;
    int32_t nclusters = 0;
    int64_t cl_nrec = 0;
    ob200::DevBuf< unsigned char > ebidx, nloc, npar, cl_recs, cl_steps;
    ob200::DevBuf< unsigned short > nbase;
    ob200::DevBuf< int32_t > cnodes, ncl, cl_begin, cl_step;
    // node-row assembly of LTRSpace (assemble_tet.cu): equation -> (node, component), per-element gradients / volume / Lame
    bool rows_ok = false, tet_fast = false;        // tet_fast: the table-driven kernel (per-node tables in row_desc / row_vtab)
    bool strips_ok = false;                        // LSpace with a general tangent: element strips (assemble_strips.cu) on the node-block schedule
    ob200::DevBuf< int32_t > eqnode;
    ob200::DevBuf< double > trec;
    ob200::DevBuf< int32_t > row_tstart, row_desc;             // strip assembly: per-node descriptors (two int4 per node) ...
    ob200::DevBuf< unsigned char > row_vtab;                   // ... and column tables (576 B per chunk of eight incidences)
};

struct ob200_elemset : ob200_sched {
    ob200_context *ctx = nullptr;
    int etype = 0, nen = 0, ngp = 0, nd = 0;
    int64_t nnode = 0, nelem = 0;
    int32_t nmat = 0, neq = 0;
    bool has_state = false;
    ob200::DevBuf< double > coords, mat, state;
    ob200::DevBuf< double > tangent;               // [nelem][36] material tangents of the LTRSpace node-row assembly (sets with a MisesMat)
    ob200::DevBuf< double > fvis;                  // [nvisit][3] nodal forces of the elements in incidence order (owner-computes vector assembly)
    ob200::DevBuf< double > kebuf;                 // [nvisit][3][24] element-matrix strips of the strip assembly, in incidence order
    ob200::DevBuf< int32_t > conn, matid, loc;
    ob200_csr *bound = nullptr;
    int64_t bound_version = -1;
    bool all_isole = true;
    bool loc_pending = false;                      // loc is still travelling on the context's copy stream
    int64_t neq_hint() const { return neq; }
    ob200::ElemSetView view() const
    {
        return ob200::ElemSetView{ coords.p, conn.p, matid.p, (const ob200::MatParams *) mat.p, loc.p,
                                   state.p, nelem * ngp, exyz.p };
    }
};

// bump the matrix version after its values changed (SparseMtrx::version)
void ob200_csr_touch(ob200_csr *A);
// perform a pending lazy zero()
int ob200_csr_materialize(ob200_csr *A);

namespace ob200 {
// make the main stream wait for the upload of loc (no-op once done)
inline int elemset_await_loc(ob200_elemset *S)
{
    if ( S->loc_pending ) {
        OB_CUDA( cudaStreamWaitEvent(S->ctx->stream, S->ctx->copy_event, 0) );
        S->loc_pending = false;
    }
    return OB200_OK;
}
int gather_prepare_mesh(ob200_elemset *S);                       // at create: incidence, nodeeq
int gather_bind(ob200_elemset *S, ob200_csr *A);                 // at bind: block schedule, group table
int gather_assemble_lspace(ob200_elemset *S, ob200_csr *A);      // the kernel
int cluster_bind(ob200_elemset *S, ob200_csr *A);                // assemble_cluster.cu: cluster schedule (after gather_bind)
int cluster_assemble_lspace(ob200_elemset *S, ob200_csr *A);     // assemble_cluster.cu: the kernel
int tet_bind(ob200_elemset *S, ob200_csr *A);                    // assemble_tet.cu: checks of the LTRSpace node-row assembly
int tet_assemble_ltrspace(ob200_elemset *S, ob200_csr *A);       // assemble_tet.cu: the kernel
int strips_element_matrices(ob200_elemset *S, double *Ke, const int32_t *vis);       // assemble_strips.cu: LSpace element matrices, general tangent (FP64 DMMA)
int strips_bind(ob200_elemset *S, ob200_csr *A);                 // assemble_strips.cu: per-node descriptors and column tables (after gather_bind)
int strips_assemble_lspace(ob200_elemset *S, ob200_csr *A);      // assemble_strips.cu: element matrices + rows from strips
}
