#!/usr/bin/env python3
"""bench.py -- elements assembled/s + PCG iterations/s on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[1], the synthetic structured 1M-hex LSpace
isotropic linear-elastic cantilever beam (250 x 64 x 64 = 1,024,000 elements, ~3.17M equations,
~250M non-zeros) per GPU; with N > 1 the beam is N times longer and element-partitioned into N
slabs (weak scaling), shared-plane halo exchange and dot-product all-reduce over NCCL.

One "step" = one pass of the hot path: zero the matrix, assemble all element stiffness
matrices into the CSR matrix (fused LSpace kernel), then `--cg-iters` PCG iterations
(SpMV + axpy + dot, diagonal preconditioner) on that matrix.
  value            elements assembled/s, inputs resident in HBM (whole job, all ranks)
  pcg_iters_per_s  PCG iterations/s in the same step
  e2e              same two numbers through the host-pointer C ABI: mesh arrays and vectors
                   start in pinned host memory, H2D/D2H copies inside the timed region
  roofline         dominant kernel (the CG SpMV) against the measured HBM copy bandwidth
  cpu_baseline     the unmodified reference (oracle/_ref/oofem_bench) on the box's host cores,
                   on a bounded sample of the same workload
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------

def slab_problem(nx, ny, nz, rank, nranks, etype="lspace"):
    """Local mesh of rank `rank`: slab [rank*nx, (rank+1)*nx] of a beam nranks*nx long
    (oofem_b200.partition.slab_partition), clamped at x = 0, with its own equation numbering."""
    from oofem_b200 import meshgen, partition
    part = partition.slab_partition(nx, ny, nz, rank, nranks, etype=etype)
    fixed_mask = np.zeros((part.coords.shape[0], 3), dtype=bool)
    plane = (ny + 1) * (nz + 1)
    if rank == 0:
        fixed_mask[:plane] = True                       # clamp x = 0
    nodeeq, neq = meshgen.equation_numbers(part.coords.shape[0], fixed_mask)
    loc = meshgen.location_arrays(part.conn, nodeeq)
    return dict(coords=part.coords, conn=part.conn, loc=loc, neq=neq, nodeeq=nodeeq, part=part, plane=plane)


def halo_arrays(pb, rank, nranks):
    from oofem_b200 import partition
    return partition.halo_arrays(pb["part"], pb["nodeeq"], pb["neq"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------

def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(nx, ny, nz, etype="lspace"):
    if etype == "ltrspace":
        return (f"synthetic structured {nx}x{ny}x{nz} cells x 6 = {6 * nx * ny * nz} LTRSpace tetrahedra, isotropic linear elastic "
                f"cantilever per GPU, FP64 PCG (BASELINE.json configs[2])")
    return (f"synthetic structured {nx}x{ny}x{nz} = {nx * ny * nz} hex LSpace isotropic linear elastic "
            f"cantilever per GPU, FP64 PCG (BASELINE.json configs[1])")


def reference_sample(dims, cg_iters, repeats, warmup=0, sample=0, prefer_omp=True):
    """Time the UNMODIFIED reference (oracle/_ref/oofem_bench[_omp]) on the mesh `dims`; with sample > 0 every
    repeat is a bounded sample (a window of `sample` elements + cg_iters CG iterations) of the work on that mesh."""
    from oofem_b200 import meshgen
    from oofem_b200.inputfile import DirichletBC, Material, NodalLoad, Problem, write_input
    nx, ny, nz = dims
    exes = [os.path.join(ROOT, "oracle", "_ref", n) for n in (("oofem_bench_omp",) if prefer_omp else ()) + ("oofem_bench",)]
    exe = next((e for e in exes if os.path.exists(e)), None)
    if exe is None:
        return None
    h = 1.0 / ny
    coords, conn = meshgen.hex_beam(nx, ny, nz, nx * h, 1.0, nz * h)
    fixed, tip = meshgen.cantilever_bcs(coords, nx * h)
    with tempfile.TemporaryDirectory() as td:
        pb = Problem(title="bench sample", outfile=os.path.join(td, "s.out"), engng="linearstatic",
                     params=dict(nsteps=1, lstype=1, smtype=2, lstol=0.5, lsiter=1000, lsprecond=1), coords=coords,
                     elem_type="lspace", conn=conn, elem_mat=np.zeros(conn.shape[0], np.int32),
                     materials=[Material("isole", 210e3, 0.3)])
        pb.ltfs[1] = ("const", 1.0)
        pb.bcs.append(DirichletBC([1, 2, 3], [0.0, 0.0, 0.0], 1, fixed))
        pb.loads.append(NodalLoad([3], [-1.0], 1, tip))
        fn = os.path.join(td, "s.in")
        write_input(fn, pb)
        env = dict(os.environ)
        # all host threads, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for N > 1)
        env["OMP_NUM_THREADS"] = str(host_cores())
        env.pop("OMP_PROC_BIND", None)
        r = subprocess.run([exe, fn, str(cg_iters), str(repeats), str(warmup), str(sample)], cwd=td, capture_output=True,
                           text=True, env=env)
        if r.returncode:
            sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode or not line:
        return None
    out = json.loads(line[-1])
    out["exe"] = os.path.basename(exe)
    return out


def run_reference(args):
    """Reference arm: the UNMODIFIED reference (oracle/_ref, its own _OPENMP build on all host threads) on the SAME
    mesh as our arm (one GPU's share of the weak-scaling workload).  Parsing the 1M-element input, building CompCol
    and one full EngngModel::assemble are done once outside the steps; each of the K timed steps is a bounded sample
    of the work on that mesh (a window of REF_SAMPLE_ELEMS elements through the loop body of EngngModel::assemble and
    REF_SAMPLE_CG IML CG iterations on the full matrix), so that the whole run ends within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dims = (args.nx, args.ny, args.nz)
    nel = dims[0] * dims[1] * dims[2]
    sample = min(args.ref_sample_elems, nel)
    cg_sample = min(args.ref_sample_cg, args.cg_iters)
    t0 = time.time()
    res = reference_sample(dims, cg_sample, args.steps, args.warmup, sample if sample < nel else 0)
    if res is None:      # reference binary not built here: time the C port of the same algorithm on a small mesh
        res = oracle_sample((40, 20, 20), cg_sample)
        kind = "port"
        same = False
    else:
        kind = "reference"
        same = True
    ms = (res["t_assemble_s"] + res["t_cg_s"]) * 1e3
    nnz = res["nnz"]
    sample_txt = (f"mesh {dims[0]}x{dims[1]}x{dims[2]} = {res['nelem']} LSpace elements, neq {res['neq']}, nnz {nnz} (the "
                  f"full configuration); per step: {res.get('sample_elements', res['nelem'])} consecutive elements through "
                  f"TangentAssembler::matrixFromElement + CompCol::assemble (loop body of EngngModel::assemble, its OpenMP "
                  f"pragmas) and {cg_sample} IML CG iterations (DiagPreconditioner) on the full matrix; the reference's own "
                  f"EngngModel::assemble over all elements, timed once: {res.get('t_assemble_full_s', float('nan')):.2f} s = "
                  f"{res.get('elements_per_s_full', float('nan')):.0f} elements/s; CompCol::buildInternalStructure "
                  f"{res.get('t_structure_s', float('nan')):.1f} s, input parse {res.get('t_parse_s', float('nan')):.1f} s "
                  f"(both outside the steps)") if same else \
                 f"oracle port on 40x20x20 = {res['nelem']} elements (reference binary absent)"
    line = {
        "impl": "reference", "metric": "elements assembled/s", "value": res["elements_per_s"], "unit": "elements/s",
        "pcg_iters_per_s": res["cg_iters_per_s"], "pcg_nnz_iters_per_s": res["cg_iters_per_s"] * nnz,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(*dims) if same else "bounded 40x20x20 sample of " + workload_name(*dims),
                   "nelem_per_gpu": res["nelem"], "neq_per_gpu": res["neq"], "nnz_per_gpu": nnz,
                   "cg_iters_per_step": args.cg_iters, "precond": "diag"},
        "cpu_baseline": {"value": res["elements_per_s"], "unit": "elements/s", "pcg_iters_per_s": res["cg_iters_per_s"],
                         "cores": res.get("threads", 1), "kind": kind, "sample": sample_txt},
        "e2e": {"value": res["elements_per_s"], "unit": "elements/s", "pcg_iters_per_s": res["cg_iters_per_s"],
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_detail": {k: res.get(k) for k in ("t_parse_s", "t_structure_s", "t_assemble_full_s", "elements_per_s_full",
                                                     "sample_elements", "t_assemble_s", "t_assemble_best_s", "t_cg_s",
                                                     "t_cg_best_s", "cg_iters", "threads", "exe")},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def oracle_sample(dims, cg_iters):
    from oofem_b200 import meshgen
    from oracle import oracle as orc
    nx, ny, nz = dims
    coords, conn = meshgen.hex_beam(nx, ny, nz)
    fixed, _ = meshgen.cantilever_bcs(coords, float(nx) / ny)
    mask = np.zeros((coords.shape[0], 3), bool)
    mask[fixed - 1] = True
    nodeeq, neq = meshgen.equation_numbers(coords.shape[0], mask)
    loc = meshgen.location_arrays(conn, nodeeq)
    colptr, rowind = orc.compcol_build(loc, neq)
    mp = np.array([[1, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    t0 = time.time()
    Ke = orc.batch_stiffness(orc.LSPACE, conn, coords, np.zeros(conn.shape[0], np.int32), mp)
    val = orc.compcol_assemble(loc, Ke, colptr, rowind)
    t_asm = time.time() - t0
    b = np.ones(neq)
    t0 = time.time()
    orc.cg(colptr, rowind, val, b, precond=1, max_iter=cg_iters, tol=1e-300)
    t_cg = time.time() - t0
    return dict(nelem=conn.shape[0], neq=neq, nnz=int(rowind.size), threads=1, t_assemble_s=t_asm, t_cg_s=t_cg,
                elements_per_s=conn.shape[0] / t_asm, cg_iters_per_s=cg_iters / t_cg)


# ---------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------

def pin_to_gpu_numa_node(device):
    """Restrict this rank to the host cores NVML lists as local to its GPU, so that the pinned host buffers of the end-to-end
    leg (first touch) and the copy threads sit on the GPU's NUMA node: with eight ranks on one host the uploads otherwise
    cross the socket interconnect.  Returns the number of cores, or None when NVML gives no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        allowed = set(os.sched_getaffinity(0))
        cores = [c for c in cores if c in allowed]
        if cores:
            os.sched_setaffinity(0, cores)
            return len(cores)
    except Exception:
        pass
    return None


def run_ours(args):
    import torch
    from oofem_b200 import capi
    from oofem_b200.capi import check, lib, ptr
    from oofem_b200.elements import ElementSet
    from oofem_b200.linsolver import CudaCG
    from oofem_b200.sparsemtrx import CudaCSR
    import ctypes as C

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: native libraries that print there (NCCL's version banner)
    # are sent to stderr for the duration of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = capi.Context(local)
    dev = torch.device("cuda", local)
    nx, ny, nz = args.nx, args.ny, args.nz
    if args.scaling == "strong":
        # the beam of --nx elements is cut into `world` x-slabs (a remainder of nx / world is dropped: say so in config)
        nx = max(1, args.nx // world)
    etype = args.etype
    pb = slab_problem(nx, ny, nz, rank, world, etype)
    nelem, neq = pb["conn"].shape[0], pb["neq"]
    matparams = np.array([[capi.MAT_ISOLE, 210e3, 0.3, 0, 0, 0, 0, 0]], dtype=np.float64)
    matid = np.zeros(nelem, np.int32)

    # ---- resident inputs (value) ---------------------------------------------------
    t = lambda a: torch.as_tensor(a, device=dev)
    d_coords, d_conn, d_loc, d_matid = t(pb["coords"]), t(pb["conn"]), t(pb["loc"]), t(matid)
    torch.cuda.synchronize()
    S = ElementSet(ctx, etype, d_coords, d_conn, d_matid, matparams, d_loc, neq)
    A = CudaCSR(ctx)
    t0 = time.time()
    A.buildInternalStructure(d_loc, neq)
    ctx.sync()
    t_structure = time.time() - t0
    S.bind(A)
    nnz = A.giveNumberOfNonzeros()
    comm = None
    if world > 1:
        from oofem_b200.comm import Comm
        comm = Comm.from_torch_distributed(ctx, dev)
        comm.set_halo(neq, *halo_arrays(pb, rank, world))
    solver = CudaCG(ctx, comm).initializeFrom(dict(lstol=0.0, lsiter=args.cg_iters, lsprecond=1))
    b = torch.ones(neq, dtype=torch.float64, device=dev)
    x = torch.zeros(neq, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()

    def step_resident(ev=None):
        if ev:
            ev[0].record(ext)
        A.zero()
        S.assembleStiffness(A)
        if ev:
            ev[1].record(ext)
        x.zero_()
        torch.cuda.current_stream().synchronize()
        solver.solve(A, b, x)
        if ev:
            ev[2].record(ext)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = ctx.launches
    sampler = ClockSampler(local)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    for k in range(args.steps):
        step_resident(evs[k])
    barrier()
    wall = time.time() - wall0
    launches = ctx.launches - launches0
    # per-kernel durations (roofline, shares): a second pass of the same K steps with two CUDA events around
    # every launch -- inside the timed pass those events cost 3.5 % of the PCG rate
    ctx.profile_reset()
    ctx.set_profiling(not args.no_kernel_events)
    for k in range(args.steps):
        step_resident()
    barrier()
    clocks = sampler.stop()
    ctx.set_profiling(False)
    prof = ctx.profile_report()
    t_asm = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps          # ms
    t_cg = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    tt = torch.tensor([t_asm, t_cg, t_asm + t_cg], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_asm, t_cg, t_step = [float(v) for v in tt.cpu()]

    # ---- end to end through the host-pointer C ABI ---------------------------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    # the mesh crosses the bus as OOFEM holds it: coordinates, connectivity, material numbers and the equation numbers of the
    # NODAL dofs; the per-element location arrays are formed on the device (ob200_elemset_create_nodal), as
    # Element::giveLocationArray forms them on the host
    h_coords, h_conn, h_nodeeq, h_matid = pin(pb["coords"]), pin(pb["conn"]), pin(pb["nodeeq"].astype(np.int32)), pin(matid)
    h_b, h_x = pin(np.ones(neq)), pin(np.zeros(neq))
    h2d_asm = h_coords.nbytes + h_conn.nbytes + h_nodeeq.nbytes + h_matid.nbytes + matparams.nbytes
    h2d_cg, d2h_cg = h_b.nbytes + h_x.nbytes, h_x.nbytes

    def step_e2e():
        t0 = time.perf_counter()
        S2 = ElementSet(ctx, etype, h_coords, h_conn, h_matid, matparams, None, neq, nodeeq=h_nodeeq)      # H2D of the mesh
        A.zero()
        S2.assembleStiffness(A)                                                           # slot map + fused assembly
        ctx.sync()
        t1 = time.perf_counter()
        h_x[:] = 0.0
        solver.solve(A, h_b, h_x)                                                          # H2D b,x ; D2H x
        t2 = time.perf_counter()
        S2.close()
        return t1 - t0, t2 - t1

    e2e_steps = max(1, min(args.steps, 3))
    if args.no_e2e:                 # profiling runs (ncu) only; never used for a reported line
        ts = [(float("nan"), float("nan"))]
        cold = ts[0]
        e2e_steps = 1
    else:
        cold = step_e2e()           # first upload of this mesh: builds the assembly schedule (kept by the context afterwards)
        barrier()
        ts = [step_e2e() for _ in range(e2e_steps)]
        barrier()
    e_asm = sum(a for a, _ in ts) / e2e_steps
    e_cg = sum(c for _, c in ts) / e2e_steps
    te = torch.tensor([e_asm, e_cg], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e_asm, e_cg = [float(v) for v in te.cpu()]

    # ---- multi-GPU parity (checker leg, after every timed region): small partitioned problems against the serial
    # oracle -- LSpace and LTRSpace, x-slabs and box partitions (dofs shared by 4 / 8 ranks), both transports
    parity, parity_ok = None, True
    if dist and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import dist_cases
        S.close()
        cases = dist_cases.run_all(ctx, dev)
        parity_ok = all(c["ok"] for c in cases)
        parity = {"ok": parity_ok, "tolerance": {"spmv_relerr": dist_cases.SPMV_TOL, "u_relerr": dist_cases.U_TOL},
                  "spmv_relerr": max(c["spmv_relerr"] for c in cases), "u_relerr": max(c["u_relerr"] for c in cases),
                  "max_sharers": max(c["max_sharers"] for c in cases),
                  "transport": sorted(set(c["transport_used"] for c in cases)),
                  "oracle": "serial C restatement of the reference (oracle/oofem_oracle.c) on the unpartitioned mesh",
                  "cases": [{k: c[k] for k in ("etype", "partition", "transport", "transport_used", "ok", "spmv_relerr", "u_relerr",
                                               "iters", "iters_equal_on_all_ranks", "max_sharers", "neq_global")} for c in cases]}

    if rank != 0:
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if parity_ok else 1

    # ---- roofline of the dominant kernel ---------------------------------------------
    peak, peak_src = load_peaks()
    # the CG product: fused p.q on one GPU (< 1 >), fused p.q + halo push over peer memory (< 2 >), plain (< 0 >, NCCL)
    spmv_name = max(("spmv_block_kernel< MODE >", "spmv_stream_kernel< 1 >", "spmv_stream_kernel< 2 >", "spmv_stream_kernel< 0 >"),
                    key=lambda k: prof.get(k, (0.0, 0))[0])
    layout = (C.c_int64 * 3)()
    check(lib().ob200_csr_spmv_layout(A.h, layout))
    blocked, nrb, nblk = int(layout[0]), int(layout[1]), int(layout[2])
    asm_name = next((k for k in ("lspace_cluster_kernel< false >", "lspace_cluster_kernel< true >",
                                 "lspace_gather_kernel< false >", "lspace_gather_kernel< true >",
                                 "lspace_stiffness_kernel< OUT_CSR >", "ltrspace_rows2_kernel< 0 >", "ltrspace_rows2_kernel< 1 >",
                                 "ltrspace_rows_kernel< 0 >", "ltrspace_rows_kernel< 1 >",
                                 "ltrspace_stiffness_kernel< OUT_CSR >") if k in prof), "lspace_cluster_kernel< false >")
    ms_spmv, n_spmv = prof.get(spmv_name, (0.0, 0))
    ms_asmk, n_asmk = prof.get(asm_name, (0.0, 0))
    # algorithmic bytes: SpMV reads val (8 B) + colind (4 B) per non-zero, and per
    # row rowptr (4 B), the operand entry once (8 B), writes the result (8 B)
    spmv_bytes = 12.0 * nnz + 20.0 * neq
    if blocked:
        # blocked index: 8 B of value per non-zero, 8 B per column block, 16 B per row block,
        # x once and y once per row
        spmv_bytes = 8.0 * nnz + 8.0 * nblk + 16.0 * nrb + 16.0 * neq
    spmv_dur = ms_spmv / max(n_spmv, 1) * 1e-3
    spmv_gbs = spmv_bytes / spmv_dur / 1e9 if spmv_dur > 0 else 0.0
    # assembly: per element conn (32 B) + matid (4 B) + location array (96 B), coordinates once per node
    # (24 B), every matrix value written once (8 B per non-zero)
    nen = pb["conn"].shape[1]
    asm_bytes = nelem * (4.0 * nen + 4.0 + 12.0 * nen) + pb["coords"].shape[0] * 24.0 + 8.0 * nnz
    asm_dur = ms_asmk / max(n_asmk, 1) * 1e-3
    asm_gbs = asm_bytes / asm_dur / 1e9 if asm_dur > 0 else 0.0
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic = {}
    full_size = (nx, ny, nz) == (250, 64, 64) and etype == "lspace"         # the captures were taken at this size
    spmv_traffic = traffic.get(spmv_name.split("<")[0].strip()) if full_size else None
    asm_traffic = traffic.get(asm_name.split("<")[0].strip()) if full_size else None
    kernel_share = {k: round(v[0] / (t_step * args.steps) , 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:8]}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and etype == "lspace":
        cpu = reference_sample((40, 20, 20), 20, 1, 0, 0)
        kind = "reference"
        if cpu is None:
            cpu = oracle_sample((40, 20, 20), 20)
            kind = "port"
    total_elems = nelem * world
    line = {
        "metric": "elements assembled/s", "value": total_elems / (t_asm * 1e-3), "unit": "elements/s",
        "pcg_iters_per_s": args.cg_iters / (t_cg * 1e-3),
        "pcg_nnz_iters_per_s": args.cg_iters / (t_cg * 1e-3) * float(nnz) * world,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step,
        "ms_assembly": t_asm, "ms_pcg": t_cg,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(nx, ny, nz, etype) if args.scaling == "weak" else
                               (f"strong scaling: one {nx * world}x{ny}x{nz} = {nx * world * ny * nz} hex LSpace beam cut into {world} x-slabs "
                                f"of {nx}x{ny}x{nz}, FP64 PCG (BASELINE.json configs[4] pattern at the configs[1] size)" if etype == "lspace" else
                                f"one {nx * world}x{ny}x{nz}-cell beam = {6 * nx * world * ny * nz} LTRSpace tetrahedra, element-partitioned into "
                                f"{world} x-slabs of {6 * nx * ny * nz}, isotropic linear elastic, FP64 PCG (BASELINE.json configs[2])"),
                   "etype": etype,
                   "nelem_per_gpu": nelem, "neq_per_gpu": neq, "nnz_per_gpu": int(nnz), "cg_iters_per_step": args.cg_iters,
                   "precond": "diag", "partition": f"{world} x-slabs, shared-plane halo" if world > 1 else "none",
                   "transport": ("peer memory (CUDA IPC mailboxes over NVLink)" if comm.p2p else "NCCL") if comm else "none",
                   "l2": "inputs larger than L2 (val+colind ~3 GB per pass vs 126 MB L2), no flush needed",
                   "structure_build_s": round(t_structure, 4), "host_cores_local_to_gpu": numa},
        "roofline": {"kernel": spmv_name, "bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s",
                     "frac": spmv_gbs / peak, "traffic": spmv_traffic, "peak_source": peak_src,
                     "traffic_source": "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per launch from the "
                                       "committed ncu --set full capture of this kernel at this size; not measured in this run",
                     "algorithmic_bytes_per_launch": spmv_bytes, "avg_launch_ms": spmv_dur * 1e3, "launches": n_spmv,
                     "timing": "CUDA events around every launch, second pass of the same steps right after the timed pass",
                     "index": ("blocked: %d row blocks, %d column blocks for %d non-zeros; plain CSR would read %.3f GB "
                               "(%.0f GB/s at this duration)" % (nrb, nblk, nnz, (12.0 * nnz + 20.0 * neq) / 1e9,
                                                                 (12.0 * nnz + 20.0 * neq) / spmv_dur / 1e9 if spmv_dur > 0 else 0.0))
                              if blocked else "CSR"},
        "roofline_assembly": {"kernel": asm_name, "bound": "hbm", "achieved": asm_gbs, "peak": peak, "unit": "GB/s",
                              "frac": asm_gbs / peak, "traffic": asm_traffic, "algorithmic_bytes_per_launch": asm_bytes,
                              "avg_launch_ms": asm_dur * 1e3, "launches": n_asmk,
                              "traffic_source": "profiles/traffic.json (committed ncu --set full capture at this size); not measured in this run",
                              "note": ("cluster assembly (assemble_cluster.cu): shared-memory / dependency-chain bound, see DESIGN.md 3.3"
                                       if etype == "lspace" else "node-row assembly (assemble_tet.cu), see DESIGN.md 3.3b")},
        "kernel_time_share": kernel_share,
        "e2e": {"value": total_elems / e_asm, "unit": "elements/s", "pcg_iters_per_s": args.cg_iters / e_cg,
                "h2d_bytes_per_step": int(h2d_asm + h2d_cg), "d2h_bytes_per_step": int(d2h_cg),
                "ms_assembly": e_asm * 1e3, "ms_pcg": e_cg * 1e3,
                "ms_assembly_first_upload": cold[0] * 1e3,
                "note": "every step uploads the mesh again (coordinates, connectivity, material numbers, nodal equation numbers: "
                        f"{h2d_asm / 1e6:.0f} MB H2D; the location arrays are formed on the device) and creates a new element set; the "
                        "assembly schedule of an unchanged mesh is recognised by content hash and reused (ms_assembly_first_upload = "
                        "the step that builds it, rank 0)"},
        "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": wall,
    }
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["elements_per_s"], "unit": "elements/s", "pcg_iters_per_s": cpu["cg_iters_per_s"],
                                "cores": cpu.get("threads", 1), "kind": kind,
                                "sample": f"{cpu['nelem']} LSpace elements (40x20x20 sample of the same beam) assembled by the "
                                          f"reference's EngngModel::assemble into CompCol; 20 IML CG iterations at nnz {cpu['nnz']}"}
    if world == 1 and not args.no_e2e_executable and not args.no_cpu_baseline and etype == "lspace":
        # OOFEM's own executable with the plugin against the unmodified reference executable, same generated input, to
        # convergence (outside every timed region above; wall-clock of whole processes)
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import e2e_executable
            line["e2e_executable"] = e2e_executable.measure((125, 32, 32), "1e-3")
        except Exception as ex:          # never let the optional leg take the bench line down
            line["e2e_executable"] = {"unavailable": repr(ex)}
    if parity is not None:
        line["parity"] = parity
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if not parity_ok:
        sys.stderr.write("bench.py: multi-GPU parity FAILED: %s\n" % json.dumps(parity))
    return 0 if parity_ok else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nx", type=int, default=250)
    ap.add_argument("--ny", type=int, default=64)
    ap.add_argument("--nz", type=int, default=64)
    ap.add_argument("--cg-iters", type=int, default=50)
    ap.add_argument("--etype", default="lspace", choices=["lspace", "ltrspace"],
                    help="ltrspace: BASELINE configs[2] (every cell of the beam split into six tetrahedra); the default line is lspace")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --nx x ny x nz elements per GPU (default, the driver's contract); strong: --nx in total, cut into N slabs")
    ap.add_argument("--ref-sample-elems", type=int, default=32000, help="reference arm: elements assembled per timed step")
    ap.add_argument("--ref-sample-cg", type=int, default=5, help="reference arm: CG iterations per timed step")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the multi-GPU parity cases after the timed region")
    ap.add_argument("--no-e2e-executable", action="store_true",
                    help="N = 1: skip the oofem_cuda -f vs oofem -f comparison on a generated 128k-hex input (about a minute)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (ncu profiling runs only)")
    ap.add_argument("--no-kernel-events", action="store_true", help="scratch: no per-kernel CUDA events in the timed region (no roofline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
