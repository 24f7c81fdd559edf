"""Reader / writer for the subset of OOFEM's text input format that the accelerated
structural path needs (data format on the caller side of the hot path).

Record grammar follows the reference's OOFEMTXTDataReader / OOFEMTXTInputRecord
(src/core/oofemtxtdatareader.C, src/core/oofemtxtinputrecord.C): one record per line,
'#' starts a comment line, keywords are case-insensitive, arrays are "n v1 .. vn",
ranges are "{a b (c d)}".  Supported records:

  engineering models   LinearStatic (src/sm/EngineeringModels/linearstatic.C:90-110),
                       StaticStructural / NonLinearStatic (Newton-Raphson, load control)
  dof managers         node <n> coords 3 x y z [bc 3 b1 b2 b3] [load k l1..lk]
  elements             lspace / ltrspace <n> nodes k ... [crossSect c] [mat m]
  cross section        SimpleCS <n> [material m] [set s]
  materials            IsoLE (E, n, tAlpha, d), MisesMat (E, n, sig0, H, omega_crit, a)
  boundary conditions  BoundaryCondition (dofs/values or prescribedvalue, set), NodalLoad
  functions            ConstantFunction, PiecewiseLinFunction
  sets                 Set <n> [nodes k ..] [noderanges {..}] [elements k ..] [elementranges {..}]

Solver selection keywords on the engineering-model record: ``lstype`` / ``smtype`` accept
the reference's integers (src/core/linsystsolvertype.h, sparsemtrxtype.h) and the two new
names this package adds, ``cudacg`` and ``cudacsr``.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

# enum values: reference ones, then ours appended after the last reference member
ST_DIRECT, ST_IML = 0, 1
ST_CUDACG = 9                      # after ST_PardisoProjectOrg = 8 (linsystsolvertype.h:56)
SMT_SKYLINE, SMT_COMPCOL = 0, 2
SMT_CUDACSR = 11                   # after SMT_DSS_unsym_LU = 10 (sparsemtrxtype.h:55)

ELEM_NEN = {"lspace": 8, "ltrspace": 4}


@dataclass
class Material:
    kind: str            # 'isole' | 'misesmat'
    E: float
    n: float
    d: float = 0.0
    talpha: float = 0.0
    sig0: float = 0.0
    H: float = 0.0
    omega_crit: float = 0.0
    a: float = 0.0

    def params(self) -> list:
        """[type, E, nu, sig0, H, omega_crit, a, 0] -- the material block the kernels take."""
        return [1.0 if self.kind == "isole" else 2.0, self.E, self.n, self.sig0, self.H,
                self.omega_crit, self.a, 0.0]


@dataclass
class DirichletBC:
    dofs: list
    values: list
    ltf: int
    nodes: np.ndarray          # 1-based node numbers


@dataclass
class NodalLoad:
    dofs: list
    components: list
    ltf: int
    nodes: np.ndarray


@dataclass
class Problem:
    title: str = ""
    outfile: str = "out.out"
    engng: str = "linearstatic"
    params: dict = field(default_factory=dict)
    coords: np.ndarray = None
    elem_type: str = "lspace"
    conn: np.ndarray = None
    elem_mat: np.ndarray = None          # 0-based material index per element
    materials: list = field(default_factory=list)
    bcs: list = field(default_factory=list)
    loads: list = field(default_factory=list)
    ltfs: dict = field(default_factory=dict)   # id -> ('const', c) | ('pwl', t[], f[])
    checks: list = field(default_factory=list)  # parsed #NODE check records

    # ---- derived data ------------------------------------------------------------
    def ltf_value(self, ltf_id: int, t: float) -> float:
        f = self.ltfs[ltf_id]
        if f[0] == "const":
            return f[1]
        ts, fs = f[1], f[2]
        # PiecewiseLinFunction::evaluateAtTime (src/core/piecewiselinfunction.C:51-83): linear
        # interpolation between dates; past the last date the last value is used.  (Before
        # the first date the reference returns dates[0] -- a quirk; we clamp to values[0].)
        return float(np.interp(t, ts, fs))

    def fixed_mask(self) -> np.ndarray:
        m = np.zeros((self.coords.shape[0], 3), dtype=bool)
        for bc in self.bcs:
            for d in bc.dofs:
                m[bc.nodes - 1, d - 1] = True
        return m

    def prescribed_values(self, t: float) -> np.ndarray:
        """u_prescribed[nnode,3] at time t (BoundaryCondition::give = value * ltf(t))."""
        u = np.zeros((self.coords.shape[0], 3))
        for bc in self.bcs:
            f = self.ltf_value(bc.ltf, t)
            for d, v in zip(bc.dofs, bc.values):
                u[bc.nodes - 1, d - 1] = v * f
        return u

    def nodal_load_vector(self, t: float) -> np.ndarray:
        """Reference load per node/dof at time t: sum of NodalLoad components * ltf(t)."""
        f = np.zeros((self.coords.shape[0], 3))
        for ld in self.loads:
            s = self.ltf_value(ld.ltf, t)
            for d, c in zip(ld.dofs, ld.components):
                np.add.at(f, (ld.nodes - 1, d - 1), c * s)
        return f

    def matparams(self) -> np.ndarray:
        return np.array([m.params() for m in self.materials], dtype=np.float64)


# ------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------

def _ranges(ids: np.ndarray) -> str:
    ids = np.asarray(ids, dtype=np.int64)
    if ids.size == 0:
        return "{}"
    out, s, p = [], int(ids[0]), int(ids[0])
    for v in ids[1:]:
        v = int(v)
        if v == p + 1:
            p = v
            continue
        out.append(f"({s} {p})" if p > s else f"{s}")
        s = p = v
    out.append(f"({s} {p})" if p > s else f"{s}")
    return "{" + " ".join(out) + "}"


def write_input(path: str, pb: Problem) -> None:
    """Write an OOFEM .in file the unmodified reference can run."""
    nn, ne = pb.coords.shape[0], pb.conn.shape[0]
    nbc = len(pb.bcs) + len(pb.loads)
    nset = 1 + nbc
    eng = {"linearstatic": "LinearStatic", "staticstructural": "StaticStructural",
           "nonlinearstatic": "NonLinearStatic"}[pb.engng]
    prm = " ".join(f"{k} {v}" for k, v in pb.params.items())
    L = [pb.outfile, pb.title or "generated by oofem_b200.inputfile", f"{eng} {prm} nmodules 0",
         "domain 3d", "OutputManager tstep_all dofman_all element_all",
         f"ndofman {nn} nelem {ne} ncrosssect {len(pb.materials)} nmat {len(pb.materials)} "
         f"nbc {nbc} nic 0 nltf {len(pb.ltfs)} nset {len(pb.materials) + nbc}"]
    for i, c in enumerate(pb.coords):
        L.append(f"node {i + 1} coords 3 {float(c[0])!r} {float(c[1])!r} {float(c[2])!r}")
    nen = ELEM_NEN[pb.elem_type]
    for e, c in enumerate(pb.conn):
        L.append(f"{pb.elem_type} {e + 1} nodes {nen} " + " ".join(str(int(v)) for v in c))
    for m in range(len(pb.materials)):
        L.append(f"SimpleCS {m + 1} material {m + 1} set {m + 1}")
    for m, mat in enumerate(pb.materials):
        if mat.kind == "isole":
            L.append(f"IsoLE {m + 1} d {mat.d!r} E {mat.E!r} n {mat.n!r} tAlpha {mat.talpha!r}")
        else:
            L.append(f"MisesMat {m + 1} d {mat.d!r} E {mat.E!r} n {mat.n!r} tAlpha {mat.talpha!r} "
                     f"sig0 {mat.sig0!r} H {mat.H!r} omega_crit {mat.omega_crit!r} a {mat.a!r}")
    k = 0
    nm = len(pb.materials)
    for bc in pb.bcs:
        k += 1
        L.append(f"BoundaryCondition {k} loadTimeFunction {bc.ltf} dofs {len(bc.dofs)} "
                 + " ".join(map(str, bc.dofs)) + f" values {len(bc.values)} "
                 + " ".join(repr(float(v)) for v in bc.values) + f" set {nm + k}")
    for ld in pb.loads:
        k += 1
        L.append(f"NodalLoad {k} loadTimeFunction {ld.ltf} dofs {len(ld.dofs)} "
                 + " ".join(map(str, ld.dofs)) + f" Components {len(ld.components)} "
                 + " ".join(repr(float(v)) for v in ld.components) + f" set {nm + k}")
    for i, f in sorted(pb.ltfs.items()):
        if f[0] == "const":
            L.append(f"ConstantFunction {i} f(t) {f[1]!r}")
        else:
            L.append(f"PiecewiseLinFunction {i} t {len(f[1])} " + " ".join(repr(float(v)) for v in f[1])
                     + f" f(t) {len(f[2])} " + " ".join(repr(float(v)) for v in f[2]))
    for m in range(nm):
        L.append(f"Set {m + 1} elementranges {_ranges(np.nonzero(pb.elem_mat == m)[0] + 1)}")
    k = 0
    for b in list(pb.bcs) + list(pb.loads):
        k += 1
        L.append(f"Set {nm + k} noderanges {_ranges(np.sort(b.nodes))}")
    with open(path, "w") as fh:
        fh.write("\n".join(L) + "\n")


# ------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------

class _Rec:
    """One input record: keyword lookup over whitespace tokens (OOFEMTXTInputRecord)."""

    def __init__(self, line: str):
        self.raw = line
        self.tok = re.findall(r"\{[^}]*\}|\"[^\"]*\"|\S+", line)
        self.low = [t.lower() for t in self.tok]

    def has(self, key):
        return key.lower() in self.low[1:]

    def _pos(self, key):
        try:
            return self.low.index(key.lower(), 1)
        except ValueError:
            raise KeyError(f"missing field '{key}' in record: {self.raw[:80]}")

    def num(self, key, default=None, typ=float):
        if not self.has(key):
            if default is None:
                raise KeyError(f"missing field '{key}' in record: {self.raw[:80]}")
            return default
        return typ(self.tok[self._pos(key) + 1])

    def arr(self, key, typ=float):
        p = self._pos(key)
        n = int(self.tok[p + 1])
        return [typ(v) for v in self.tok[p + 2:p + 2 + n]]

    def rng(self, key):
        body = self.tok[self._pos(key) + 1].strip("{}")
        out = []
        for m in re.finditer(r"\(\s*(\d+)\s+(\d+)\s*\)|(\d+)", body):
            if m.group(3):
                out.append(int(m.group(3)))
            else:
                out.extend(range(int(m.group(1)), int(m.group(2)) + 1))
        return out


def _solver_kw(v: str, names: dict) -> int:
    v = v.strip('"').lower()
    return names[v] if v in names else int(v)


def read_input(path: str) -> Problem:
    lines = []
    checks = []
    in_check = False
    with open(path) as fh:
        for ln in fh:
            s = ln.strip()
            if s.startswith("#%BEGIN_CHECK%"):
                in_check = True
                continue
            if s.startswith("#%END_CHECK%"):
                in_check = False
                continue
            if in_check and s.startswith("#NODE"):
                r = _Rec(s)
                checks.append(dict(tstep=r.num("tStep", typ=int), number=r.num("number", typ=int),
                                   dof=r.num("dof", typ=int), value=r.num("value"),
                                   tol=r.num("tolerance", 1e-6)))
            if s and not s.startswith("#"):
                lines.append(s)
    pb = Problem(outfile=lines[0], title=lines[1], checks=checks)
    eng = _Rec(lines[2])
    name = eng.low[0]
    if name not in ("linearstatic", "staticstructural", "nonlinearstatic"):
        raise ValueError(f"unsupported engineering model '{eng.tok[0]}'")
    pb.engng = name
    p = {"nsteps": eng.num("nsteps", 1, int)}
    if eng.has("lstype"):
        p["lstype"] = _solver_kw(eng.tok[eng._pos("lstype") + 1], {"cudacg": ST_CUDACG, "iml": ST_IML})
    if eng.has("smtype"):
        p["smtype"] = _solver_kw(eng.tok[eng._pos("smtype") + 1], {"cudacsr": SMT_CUDACSR, "compcol": SMT_COMPCOL})
    for k, t in (("lstol", float), ("lsiter", int), ("lsprecond", int), ("stype", int),
                 ("rtolf", float), ("rtolv", float), ("rtold", float), ("maxiter", int),
                 ("manrmsteps", int)):
        if eng.has(k):
            p[k] = eng.num(k, typ=t)
    pb.params = p
    idx = 3
    while not lines[idx].lower().startswith("ndofman"):
        idx += 1
    cnt = _Rec("x " + lines[idx])
    nn, ne = cnt.num("ndofman", typ=int), cnt.num("nelem", typ=int)
    ncs, nmat = cnt.num("ncrosssect", typ=int), cnt.num("nmat", typ=int)
    nbc, nltf = cnt.num("nbc", typ=int), cnt.num("nltf", typ=int)
    nset = cnt.num("nset", 0, int)
    idx += 1
    coords = np.zeros((nn, 3))
    node_bc, node_load = {}, {}
    for i in range(nn):
        r = _Rec(lines[idx + i])
        num = int(r.tok[1])
        c = r.arr("coords")
        coords[num - 1, :len(c)] = c
        if r.has("bc"):
            node_bc[num] = r.arr("bc", int)
        if r.has("load"):
            node_load[num] = r.arr("load", int)
    idx += nn
    etype = None
    conn, elem_cs, elem_mat_direct = [], [], []
    for i in range(ne):
        r = _Rec(lines[idx + i])
        t = r.low[0]
        if t not in ELEM_NEN:
            raise ValueError(f"unsupported element type '{r.tok[0]}'")
        if etype is None:
            etype = t
        elif etype != t:
            raise ValueError("mixed element types are not supported by the batched path")
        conn.append(r.arr("nodes", int))
        elem_cs.append(r.num("crosssect", 0, int))
        elem_mat_direct.append(r.num("mat", 0, int))
    idx += ne
    cs = {}
    for i in range(ncs):
        r = _Rec(lines[idx + i])
        cs[int(r.tok[1])] = dict(material=r.num("material", 0, int), set=r.num("set", 0, int))
    idx += ncs
    mats = {}
    for i in range(nmat):
        r = _Rec(lines[idx + i])
        kind = r.low[0]
        if kind == "isole":
            mats[int(r.tok[1])] = Material("isole", r.num("E"), r.num("n"), r.num("d", 0.0), r.num("talpha", 0.0))
        elif kind == "misesmat":
            mats[int(r.tok[1])] = Material("misesmat", r.num("E"), r.num("n"), r.num("d", 0.0), r.num("talpha", 0.0),
                                           r.num("sig0"), r.num("H", 0.0), r.num("omega_crit", 0.0), r.num("a", 0.0))
        else:
            raise ValueError(f"unsupported material '{r.tok[0]}'")
    idx += nmat
    bc_recs = {}
    for i in range(nbc):
        r = _Rec(lines[idx + i])
        bc_recs[int(r.tok[1])] = r
    idx += nbc
    for i in range(nltf):
        r = _Rec(lines[idx + i])
        kind = r.low[0]
        if kind == "constantfunction":
            pb.ltfs[int(r.tok[1])] = ("const", r.num("f(t)"))
        elif kind == "piecewiselinfunction":
            pb.ltfs[int(r.tok[1])] = ("pwl", np.array(r.arr("t")), np.array(r.arr("f(t)")))
        else:
            raise ValueError(f"unsupported function '{r.tok[0]}'")
    idx += nltf
    sets = {}
    for i in range(nset):
        r = _Rec(lines[idx + i])
        nodes, elems = [], []
        if r.has("nodes"):
            nodes += r.arr("nodes", int)
        if r.has("noderanges"):
            nodes += r.rng("noderanges")
        if r.has("elements"):
            elems += r.arr("elements", int)
        if r.has("elementranges"):
            elems += r.rng("elementranges")
        sets[int(r.tok[1])] = dict(nodes=np.array(sorted(set(nodes)), dtype=np.int32),
                                   elems=np.array(sorted(set(elems)), dtype=np.int32))
    # resolve element -> material
    mat_ids = sorted(mats)
    mat_index = {m: k for k, m in enumerate(mat_ids)}
    pb.materials = [mats[m] for m in mat_ids]
    elem_mat = np.zeros(ne, dtype=np.int32)
    for c, info in cs.items():
        if info["set"] and info["material"]:
            elem_mat[sets[info["set"]]["elems"] - 1] = mat_index[info["material"]]
    for e in range(ne):
        if elem_mat_direct[e]:
            elem_mat[e] = mat_index[elem_mat_direct[e]]
        elif elem_cs[e] and cs[elem_cs[e]]["material"]:
            elem_mat[e] = mat_index[cs[elem_cs[e]]["material"]]
    pb.coords, pb.elem_type, pb.conn, pb.elem_mat = coords, etype, np.array(conn, dtype=np.int32), elem_mat
    # boundary conditions and loads
    for num, r in bc_recs.items():
        kind = r.low[0]
        ltf = r.num("loadtimefunction", typ=int)
        nodes = sets[r.num("set", typ=int)]["nodes"] if r.num("set", 0, int) else np.zeros(0, np.int32)
        if kind == "boundarycondition":
            if r.has("dofs"):
                dofs = r.arr("dofs", int)
                vals = r.arr("values") if r.has("values") else [r.num("prescribedvalue", 0.0)] * len(dofs)
                if nodes.size:
                    pb.bcs.append(DirichletBC(dofs, vals, ltf, nodes))
            # old style: node records carry "bc 3 i j k" referring to this bc number
            v = r.num("prescribedvalue", 0.0) if r.has("prescribedvalue") else None
            for d in (1, 2, 3):
                nd = np.array([n for n, b in node_bc.items() if len(b) >= d and b[d - 1] == num], dtype=np.int32)
                if nd.size:
                    val = v if v is not None else (r.arr("values")[0] if r.has("values") else 0.0)
                    pb.bcs.append(DirichletBC([d], [val], ltf, nd))
        elif kind == "nodalload":
            comps = r.arr("components")
            dofs = r.arr("dofs", int) if r.has("dofs") else list(range(1, len(comps) + 1))
            nd = list(nodes) + [n for n, l in node_load.items() if num in l]
            if nd:
                pb.loads.append(NodalLoad(dofs, comps, ltf, np.array(nd, dtype=np.int32)))
        else:
            raise ValueError(f"unsupported boundary condition '{r.tok[0]}'")
    return pb
