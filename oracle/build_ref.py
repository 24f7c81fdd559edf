#!/usr/bin/env python3
"""Build the UNMODIFIED reference (OOFEM core + sm module, IML solvers) into oracle/_ref/oofem.

TEST INFRASTRUCTURE ONLY.  Nothing under oofem_b200/ may import or execute this.

The reference's own build system (cmake) is NOT run.  This recipe reads the source
lists out of the reference's CMakeLists.txt files (so that no file list is copied
into this repo), compiles those sources where they lie under /root/reference with
g++ and links one executable.  The only hand-supplied file is a 6-line oofemenv.h
config stub (what cmake's configure_file would emit for a static, non-exported
build).  Objects go to a scratch directory, the binary to oracle/_ref/.

Usage: python oracle/build_ref.py [--jobs N] [--openmp]
"""
import argparse, os, re, subprocess, sys, concurrent.futures as cf

REF = os.environ.get("OOFEM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

ON = {"USE_IML", "USE_SM"}          # every other option OFF (no TM/FM/MPM/PETSc/...)


def cmake_sources(cmakelists, target):
    """Resolve the source list of add_library(<target> ...): evaluate set()/list(APPEND)
    into variables, honouring if(USE_X)/else/endif, then expand ${var} recursively."""
    txt = open(cmakelists).read()
    txt = re.sub(r"#.*", "", txt)
    active = [True]
    var = {}
    for m in re.finditer(r"(\w+)\s*\(([^()]*(?:\([^()]*\)[^()]*)*)\)", txt):
        cmd, body = m.group(1).lower(), m.group(2)
        if cmd == "if":
            toks = body.split()
            if "OR" in toks:
                val = any(t in ON for t in toks if t != "OR")
            else:
                val = len(toks) == 1 and toks[0] in ON
            active.append(active[-1] and val)
        elif cmd == "elseif":
            active[-1] = False
        elif cmd == "else":
            active[-1] = active[-2] and not active[-1]
        elif cmd == "endif":
            active.pop()
        elif cmd in ("set", "list") and active[-1]:
            toks = body.split()
            if cmd == "list":
                if toks[0] != "APPEND":
                    continue
                toks = toks[1:]
                var.setdefault(toks[0], []).extend(toks[1:])
            else:
                var[toks[0]] = toks[1:]

    def expand(name, seen=()):
        out = []
        for t in var.get(name, []):
            m = re.fullmatch(r"\$\{(\w+)\}", t)
            if m:
                if m.group(1) not in seen:
                    out += expand(m.group(1), seen + (name,))
            elif t.endswith(".C"):
                out.append(t)
        return out
    return expand(target)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    ap.add_argument("--openmp", action="store_true")
    ap.add_argument("--objdir", default="/tmp/oofem_ref_obj")
    a = ap.parse_args()
    if not os.path.isdir(REF):
        print("reference tree absent; keeping prebuilt oracle/_ref as is")
        return 0
    name = "oofem_omp" if a.openmp else "oofem"
    objdir = a.objdir + ("_omp" if a.openmp else "")
    os.makedirs(objdir, exist_ok=True)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(objdir, "oofemenv.h"), "w") as f:
        f.write("#include <cstddef>\n#define HAVE_EXECINFO_H 1\n#define HAVE_CBRT 1\n"
                "#define HAVE_ISNAN 1\ntypedef std::size_t indexType;\n"
                "#define OOFEM_EXPORT\n#define OOFEM_NO_EXPORT\n")
    srcs = []
    for sub, target in (("core", "core"), ("sm", "sm"), ("sm/Elements", "sm_elements"),
                        ("sm/Materials", "sm_materials")):
        d = os.path.join(REF, "src", sub)
        for s in cmake_sources(os.path.join(d, "CMakeLists.txt"), target):
            p = os.path.join(d, s)
            if os.path.exists(p) and p not in srcs:
                srcs.append(p)
    srcs.append(os.path.join(REF, "src/main/main.C"))
    incs = ["-I" + objdir, "-I" + REF, "-I" + os.path.join(REF, "src"),
            "-I" + os.path.join(REF, "src/core"), "-I" + os.path.join(REF, "src/sm"),
            "-I" + os.path.join(REF, "src/core/iml"), "-I" + os.path.join(REF, "src/core/xfem")]
    cfgdefs = ['-D__OOFEM_VERSION="ref"', '-D__OOFEM_MAJOR_VERSION="0"', '-D__OOFEM_MINOR_VERSION="0"',
               '-D__OOFEM_GIT_HASH="none"', '-D__OOFEM_GIT_REPOURL="none"', '-D__OOFEM_GIT_BRANCH="none"',
               '-D__HOST_TYPE="x86_64-Linux"', '-D__HOST_NAME="oracle"', '-D__OOFEM_COPYRIGHT="see reference"',
               '-D__MODULE_LIST="sm iml"']
    flags = ["-O2", "-std=c++17", "-w", "-fPIC", "-D__SM_MODULE", "-D__IML_MODULE"] + cfgdefs
    if a.openmp:
        flags += ["-fopenmp", "-D_OPENMP_PARALLEL" ]
    print(f"{len(srcs)} sources")

    def cc(src):
        obj = os.path.join(objdir, os.path.relpath(src, REF).replace("/", "_") + ".o")
        if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src):
            return obj, 0, ""
        r = subprocess.run(["g++", "-c", src, "-o", obj] + flags + incs, capture_output=True, text=True)
        return obj, r.returncode, r.stderr[-2000:]

    objs, bad = [], []
    with cf.ThreadPoolExecutor(a.jobs) as ex:
        for i, (obj, rc, err) in enumerate(ex.map(cc, srcs)):
            if rc:
                bad.append((obj, err))
            else:
                objs.append(obj)
            if i % 50 == 0:
                print(f"  [{i}/{len(srcs)}]", flush=True)
    for obj, err in bad:
        print("FAILED", obj, "\n", err)
    if bad:
        return 1
    exe = os.path.join(OUT, name)
    r = subprocess.run(["g++", "-o", exe] + objs + (["-fopenmp"] if a.openmp else []) + ["-ldl", "-lpthread"],
                       capture_output=True, text=True)
    print(r.stderr[-6000:])
    print("built" if r.returncode == 0 else "LINK FAILED", exe)
    if r.returncode:
        return r.returncode
    # our own drivers (main() only) over the same unmodified reference objects:
    # full-precision dump (fixtures) and the timing harness of the reference arm
    lib_objs = [o for o in objs if not o.endswith("src_main_main.C.o")]
    drivers = [("ref_bench.cpp", "oofem_bench_omp" if a.openmp else "oofem_bench")]
    if not a.openmp:
        drivers.append(("ref_dump.cpp", "oofem_dump"))
    for src, exe_name in drivers:
        exe = os.path.join(OUT, exe_name)
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-w", os.path.join(HERE, src), "-o", exe]
                           + flags + incs + lib_objs + ["-ldl", "-lpthread"], capture_output=True, text=True)
        print(r.stderr[-6000:])
        print("built" if r.returncode == 0 else "LINK FAILED", exe)
        if r.returncode:
            return r.returncode
    return 0


if __name__ == "__main__":
    sys.exit(main())
