#!/usr/bin/env python3
"""Build OOFEM with the cudacsr / cudacg plugin: plugin/_build/oofem_cuda.

The plugin sources (plugin/*.C) are compiled against the reference's headers; the hooks
(plugin/engngm_hook.patch, plugin/structengngmodel_hook.patch) are applied to scratch copies of src/core/engngm.C and
src/sm/EngineeringModels/structengngmodel.C; everything else is the
reference's own objects as compiled by oracle/build_ref.py (same flags, same oofemenv.h stub), linked
with liboofem_b200.so.  Nothing is written into /root/reference, no reference source enters the repo.

    python plugin/build_plugin.py            # needs /root/reference and a finished oracle/build_ref.py
Outputs (git-ignored, they travel to the GPU box): plugin/_build/oofem_cuda (the stock main.C
executable with the two new types registered) and plugin/_build/oofem_dump_cuda (plugin/plugin_dump.cpp,
the full-precision dump driver of the plugin parity tests).
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("OOFEM_REFERENCE", "/root/reference")
OBJDIR = os.environ.get("OOFEM_REF_OBJDIR", "/tmp/oofem_ref_obj")
OUT = os.path.join(HERE, "_build")
SCRATCH = "/tmp/oofem_b200_plugin"


def main():
    if not os.path.isdir(REF):
        print("reference tree absent; keeping prebuilt plugin/_build as is")
        return 0
    ref_objs = sorted(glob.glob(os.path.join(OBJDIR, "*.o")))
    if len(ref_objs) < 100:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")])
        ref_objs = sorted(glob.glob(os.path.join(OBJDIR, "*.o")))
    from importlib import import_module
    sys.path.insert(0, ROOT)
    lib = import_module("oofem_b200.build").build()
    os.makedirs(OUT, exist_ok=True)
    os.makedirs(SCRATCH, exist_ok=True)
    # the hooks: patch scratch copies of engngm.C and structengngmodel.C
    shutil.copy(os.path.join(REF, "src/core/engngm.C"), os.path.join(SCRATCH, "engngm.C"))
    subprocess.check_call(["patch", "-s", os.path.join(SCRATCH, "engngm.C"), os.path.join(HERE, "engngm_hook.patch")])
    shutil.copy(os.path.join(REF, "src/sm/EngineeringModels/structengngmodel.C"), os.path.join(SCRATCH, "structengngmodel.C"))
    subprocess.check_call(["patch", "-s", os.path.join(SCRATCH, "structengngmodel.C"), os.path.join(HERE, "structengngmodel_hook.patch")])
    incs = ["-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-I" + OBJDIR, "-I" + REF, "-I" + os.path.join(REF, "src"),
            "-I" + os.path.join(REF, "src/core"), "-I" + os.path.join(REF, "src/sm"),
            "-I" + os.path.join(REF, "src/core/iml"), "-I" + os.path.join(REF, "src/core/xfem")]
    cfgdefs = ['-D__OOFEM_VERSION="ref"', '-D__OOFEM_MAJOR_VERSION="0"', '-D__OOFEM_MINOR_VERSION="0"',
               '-D__OOFEM_GIT_HASH="none"', '-D__OOFEM_GIT_REPOURL="none"', '-D__OOFEM_GIT_BRANCH="none"',
               '-D__HOST_TYPE="x86_64-Linux"', '-D__HOST_NAME="oracle"', '-D__OOFEM_COPYRIGHT="see reference"',
               '-D__MODULE_LIST="sm iml cuda"']
    flags = ["-O2", "-std=c++17", "-w", "-fPIC", "-D__SM_MODULE", "-D__IML_MODULE"] + cfgdefs
    objs = []
    for src in [os.path.join(HERE, f) for f in ("cudacontext.C", "cudacsr.C", "cudacg.C")] + [os.path.join(SCRATCH, "engngm.C"), os.path.join(SCRATCH, "structengngmodel.C")]:
        obj = os.path.join(SCRATCH, os.path.basename(src) + ".o")
        r = subprocess.run(["g++", "-c", src, "-o", obj] + flags + incs, capture_output=True, text=True)
        if r.returncode:
            print(r.stderr[-6000:])
            return 1
        objs.append(obj)
    base = [o for o in ref_objs if not o.endswith(("src_core_engngm.C.o", "src_sm_EngineeringModels_structengngmodel.C.o"))]
    nomain = [o for o in base if not o.endswith("src_main_main.C.o")]
    link = ["-L" + os.path.dirname(lib), "-loofem_b200", "-Wl,-rpath,$ORIGIN/../../oofem_b200", "-ldl", "-lpthread", "-rdynamic"]
    exe = os.path.join(OUT, "oofem_cuda")
    r = subprocess.run(["g++", "-o", exe] + objs + base + link, capture_output=True, text=True)
    print(r.stderr[-4000:])
    if r.returncode:
        return 1
    print("built", exe)
    exe = os.path.join(OUT, "oofem_dump_cuda")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-w", os.path.join(HERE, "plugin_dump.cpp"), "-o", exe]
                       + flags + incs + objs + nomain + link, capture_output=True, text=True)
    print(r.stderr[-4000:])
    if r.returncode:
        return 1
    print("built", exe)
    return 0


if __name__ == "__main__":
    sys.exit(main())
