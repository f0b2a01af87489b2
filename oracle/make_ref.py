"""Recipe for oracle/_ref/: a verbatim, git-ignored copy of the three source files of the reference package
(/root/reference/rayen/{__init__,utils,constraints,constraint_module}.py), made in the build container so that the
UNMODIFIED reference can travel to the GPU box with the repo snapshot and be timed there as the CPU baseline
(`bench.py --impl reference`, `cpu_baseline.kind = "reference"`).  Test / measurement infrastructure only: nothing under
rayen_b200/ imports it, oracle/_ref/ is listed in .gitignore (never in history) and nothing is edited on the way.

    python oracle/make_ref.py            # no-op when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RAYEN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "utils.py", "constraints.py", "constraint_module.py")


def make(verbose=False):
    src_pkg = os.path.join(SRC, "rayen")
    if not os.path.isfile(os.path.join(src_pkg, "constraint_module.py")):
        if verbose:
            print(f"make_ref: no reference tree under {SRC}; oracle/_ref left as it is")
        return False
    dst_pkg = os.path.join(DST, "rayen")
    os.makedirs(dst_pkg, exist_ok=True)
    for name in FILES:
        p = os.path.join(src_pkg, name)
        if os.path.isfile(p):
            shutil.copyfile(p, os.path.join(dst_pkg, name))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as fh:
        fh.write(f"verbatim copy of {src_pkg}/*.py made by oracle/make_ref.py; not tracked by git\n")
    if verbose:
        print(f"make_ref: copied {src_pkg} -> {dst_pkg}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make(verbose=True) or True else 1)
