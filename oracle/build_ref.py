"""TEST INFRASTRUCTURE — recipe that stages the UNMODIFIED reference modules for the GPU box.

The reference is pure Python: there is nothing to compile, so "building" `oracle/_ref` means copying the two packages
its model path imports (`prototype/`, `linklink/`; *.py only) from /root/reference into `oracle/_ref/` — git-ignored
(never part of the history), not gpurun-ignored (it travels with the snapshot, like the built `_C.so`).  On the GPU box
`/root/reference` does not exist; `oracle/ref_harness.py` then imports the staged copy, so `bench.py --impl reference`
and the `cpu_baseline` leg time the reference's own modules (`kind: "reference"`), not the restatement.

    python oracle/build_ref.py        (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("DECLIP_REFERENCE_SRC", "/root/reference")
PACKAGES = ("prototype", "linklink")


def build(verbose=False):
    if not os.path.isdir(os.path.join(SRC, "prototype")):
        return None                      # GPU box: use whatever was staged in the build container
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(SRC, pkg)):
            dirs[:] = [d for d in dirs if d != "__pycache__"]
            rel = os.path.relpath(root, SRC)
            for f in files:
                if not f.endswith(".py"):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                dst = os.path.join(DST, rel, f)
                src = os.path.join(root, f)
                if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
                    shutil.copyfile(src, dst)
                n += 1
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("staged by oracle/build_ref.py from %s (%d files, unmodified); not tracked by git\n" % (SRC, n))
    if verbose:
        print("oracle/_ref: %d files" % n)
    return DST


if __name__ == "__main__":
    print(build(verbose=True))
    sys.exit(0)
