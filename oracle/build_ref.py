"""Recipe that puts the UNMODIFIED reference next to the oracle so it can travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (cremebrule/rgb-proprioceptive-pose-estimator) is pure Python
with no setup.py, so there is nothing to compile or pip-install: this copies its `models/`, `util/` and `scripts/`
Python files from where they lie (/root/reference, or $PE_REFERENCE_ROOT) into `oracle/_ref/`, which is listed in
.gitignore (never committed -- the history stays free of reference sources) but NOT in .gpurunignore, so it ships
with the gpurun snapshot exactly like the built libpe_b200.so does.  On the GPU box it is what
  * `bench.py --impl reference` and the `cpu_baseline` leg time (the reference's own nn.Modules on the host cores),
  * tests/test_dropin_scripts.py runs UNCHANGED (scripts/train_model.py, scripts/rollout.py) against the mirrors,
  * tests/test_oracle_cpu.py pins the oracle restatement to.
Nothing under rgb-proprioceptive-pose-estimator_b200/ imports it.

Run:  python oracle/build_ref.py      (__graft_entry__.build() calls build_ref())
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "oracle", "_ref")
SOURCE = os.environ.get("PE_REFERENCE_ROOT", "/root/reference")
PARTS = ("models", "util", "scripts")


def build_ref(verbose=True):
    """Copy the reference's Python files; returns DEST, or None when no reference tree is reachable (GPU box: the
    copy made in the build container travels with the snapshot and is used as is)."""
    if not os.path.isdir(os.path.join(SOURCE, "models")):
        if verbose:
            print("oracle/_ref: no reference tree at %s; %s" % (
                SOURCE, "using the shipped copy" if os.path.isdir(os.path.join(DEST, "models")) else "none available"),
                file=sys.stderr)
        return DEST if os.path.isdir(os.path.join(DEST, "models")) else None
    n = 0
    for part in PARTS:
        src_dir, dst_dir = os.path.join(SOURCE, part), os.path.join(DEST, part)
        if not os.path.isdir(src_dir):
            continue
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if name.endswith(".py"):
                shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
                n += 1
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Verbatim copy of %s/{%s}/*.py made by oracle/build_ref.py.\nGit-ignored; test and baseline "
                "infrastructure only; never imported by the product.\n" % (SOURCE, ",".join(PARTS)))
    if verbose:
        print("oracle/_ref: copied %d reference files from %s" % (n, SOURCE), file=sys.stderr)
    return DEST


def ref_root():
    """Where the reference's own sources can be imported from: the live tree in the build container, else the copy."""
    if os.path.isdir(os.path.join(SOURCE, "models")):
        return SOURCE
    if os.path.isdir(os.path.join(DEST, "models")):
        return DEST
    return None


if __name__ == "__main__":
    print(build_ref())
