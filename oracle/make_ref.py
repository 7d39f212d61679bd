#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: the UNMODIFIED reference's hot-path Python modules, placed (not tracked: `oracle/_ref/` is in
.gitignore, NOT in .gpurunignore) where they travel to the GPU box, so that `bench.py --impl reference` can time the reference's
OWN CPU implementation there (`cpu_baseline.kind = "reference"`) instead of the oracle port.  The reference is Python with no
build step and no installer (no setup.py / pyproject.toml): "building" it is copying the modules the path imports.

    python oracle/make_ref.py            (run in the build container, where /root/reference exists; __graft_entry__.build() does)

Only `bench.py --impl reference`, `bench.py`'s cpu_baseline legs and tests may execute anything under `oracle/`; nothing in
`indm_b200/` does (tests/test_isolation_cpu.py).  Nothing is modified: files are copied byte for byte; the import stubs the
reference needs on a modern toolchain live in oracle/ref_stubs/ (SURVEY.md 8c)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("INDM_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
# what the sampling / training / likelihood path imports; datasets, FID, run_lib, main stay out
KEEP_DIRS = ("models", "op", "flow_models", "configs")
KEEP_FILES = ("sde_lib.py", "sampling.py", "losses.py", "likelihood.py", "LICENSE")
EXTS = (".py", ".json")


def main():
    if not os.path.isdir(os.path.join(SRC, "models")):
        print(f"oracle/make_ref.py: no reference tree at {SRC}; nothing to do")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for d in KEEP_DIRS:
        for root, _, files in os.walk(os.path.join(SRC, d)):
            for f in files:
                if f.endswith(EXTS):
                    rel = os.path.relpath(os.path.join(root, f), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
                    n += 1
    for f in KEEP_FILES:
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    print(f"oracle/make_ref.py: {n} files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
