#!/usr/bin/env python
"""Copy recipe for the reference arm of bench.py (`--impl reference`, `cpu_baseline.kind == "reference"`).

The reference (jokofa/JAMPR_plus) is 100 % Python with no build step, so "installing" it is a verbatim copy of its
`lib/` package and the one policy config the path uses into `baseline/_ref/` (git-ignored: reference sources never enter
this repository's history; NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot).  Its third-party
imports that are absent from the image (torch_geometric, torch_cluster, torch_scatter, matplotlib, seaborn, omegaconf,
hydra) are satisfied by the import shims in `oracle/ref_shims/` (see its README for the restated semantics).

    python baseline/make_ref.py            # run in the build container (needs /root/reference)

`__graft_entry__.build()` calls this when /root/reference is present.  Nothing is modified: tests/test_reference_arm.py
compares the copy with the mount file by file when both exist."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JAMPR_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["LICENSE", "config/policy/gnn_attn.yaml", "config_lns/search/default.yaml"]


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(os.path.join(REF, "lib")):
        if verbose:
            print(f"{REF}/lib not found: nothing copied")
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(os.path.join(REF, "lib"), os.path.join(DST, "lib"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for f in FILES:
        src = os.path.join(REF, f)
        if os.path.exists(src):
            os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
            shutil.copy2(src, os.path.join(DST, f))
    if verbose:
        n = sum(len(fs) for _, _, fs in os.walk(DST))
        print(f"copied {n} files of the unmodified reference into {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
