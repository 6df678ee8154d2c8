"""Builds measurement variants of libdynoed_cuda.so into dynamicoed.jl_b200/_variants/ (git-ignored,
shipped to the GPU box): `python scripts/build_variants.py name=DEFINE[,DEFINE...] ...`; select one
with DYNOED_LIB=<path>.  `old=<git rev>` builds the library of another revision (A/B on one box)."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamicoed.jl_b200 import build

out = os.path.join(ROOT, "dynamicoed.jl_b200", "_variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, val = spec.split("=", 1)
    dst = os.path.join(out, f"libdynoed_{name}.so")
    if name == "old":
        with tempfile.TemporaryDirectory() as tmp:
            subprocess.run(["git", "-C", ROOT, "worktree", "add", "--detach", tmp + "/w", val], check=True)
            try:
                subprocess.run([sys.executable, "-c",
                                "import sys; sys.path.insert(0, '.'); from dynamicoed.jl_b200 import build as B; "
                                f"B.build_library({dst!r}, force=True)"], cwd=tmp + "/w", check=True)
            finally:
                subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", tmp + "/w"], check=True)
    else:
        os.environ["DYNOED_NVCC_DEFINES"] = " ".join(val.split(","))
        build.build_library(dst, force=True)
    print("built", dst)
