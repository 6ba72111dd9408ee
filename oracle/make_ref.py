"""Recipe for oracle/_ref: the UNMODIFIED reference sources the hot path needs, staged from /root/reference so that
`bench.py --impl reference` (and tests that want the live reference) can import them on the GPU box, where
/root/reference does not exist.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing in the product path (intrinsic-neural-fields_b200/) imports oracle/.

The reference is pure Python (no native code, no packaging), so "building" it is staging the
files its path imports: model.py, layers.py (the MLP + view branch), mesh.py (get_k_eigenfunc_vec_vals), the two modules
mesh.py imports at module scope (cameras.py, utils.py) and the neutex/ package model.py imports.  oracle/_ref/ is
git-ignored (like a compiled .so: it never enters history) but not gpurun-ignored (it travels with the snapshot).
Missing third-party imports of those modules (igl, trimesh, torchinfo, imageio) are NOT on the path and are replaced by
empty stub modules at import time (oracle/ref_path.py), exactly as tests/golden/make_golden.py does.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["model.py", "layers.py", "mesh.py", "cameras.py", "utils.py"]
PACKAGES = ["neutex"]


def main() -> bool:
    if not os.path.isdir(REF):
        print(f"[make_ref] {REF} not present: keeping whatever oracle/_ref holds", file=sys.stderr)
        return os.path.isdir(OUT)
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(OUT, f))
    for p in PACKAGES:
        dst = os.path.join(OUT, p)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, p), dst, ignore=shutil.ignore_patterns("*.ipynb", "pretrained", "__pycache__", "*.pt", "*.pth"))
    with open(os.path.join(OUT, "STAGED_FROM"), "w") as fh:
        fh.write(f"{REF}: " + " ".join(FILES + [p + "/" for p in PACKAGES]) + "\n")
    return True


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
