"""Builds an experimental variant of the library next to the product one:
    python tools/build_variant.py <name> [-DSWITCH=value ...]   ->  multiview_inpaint_b200/variants/libgsrast_b200_<name>.so
and `GSR_LIB_VARIANT=<name>` makes multiview_inpaint_b200._C load it (tools/gpu_variants.sh benches several in one GPU call).
The variants directory is git-ignored; the product path never sets the variable."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multiview_inpaint_b200 import build as B  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.HERE, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libgsrast_b200_{name}.so")
cmd = [B.NVCC] + B.FLAGS + defs + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-o", out]
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode != 0:
    sys.stderr.write(r.stdout + r.stderr)
    raise SystemExit(1)
print(out)
