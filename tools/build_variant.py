"""Build a measurement variant of liblsvs_b200.so into variants/<name>/ (git-ignored, travels to the GPU box):

    python tools/build_variant.py phases -DLSVS_ATTN_PHASES
    LSVS_B200_LIB=variants/phases/liblsvs_b200.so python tools/attn_phases.py

The product library (large-scale-vit-slam_b200/lib) is untouched."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "large-scale-vit-slam_b200"))
from lsvs_b200 import build as B

name, defines = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "variants", name)
os.makedirs(out, exist_ok=True)
procs = []
for src in B.sources():
    obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
    procs.append((obj, subprocess.Popen([B.NVCC, *B.FLAGS, *defines, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
log = []
for obj, p in procs:
    o, _ = p.communicate()
    log.append(f"== {os.path.basename(obj)}\n{o}")
    if p.returncode:
        sys.exit("\n".join(log))
open(os.path.join(out, "ptxas.log"), "w").write("\n".join(log))
lib = os.path.join(out, "liblsvs_b200.so")
subprocess.check_call([B.NVCC, "-shared", "-o", lib, *[o for o, _ in procs], "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"])
for o, _ in procs:
    os.remove(o)
print(lib)
