"""Development tool: build libc2m_warp.so variants with different -D switches into c2m_b200/variants/<name>.so
(selected at run time with C2M_WARP_LIB=<path>), for A/B timing of kernel tuning switches on the GPU box.

    python tools/build_variants.py name1:-DC2M_X=1,-DC2M_Y=0 name2:...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from c2m_b200 import _build  # noqa: E402

out_dir = os.path.join(_build.PKG, "variants")
os.makedirs(out_dir, exist_ok=True)
nvcc = _build._nvcc()
_build.build()  # the default objects exist
jobs = []
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    flags = [d for d in defs.split(",") if d]
    obj = os.path.join(out_dir, name + "_gather.o")
    cmd = [nvcc, *_build.NVCC_FLAGS, *flags, "-c", os.path.join(_build.CSRC, "warp_bwd_gather.cu"), "-o", obj]
    jobs.append((name, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, obj, pr in jobs:
    out, _ = pr.communicate()
    if pr.returncode:
        raise SystemExit(f"{name}: nvcc failed\n{out}")
    objs = [os.path.join(_build.PKG, "build", s.replace(".cu", ".o")) for s in _build.SOURCES if s != "warp_bwd_gather.cu"]
    lib = os.path.join(out_dir, name + ".so")
    subprocess.run([nvcc, "-shared", "-o", lib, obj, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                    "-fPIC"], check=True)
    os.remove(obj)
    print("built", lib)
