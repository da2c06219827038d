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
    procs = []
    for src in _build.SOURCES:  # every source is rebuilt with the switches (a switch may live in any of them)
        obj = os.path.join(out_dir, name + "_" + src.replace(".cu", ".o"))
        cmd = [nvcc, *_build.NVCC_FLAGS, *flags, "-c", os.path.join(_build.CSRC, src), "-o", obj]
        procs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    jobs.append((name, procs))
for name, procs in jobs:
    objs = []
    for obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode:
            raise SystemExit(f"{name}: nvcc failed\n{out}")
        objs.append(obj)
    lib = os.path.join(out_dir, name + ".so")
    subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                    "-fPIC"], check=True)
    for obj in objs:
        os.remove(obj)
    print("built", lib)
