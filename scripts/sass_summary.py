#!/usr/bin/env python
"""SASS evidence of the in-tree build: per hot kernel, the count of the Blackwell-specific mnemonics
(bulk TMA copies, mbarrier transactions, DSMEM st.async, cluster barriers, packed fp32x2 math) in
fuser_b200/lib/obj/*.o, read with cuobjdump -sass.   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = [  # (object, substring of the demangled kernel name)
    ("inst_fwd_bf16.o", "rmsnorm_fwd_ring_kernel<__nv_bfloat16, 128, 4, false, false>"),
    ("inst_bwd_bf16.o", "norm_bwd_ring_kernel<__nv_bfloat16, 256, 2, false, true, true, false, false>"),
    ("inst_fwd_bf16.o", "softmax_fwd_ring_kernel<__nv_bfloat16, 256, 4, false>"),
    ("inst_fwd_bf16.o", "softmax_bwd_ring_kernel<__nv_bfloat16, 256, 4, false>"),
    ("inst_bwd_bf16.o", "norm_bwd_ring_kernel<__nv_bfloat16, 512, 3, true, false, true, false, false>"),
    ("inst_pipe_bf16.o", "rmsnorm_fwd_pipe_kernel<__nv_bfloat16, 512, 4>"),
    ("inst_pipe_bf16.o", "rms_bwd_pipe_lean_kernel<__nv_bfloat16, 512, 4>"),
    ("inst_pipe_bf16.o", "norm_bwd_pipe_kernel<__nv_bfloat16, 512, 4, false>"),
    ("inst_lnpipe_bf16.o", "layernorm_fwd_pipe_kernel<__nv_bfloat16, 512, 3, false>"),
    ("inst_lnpipe_bf16.o", "layernorm_fwd_pipe_kernel<__nv_bfloat16, 128, 4, false>"),
    ("inst_lnpipe_f32.o", "layernorm_fwd_pipe_kernel<float, 256, 4, false>"),
    ("inst_reduce_bf16.o", "reduce_inner_ring_kernel<__nv_bfloat16, 128, 0>"),
    ("inst_reduce_bf16.o", "colsum_ring_kernel<__nv_bfloat16, 512, 3>"),
    ("inst_reduce_bf16.o", "reduce_inner_ring_kernel<__nv_bfloat16, 512, 0>"),
    ("outer.o", "outer_rows_kernel<__nv_bfloat16, false>"),
    ("outer.o", "outer_rows_kernel<__nv_bfloat16, true>"),
    ("inst_drop_bf16.o", "layernorm_fwd_ring_kernel<__nv_bfloat16, 128, 4, false, true, true>"),
    ("inst_drop_bf16.o", "rmsnorm_fwd_ring_kernel<__nv_bfloat16, 128, 4, true, true>"),
    ("inst_bwd_bf16.o", "norm_bwd_ring_kernel<__nv_bfloat16, 256, 2, true, true, true, true, false>"),
    ("b200norm.o", "cross_entropy_bwd_kernel<__nv_bfloat16, true>"),
    ("inst_smpipe_bf16.o", "softmax_fwd_pipe_kernel<__nv_bfloat16, 512, 8, 0>"),
    ("inst_smpipe_bf16.o", "softmax_fwd_pipe_kernel<__nv_bfloat16, 512, 8, 2>"),
    ("inst_lnpipe_bf16.o", "layernorm_fwd_pipe_kernel<__nv_bfloat16, 512, 6, false>"),
    ("inst_pipe_bf16.o", "rmsnorm_fwd_pipe_kernel<__nv_bfloat16, 512, 6>"),
]
KEYS = ["UBLKCP", "UTMALDG", "SYNCS.ARRIVE", "SYNCS.PHASECHK", "STAS", "UCGABAR", "MAPA", "FFMA2", "FADD2", "FMUL2", "IMAD.WIDE",
        "F2FP", "MUFU", "LDS.128", "STG.E.EF.128", "ATOMG", "ACQBULK", "LDL", "STL"]


def main():
    objdir = os.path.join(ROOT, "fuser_b200", "lib", "obj")
    cache = {}
    print("# cuobjdump -sass of fuser_b200/lib/obj/*.o (sm_100a): mnemonic counts per hot kernel")
    print("# (static counts over the whole kernel, not per iteration; LDL / STL = register spills)")
    for obj, name in HOT:
        path = os.path.join(objdir, obj)
        if path not in cache:
            sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
            dem = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout or sass
            cache[path] = dem
        text = cache[path]
        blocks = re.split(r"\n\s*Function : ", text)
        hit = [b for b in blocks if name.replace(" ", "") in b.split("\n", 1)[0].replace(" ", "").replace("(int)", "").replace("(bool)0", "false").replace("(bool)1", "true")]
        if not hit:
            print(f"{name}: not found in {obj}")
            continue
        body = hit[0]
        ops = collections.Counter()
        total = 0
        for line in body.splitlines():
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if not m:
                continue
            total += 1
            for k in KEYS:
                if m.group(1).startswith(k):
                    ops[k] += 1
        print(f"{name}\n    {total} instructions: " + ", ".join(f"{k} {ops[k]}" for k in KEYS if ops[k]))
    # whole library
    lib = os.path.join(ROOT, "fuser_b200", "lib", "libb200norm.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts = {k: len(re.findall(r"\b" + re.escape(k), sass)) for k in ["UBLKCP", "UTMALDG", "SYNCS", "STAS", "UCGABAR", "FFMA2"]}
    print(f"\n# whole library ({os.path.getsize(lib) / 1e6:.1f} MB): " + ", ".join(f"{k} {v}" for k, v in counts.items()))
    print("# tensor-core instructions (HMMA / UTCMMA / tcgen05): " + str(len(re.findall(r"\b(HMMA|UTC[A-Z]*MMA|QMMA)", sass))) +
          " — nothing on this path is a contraction")


if __name__ == "__main__":
    main()
