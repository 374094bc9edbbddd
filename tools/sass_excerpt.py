"""SASS evidence for the hot kernels: mnemonic histogram + the memory-instruction skeleton of each, taken with
`cuobjdump -sass` from the in-tree libdynscaler_sm100.so.  usage: python tools/sass_excerpt.py <round-label> [planar]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "dynamicscaler_b200", "libdynscaler_sm100.so")
label = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = [("pm_gather_kernel<float, 4, 16, 4, true>", r"pm_gather_kernelIfLi4ELi16ELi4ELb1E"),
        ("pm_gather_kernel<float, 1, 16, 8, false>", r"pm_gather_kernelIfLi1ELi16ELi8ELb0E"),
        ("pm_merge_nearest_kernel<float, 16, false, false>", r"pm_merge_nearest_kernelIfLi16ELb0ELb0E"),
        ("pm_merge_weighted_tiled_kernel<float, 16, 2, false, false>", r"pm_merge_weighted_tiled_kernelIfLi16ELi2ELb0ELb0E"),
        ("pm_merge_weighted_tiled_kernel<float, 16, 2, false, true> (exchange mode)", r"pm_merge_weighted_tiled_kernelIfLi16ELi2ELb0ELb1E"),
        ("pm_merge_nearest_kernel<float, 16, true, false> (fused CFG + ddim_step)", r"pm_merge_nearest_kernelIfLi16ELb1ELb0E"),
        ("ring_copy_kernel<float, 4, true, 16>", r"ring_copy_kernelIfLi4ELb1ELi16E")]
if len(sys.argv) > 2 and sys.argv[2] == "planar":        # the reference-order kernels of the planar loops
    WANT = [("merge_nearest_ddim_kernel<float, 4, 2, false, false> (planar merge + CFG + ddim_step, 16-byte vectors)",
             r"merge_nearest_ddim_kernelIfLi4ELi2ELb0ELb0E"),
            ("merge_nearest_ddim_kernel<float, 4, 2, false, true> (+ merge_prev mix)", r"merge_nearest_ddim_kernelIfLi4ELi2ELb0ELb1E"),
            ("merge_nearest_ddim_kernel<float, 4, 2, true, false> (exchange mode: peer stores)", r"merge_nearest_ddim_kernelIfLi4ELi2ELb1ELb0E"),
            ("merge_nearest_kernel<float, 4, 4, 4, false> (plain planar merge)", r"merge_nearest_kernelIfLi4ELi4ELi4ELb0E"),
            ("renoise_mix_kernel<float, 4>", r"renoise_mix_kernelIfLi4E"),
            ("ring_copy_kernel<float, 4, true, 16>", r"ring_copy_kernelIfLi4ELb1ELi16E")]
funcs = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", funcs)
KEY = ("UBLKCP", "LDGSTS", "LDG", "STG", "LDS", "STS", "MULTIMEM", "MEMBAR", "BAR", "ACQBULK", "DEPBAR", "FFMA", "FMUL", "FADD",
       "MUFU", "ATOM", "RED", "UTMALDG", "UTMASTG", "SYNCS", "LDL", "STL")
out = [f"# {label} -- SASS of the hot kernels (`cuobjdump -sass dynamicscaler_b200/libdynscaler_sm100.so`, sm_100a)\n",
       "Counts are static instruction counts of the function body (all paths); `UBLKCP` = TMA bulk copy shared->global, "
       "`LDGSTS` = cp.async global->shared, `LDG.E.128` = 16-byte loads, `ACQBULK`/`DEPBAR` = griddepcontrol / bulk-group waits, "
       "`MULTIMEM` = NVLS multicast stores.  " + ("" if len(sys.argv) <= 2 else "Planar (reference-order) kernels: plain 16-byte `LDG`/`STG`, no shared memory.  ") + "The headline fp32 / 16-frame kernels below have no `LDL`/`STL` (no local-memory spills); `-Xptxas -v` shows 4-168 bytes of spill only in rarely used variants (runtime frame count, fused denoise at the 48-register budget).\n"]
for title, pat in WANT:
    blk = next((b for b in blocks if re.match(r"_ZN2ds\d+" + pat, b)), None)
    if blk is None:
        out.append(f"\n## `{title}`\n\nnot found\n")
        continue
    ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", blk)
    hist = collections.Counter()
    for i in ins:
        root = i.split(".")[0]
        if root in KEY:
            hist[i if root in ("LDG", "STG", "LDS", "STS", "MEMBAR", "UBLKCP", "LDGSTS", "MULTIMEM") else root] += 1
    out.append(f"\n## `{title}`\n\n{len(ins)} instructions.  " + ", ".join(f"`{k}` x{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])) + "\n")
    skel = [l.strip() for l in blk.splitlines() if re.search(r"\b(UBLKCP|LDGSTS|LDG\.E\.128|MULTIMEM|ACQBULK|BAR\.SYNC|STS\.128)", l)][:14]
    out.append("\n```\n" + "\n".join(re.sub(r"\s+/\*.*$", "", l)[:110] for l in skel) + "\n```\n")
open(os.path.join(ROOT, "profiles", f"{label}_sass_excerpts.md"), "w").write("".join(out))
print("".join(out)[:3000])
