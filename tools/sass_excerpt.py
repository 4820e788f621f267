#!/usr/bin/env python
"""Write profiles/sass_<kernel>.txt: opcode histogram + the first occurrences of the instructions DESIGN.md cites
(FFMA2 / LDG.E.128 / UBLKCP / SYNCS ...) from `cuobjdump -sass` of the in-tree library.  No GPU needed.

  python tools/sass_excerpt.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "godot-volumetric-cloud-demo-v2_b200", "csrc", "libcloudsky_b200.so")
TARGETS = {  # output file -> (substring of the mangled name, mnemonics to show in context)
    "sass_clouds_fast.txt": ("clouds_fast_kernelILb0ELb1ELi7ELb0EE", ["FFMA2", "FADD2", "FMUL2", "HADD2.F32", "LDG.E.128", "LDS.128", "MUFU", "SHFL", "VOTE", "STG"]),  # <COUNT=0, TYPE_HI=1, FMT=7 (fp16 records), EARLY=0>: the headline kernel
    "sass_clouds_fast_sunbatch.txt": ("clouds_fast_sunbatch_kernelILb1ELi7EE", ["FFMA2", "LDG.E.128", "LDS", "STS", "STG"]),
    "sass_sky_lut.txt": ("sky_lut_kernelILb0EE", ["UBLKCP", "SYNCS", "LDS", "SHFL", "MUFU", "STG"]),
    "sass_transmittance_lut.txt": ("transmittance_lut_kernel", ["MUFU", "STG"]),
}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    blocks = re.split(r"\n\s*Function : ", sass)
    for out, (needle, show) in TARGETS.items():
        blk = next(b for b in blocks[1:] if needle in b.split("\n", 1)[0])
        name = blk.split("\n", 1)[0].strip()
        lines = [l for l in blk.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1)] += 1
        fam = collections.Counter()
        for op, n in ops.items():
            fam[op.split(".")[0]] += n
        with open(os.path.join(ROOT, "profiles", out), "w") as f:
            f.write(f"# cuobjdump -sass {os.path.relpath(SO, ROOT)}   (arch in the fatbin: {', '.join(arch)})\n# kernel: {name}\n")
            f.write(f"# {len(lines)} SASS instructions.  Opcode families, most frequent first:\n")
            f.write("#   " + ", ".join(f"{k} {v}" for k, v in fam.most_common(40)) + "\n")
            f.write("# full opcodes of the families cited in DESIGN.md:\n")
            for s in show:
                hits = {k: v for k, v in ops.items() if k.startswith(s)}
                f.write(f"#   {s}: " + (", ".join(f"{k} x{v}" for k, v in sorted(hits.items(), key=lambda kv: -kv[1])) or "none") + "\n")
            f.write("\n")
            for s in show:
                idx = [i for i, l in enumerate(lines) if re.search(r"\s" + re.escape(s), l)]
                if not idx:
                    continue
                f.write(f"---- first {min(4, len(idx))} of {len(idx)} x {s}* ----\n")
                for i in idx[:4]:
                    f.write(lines[i].rstrip() + "\n")
                f.write("\n")
        print(out, name[:70], len(lines), {s: sum(v for k, v in ops.items() if k.startswith(s)) for s in show})


if __name__ == "__main__":
    main()
