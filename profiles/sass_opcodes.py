"""Static SASS opcode histogram of every kernel in libmaf_b200.so (cuobjdump -sass on the object files of the build).

    python profiles/sass_opcodes.py > profiles/r02_sass_opcodes.txt

One line per kernel: instruction count, then the data-movement / synchronisation opcodes that tell HOW it moves data
(LDG / STG = register-staged global access, LDGSTS = cp.async global->shared, UBLKCP / UTMALDG = bulk / tensor TMA copies,
SYNCS = mbarrier, BAR = CTA barrier, SHFL / LDS / STS, MUFU = special-function unit, the FP64 opcodes) and the ten most
frequent opcodes.  Static counts, not executed counts: the executed mix is in the ncu summaries (`*_ncu.md`)."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["LDG", "STG", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "BAR", "LDS", "STS", "SHFL", "MUFU", "DFMA", "DADD", "DMUL",
         "ATOMG", "ATOMS", "RED", "FRND", "I2F", "F2I", "CCTL"]


def main():
    for obj in sorted(glob.glob(os.path.join(ROOT, "modular_audio_pipeline_b200", "csrc", "*.o"))):
        txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, ops = None, None
        out = []
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    out.append((name, ops))
                name, ops = m.group(1), collections.Counter()
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if m and name:
                ops[m.group(1)] += 1
        if name:
            out.append((name, ops))
        print(f"== {os.path.basename(obj)}")
        for name, ops in out:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = re.sub(r"\(.*", "", dem).replace("maf::", "").replace("void ", "")
            total = sum(ops.values())
            watch = " ".join(f"{k}={ops[k]}" for k in WATCH if ops.get(k))
            top = " ".join(f"{k}:{v}" for k, v in ops.most_common(10))
            print(f"{dem:40s} {total:6d} inst | {watch} | top: {top}")


if __name__ == "__main__":
    main()
