"""Regenerates profiles/sass/*.sass and MNEMONICS.txt from the objects of the current in-tree build (no GPU needed).

The listings are the evidence that the hot kernels use the Blackwell paths they claim: UTCHMMA (tcgen05.mma), UTMALDG (TMA),
LDTM (tcgen05.ld), LDGMC / multimem (NVLS), REDG (vector red.add), LDG/STG .STRONG.SYS (peer flags and flag-in-data lines).
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dist_tuto.pth_b200", "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
INTERESTING = re.compile(r"^(UTC|UTMA|LDTM|STTM|LDGMC|LDG|STG|REDG|RED|ATOMG|LDS|STS|FFMA|HFMA2|MEMBAR|BAR|CCTL|SYNCS|UCGABAR|ACQBULK|ELECT|FENCE|ERRBAR)")


KEY = re.compile(r"^(UTC|UTMA|LDTM|STTM|LDGMC|REDG|ATOMG|UCGABAR|ACQBULK|.*STRONG\.SYS)")


def newest(stem):
    objs = sorted(glob.glob(os.path.join(BUILD, stem + ".cu.*.o")), key=os.path.getmtime)
    return objs[-1] if objs else None


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = ["# SASS evidence (cuobjdump -sass of the objects linked into dist_tuto.pth_b200/_C.so, sm_100a) -- mnemonic counts per file"]
    for stem in ("allreduce", "sgd", "convnet", "convnet_cluster", "gemm_tcgen05", "convnet_batched"):
        obj = newest(stem)
        if obj is None:
            continue
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        # keep the listing readable and small: drop the hex encodings (second comment column and encoding-only lines)
        slim = []
        for ln in sass.splitlines():
            ln = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/\s*$", "", ln)
            if ln.strip():
                slim.append(ln.rstrip())
        open(os.path.join(OUT, stem + ".sass"), "w").write("\n".join(slim) + "\n")
        per_fn, cur = collections.OrderedDict(), None
        for ln in sass.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                cur = per_fn.setdefault(m.group(1), collections.Counter())
                continue
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if m and cur is not None:
                op = m.group(1)
                if INTERESTING.match(op):
                    cur[op] += 1
        lines.append(f"## {stem}")
        for fn, cnt in per_fn.items():
            demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
            keys = [k for k, _ in cnt.most_common(12)]
            keys += [k for k in cnt if KEY.match(k) and k not in keys]      # the Blackwell-specific ones are always listed
            top = ", ".join(f"{k} x{cnt[k]}" for k in keys)
            lines.append(f"* `{demangled}`: {top}")
    open(os.path.join(OUT, "MNEMONICS.txt"), "w").write("\n".join(lines) + "\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
