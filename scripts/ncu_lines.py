"""Join an ncu SASS-level source page (csv) with nvdisasm line info -> per-source-line instruction/stall table.
usage: ncu_lines.py <ncu-rep> <cubin> <kernel-substr> [topN]"""
import csv, re, subprocess, sys, collections
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ".text." in ln and ln.strip().startswith(".section"):
        infn = kern in ln
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel only
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0])
base = None
for r in rows[start + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    a = int(r[ia], 16)
    base = a if base is None else base
    key = line_of.get(a - base, ("?", 0))
    agg[key][0] += int(r[ii] or 0)
    agg[key][1] += int(r[isamp] or 0)
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot_i}  samples {tot_s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:<5} instr {v[0]:>9} ({100*v[0]/tot_i:5.1f}%)  samples {v[1]:>6} ({100*v[1]/max(tot_s,1):5.1f}%)")

# ---- phase summary (markers: comment lines containing '-- P0:' / '-- S1:' etc.)
import os
src = sorted({p for p in ("dist_tuto.pth_b200/csrc/" + k[0] for k in agg if k[0].endswith(".cu")) if os.path.isfile(p)}, key=lambda p: -sum(v[1] for k, v in agg.items() if p.endswith(k[0])))
if src:
    marks = []
    for n, ln in enumerate(open(src[0]), 1):
        m = re.search(r"// -+ ((?:P|S)[\w/]+):? (.*)", ln)
        if m:
            marks.append((n, m.group(1) + " " + m.group(2)[:40]))
        elif "// ---" in ln and "flush" in ln:
            marks.append((n, "flush"))
    marks.append((10 ** 9, "end"))
    print("\nphase summary")
    for (a, name), (b, _) in zip(marks, marks[1:]):
        ins = sum(v[0] for k, v in agg.items() if k[0].endswith(".cu") and a <= k[1] < b)
        smp = sum(v[1] for k, v in agg.items() if k[0].endswith(".cu") and a <= k[1] < b)
        print(f"  {name:<48} instr {100*ins/tot_i:5.1f}%  samples {100*smp/max(tot_s,1):5.1f}%")
