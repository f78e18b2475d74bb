"""Aggregate an ncu SASS source page (csv) by CUDA source line using nvdisasm line info.
usage: ncu_by_line.py <report.ncu-rep> <cubin> <kernel-substring> [topN]"""
import csv, collections, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]
ia, isamp, iexec = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = rows[2:]
base = int(body[0][ia], 16)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# walk the disassembly of the kernel: track current line annotation; instruction offsets /*0010*/
infn = False; cur = ("?", 0); off2line = {}
for ln in dis:
    if ln.startswith("\t.section") or ".text." in ln:
        infn = (kname in ln)
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
agg = collections.Counter(); ex = collections.Counter(); tot = 0
for r in body:
    try:
        off = int(r[ia], 16) - base; s = int(r[isamp]); e = int(r[iexec])
    except ValueError:
        continue
    key = off2line.get(off, (("?", 0), ""))[0]
    agg[key] += s; ex[key] += e; tot += s
print("total samples", tot)
src = {}
for (f, l), s in agg.most_common(top):
    if f not in src:
        try: src[f] = open("/root/repo/mpc-sensorlessao_b200/csrc/" + f).read().splitlines()
        except OSError: src[f] = []
    text = src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
    print(f"{100*s/tot:5.1f}%  {ex[(f,l)]:>12}  {f}:{l}  {text}")

# optional phase summary: env NCU_PHASES="name:file:lo-hi,..."
import os
ph = os.environ.get("NCU_PHASES")
if ph:
    print("---- phases ----")
    for item in ph.split(","):
        name, f, rng = item.split(":")
        lo, hi = map(int, rng.split("-"))
        s = sum(v for (ff, l), v in agg.items() if ff == f and lo <= l <= hi)
        e = sum(v for (ff, l), v in ex.items() if ff == f and lo <= l <= hi)
        print(f"{name:28s} samples {100*s/tot:5.1f}%   instr {e/1e6:9.1f} M")
