"""Aggregate an ncu source page by the OUTERMOST call-site line (inline chains from nvdisasm -gi), so that heavily
inlined device functions are attributed to the phase that called them.
usage: ncu_phases.py <report.ncu-rep> <cubin> <kernel-substring> [depth]   (depth 1 = kernel-level line, 2 = one level deeper)"""
import csv, collections, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]
ia, isamp, iexec, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[2:] if r[ia].startswith("0x")]
base = int(body[0][ia], 16)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
infn = False; chain = []; off2chain = {}; fresh = True
for ln in dis:
    if ".text." in ln and ":" in ln:
        infn = (kname in ln); continue
    if not infn: continue
    m = re.search(r'//## File "[^"]*?([^/"]+)", line (\d+)(?: inlined at "[^"]*?([^/"]+)", line (\d+))?', ln)
    if m:
        if fresh: chain = []; fresh = False
        chain.append((int(m.group(2)), int(m.group(4)) if m.group(4) else None))
        continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m:
        off2chain[int(m.group(1), 16)] = list(chain); fresh = True
def key_of(ch):
    # chain entries: (line, inlined_at_line); outermost = entry whose inlined_at is None -> its line is in the kernel body
    if not ch: return ("?",)
    lines = [c[0] for c in ch]           # inner ... outer
    outer = lines[::-1]
    return tuple(outer[:depth])
agg = collections.Counter(); ex = collections.Counter(); dm = collections.Counter(); tot = 0; tex = 0
for r in body:
    off = int(r[ia], 16) - base
    try: s = int(r[isamp]); e = int(r[iexec])
    except ValueError: continue
    k = key_of(off2chain.get(off, []))
    agg[k] += s; ex[k] += e; tot += s; tex += e
    if "DMMA" in r[isrc]: dm[k] += e
src = open("/root/repo/mpc-sensorlessao_b200/csrc/fmpc_kernel_warp.cu").read().splitlines()
print(f"total samples {tot}, warp instructions {tex}")
for k, s in agg.most_common(45):
    l = k[-1]
    text = src[l - 1].strip()[:90] if isinstance(l, int) and 0 < l <= len(src) else ""
    print(f"{100*s/tot:5.1f}% samp  {100*ex[k]/tex:5.1f}% instr ({ex[k]:>10}) dmma {dm[k]:>9}  {k}  {text}")
