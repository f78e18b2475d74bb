"""Opcode mix (executed warp instructions per solve, stall samples) inside source-line regions of a kernel.
usage: ncu_region_ops.py rep cubin kernel nsolves name:lo-hi [name:lo-hi ...]"""
import csv, collections, re, subprocess, sys
rep, cubin, kname, ns = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
regions = []
for a in sys.argv[5:]:
    nm, rng = a.split(":"); lo, hi = map(int, rng.split("-")); regions.append((nm, lo, hi))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines())); hdr = rows[1]
ia, isamp, iexec, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = [r for r in rows[2:] if r[ia].startswith("0x")]; base = int(body[0][ia], 16)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
infn = False; chain = []; off2chain = {}; fresh = True
for ln in dis:
    if ".text." in ln and ":" in ln: infn = (kname in ln); continue
    if not infn: continue
    m = re.search(r'//## File "[^"]*?([^/"]+)", line (\d+)', ln)
    if m:
        if fresh: chain = []; fresh = False
        chain.append(int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m: off2chain[int(m.group(1), 16)] = list(chain); fresh = True
alls = sum(int(r[isamp]) for r in body if r[isamp].isdigit())
for nm, lo, hi in regions:
    ops = collections.Counter(); samp = collections.Counter()
    for r in body:
        try: s = int(r[isamp]); e = int(r[iexec])
        except ValueError: continue
        ch = off2chain.get(int(r[ia], 16) - base, [])
        if not any(lo <= l <= hi for l in ch): continue
        m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', r[isrc]); op = m.group(1) if m else '?'
        ops[op] += e; samp[op] += s
    te, ts = sum(ops.values()), sum(samp.values())
    print(f"== {nm} lines {lo}-{hi}: {te/ns:.0f} instr/solve, {100*ts/alls:.1f}% of samples")
    print("   " + "  ".join(f"{op} {e/ns:.0f} ({100*samp[op]/max(ts,1):.0f}%)" for op, e in ops.most_common(16)))
