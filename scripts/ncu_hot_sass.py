"""Hottest SASS instructions (by stall samples) whose inline chain contains a given source line range.
usage: ncu_hot_sass.py rep cubin kernel lo hi [top]"""
import csv, collections, re, subprocess, sys
rep, cubin, kname, lo, hi = sys.argv[1:6]; lo = int(lo); hi = int(hi); top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines())); hdr = rows[1]
ia, isamp, iexec, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
reasons = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_no_inst", "stall_selected", "stall_branch_resolving", "stall_dispatch"]
ridx = [hdr.index(r) for r in reasons]
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
sel = []; tot = 0; alls = 0
for r in body:
    off = int(r[ia], 16) - base
    try: s = int(r[isamp]); e = int(r[iexec])
    except ValueError: continue
    alls += s
    ch = off2chain.get(off, [])
    if any(lo <= l <= hi for l in ch):
        sel.append((s, e, off, r[isrc].strip(), [r[i] for i in ridx], ch[0] if ch else 0)); tot += s
print(f"region samples {tot} = {100*tot/alls:.1f}% of kernel; instructions {len(sel)}; executed {sum(x[1] for x in sel)}")
print("  samp%   exec      off   line  " + " ".join(r[6:12] for r in reasons))
for s, e, off, src, rs, l in sorted(sel, key=lambda x: -x[0])[:top]:
    print(f"{100*s/max(tot,1):6.2f} {e:>9} {off:6x} {l:5d}  {' '.join(f'{x:>6s}' for x in rs)}  {src[:70]}")
