"""Top SASS instructions by one stall reason, with source line chain.  usage: ncu_top_stall.py rep cubin kernel stall_column [top]"""
import csv, re, subprocess, sys
rep, cubin, kname, col = sys.argv[1:5]; top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines())); hdr = rows[1]
ia, isamp, iexec, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
ic = hdr.index(col)
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
sel = []
tot = 0; alls = 0
for r in body:
    try: v = int(r[ic]); s = int(r[isamp])
    except ValueError: continue
    alls += s; tot += v
    sel.append((v, int(r[iexec]), int(r[ia], 16) - base, r[isrc].strip()))
print(f"{col}: {tot} of {alls} samples ({100*tot/alls:.1f}%)")
src = open("/root/repo/mpc-sensorlessao_b200/csrc/fmpc_kernel_warp.cu").read().splitlines()
for v, e, off, txt in sorted(sel, key=lambda x: -x[0])[:top]:
    ch = off2chain.get(off, [])
    print(f"{100*v/max(tot,1):5.1f}% {e:>9} {off:6x} lines {ch[:3]}  {txt[:60]}   | {src[ch[0]-1].strip()[:70] if ch and 0 < ch[0] <= len(src) else ''}")
