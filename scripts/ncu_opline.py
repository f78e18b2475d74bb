"""Executed count of one opcode family by (outermost call line, innermost line).  usage: ncu_opline.py rep cubin kernel OPCODE [top]"""
import csv, collections, re, subprocess, sys
rep, cubin, kname, opc = sys.argv[1:5]; top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines())); hdr = rows[1]
ia, iexec, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
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
agg = collections.Counter(); tot = 0
for r in body:
    if not re.match(r'\s*(?:@!?U?P\w+\s+)?' + opc + r'\b', r[isrc]): continue
    try: e = int(r[iexec])
    except ValueError: continue
    ch = off2chain.get(int(r[ia], 16) - base, [])
    key = (ch[-1] if ch else 0, ch[0] if ch else 0)
    agg[key] += e; tot += e
src = open("/root/repo/mpc-sensorlessao_b200/csrc/fmpc_kernel_warp.cu").read().splitlines()
print(opc, "total", tot)
for (o, i), e in agg.most_common(top):
    print(f"{100*e/tot:5.1f}% {e:>10}  outer {o} inner {i}: {src[i-1].strip()[:100] if 0 < i <= len(src) else ''}")
