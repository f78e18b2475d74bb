"""Stall-reason samples per phase (outermost call line [and second level for the forward sweep]).
usage: ncu_stalls.py rep cubin kernel [depth]"""
import csv, collections, re, subprocess, sys
rep, cubin, kname = sys.argv[1:4]; depth = int(sys.argv[4]) if len(sys.argv) > 4 else 1
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines())); hdr = rows[1]
ia = hdr.index("Address"); isamp = hdr.index("# Samples")
reasons = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_no_inst", "stall_selected", "stall_branch_resolving", "stall_mio", "stall_lg", "stall_dispatch", "stall_not_selected"]
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
agg = collections.defaultdict(lambda: [0] * (len(reasons) + 1)); tot = 0
for r in body:
    ch = off2chain.get(int(r[ia], 16) - base, [])
    key = tuple(ch[::-1][:depth]) if ch else ("?",)
    try: s = int(r[isamp])
    except ValueError: continue
    agg[key][0] += s; tot += s
    for k, i in enumerate(ridx):
        try: agg[key][k + 1] += int(r[i])
        except ValueError: pass
src = open("/root/repo/mpc-sensorlessao_b200/csrc/fmpc_kernel_warp.cu").read().splitlines()
print("samples", tot); print(" " * 34 + " ".join(f"{r[6:12]:>7s}" for r in reasons))
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
    l = key[-1]
    print(f"{100*v[0]/tot:5.1f}% {str(key):26s} " + " ".join(f"{100*x/max(v[0],1):6.0f}%" for x in v[1:]) + "  " + (src[l-1].strip()[:60] if isinstance(l, int) and 0 < l <= len(src) else ""))
