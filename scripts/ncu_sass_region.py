"""Print SASS instructions (with stall samples) of a kernel whose source line is within [lo,hi] of a file.
usage: ncu_sass_region.py rep cubin kernel-substr file lo hi [maxlines]"""
import csv, re, subprocess, sys
rep, cubin, kname, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
mx = int(sys.argv[7]) if len(sys.argv) > 7 else 400
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr = rows[1]; ia, isamp, iexec, isrc = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
body = rows[2:]; base = int(body[0][ia], 16)
samp = {int(r[ia],16)-base: (int(r[isamp]), int(r[iexec]), r[isrc].strip()) for r in body if r[ia].startswith("0x")}
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
infn=False; cur=("?",0); n=0
for ln in dis:
    if ".text." in ln: infn = (kname in ln)
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(\S.*?);', ln)
    if m and cur[0]==fname and lo<=cur[1]<=hi:
        off=int(m.group(1),16); s=samp.get(off,(0,0,""))
        print(f"{cur[1]:4d} {off:6x} {s[0]:6d} {s[1]:9d}  {m.group(2)[:90]}")
        n+=1
        if n>=mx: break
