"""Aggregate scripts/ncu_by_line.py output of the CTA DMMA kernel by code region. usage: ncu_regions_c5.py lines.txt"""
import collections, re, sys
rows = []
for l in open(sys.argv[1]):
    m = re.match(r"\s*([\d.]+)%\s+(\d+)\s+(\S+):(\d+)", l)
    if m:
        rows.append((float(m.group(1)), m.group(3), int(m.group(4))))
src = open("mpc-sensorlessao_b200/csrc/fmpc_kernel_mma.cu").read().splitlines()
def find(s):
    for i, l in enumerate(src):
        if s in l:
            return i + 1
    raise KeyError(s)
marks = [("dmma asm", 40, 45), ("tile_nt", find("void tile_nt"), find("void tile_nt") + 15),
         ("mma_gA_sB", find("void mma_gA_sB"), find("void mma_gA_sB") + 25),
         ("cta_potrf_inverse", find("Whole CTA (n > 32)"), find("Everything a phase needs")),
         ("apply_C / apply_Ct helpers", find("Everything a phase needs"), find("fmpc_solve_kernel_mma(const DevSys")),
         ("kernel: init .. residual pass", find("fmpc_solve_kernel_mma(const DevSys"), find("D_t = B diag(w_t) B' for all stages")),
         ("kernel: G GEMM + zero ops", find("D_t = B diag(w_t) B' for all stages"), find("band-2 block Cholesky of Y fused")),
         ("kernel: phase 1 (S / M1 tiles, rhs)", find("band-2 block Cholesky of Y fused"), find("if constexpr (NP > 32)")),
         ("kernel: potrf call + sync", find("if constexpr (NP > 32)"), find("phase 3: L1_i = M1 inv(L)'")),
         ("kernel: phase 3 + y + rotate", find("phase 3: L1_i = M1 inv(L)'"), find("backward solve  dnu_i")),
         ("kernel: backward", find("backward solve  dnu_i"), find("dz = inv(Phi)(-r_d - C' dnu)")),
         ("kernel: dz + line search + rest", find("dz = inv(Phi)(-r_d - C' dnu)"), len(src) + 1)]
agg = collections.OrderedDict((m[0], 0.0) for m in marks)
other = 0.0
for pct, f, ln in rows:
    if f != "fmpc_kernel_mma.cu":
        other += pct
        continue
    for name, a, b in marks:
        if a <= ln < b:
            agg[name] += pct
            break
    else:
        other += pct
for k, v in agg.items():
    print(f"{v:6.1f}%  {k}")
print(f"{other:6.1f}%  other files / lines")
