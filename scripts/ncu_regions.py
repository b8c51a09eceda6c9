"""Per-region wall-time estimate of k_fused from an ncu source-page CSV (cuda,sass view).
usage: ncu_regions.py file.csv n_ctas warps_per_cta"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
nctas = int(sys.argv[2]); wpc = int(sys.argv[3])
src = open('/root/repo/yalrf_b200/csrc/yhb_kernels.cuh').read().splitlines()
def line_of(pat, after=0):
    for i, l in enumerate(src):
        if i >= after and pat in l: return i + 1
    raise KeyError(pat)
marks = [('eval_pass', line_of('__device__ inline int eval_pass')), ('advance/apply', line_of('__device__ inline int advance')),
         ('k_eval..jelem', line_of('k_eval(PlanDev')), ('jelem', line_of('__device__ __forceinline__ double th_elem')),
         ('assemble/ge', line_of('constexpr int kAsmRows')), ('newton pre/post', line_of('__device__ __forceinline__ double lin_dot')),
         ('staged', line_of('k_assemble_core(PlanDev')), ('gj factor', line_of('__device__ __forceinline__ void gj_factor_panel')),
         ('gj loop', line_of('__device__ __forceinline__ void gj_solve_mma')), ('gj stage', line_of('k_gj_stage(int batch')),
         ('fused body', line_of('k_fused(const __grid_constant__')), ('end', len(src) + 1)]
hdr = None; cur = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]
    if len(r) > 8 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] not in ('', 'Line No'):
        d = dict(zip(hdr, r))
        try: out.append((cur, int(r[0]), int(d['Instructions Executed']), int(d['# Samples'])))
        except ValueError: pass
tot = sum(o[3] for o in out); toti = sum(o[2] for o in out)
print('samples %d  instr %d  samples per warp-slot %.1f' % (tot, toti, tot / nctas / wpc))
for k in range(len(marks) - 1):
    lo, hi = marks[k][1], marks[k + 1][1] - 1
    s = sum(o[3] for o in out if o[0] == 'yhb_kernels.cuh' and lo <= o[1] <= hi)
    i = sum(o[2] for o in out if o[0] == 'yhb_kernels.cuh' and lo <= o[1] <= hi)
    print('%-16s %5d-%5d  samples %5.1f%%  instr %5.1f%%' % (marks[k][0], lo, hi, 100. * s / tot, 100. * i / toti))
oth = sum(o[3] for o in out if o[0] != 'yhb_kernels.cuh')
print('other files      samples %5.1f%%' % (100. * oth / tot))
