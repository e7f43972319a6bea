#!/usr/bin/env python3
"""Per-source-line instruction / stall-sample totals of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 8 and r[0] not in ('', '-'):
        try:
            i_inst = hdr.index('Instructions Executed'); i_smp = hdr.index('# Samples'); i_thr = hdr.index('Thread Instructions Executed')
            out.append((int(r[i_inst]), int(r[i_smp]), int(r[i_thr]), fname, r[0], r[1].strip()[:110]))
        except Exception:
            pass
tot_i = sum(o[0] for o in out) or 1; tot_s = sum(o[1] for o in out) or 1
print(f'total warp-instr {tot_i}  samples {tot_s}')
print('--- by stall samples (time)')
for o in sorted(out, key=lambda x: -x[1])[:top]:
    print(f'{100*o[1]/tot_s:5.1f}% smp {100*o[0]/tot_i:5.1f}% inst lanes {o[2]/max(o[0],1):4.1f}  {o[3]}:{o[4]}  {o[5]}')
