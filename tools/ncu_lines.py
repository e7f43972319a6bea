#!/usr/bin/env python3
"""Per source line: warp instructions executed and stall samples, from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py X.ncu-rep [top N]"""
import csv, io, subprocess, sys, collections
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
inst = collections.Counter(); samp = collections.Counter(); srcs = {}
cur_file = None; hdr = None
tot_i = tot_s = 0
for row in csv.reader(io.StringIO(raw)):
    if not row: continue
    if row[0] == 'File Path': cur_file = row[1].split('/')[-1]; continue
    if row[0] == 'Function Name': continue
    if row[0] == 'Line No': hdr = row; continue
    if hdr is None: continue
    try:
        ln = int(row[0])
    except ValueError:
        continue
    d = dict(zip(hdr, row))
    # rows with an Address are SASS rows of that line
    i_ = d.get('Instructions Executed', ''); s_ = d.get('# Samples', '')
    if row[2] not in ('', '-'):
        continue
    key = (cur_file, ln)
    srcs[key] = row[1].strip()[:110]
    try: inst[key] += int(i_); tot_i += int(i_)
    except ValueError: pass
    try: samp[key] += int(s_); tot_s += int(s_)
    except ValueError: pass
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("== by instructions")
for k, v in inst.most_common(top): print(f"{v:12d} {100*v/max(tot_i,1):5.1f}%  samp {100*samp[k]/max(tot_s,1):5.1f}%  {k[0]}:{k[1]}  {srcs[k]}")
print("== by stall samples")
for k, v in samp.most_common(top): print(f"{v:12d} {100*v/max(tot_s,1):5.1f}%  {k[0]}:{k[1]}  {srcs[k]}")
