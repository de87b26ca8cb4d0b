"""Summarise an ncu report per CUDA source line: instructions executed and stall samples.
usage: python scratch/ncu_lines.py report.ncu-rep [launch_index] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; li = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--launch-skip',str(li),'--launch-count','1'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and r[0].isdigit():
        off = len(r) - len(hdr)   # broken quoting in source text shifts columns: align from the right
        def col(name):
            return r[hdr.index(name) + off]
        try:
            lines.append((cur_file, int(r[0]), ','.join(r[1:2+off]).strip(), int(col('Instructions Executed') or 0), int(col('# Samples') or 0), col('Avg. Threads Executed')))
        except ValueError:
            pass
tot_i = sum(l[3] for l in lines); tot_s = sum(l[4] for l in lines)
print(f"total instr {tot_i:,}  samples {tot_s:,}")
key = 4 if len(sys.argv) > 4 and sys.argv[4] == "smp" else 3
lines.sort(key=lambda l: -l[key])
for f, n, src, ins, smp, thr in lines[:top]:
    print(f"{f}:{n:4d} {100*ins/tot_i:5.1f}% ins {100*smp/max(tot_s,1):5.1f}% smp thr={thr:>4} | {src[:110]}")
