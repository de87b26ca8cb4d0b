"""Bucket an ncu source page by line ranges of render.cu. usage: ncu_phases.py rep 'name:lo-hi,...'"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ranges = []
for part in sys.argv[2].split(','):
    name, r = part.split(':'); lo, hi = r.split('-'); ranges.append((name, int(lo), int(hi)))
out = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file=None; hdr=None; acc={}
tot_i=tot_s=0
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0].isdigit():
        off=len(r)-len(hdr)
        try:
            ins=int(r[hdr.index('Instructions Executed')+off] or 0); smp=int(r[hdr.index('# Samples')+off] or 0)
        except ValueError: continue
        ln=int(r[0]); key='other:'+cur_file
        if cur_file=='render.cu':
            key='render.cu:unbucketed'
            for name,lo,hi in ranges:
                if lo<=ln<=hi: key=name; break
        a=acc.setdefault(key,[0,0]); a[0]+=ins; a[1]+=smp; tot_i+=ins; tot_s+=smp
for k,(i,s) in sorted(acc.items(), key=lambda kv:-kv[1][1]):
    print(f"{k:28s} {100*i/tot_i:5.1f}% ins {100*s/tot_s:5.1f}% smp")
