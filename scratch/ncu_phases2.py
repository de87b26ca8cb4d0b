"""Bucket an ncu source page by (file, line range). usage: ncu_phases2.py rep launch 'name:file:lo-hi,...'"""
import csv, io, subprocess, sys
rep=sys.argv[1]; li=sys.argv[2]
ranges=[]
for part in sys.argv[3].split(','):
    name,f,r=part.split(':'); lo,hi=r.split('-'); ranges.append((name,f,int(lo),int(hi)))
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--launch-skip',li,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur=None; hdr=None; acc={}; ti=ts=0
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0].isdigit():
        off=len(r)-len(hdr)
        try: ins=int(r[hdr.index('Instructions Executed')+off] or 0); smp=int(r[hdr.index('# Samples')+off] or 0)
        except ValueError: continue
        ln=int(r[0]); key='other:'+cur
        for name,f,lo,hi in ranges:
            if cur==f and lo<=ln<=hi: key=name; break
        a=acc.setdefault(key,[0,0]); a[0]+=ins; a[1]+=smp; ti+=ins; ts+=smp
print(f"total instr {ti:,} samples {ts:,}")
for k,(i,s) in sorted(acc.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:28s} {100*i/ti:5.1f}% ins {100*s/ts:5.1f}% smp")
