"""Per-source-line memory traffic from an ncu report: usage ncu_mem.py rep launch_index [top]"""
import csv, io, subprocess, sys
rep=sys.argv[1]; li=int(sys.argv[2]); top=int(sys.argv[3]) if len(sys.argv)>3 else 30
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--launch-skip',str(li),'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
cur=None; hdr=None; L=[]
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if r[0]=='Line No': hdr=r; continue
    if hdr and r[0].isdigit():
        off=len(r)-len(hdr)
        def col(n):
            v=r[hdr.index(n)+off]
            try: return int(v or 0)
            except ValueError: return 0
        L.append((cur,int(r[0]),','.join(r[1:2+off]).strip(),col('L1 Tag Requests Global'),col('L1 Wavefronts Shared'),col('L2 Theoretical Sectors Global'),col('L2 Theoretical Sectors Local'),col('Instructions Executed')))
tg=sum(l[3] for l in L); ts=sum(l[4] for l in L); t2=sum(l[5] for l in L); tl=sum(l[6] for l in L)
print(f"L1 tag requests global {tg:,}  smem wavefronts {ts:,}  L2 theoretical sectors global {t2:,} local {tl:,}")
key=(lambda l:-l[6]) if len(sys.argv)>4 and sys.argv[4]=="loc" else (lambda l:-(l[3]+l[4]))
L.sort(key=key)
for f,n,src,g,s,l2,ll,ins in L[:top]:
    print(f"{f[:17]:17s}:{n:4d} tagG {100*g/max(tg,1):5.1f}% smemW {100*s/max(ts,1):5.1f}% L2sec {100*l2/max(t2,1):5.1f}% loc {ll:9d} | {src[:90]}")
