#!/usr/bin/env python
"""Map of a kernel's SASS by how often it runs: usage  region_map.py <ncu --page source --csv export> [min run length]
Consecutive instructions are grouped by the fraction of warps that execute them (dead, <1 %, 1-20 %, 20-90 %, once
per warp, loop); per group: length, address, warp instructions per warp, active lanes, share of the warp-state
samples and the share that were instruction-fetch stalls (stall_no_inst).  The last line is the hot static
footprint (groups executed by at least 20 % of the warps), the number that has to fit the instruction caches."""
import csv,sys
f=sys.argv[1]; minlen=int(sys.argv[2]) if len(sys.argv)>2 else 6
rows=list(csv.reader(open(f))); hdr=rows[1]; data=rows[2:]
ia=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isrc=hdr.index("Source"); iaddr=hdr.index("Address"); isamp=hdr.index("# Samples"); ino=hdr.index("stall_no_inst")
W=int(data[0][ia]); base=int(data[0][iaddr],16)
def bucket(n):
    f=n/W
    if n==0: return 'dead'
    if f<0.01: return '<1%'
    if f<0.2: return '1-20%'
    if f<0.9: return '20-90%'
    if f<1.5: return '100%'
    return 'loop'
runs=[];cur=None
for i,r in enumerate(data):
    b=bucket(int(r[ia]))
    if b!=cur: runs.append([b,i,i,0,0,0,0]); cur=b
    x=runs[-1]; x[2]=i; x[3]+=int(r[ia]); x[4]+=int(r[it]); x[5]+=int(r[isamp]); x[6]+=int(r[ino])
ts=sum(int(r[isamp]) for r in data)
hot=0
for b,s,e,n,t,sm,no in runs:
    if b in('100%','loop','20-90%'): hot+=e-s+1
    if b!='dead' and (e-s+1)>=minlen:
        print(f"{b:7s} {s:5d}-{e:5d} len {e-s+1:4d} addr {int(data[s][iaddr],16)-base:#7x} inst/warp {n/W:7.1f} lanes {t/max(n,1):5.1f} samples {100*sm/ts:5.2f}% no_inst {100*no/ts:5.2f}%  {data[s][isrc].strip()[:36]}")
print("hot static",hot, "KB",hot*16/1024)
