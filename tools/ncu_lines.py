#!/usr/bin/env python
"""Per-source-line shares of executed instructions and stall samples of one kernel.

ncu's CSV export of the source page is per SASS address; this joins it with the line table
of the same build (kernels are compiled with -lineinfo):

    cuobjdump -xelf all rfsi_b200/librfsi_b200.so            # -> engine.sm_100a.cubin
    nvdisasm -g -c engine.sm_100a.cubin > /tmp/all_dis.txt
    ncu -i capture.ncu-rep --page source --csv > /tmp/src.csv
    python tools/ncu_lines.py qrf_quantile_kernelILi16E /tmp/src.csv 30 [/tmp/all_dis.txt]
"""
import re,csv,collections,sys
kern=sys.argv[1]; srccsv=sys.argv[2]
txt=open(sys.argv[4] if len(sys.argv)>4 else '/tmp/all_dis.txt').read().split('\n')
start=None
for i,l in enumerate(txt):
    if l.startswith('.text.') and kern in l: start=i;break
line=None; addr2line={}
for l in txt[start+1:]:
    if l.startswith('.text.') or l.startswith('.section'):
        if addr2line: break
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: line=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: addr2line[int(m.group(1),16)]=(line,m.group(2))
rows=list(csv.reader(open(srccsv)))
h=rows[1]; ia=h.index('Address'); ie=h.index('Instructions Executed'); iss=h.index('Warp Stall Sampling (All Samples)')
base=None
per=collections.Counter(); stall=collections.Counter(); tot=0
for r in rows[2:]:
    try: a=int(r[ia],16)
    except: continue
    if base is None: base=a
    off=a-base
    ex=int(r[ie] or 0); st=int(r[iss] or 0)
    ln=addr2line.get(off,(None,''))[0]
    per[ln]+=ex; stall[ln]+=st; tot+=ex
import glob
src={}
for f in glob.glob(__import__('os').path.join(__import__('os').path.dirname(__import__('os').path.abspath(__file__)), '..', 'rfsi_b200', 'csrc', '*.cuh')):
    src[f.split('/')[-1]]=open(f).read().split('\n')
print('total',tot)
ts=sum(stall.values())
for ln,c in per.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 30):
    s=src.get(ln[0],[''])[ln[1]-1].strip()[:90] if ln and ln[0] in src else ''
    print(f"{c/tot*100:5.1f}% inst {stall[ln]/ts*100:5.1f}% stall  {ln}  {s}")
