"""Warp-stall samples of one kernel aggregated by SOURCE LINE and stall reason, from an `ncu --set full --import-source on`
report.  Needs the SASS->line map of the same build:
  nvcc ... -lineinfo -c jorldy_b200/csrc/ppo_fused.cu -o /tmp/pf.o && cuobjdump -xelf all /tmp/pf.o && nvdisasm -g -c ppo_fused.sm_100a.cubin > /tmp/pf.dis
usage: python scripts/ncu_stall_by_source.py <report.ncu-rep> [line bucket=5] [rows=40]"""
import csv,re,collections,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.split('\n')))
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr)]
ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
cur=None; amap={}
for l in open('/tmp/pf.dis'):
    m=re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)), (m.group(3) or '').split('/')[-1], int(m.group(4) or 0)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/',l)
    if m: amap[int(m.group(1),16)]=cur
addrs=[int(r[ix['Address']],16) if r[ix['Address']].startswith('0x') else int(r[ix['Address']]) for r in data]
base=min(addrs)
G=int(sys.argv[2]) if len(sys.argv)>2 else 5
bysrc=collections.Counter(); bystall=collections.defaultdict(collections.Counter); tot=collections.Counter(); S=0
for r,a in zip(data,addrs):
    try: n=int(r[ix['# Samples']])
    except: continue
    S+=n
    src=amap.get(a-base)
    if src is None: key=('?',0)
    else:
        f,ln,f2,ln2=src
        key=(f2,ln2//G*G) if f2=='ppo_fused.cu' else (f,ln//G*G) if f=='ppo_fused.cu' else (f+'<-'+f2, ln2//G*G)
    bysrc[key]+=n
    for h in stalls:
        try: c=int(r[ix[h]])
        except: c=0
        bystall[key][h]+=c; tot[h]+=c
print('samples',S)
print(', '.join(f'{k[6:]}:{100*v/S:.1f}%' for k,v in tot.most_common(10)))
for k,v in bysrc.most_common(int(sys.argv[3]) if len(sys.argv)>3 else 40):
    top=', '.join(f'{h[6:]}:{c}' for h,c in bystall[k].most_common(3))
    print(f'{str(k):42s}{v:7d} {100*v/S:5.1f}%  {top}')
