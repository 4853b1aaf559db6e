import sys
ev=[tuple(map(int,l.split())) for l in open(sys.argv[1])]
t0=min(t for r,e,t in ev)
ev=sorted(((t-t0)&0xffffffff,r,e) for r,e,t in ev)
def nm(e):
    if e in (250,251): return 'MMA  qfull %s'%'AB'[e-250]
    if e in (600,601): return 'SM%s  ofull'%'AB'[e-600]
    base=e//100*100
    if base==700: return ['    ld done','    max done','    turn ok','    exp done'][e-700]
    if base==100: return f"PROD kv_empty ok, load j={e-100}"
    if base==200: return f"MMA  kvfull j={e-200}"
    x=(e-base)//10; j=(e-base)%10
    return {300:f"MMA  Pready {'AB'[x]} j={j}",400:f"SM{'AB'[x]}  Sfull j={j}",500:f"SM{'AB'[x]}  Parrive j={j}"}[base]
lo=int(sys.argv[2]) if len(sys.argv)>2 else 0
hi=int(sys.argv[3]) if len(sys.argv)>3 else 100
prev=None
for t,r,e in ev[lo:hi]:
    print(f"{t:8d} (+{0 if prev is None else t-prev:5d}) {'  '*r}{nm(e)}")
    prev=t
print("total span", ev[-1][0], "events", len(ev))
