"""Brute-force search over dense-block schedules (which launch computes which (conv k, input chunk c) product).

Cost model per launch (two resources, profiles/r2_summary.md): MMA-issue time = chunk passes x 18 MMAs x c2(N) cycles with the
measured CTA-pair cost c2(N) = max(64, N/2); HBM time = 64 B-per-pixel "slabs" moved (A chunks + partial-sum reads in 64-channel
blocks + writes + residual) x SLAB_US; launch time = max of the two + 4 us.  34,560 assignments; prints the best ones and where
engine.SCHED2 / SCHED3 stand.  Constants: UNIT_US = us per chunk pass at c2 = 64 (33.6 at 1.9 GHz, 43 at the power-capped 1.48 GHz),
SLAB_US = 14 (67 MB per slab at ~4.8 TB/s)."""
import itertools
# products: (k, c): conv k in 1..5, chunk c in 0..k-1 (c=0: x, two 32-ch units; c>=1: x_c one unit)
prods=[(k,c) for k in range(1,6) for c in range(0,k)]
choices={(k,c): list(range(c+1 if c>0 else 1, k+1)) for (k,c) in prods}
width={1:1,2:1,3:1,4:1,5:2}
units={0:2,1:1,2:1,3:1,4:1}
def c2(N, pair=True):
    if pair: return max(64, N/2)
    return max(84, N/2+5) if N<=256 else 1e9
SLAB_US=14.0   # us per 64B/px slab per RDB-stage at ~4.8 TB/s
UNIT_US=33.6/64  # us per chunk-unit per cycle of MMA cost
def cost(assign, pair=True, verbose=False):
    tot=0; det=[]
    first={}  # conv k -> first stage touching it
    for s in range(1,6):
        ps=[p for p in prods if assign[p]==s]
        chunks=sorted({c for (_,c) in ps}); convs=sorted({k for (k,_) in ps})
        if s not in convs: convs=sorted(set(convs)|{s})  # conv s completes here (always has product (s,s-1) anyway)
        N=32*sum(width[k] for k in convs); nu=sum(units[c] for c in chunks)
        # smem limit for resident filters
        if pair:
            if nu*N>384 and not (nu*N<=400): return 1e9,None
        else:
            if nu*N>256: return 1e9,None
        mma=nu*c2(N,pair)*UNIT_US
        rd=nu; wr=0
        for k in convs:
            if k in first: rd+=width[k]     # partial read
            else: first[k]=s
            wr+=width[k]
        if s==5: rd+=2   # residual x
        hbm=(rd+wr)*SLAB_US
        t=max(mma,hbm)+4.0
        tot+=t; det.append((s,chunks,convs,N,nu,round(mma),round(hbm)))
    return tot,det
best=[]
keys=prods
for combo in itertools.product(*[choices[p] for p in keys]):
    a=dict(zip(keys,combo))
    t,d=cost(a,True)
    if t<1e8: best.append((t,combo))
best.sort()
print(len(best))
for t,combo in best[:6]:
    a=dict(zip(keys,combo)); print(round(t,1)); 
    for row in cost(a)[1]: print('   ',row)
# reference schedules
s2={ (k,0):1 for k in range(1,6)}
s2.update({(2,1):2,(3,1):2,(3,2):3,(4,2):3,(4,1):4,(5,1):4,(4,3):4,(5,3):4,(5,2):5,(5,4):5})
print('sched2',cost(s2))
s1={(k,c):(c+1) for (k,c) in prods}
print('sched1',cost(s1))

print('---- block-granular pre model')
def cost2(assign, verbose=False):
    tot=0; det=[]; first={}; slabs=0
    for s in range(1,6):
        ps=[p for p in prods if assign[p]==s]
        chunks=sorted({c for (_,c) in ps}); convs=sorted({k for (k,_) in ps})
        if s not in convs: convs=sorted(set(convs)|{s})
        if convs!=list(range(convs[0],convs[-1]+1)) or convs[0]!=s: return 1e9,None   # contiguous slots starting at conv s
        N=32*sum(width[k] for k in convs); nu=sum(units[c] for c in chunks)
        if nu*N>384: return 1e9,None
        mma=nu*c2(N,True)*UNIT_US
        # column layout in 32-ch units
        cols=[]
        for k in convs: cols+= [k]*width[k]
        rd=nu; wr=len(cols)
        # 64-ch blocks
        for b0 in range(0,len(cols),2):
            blk=cols[b0:b0+2]
            if any(k in first for k in blk): rd+=len(blk)
        for k in convs:
            if k not in first: first[k]=s
        if s==5: rd+=2
        hbm=(rd+wr)*SLAB_US
        t=max(mma,hbm)+4.0
        tot+=t; slabs+=rd+wr; det.append((s,chunks,convs,N,nu,round(mma),round(hbm)))
    return tot,(det,slabs)
best=[]
for combo in itertools.product(*[choices[p] for p in keys]):
    a=dict(zip(keys,combo))
    t,d=cost2(a)
    if t<1e8: best.append((t,combo))
best.sort()
print(len(best))
for t,combo in best[:5]:
    a=dict(zip(keys,combo)); print(round(t,1), 'slabs',cost2(a)[1][1]); 
    for row in cost2(a)[1][0]: print('   ',row)
print('sched2',cost2(s2)[0],cost2(s2)[1][1])

print('---- pair-only (N % 64 == 0), block-granular')
best=[]
for combo in itertools.product(*[choices[p] for p in keys]):
    a=dict(zip(keys,combo))
    ok=True
    for s_ in range(1,6):
        ps=[p for p in prods if a[p]==s_]
        convs=sorted({k for (k,_) in ps}|{s_})
        N=32*sum(width[k] for k in convs)
        if N%64: ok=False;break
    if not ok: continue
    t,d=cost2(a)
    if t<1e8: best.append((t,combo))
best.sort()
print(len(best))
for t,combo in best[:4]:
    a=dict(zip(keys,combo)); print(round(t,1), 'slabs',cost2(a)[1][1]); 
    for row in cost2(a)[1][0]: print('   ',row)
