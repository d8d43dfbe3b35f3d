"""Cost-model search over dense-block schedules (which launch computes which (conv, input-chunk) product): every
partition of the triangle {(k, c): c < k} into five rectangles (launch j completes conv j), MMA time from the measured
tcgen05.mma cost (72 cycles for N <= 96, 18 + N/2 above), HBM time at 4.5 TB/s, shared-memory feasibility of the
resident filters.  engine.SCHED2 is the best assignment with non-adjacent input chunks (second block of output)."""
import itertools
cout = {1:32,2:32,3:32,4:32,5:64}
chunkB = {0:128,1:64,2:64,3:64,4:64}   # bytes/pixel of input chunk (bf16)
nch = {0:2,1:1,2:1,3:1,4:1}
def mma_cost(N):
    return 72 if N <= 96 else 18 + N/2
TILES=55.35; GHZ=1.8; BW=4.5e12; PIX=16*256*256; INEFF=1.29
best=[]
# assignment a(k,c) in [c+1..k]; require rectangles
pairs=[(k,c) for k in range(1,6) for c in range(0,k)]
ranges=[range(c+1,k+1) for (k,c) in pairs]
cnt=0
for assign in itertools.product(*ranges):
    a=dict(zip(pairs,assign))
    ok=True
    L={}
    for l in range(1,6):
        Cs=sorted({c for (k,c),v in a.items() if v==l}); Ks=sorted({k for (k,c),v in a.items() if v==l})
        # rectangle check
        for k in Ks:
            for c in Cs:
                if c<k and a[(k,c)]!=l: ok=False
                if c>=k: ok=False
        L[l]=(Cs,Ks)
        if not ok: break
    if not ok: continue
    # every conv k must have its own chunk c_{k-1} at launch k (x_{k-1} only available then) - implied by range
    tot=0; detail=[]
    for l in range(1,6):
        Cs,Ks=L[l]
        if not Ks:
            ok=False;break
        N=sum(cout[k] for k in Ks)
        # n-tiling if weights don't fit: weights bytes = 9*chunks*N*64
        chunks=sum(nch[c] for c in Cs)
        wbytes=9*chunks*N*64
        ntile=1
        while wbytes/ntile>150*1024: ntile+=1
        if N%ntile or (N//ntile)%16: 
            # find next valid
            while N%ntile or (N//ntile)%16: ntile+=1
        nt=N//ntile
        mma=TILES*ntile*chunks*18*mma_cost(nt)/ (GHZ*1e9) * (1 if ntile==1 else 1)  # each ntile has 148/ntile CTAs -> tiles per CTA x ntile
        # bytes: A read per n-tile pass
        rd=sum(chunkB[c] for c in Cs)*ntile
        pre=sum(cout[k]*2 for k in Ks if any(a[(k,c)]<l for c in range(0,k)))
        wr=N*2
        res=128 if l==5 else 0
        hbm=(rd+pre+wr+res)*PIX/BW
        t=max(mma*INEFF,hbm)
        tot+=t; detail.append((Cs,Ks,nt,ntile,round(mma*1e6),round(hbm*1e6)))
    if ok: best.append((tot,detail)); cnt+=1
best.sort(key=lambda x:x[0])
print(cnt,'valid')
for tot,d in best[:6]:
    print(round(tot*1e6),'us')
    for x in d: print('   ',x)
# current fused and per-layer for reference

print('==== with smem + contiguity constraints')
def evaluate(L, a, contiguous_chunks):
    tot=0; detail=[]
    for l in range(1,6):
        Cs,Ks=L[l]
        if not Ks: return None
        if Ks != list(range(Ks[0], Ks[-1]+1)): return None
        if contiguous_chunks and Cs != list(range(Cs[0], Cs[-1]+1)): return None
        N=sum(cout[k] for k in Ks)
        chunks=sum(nch[c] for c in Cs)
        nres = 1 if l==5 else 0
        bestt=None
        for ntile in (1,2,3,4):
            if N%ntile or (N//ntile)%32: continue
            nt=N//ntile
            w=9*chunks*nt*64
            smem = w + 2*nt*256*(1+nres) + 4*12288 + 3000
            if smem > 225*1024: continue
            mma=TILES*ntile*chunks*18*mma_cost(nt)/(GHZ*1e9)
            rd=sum(chunkB[c] for c in Cs)*ntile
            pre=sum(cout[k]*2 for k in Ks if any(a[(k,c)]<l for c in range(0,k)))
            wr=N*2
            res=128 if l==5 else 0
            hbm=(rd+pre+wr+res)*PIX/BW
            t=max(mma*INEFF,hbm)
            if bestt is None or t<bestt[0]: bestt=(t,(Cs,Ks,nt,ntile,round(mma*INEFF*1e6),round(hbm*1e6)))
        if bestt is None: return None
        tot+=bestt[0]; detail.append(bestt[1])
    return tot,detail
for contig in (True, False):
    res=[]
    for assign in itertools.product(*ranges):
        a=dict(zip(pairs,assign))
        ok=True; L={}
        for l in range(1,6):
            Cs=sorted({c for (k,c),v in a.items() if v==l}); Ks=sorted({k for (k,c),v in a.items() if v==l})
            for k in Ks:
                for c in Cs:
                    if c>=k or a[(k,c)]!=l: ok=False
            L[l]=(Cs,Ks)
            if not ok: break
        if not ok: continue
        r=evaluate(L,a,contig)
        if r: res.append(r)
    res.sort(key=lambda x:x[0])
    print('contiguous chunks' if contig else 'any chunk sets', len(res))
    for tot,d in res[:4]:
        print(' ',round(tot*1e6),'us')
        for x in d: print('     ',x)
