"""Prototype: warm start of candidate k+1 from candidate k's optimal flow + potentials (same scenario).
Counts augmentations / dual updates / searches of the repair; checks the objective against a cold solve."""
import os, sys, numpy as np
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0,HERE); sys.path.insert(0,ROOT)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    from reduce import graph, I
from dfs_proto import ssp_value
INF=1<<29; M=1<<19
class G:
    def __init__(s,sv,ev,r,nc,chains):
        s.sv=list(map(int,sv)); s.ev=list(map(int,ev)); s.r=list(map(int,r)); s.nc=nc; s.n=len(sv); s.chains=chains
    def arcs(s,c):  # residual arcs of chain c: (tail, head, dir, cost)
        hf = s.nc if s.ev[c]==0 else s.ev[c]; hb = s.nc if s.sv[c]==0 else s.sv[c]
        return ((s.sv[c],hf,0,-s.r[c]),(s.ev[c],hb,1,s.r[c]))
def resid(g,cap,x,c,d): return cap[c]-x[c] if d==0 else x[c]
def search(g,cap,x,lab,src,extra=()):
    R={src}|set(extra); pred={}; grew=True
    while grew:
        grew=False
        for c in range(g.n):
            for (t,h,d,cost) in g.arcs(c):
                if t in R and h not in R and resid(g,cap,x,c,d)>0 and lab[t]+cost==lab[h]:
                    R.add(h); pred[h]=(t,c,d); grew=True
    return R,pred
def dual_update(g,cap,x,lab,R):
    best=INF
    for c in range(g.n):
        for (t,h,d,cost) in g.arcs(c):
            if t in R and h not in R and resid(g,cap,x,c,d)>0: best=min(best,lab[t]+cost-lab[h])
    return best
def route(g,cap,x,lab,src,dst,amount,st,extra=()):
    """push `amount` from src to dst along tight paths with dual updates; returns False if stuck"""
    while amount>0:
        R,pred=search(g,cap,x,lab,src,extra); st['searches']+=1
        while dst not in R:
            d=dual_update(g,cap,x,lab,R); st['duals']+=1
            if d>=INF: return False
            assert d>=0
            for v in range(g.nc+1):
                if v not in R: lab[v]+=d
            R,pred=search(g,cap,x,lab,src,extra); st['searches']+=1
        v=dst; bn=amount; path=[]
        while v!=src:
            t,c,d=pred[v]; bn=min(bn,resid(g,cap,x,c,d)); path.append((c,d)); v=t
        for c,d in path: x[c]+= bn if d==0 else -bn
        amount-=bn; st['pushes']+=1
    return True
def cold(g,cap,st):
    x=[0]*g.n; lab=[M]*(g.nc+1); lab[0]=0
    ch=True
    while ch:
        ch=False
        for c in range(g.n):
            t,h,d,cost=g.arcs(c)[0]
            if cap[c]>0 and lab[t]+cost<lab[h]: lab[h]=lab[t]+cost; ch=True
    while lab[g.nc]<0:
        R,pred=search(g,cap,x,lab,0); st['searches']+=1
        if g.nc in R:
            v=g.nc; bn=INF; path=[]
            while v!=0:
                t,c,d=pred[v]; bn=min(bn,resid(g,cap,x,c,d)); path.append((c,d)); v=t
            for c,d in path: x[c]+= bn if d==0 else -bn
            st['pushes']+=1
        else:
            d=dual_update(g,cap,x,lab,R); st['duals']+=1
            if d>=INF: d=-lab[g.nc]
            d=min(d,-lab[g.nc])
            for v in range(g.nc+1):
                if v not in R: lab[v]+=d
    return x,lab
def canon(g,cap,x):
    """feasible potentials for the final residual graph: BF distances from merged root, unreachable nodes M-consistent"""
    lab=[M]*(g.nc+1); lab[0]=0
    ch=True
    while ch:
        ch=False
        for c in range(g.n):
            for (t,h,d,cost) in g.arcs(c):
                if h==g.nc: continue
                if resid(g,cap,x,c,d)>0 and lab[t]+cost<lab[h]: lab[h]=lab[t]+cost; ch=True
    lab[g.nc]=0
    return lab
def warm(gp,xp,labp,g,cap,st):
    e=[0]*(g.nc+1); x=[0]*g.n
    prev={ch:i for i,ch in enumerate(gp.chains)}
    kept=set()
    for c,chn in enumerate(g.chains):
        if chn in prev: x[c]=xp[prev[chn]]; kept.add(prev[chn])
    for c in range(gp.n):
        if c not in kept and xp[c]>0:
            st['removed_flow']+=1
            if gp.sv[c]!=0: e[gp.sv[c]]+=xp[c]
            if gp.ev[c]!=0: e[gp.ev[c]]-=xp[c]
    st['newchains']+=g.n-len(kept)
    lab=list(labp)
    for c in range(g.n):
        (t,h,_,cost),(t2,h2,_,cost2)=g.arcs(c)
        x[c]=min(x[c],cap[c])
        if x[c]<cap[c] and lab[t]+cost<lab[h]:
            d=cap[c]-x[c]; x[c]=cap[c]; st['sat']+=1
            if g.sv[c]!=0: e[g.sv[c]]-=d
            if g.ev[c]!=0: e[g.ev[c]]+=d
        elif x[c]>0 and lab[t2]+cost2<lab[h2]:
            d=x[c]; x[c]=0; st['sat']+=1
            if g.sv[c]!=0: e[g.sv[c]]+=d
            if g.ev[c]!=0: e[g.ev[c]]-=d
    for v in range(1,g.nc):
        if e[v]>0:
            st['imb']+=1
            if not route(g,cap,x,lab,v,g.nc,e[v],st): return None
    for v in range(1,g.nc):
        if e[v]<0:
            st['imb']+=1
            if not route(g,cap,x,lab,0,v,-e[v],st,extra=(g.nc,)): return None
    assert lab[0]==lab[g.nc]
    return x,lab
if __name__=="__main__":
    name=sys.argv[1]; S=int(sys.argv[2]); K=int(sys.argv[3])
    inst = I.config2(S=S) if name=='c2' else I.config4(S=S)
    paths=np.load(os.path.join(ROOT,'sgufp_solver_b200','data','bench_candidates.npz'))['config2' if name=='c2' else 'config4']
    gs=[]
    for k in range(K):
        sv,ev,r,nc,ac=graph(inst,paths[k])
        chains=[[] for _ in range(len(sv))]
        for a in range(inst.m):
            if ac[a]>=0: chains[ac[a]].append(a)
        gs.append((G(sv,ev,r,nc,[tuple(c) for c in chains]),ac))
    cs=dict(searches=0,pushes=0,duals=0); ws=dict(searches=0,pushes=0,duals=0,removed_flow=0,newchains=0,sat=0,imb=0); nw=0; nc_=0
    for s in range(S):
        prevstate=None
        for k in range(K):
            g,ac=gs[k]
            cap=[INF]*g.n
            for a in range(inst.m):
                c=ac[a]
                if c>=0: cap[c]=min(cap[c],int(inst.upper[a,s]))
            x,lab=cold(g,cap,cs); nc_+=1
            obj=sum(g.r[c]*x[c] for c in range(g.n))
            if s==0 and k<2: assert obj==ssp_value(g.sv,g.ev,g.r,g.nc,cap)
            if prevstate is not None:
                gp,xp,labp=prevstate
                res=warm(gp,xp,labp,g,cap,ws); nw+=1
                assert res is not None
                xw,labw=res
                objw=sum(g.r[c]*xw[c] for c in range(g.n))
                assert objw==obj,(objw,obj)
                x=xw
            prevstate=(g,x,canon(g,cap,x))
    print(name,"cold per eval",{k:round(v/nc_,1) for k,v in cs.items()})
    print(name,"warm per eval",{k:round(v/nw,1) for k,v in ws.items()})
