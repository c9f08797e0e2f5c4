"""Prototype: primal-dual max-reward flow with a DFS over tight residual arcs (counts steps)."""
import os, sys, numpy as np
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0,HERE); sys.path.insert(0,ROOT)
from reduce import graph, I
INF=1<<29
def solve(sv,ev,r,nc,cap,revive=True,stats=None):
    n=len(sv); x=[0]*n
    # slots by tail: (head, c, dir, cost). tail index: root as source 0; arcs entering root -> head nc
    out=[[] for _ in range(nc+1)]
    for c in range(n):
        hf = nc if ev[c]==0 else ev[c]
        out[sv[c]].append((hf,c,0,-r[c]))
        hb = nc if sv[c]==0 else sv[c]
        out[ev[c]].append((hb,c,1,r[c]))   # tail ev (if ev==0 root: index 0 as source)
    # initial labels: Bellman-Ford from 0 on forward arcs (x=0)
    lab=[INF]*(nc+1); lab[0]=0
    ch=True
    while ch:
        ch=False
        for c in range(n):
            if cap[c]>0 and lab[sv[c]]<INF:
                hf = nc if ev[c]==0 else ev[c]
                if lab[sv[c]]-r[c]<lab[hf]: lab[hf]=lab[sv[c]]-r[c]; ch=True
    st=dict(steps=0,enters=0,pushes=0,phases=0,fresh=0,hops=0,unmarked=0)
    if lab[nc]>=0 or lab[nc]==INF: return x,st
    marked=[False]*(nc+1); stamp=[0]*(nc+1); par=[-1]*(nc+1); cnt=1
    def resid(c,d): return (cap[c]-x[c]) if d==0 else x[c]
    while True:
        # DFS from root until failure
        stack=[(0,-1,-1)]; marked[0]=True; stamp[0]=0
        while stack:
            v=stack[-1][0]
            st['steps']+=1
            nxt=None
            for (h,c,d,cost) in out[v]:
                if resid(c,d)>0 and lab[v]+cost==lab[h] and not marked[h]: nxt=(h,c,d); break
            if nxt is None:
                stack.pop(); continue
            h,c,d=nxt
            if h==nc:
                path=stack[1:]+[(h,c,d)]
                bn=min(resid(cc,dd) for _,cc,dd in path)
                first=None
                for i,(hh,cc,dd) in enumerate(path):
                    x[cc]+= bn if dd==0 else -bn
                    if first is None and resid(cc,dd)==0: first=i
                st['pushes']+=1; st['hops']+=len(path)
                # pop to tail of first saturated arc: path[i] enters node path[i][0]; stack index i+1
                c1=path[first][0]
                if c1!=nc:
                    s1=stamp[c1]
                    for u in range(1,nc+1):
                        if marked[u] and stamp[u]>=s1: marked[u]=False; st['unmarked']+=1
                del stack[first+1:]
                continue
            marked[h]=True; stamp[h]=cnt; cnt+=1; par[h]=v; st['enters']+=1
            stack.append((h,c,d))
        # failure: R = marked
        st['phases']+=1
        best=INF; newt=[]
        for c in range(n):
            for (t,h,d,cost) in ((sv[c], nc if ev[c]==0 else ev[c],0,-r[c]),(ev[c], nc if sv[c]==0 else sv[c],1,r[c])):
                if resid(c,d)>0 and marked[t] and not marked[h] and lab[h]<INF:
                    s=lab[t]+cost-lab[h]
                    if s<best: best=s; newt=[t]
                    elif s==best: newt.append(t)
        if best==INF: break
        if best==0:
            st['fresh']+=1
            for u in range(nc+1): marked[u]=False
            continue
        for u in range(nc+1):
            if not marked[u] and lab[u]<INF: lab[u]+=best
        if lab[nc]>=0: break
        if revive:
            for t in newt:
                u=t
                while u>0 and marked[u]: marked[u]=False; u=par[u]
            marked[0]=False
        else:
            for u in range(nc+1): marked[u]=False
    return x,st

def ssp_value(sv,ev,r,nc,cap):
    # reference: objective via simple Bellman-Ford SSP
    n=len(sv); x=[0]*n
    while True:
        lab=[INF]*(nc+1); lab[0]=0; pr=[None]*(nc+1)
        for _ in range(nc+2):
            ch=False
            for c in range(n):
                hf = nc if ev[c]==0 else ev[c]; hb = nc if sv[c]==0 else sv[c]
                if x[c]<cap[c] and lab[sv[c]]<INF and lab[sv[c]]-r[c]<lab[hf] and hf!=0: lab[hf]=lab[sv[c]]-r[c]; pr[hf]=(c,0,sv[c]); ch=True
                te = ev[c]
                if x[c]>0 and te!=0 and lab[te]<INF and lab[te]+r[c]<lab[hb]: lab[hb]=lab[te]+r[c]; pr[hb]=(c,1,te); ch=True
            if not ch: break
        if lab[nc]>=0 or lab[nc]==INF: break
        v=nc; bn=INF; path=[]
        while v!=0:
            c,d,t=pr[v]; bn=min(bn,(cap[c]-x[c]) if d==0 else x[c]); path.append((c,d)); v=t
        for c,d in path: x[c]+= bn if d==0 else -bn
    return sum(r[c]*x[c] for c in range(n))

if __name__=="__main__":
    name=sys.argv[1]; S=int(sys.argv[2]); K=int(sys.argv[3])
    inst = I.config2(S=S) if name=='c2' else I.config4(S=S)
    paths=np.load(os.path.join(ROOT,'sgufp_solver_b200','data','bench_candidates.npz'))['config2' if name=='c2' else 'config4']
    for rev in (True,False):
        tot={}
        for k in range(K):
            sv,ev,r,nc,ac=graph(inst,paths[k])
            n=len(sv)
            for s in range(S):
                cap=[INF]*n
                for a in range(inst.m):
                    c=ac[a]
                    if c>=0: cap[c]=min(cap[c],int(inst.upper[a,s]))
                x,st=solve(list(map(int,sv)),list(map(int,ev)),list(map(int,r)),nc,cap,revive=rev)
                if rev and s<3:
                    ref=ssp_value(list(map(int,sv)),list(map(int,ev)),list(map(int,r)),nc,cap)
                    got=sum(int(r[c])*x[c] for c in range(n))
                    assert ref==got,(ref,got)
                for kk,v in st.items(): tot[kk]=tot.get(kk,0)+v
        ne=K*S
        print(name,"revive" if rev else "fresh each phase",{k:round(v/ne,1) for k,v in tot.items()})
