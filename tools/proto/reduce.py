import os
import ctypes as C, numpy as np, sys
HERE=os.path.dirname(os.path.abspath(__file__)); ROOT=os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0,ROOT)
from sgufp_solver_b200 import instances as I
L=C.CDLL(os.path.join(HERE,'libpg.so'))
ip=C.POINTER(C.c_int32)
def graph(inst,path):
    arr=lambda a: np.ascontiguousarray(a,dtype=np.int32)
    t,h,r0,vb=arr(inst.tail),arr(inst.head),arr(inst.reward[:,0]),arr(inst.vbar)
    sv=np.zeros(inst.m,np.int32);ev=np.zeros(inst.m,np.int32);r=np.zeros(inst.m,np.int32);nc=C.c_int32();ac=np.zeros(inst.m,np.int32)
    p=np.ascontiguousarray(path,dtype=np.int16)
    n=L.plan_graph(inst.n,inst.m,t.ctypes.data_as(ip),h.ctypes.data_as(ip),r0.ctypes.data_as(ip),vb.ctypes.data_as(ip),len(vb),p.ctypes.data_as(C.POINTER(C.c_int16)),len(p),sv.ctypes.data_as(ip),ev.ctypes.data_as(ip),r.ctypes.data_as(ip),C.byref(nc),ac.ctypes.data_as(ip))
    assert n>=0,n
    return sv[:n].copy(),ev[:n].copy(),r[:n].copy(),nc.value,ac
def reduce(sv,ev,r,nc):
    # nodes: 0 root (source and sink), -1 dangling? open chains should have both ends
    ch=[(int(a),int(b),int(c),[i]) for i,(a,b,c) in enumerate(zip(sv,ev,r))]
    alive=[True]*len(ch)
    changed=True
    while changed:
        changed=False
        ins={};outs={}
        for i,(a,b,c,_) in enumerate(ch):
            if not alive[i]: continue
            outs.setdefault(a,[]).append(i); ins.setdefault(b,[]).append(i)
        nodes=set(ins)|set(outs)
        for v in nodes:
            if v==0: continue
            if v<0:  # dangling
                for i in ins.get(v,[])+outs.get(v,[]):
                    if alive[i]: alive[i]=False; changed=True
                continue
            if not ins.get(v) or not outs.get(v):
                for i in ins.get(v,[])+outs.get(v,[]):
                    if alive[i]: alive[i]=False; changed=True
        if changed: continue
        for v in nodes:
            if v<=0: continue
            if len(ins.get(v,[]))==1 and len(outs.get(v,[]))==1:
                i=ins[v][0]; j=outs[v][0]
                if i==j or not alive[i] or not alive[j]: continue
                a,_,c1,m1=ch[i]; _,b,c2,m2=ch[j]
                ch[i]=(a,b,c1+c2,m1+m2); alive[j]=False; changed=True
                break
    res=[ch[i] for i in range(len(ch)) if alive[i]]
    nodes=set()
    for a,b,_,_ in res: nodes.add(a);nodes.add(b)
    return res,nodes
for name,inst in (("c2",I.config2(S=4)),("c4",I.config4(S=4))):
    paths=np.load(os.path.join(ROOT,'sgufp_solver_b200','data','bench_candidates.npz'))['config2' if name=='c2' else 'config4']
    for k in range(min(6,len(paths))):
        sv,ev,r,nc,ac=graph(inst,paths[k])
        res,nodes=reduce(sv,ev,r,nc)
        used=set(sv)|set(ev)
        indeg={};outdeg={}
        for a,b,_,_ in res: outdeg[a]=outdeg.get(a,0)+1; indeg[b]=indeg.get(b,0)+1
        # parallel chains
        from collections import Counter
        par=Counter((a,b) for a,b,_,_ in res)
        npar=sum(v-1 for v in par.values())
        in1=sum(1 for v in nodes if v and indeg.get(v,0)==1); out1=sum(1 for v in nodes if v and outdeg.get(v,0)==1)
        print(name,k,"nc",nc,"nopen",len(sv),"nodes used",len(used),"-> chains",len(res),"nodes",len(nodes),"parallel extra",npar,"indeg1",in1,"outdeg1",out1,"root out",outdeg.get(0),"root in",indeg.get(0), "min/max r", min(c for _,_,c,_ in res), max(c for _,_,c,_ in res))
