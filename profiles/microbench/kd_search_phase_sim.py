"""CPU replay of the kd-tree traversal of kernel 2a (coarse word search) on a bench-like vocabulary: records
per leaf visit the (descend, bucket, unwind) micro-step counts of every search and evaluates the warp-level cost
of the three-phase schedule of kd_search_kernel under different lane groupings / refill policies.
Result (1 536 searches): 8.3 leaf visits, 116 micro-steps per search; phase efficiency 0.43 with immediate refill
(the kernel measures 0.38); sorting by first leaf does not help, batched refill is worse.
    python profiles/microbench/kd_search_phase_sim.py"""
import sys; import os; R=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0,R)
import numpy as np
from maplab_b200 import synthetic
from oracle import pyoracle as po
# bench-like world at reduced size for the vocabulary (1000 words), queries 2000 descriptors
m = synthetic.make_map(60000, seed=1)
blob, voc = synthetic.make_vocabulary(m["bits"][::2][:100000], num_words=1000, seed=7)
q = synthetic.make_queries(m, 8, seed=11)
ora = po.Engine(blob)
qp = ora.project(q["bits"])[:, :5]   # first half
W1 = voc["W1"]  # [5][1000]
tree = po.KdTree(po.colmajor(W1.tolist()), 5, 1000)
nodes, buckets = tree.export()
cloud = np.ascontiguousarray(W1.T, np.float32)  # [n][5]
D=5; K=10; eps2=np.float32(9.0); rad2=np.float32(400.0)
f32=np.float32
def search(qv):
    trace=[]  # per leaf visit: (descend_steps, bucket_size, pop_steps_before_next)
    heap_v=[np.inf]*K; heap_i=[-1]*K
    off=[f32(0)]*D
    stack=[]
    node=0; rd=f32(0)
    firstleaf=None
    while True:
        dsteps=0
        while True:
            dim,cs,cb=nodes[node]
            dsteps+=1
            if dim==D:
                pos=int(cb); end=pos+int(cs); break
            cut=np.array([cb],np.uint32).view(np.float32)[0]
            old=off[dim]; new=f32(qv[dim]-cut)
            frd=f32(rd+f32(f32(-old*old)+f32(new*new)))
            stack.append((int(node),frd,None))
            node = int(cs) if new>0 else node+1
        if firstleaf is None: firstleaf=node
        bs=end-pos
        for p in range(pos,end):
            pi=buckets[p]
            d=f32(0)
            for j in range(D):
                diff=f32(qv[j]-cloud[pi][j]); d=f32(d+f32(diff*diff))
            if d<=rad2 and d<heap_v[K-1]:
                i=K-1
                while i>0 and heap_v[i-1]>d:
                    heap_v[i]=heap_v[i-1]; heap_i[i]=heap_i[i-1]; i-=1
                heap_v[i]=d; heap_i[i]=pi
        psteps=0; done=False
        while True:
            psteps+=1
            if not stack: done=True; break
            tag,val,extra=stack.pop()
            if extra is not None:
                off[extra]=val; continue
            frd=val
            if frd<=rad2 and f32(frd*eps2)<heap_v[K-1]:
                dim,cs,cb=nodes[tag]
                cut=np.array([cb],np.uint32).view(np.float32)[0]
                new=f32(qv[dim]-cut)
                stack.append((0,off[dim],int(dim)))
                off[dim]=new
                node = tag+1 if new>0 else int(cs)
                rd=frd
                break
        trace.append((dsteps,bs,psteps))
        if done: break
    return trace, firstleaf, heap_i
import time
t=time.time()
N=1536
traces=[];leafs=[]
for i in range(N):
    tr,fl,hi=search(qp[i])
    traces.append(tr); leafs.append(fl)
# verify one against oracle
idx,_=tree.knn(qp[0],K,2.0,20.0)
print("check", list(idx)==search(qp[0])[2], time.time()-t)
leafv=[len(t) for t in traces]
print("leaf visits mean",np.mean(leafv),"max",max(leafv))
own=[sum(a+b+c for a,b,c in t) for t in traces]
print("own micro-steps mean",np.mean(own))
def warp_cost(group):
    L=max(len(traces[i]) for i in group); cost=0
    for it in range(L):
        ds=[traces[i][it][0] for i in group if it<len(traces[i])]
        bs=[traces[i][it][1] for i in group if it<len(traces[i])]
        ps=[traces[i][it][2] for i in group if it<len(traces[i])]
        cost+=max(ds)+max(bs)+max(ps)
    return cost
def total(order):
    return sum(warp_cost(order[w:w+32]) for w in range(0,N,32))
rnd=list(range(N))
print("random grouping cost/item",total(rnd)/N, "efficiency", np.mean(own)/(total(rnd)/N*1.0))
srt=sorted(range(N),key=lambda i:(leafs[i],))
print("sorted by first leaf cost/item",total(srt)/N)
# sort by first leaf then by number of visits (oracle-ish upper bound)
srt2=sorted(range(N),key=lambda i:(len(traces[i]),leafs[i]))
print("sorted by #visits cost/item",total(srt2)/N)
# --- with persistent-lane refill: every outer iteration each lane works on one leaf visit
visits=[v for t in traces for v in t]
import random
random.seed(0)
def sim(vlist, trials=400):
    own=0; cost=0
    for _ in range(trials):
        g=random.sample(vlist,32)
        own+=sum(a+b+c for a,b,c in g)
        cost+=32*(max(a for a,b,c in g)+max(b for a,b,c in g)+max(c for a,b,c in g))
    return own/cost
print("visit stats mean d,b,p:",np.mean([v[0] for v in visits]),np.mean([v[1] for v in visits]),np.mean([v[2] for v in visits]))
print("max-ish d,b,p (95pct):",np.percentile([v[0] for v in visits],95),np.percentile([v[1] for v in visits],95),np.percentile([v[2] for v in visits],95))
print("efficiency with refill (3 phases):",sim(visits))
# alternative: combined D+P loop (2 bodies per iteration) + L
def sim2(vlist, trials=400):
    own=0; cost=0
    for _ in range(trials):
        g=random.sample(vlist,32)
        own+=sum(a+b+c for a,b,c in g)
        cost+=32*(2*max(a+c for a,b,c in g)+max(b for a,b,c in g))
    return own/cost
print("efficiency combined D+P (x2 body):",sim2(visits))
# --- lane-stream simulation with batched refill threshold T
def lane_sim(T, items=traces, reps=3):
    tot_cost=0; tot_items=0
    for rep in range(reps):
        order=list(range(len(items))); random.shuffle(order)
        queue=order[:]  # warp's item queue
        cur=[None]*32; pos=[0]*32
        cost=0
        def refill():
            for l in range(32):
                if cur[l] is None and queue:
                    cur[l]=items[queue.pop()]; pos[l]=0
        refill()
        while any(c is not None for c in cur):
            act=[l for l in range(32) if cur[l] is not None]
            vs=[cur[l][pos[l]] for l in act]
            cost+=max(v[0] for v in vs)+max(v[1] for v in vs)+max(v[2] for v in vs)
            for l in act:
                pos[l]+=1
                if pos[l]>=len(cur[l]): cur[l]=None
            idle=sum(1 for c in cur if c is None)
            if idle>=T or idle==32: refill()
        tot_cost+=cost; tot_items+=len(items)
    return tot_cost/tot_items
base=lane_sim(1)
for T in (1,2,4,8,12,16,24,32):
    c=lane_sim(T); print("T",T,"warp-steps/item",round(c,2),"rel",round(c/base,3))
