"""Scheduling model of the trace kernels: replays per-item step sequences (recorded from an instrumented copy of the
oracle over binned rays of the bench workload) under different warp scheduling policies and counts warp instructions
per item.  Policy A = traceFastKernel (one item per lane, gangs), run_multi = tracePoolKernel (K items per lane).
Input: /tmp/sim/seqs.txt, one line per item, letters N node, S split, L leaf entry, l entry passing the bounds test,
B brush clip, P pop, D done.  Kept for the record of how the pool design was chosen."""
import sys, collections
seqs=open('/tmp/sim/seqs.txt').read().split('\n')
seqs=[s for s in seqs if s][:49152]   # short-class part
COST=dict(N=32,L=28,l=66,P=18,S=85,B=260,D=220)
CENSUS=25; REP=8
LIGHT='NLlP'
def run(policy, nrays=20000, **kw):
    it=iter(seqs[:nrays])
    lanes=[None]*32   # (seq, pos)
    def refill(i):
        try: lanes[i]=[next(it),0]
        except StopIteration: lanes[i]=None
    for i in range(32): refill(i)
    total=0; done=0; execs=collections.Counter(); lanesum=collections.Counter()
    def cur(i):
        l=lanes[i]; return l[0][l[1]] if l else '-'
    def step(i):
        l=lanes[i]; l[1]+=1
    def runblock(types, cost=None):
        nonlocal total, done
        idx=[i for i in range(32) if cur(i) in types]
        if not idx: return 0
        c=max(COST[cur(i)] for i in idx) if cost is None else cost
        total+=c; execs[types]+=1; lanesum[types]+=len(idx)
        for i in idx:
            if cur(i)=='D': done+=1; refill(i)
            else: step(i)
        return len(idx)
    while any(l is not None for l in lanes):
        cnt=collections.Counter(cur(i) for i in range(32) if lanes[i])
        nL=sum(cnt[t] for t in LIGHT)
        total+=CENSUS
        if policy=='A':
            g=kw.get('g',4); gf=kw.get('gf',4); R=kw.get('R',3)
            if cnt['D']>0 and cnt['D']>=min(gf,nL): runblock('D')
            for r in range(R):
                total+=REP
                runblock('Ll'); runblock('P'); runblock('N')
            cnt=collections.Counter(cur(i) for i in range(32) if lanes[i]); nL=sum(cnt[t] for t in LIGHT)
            if cnt['S']>0 and cnt['S']>=min(g,nL): runblock('S')
            if cnt['B']>0 and cnt['B']>=min(g,nL): runblock('B')
        elif policy=='M':
            # run only the mode whose (count) is largest, weighted thresholds
            th=kw.get('th',dict(N=1,L=1,P=1,S=1,B=1,D=1))
            groups={'N':'N','L':'Ll','P':'P','S':'S','B':'B','D':'D'}
            best=None;bv=-1
            for k,t in groups.items():
                c=sum(cnt[x] for x in t)
                if c==0: continue
                v=c*th[k]
                if v>bv: bv=v;best=k
            t=groups[best]
            n0=sum(cnt[x] for x in t)
            # repeat while still >= frac of initial lanes
            frac=kw.get('frac',0.6); maxrep=kw.get('maxrep',4)
            for r in range(maxrep):
                total+=REP
                n=runblock(t)
                c=sum(1 for i in range(32) if cur(i) in t)
                if c<frac*n0 or c==0: break
        elif policy=='T':
            # all blocks gang-thresholded: run block if count >= min(T[k], others)
            T=kw['T']
            groups={'D':'D','L':'Ll','P':'P','N':'N','S':'S','B':'B'}
            ran=False
            R=kw.get('R',2)
            for r in range(R):
              for k,t in groups.items():
                cnt=collections.Counter(cur(i) for i in range(32) if lanes[i])
                c=sum(cnt[x] for x in t)
                if c==0: continue
                act=sum(cnt.values())
                if c>=min(T[k], act) or c==max(sum(cnt[x] for x in tt) for tt in groups.values()):
                    total+=REP if k in 'LPN' else 0
                    runblock(t); ran=True
    util={k:lanesum[k]/execs[k] for k in execs}
    return total/done, util
if __name__=='__main__':
    print('A g4 R3', run('A'))
    print('A g8 R3', run('A',g=8,gf=8))
    print('M frac.6', run('M'))
    print('M frac.75 maxrep8', run('M',frac=0.75,maxrep=8))
    print('M weighted', run('M',th=dict(N=1,L=1.2,P=2,S=1.5,B=3,D=2),frac=0.6,maxrep=6))
    print('T 8', run('T',T=dict(D=8,L=8,P=6,N=8,S=8,B=8)))
    print('T 12', run('T',T=dict(D=8,L=10,P=8,N=12,S=8,B=8)))

def run_pool(P=128, nrays=20000, qcost=30, mult=1.3, maxrep=4, frac=0.7, prio=None):
    it=iter(seqs[:nrays])
    pool=[]
    def newray():
        try: return [next(it),0]
        except StopIteration: return None
    for i in range(P):
        r=newray()
        if r: pool.append(r)
    total=0; done=0; execs=collections.Counter(); lanesum=collections.Counter()
    groups={'N':'N','L':'Ll','P':'P','S':'S','B':'B','D':'D'}
    while pool:
        cnt=collections.Counter()
        for r in pool: cnt[r[0][r[1]]]+=1
        best=None;bv=-1
        for k,t in groups.items():
            c=sum(cnt[x] for x in t)
            if c==0: continue
            v=min(c,32)*(prio[k] if prio else 1)+ (0.001*c)
            if v>bv: bv=v;best=k
        t=groups[best]
        batch=[r for r in pool if r[0][r[1]] in t][:32]
        total+=qcost
        n0=len(batch)
        for rep in range(maxrep):
            act=[r for r in batch if r is not None and r[0][r[1]] in t]
            if not act or len(act)<frac*n0: break
            c=max(COST[r[0][r[1]]] for r in act)
            total+=c*mult; execs[best]+=1; lanesum[best]+=len(act)
            for r in act:
                if r[0][r[1]]=='D':
                    done+=1
                    nr=newray()
                    if nr: r[0],r[1]=nr[0],nr[1]
                    else: pool.remove(r); batch[batch.index(r)]=None
                else: r[1]+=1
    util={k:lanesum[k]/execs[k] for k in execs}
    return total/done, util
if __name__=='__main__':
    for P in (64,128,256,512):
        print('pool',P, run_pool(P))
    print('pool 256 maxrep8', run_pool(256,maxrep=8))
    print('pool 256 mult1.5 q40', run_pool(256,mult=1.5,qcost=40))

def run_multi(K=4, nrays=20000, sched=25, mult=1.3, maxrep=3, frac=0.7, prio=None):
    it=iter(seqs[:nrays])
    def newray():
        try: return [next(it),0]
        except StopIteration: return None
    lanes=[[newray() for j in range(K)] for i in range(32)]
    total=0; done=0; execs=collections.Counter(); lanesum=collections.Counter()
    groups={'N':'N','L':'Ll','P':'P','S':'S','B':'B','D':'D'}
    def mode(r): return r[0][r[1]]
    while any(r is not None for l in lanes for r in l):
        best=None;bv=-1
        for k,t in groups.items():
            c=sum(1 for l in lanes if any(r is not None and mode(r) in t for r in l))
            if c==0: continue
            v=c*(prio[k] if prio else 1)
            if v>bv: bv=v;best=k
        t=groups[best]
        total+=sched
        pick=[]
        for l in lanes:
            p=None
            for r in l:
                if r is not None and mode(r) in t: p=r;break
            pick.append(p)
        n0=sum(1 for p in pick if p is not None)
        for rep in range(maxrep):
            act=[(i,p) for i,p in enumerate(pick) if p is not None and mode(p) in t]
            if not act or len(act)<frac*n0: break
            c=max(COST[mode(p)] for i,p in act)
            total+=c*mult; execs[best]+=1; lanesum[best]+=len(act)
            for i,p in act:
                if mode(p)=='D':
                    done+=1
                    nr=newray()
                    j=lanes[i].index(p)
                    lanes[i][j]=nr
                    pick[i]=None
                else: p[1]+=1
    util={k:round(lanesum[k]/execs[k],1) for k in execs}
    return round(total/done,1), util
if __name__=='__main__':
    for K in (1,2,3,4,6,8):
        print('multi K',K, run_multi(K))
    print('multi K4 prio', run_multi(4,prio=dict(N=1,L=1,P=1,S=1.3,B=2,D=1.5)))
    print('multi K4 mult1.5 sched40', run_multi(4,mult=1.5,sched=40))
