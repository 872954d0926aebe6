import sys
def next_radix(rem,E):
    if rem>=64 or rem==E: return E
    if rem==32: return 4
    if rem==16: return 4
    return rem
def plan(L,E):
    if E==32:
        if L==512: return [16,32]
        if L==1024: return [32,32]
        if L==8192: return [32,16,16]
        return [32,16,32]
    if L<=E: return [L]
    r=[E]; rem=L//E
    while rem>1:
        x=next_radix(rem,E); r.append(x); rem//=x
    return r
def conflicts(L,E,B,LP,bf_load=False,bf_store=False):
    R=plan(L,E); NP=len(R); T=L*B//E
    PADW=R[0]; sh=PADW.bit_length()-1
    def phys(ex,batch,pos): return batch*LP+pos+((pos>>sh) if ex==0 else 0)
    worst={}
    NS=1
    for p in range(NP):
        Rp=R[p]; NBF=L//Rp; U=E//Rp
        BF=(p==0 and bf_load) or (p==NP-1 and bf_store)
        def mp(i):
            return (i% B, i//B) if BF else (i//NBF, i%NBF)
        for kind in ('read','write'):
            if kind=='read' and p==0: continue
            if kind=='write' and p==NP-1: continue
            mx=0
            for u in range(U):
              for t in range(Rp):
                for w0 in range(0,T,32):
                    for half in range(2):
                        banks={}
                        for lane in range(16):
                            i=w0+half*16+lane+u*T
                            batch,j=mp(i)
                            if kind=='read': a=phys(p-1,batch,j+t*NBF)
                            else:
                                k=j%NS; o=(j-k)*Rp+k; a=phys(p,batch,o+t*NS)
                            for w in (2*a,2*a+1):
                                banks[w%32]=banks.get(w%32,0)+1
                        mx=max(mx,max(banks.values()))
            worst[(p,kind)]=mx
        NS*=Rp
    return worst
if __name__=='__main__':
    cases=[(64,8,32),(128,8,16),(256,8,8),(512,8,4),(64,16,64),(128,16,32),(256,16,16),(512,16,8),(1024,16,4),(512,32,8),(1024,32,4),(8192,32,1),(8192,16,1),(4096,16,1),(2048,16,2),(16,16,256),(32,16,128)]
    for L,E,B in cases:
        R=plan(L,E)
        if len(R)==1: continue
        PADW=R[0]
        LPdef=((L+L//PADW)|1)
        res=conflicts(L,E,B,LPdef)
        best=None
        for LP in range(L+L//PADW, L+L//PADW+40):
            r=conflicts(L,E,B,LP); m=max(r.values())
            if best is None or m<best[0]: best=(m,LP,r)
        print(L,E,B,R,'LP default',LPdef,'worst',max(res.values()),res,'| best LP',best[1],'worst',best[0])

def conflicts2(L,E,B,LP,G,bf_load=False,bf_store=False):
    R=plan(L,E); NP=len(R); T=L*B//E
    sh=G.bit_length()-1
    def phys(ex,batch,pos): return batch*LP+pos+((pos>>sh) if ex==0 else 0)
    worst={}; NS=1
    for p in range(NP):
        Rp=R[p]; NBF=L//Rp; U=E//Rp
        BF=(p==0 and bf_load) or (p==NP-1 and bf_store)
        def mp(i): return (i% B, i//B) if BF else (i//NBF, i%NBF)
        for kind in ('read','write'):
            if kind=='read' and p==0: continue
            if kind=='write' and p==NP-1: continue
            mx=0
            for u in range(U):
              for t in range(Rp):
                for w0 in range(0,T,32):
                    for half in range(2):
                        banks={}
                        for lane in range(16):
                            i=w0+half*16+lane+u*T
                            batch,j=mp(i)
                            if kind=='read': a=phys(p-1,batch,j+t*NBF)
                            else:
                                k=j%NS; o=(j-k)*Rp+k; a=phys(p,batch,o+t*NS)
                            for w in (2*a,2*a+1): banks[w%32]=banks.get(w%32,0)+1
                        mx=max(mx,max(banks.values()))
            worst[(p,kind)]=mx
        NS*=Rp
    return worst
print('--- search pad granularity G and LP')
for L,E,B in [(64,8,32),(128,8,16),(256,8,8),(512,8,4),(1024,8,2),(2048,8,1),(32,16,128),(64,16,64),(128,16,32)]:
    R=plan(L,E); best=None
    for G in (R[0],16,32):
        if G<R[0]: continue
        base=L+L//G
        for LP in range(base, base+34):
            r=conflicts2(L,E,B,LP,G); m=sum(r.values())
            if best is None or m<best[0]: best=(m,G,LP,r)
    print(L,E,B,R,'best G',best[1],'LP',best[2],'(L+L/G =',L+L//best[1],')',best[3])
