"""Developer tool: search for the conflict-free packed layout of a symmetric 12x12 block used by
srbd_pdipm_fast.h (pkf): a proper edge colouring of K12 with loops into 16 bank pairs with class sizes
5 (x14) and 4 (x2), turned into addresses colour + 16 * slot.  Usage: python tools/pack_layout.py [seed]"""
import random, sys
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
n=12
entries=[(i,j) for i in range(n) for j in range(i+1)]
cap=[5]*14+[4]*2
def solve():
    color={}
    used=[set() for _ in range(n)]   # colors used in star of vertex
    cnt=[0]*16
    order=entries[:]; random.shuffle(order)
    # order by constraint: do edges first
    def assign(idx):
        if idx==len(order): return True
        i,j=order[idx]
        cols=list(range(16)); random.shuffle(cols)
        for c in cols:
            if cnt[c]>=cap[c]: continue
            if c in used[i] or (i!=j and c in used[j]): continue
            color[(i,j)]=c; cnt[c]+=1; used[i].add(c); 
            if i!=j: used[j].add(c)
            if assign(idx+1): return True
            del color[(i,j)]; cnt[c]-=1; used[i].discard(c)
            if i!=j: used[j].discard(c)
        return False
    sys.setrecursionlimit(10000)
    import time
    return color if assign(0) else None
col=solve()
assert col
# addresses
slot=[0]*16; addr={}
for e in entries:
    c=col[e]; addr[e]=c+16*slot[c]; slot[c]+=1
assert sorted(addr.values())==list(range(78)), sorted(addr.values())
# verify
for m in range(n):
    res=set()
    for r in range(n):
        i,j=(r,m) if r>=m else (m,r)
        res.add(addr[(i,j)]%16)
    assert len(res)==12
table=[[addr[(max(i,j),min(i,j))] for j in range(12)] for i in range(12)]
print("OK perfect layout"); 
for row in table: print(row)
