"""Randomised parity of cosma_b200_strategy against the unmodified reference Strategy (oracle/_ref): python tests/fuzz/fuzz_strategy_vs_reference.py SEED N.
Last run: 4000 cases (dims 1..300000, P 1..1024, 35 % with memory limits), 0 mismatches in steps, ranks used and memory_used."""
import sys, ctypes, random, io, contextlib, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from cosma_b200 import _lib
from oracle import oracle as orc
lib=_lib.load(); R=orc.ref()
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
N=int(sys.argv[2]) if len(sys.argv)>2 else 3000
def ours(m,n,k,P,mem):
    out=ctypes.create_string_buffer(8192); Po=ctypes.c_int(0); mu=ctypes.c_longlong(0)
    rc=lib.cosma_b200_strategy(m,n,k,P,ctypes.c_longlong(mem),b"",out,8192,ctypes.byref(Po),ctypes.byref(mu))
    return (rc==0, out.value.decode(), Po.value, mu.value)
def theirs(m,n,k,P,mem):
    out=ctypes.create_string_buffer(8192); Po=ctypes.c_int(0); mu=ctypes.c_longlong(0)
    rc=R.ref_strategy(m,n,k,P,ctypes.c_longlong(mem if mem>0 else 0),b"",out,8192,ctypes.byref(Po),ctypes.byref(mu))
    return (rc>=0, out.value.decode(), Po.value, mu.value)
bad=0; thrown=0
devnull=os.open(os.devnull, os.O_WRONLY); saved=os.dup(1); os.dup2(devnull,1)
res=[]
for it in range(N):
    def dim():
        c=random.random()
        if c<0.3: return random.randint(1,400)
        if c<0.7: return random.randint(200,20000)
        return random.choice([2**random.randint(8,17), random.randint(10000,300000)])
    m,n,k=dim(),dim(),dim()
    P=random.choice([1,2,3,4,5,6,7,8,9,12,16,24,32,36,48,64,100,128,256,500,1024])
    mem=0
    if random.random()<0.35:
        base=(m*k+k*n+m*n)//P
        mem=int(base*random.uniform(1.0,4.0))+1
    a=ours(m,n,k,P,mem); b=theirs(m,n,k,P,mem)
    if a[0]!=b[0] or (a[0] and (a[1]!=b[1] or a[2]!=b[2] or a[3]!=b[3])):
        bad+=1; res.append((m,n,k,P,mem,a,b))
    if not a[0]: thrown+=1
os.dup2(saved,1)
print("cases",N,"mismatches",bad,"both threw",thrown)
for r in res[:10]: print(r)
