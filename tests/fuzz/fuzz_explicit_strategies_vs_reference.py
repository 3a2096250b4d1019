"""Randomised parity of EXPLICIT step lists (random mixes of sequential and parallel steps, possibly invalid) through cosma_b200_strategy and the
unmodified reference Strategy: acceptance / rejection and the validated strategy must agree. Last run: 3000 lists, 2713 accepted, 287 rejected by
both, 0 disagreements.  python tests/fuzz/fuzz_explicit_strategies_vs_reference.py"""
import sys, ctypes, random, os
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cosma_b200 import _lib
from oracle import oracle as orc
from fuzz_schedule_vs_reference import random_steps
lib=_lib.load(); R=orc.ref()
rnd=random.Random(5)
devnull=os.open(os.devnull, os.O_WRONLY); saved=os.dup(1); os.dup2(devnull,1)
bad=[]; acc=rej=0
for it in range(3000):
    P=rnd.choice([1,2,3,4,6,8,12,16]); m,n,k=(rnd.randint(1,150) for _ in range(3))
    steps=random_steps(rnd,P)
    out=ctypes.create_string_buffer(8192); Po=ctypes.c_int(0); mu=ctypes.c_longlong(0)
    rc1=lib.cosma_b200_strategy(m,n,k,P,ctypes.c_longlong(0),steps.encode(),out,8192,ctypes.byref(Po),ctypes.byref(mu)); a=(rc1==0,out.value.decode(),Po.value,mu.value)
    out2=ctypes.create_string_buffer(8192); Po2=ctypes.c_int(0); mu2=ctypes.c_longlong(0)
    rc2=R.ref_strategy(m,n,k,P,ctypes.c_longlong(0),steps.encode(),out2,8192,ctypes.byref(Po2),ctypes.byref(mu2)); b=(rc2>=0,out2.value.decode(),Po2.value,mu2.value)
    if a[0]!=b[0] or (a[0] and a!=b): bad.append((m,n,k,P,steps,a,b))
    acc+=a[0]; rej+=not a[0]
os.dup2(saved,1)
print("explicit strategies:",3000,"accepted",acc,"rejected",rej,"disagreements",len(bad))
for x in bad[:8]: print(x)
