import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oov_b200 import ops
dev="cuda:0"
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
M=1<<18
for N in (512, 256):
  for K in (256,512,1024):
    A=torch.randn(M,K,device=dev).to(torch.bfloat16); W=torch.randn(N,K,device=dev).to(torch.bfloat16); b=torch.zeros(N,device=dev)
    row=[]
    for act,dbg,bias in (("none",0,b),("none",0,None),("none",1,b),("none",2,b),("gelu",0,b),("gelu",1,b)):
        ms=t(lambda: ops.tc_linear(A,W,bias,act=act,out_dtype=torch.bfloat16,_debug=dbg))
        row.append(f"{act}/dbg{dbg}/{'b' if bias is not None else 'nob'}={ms*1e3:.0f}us")
    print(f"N={N} K={K}: "+"  ".join(row), f"| MMA-bound {2*M*N*K/1.65e15*1e6:.0f}us")
