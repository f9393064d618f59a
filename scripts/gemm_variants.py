import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from surmise_b200 import ops
dev = torch.device('cuda')
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e-3)
    return float(np.median(ts))
out={}
n=4096
A=torch.randn((n,n),dtype=torch.float64,device=dev); B=torch.randn((n,n),dtype=torch.float64,device=dev); C=torch.zeros((n,n),dtype=torch.float64,device=dev)
for tile in (0,1,4):
    for akc,bkc in ((True,True),(True,False),(False,False)):
        t=timeit(lambda: ops.dgemm(A,B,C,akc,bkc,n,n,n,tile=tile,lds=(n,n,n)))
        out[f'sq4096_tile{tile}_{int(akc)}{int(bkc)}']=2*n**3/t/1e12
nb,M,K=20,1000,1024
P=torch.randn((nb,M,K),dtype=torch.float64,device=dev); Cb=torch.zeros((nb,M,M),dtype=torch.float64,device=dev)
for tile in (0,1,4):
    t=timeit(lambda: ops.dgemm(P,P,Cb,True,True,M,M,K,alpha=-1.0,beta=1.0,flags=16,batch=nb,strides=(M*K,M*K,M*M),lds=(K,K,M),tile=tile))
    out[f'syrk_1000x1024_b20_tile{tile}']=nb*M*M*K/t/1e12
# skinny N=64 long-K (left-looking column update): M=1500, N=64, K=1000, batch 20
M2,N2,K2=1500,64,1000
A2=torch.randn((nb,M2,K2),dtype=torch.float64,device=dev); B2=torch.randn((nb,N2,K2),dtype=torch.float64,device=dev); C2=torch.zeros((nb,M2,N2),dtype=torch.float64,device=dev)
for tile in (1,):
    t=timeit(lambda: ops.dgemm(A2,B2,C2,True,True,M2,N2,K2,alpha=-1.0,beta=1.0,batch=nb,strides=(M2*K2,N2*K2,M2*N2),lds=(K2,K2,N2),tile=tile))
    out[f'skinny_1500x64x1000_b20_tile{tile}']=2*nb*M2*N2*K2/t/1e12
N3=128
B3=torch.randn((nb,N3,K2),dtype=torch.float64,device=dev); C3=torch.zeros((nb,M2,N3),dtype=torch.float64,device=dev)
for tile in (0,1,4):
    t=timeit(lambda: ops.dgemm(A2,B3,C3,True,True,M2,N3,K2,alpha=-1.0,beta=1.0,batch=nb,strides=(M2*K2,N3*K2,M2*N3),lds=(K2,K2,N3),tile=tile))
    out[f'skinny_1500x128x1000_b20_tile{tile}']=2*nb*M2*N3*K2/t/1e12
# predict shape: M=2000,N=264,K=2000 lower-tri A, batch 20, B operand row-contiguous
M4,N4=2000,264
L=torch.tril(torch.randn((nb,M4,M4),dtype=torch.float64,device=dev)); R=torch.randn((M4,N4),dtype=torch.float64,device=dev); T=torch.zeros((nb,M4,N4),dtype=torch.float64,device=dev)
for tile in (0,1,4):
    t=timeit(lambda: ops.dgemm(L,R,T,True,False,M4,N4,M4,flags=1,batch=nb,strides=(M4*M4,0,M4*N4),lds=(M4,N4,N4),tile=tile))
    out[f'trmm_2000x264_b20_tile{tile}']=nb*M4*M4*N4/t/1e12
print(json.dumps(out,indent=1))
json.dump(out,open(os.path.join(ROOT,'gpurun_out','gemm_variants.json'),'w'),indent=1)
