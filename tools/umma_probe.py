import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mmdfn_b200 import _lib as L
dev="cuda"
def probe(N, lbo, sbo, mn, pa):
    out = torch.zeros(128, N, device=dev)
    L.call("mmdfn_umma_probe", L.ptr(out), N, lbo, sbo, mn, pa, L.stream()); torch.cuda.synchronize()
    return out.cpu()
N=32
print("== B operand, K-major reference (lbo=128,sbo=528): word offset of B(k,n) for k=0..7 (rows), n=0..11")
o=probe(N,128,528,0,0); print(o[:8,:12].int())
for lbo,sbo in ((4608,144),(144,4608),(128,1024),(1024,128)):
    print(f"== B operand MN-major lbo={lbo} sbo={sbo}: word offset of B(k,n), k=0..7, n=0..N")
    o=probe(N,lbo,sbo,1,0); print(o[:8,:N].int())
for lbo,sbo in ((4608,144),(144,4608)):
    print(f"== A operand MN-major lbo={lbo} sbo={sbo}: word offset of A(m,k), m=0..15 (rows), k=0..7")
    o=probe(N,lbo,sbo,1,1); print(o[:16,:8].int()); print("rows 120..127:", o[120:128,:8].int())
