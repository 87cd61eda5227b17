import sys; sys.path.insert(0,'.')
import torch, vfidkr_b200 as V
from vfidkr_b200 import _lib
from vfidkr_b200._common import ptr, stream_ptr
dev=torch.device('cuda',0); sp=stream_ptr(dev)
def t(fn,n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/n*1e3
Bs,Hs,Ws,Fs=8,256,448,51
Is=torch.rand(Bs,3,Hs,Ws,device=dev); vs=torch.rand(Bs,Fs,Hs-Fs+1,Ws-Fs+1,device=dev)/Fs; hs=torch.rand_like(vs)/Fs
os_=torch.empty(Bs,3,Hs-Fs+1,Ws-Fs+1,device=dev); gs=torch.randn_like(os_)
g1,g2,g3=torch.empty_like(Is),torch.empty_like(vs),torch.empty_like(hs)
print("fwd us", t(lambda:_lib.call("vfidkr_separableconv_forward",ptr(Is),ptr(vs),ptr(hs),ptr(os_),Bs,3,Hs,Ws,Fs,sp)))
print("bwd us", t(lambda:_lib.call("vfidkr_separableconv_backward",ptr(Is),ptr(vs),ptr(hs),ptr(gs),ptr(g1),ptr(g2),ptr(g3),Bs,3,Hs,Ws,Fs,sp)))
