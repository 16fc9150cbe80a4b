import torch, time
n=8192
a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
for _ in range(2): c=a@b
torch.cuda.synchronize()
best=1e9
for _ in range(5):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); c=a@b; e1.record(); torch.cuda.synchronize()
    best=min(best,e0.elapsed_time(e1))
print(f"cuBLAS DGEMM {n}^3: {2*n**3/best/1e9:.2f} TFLOP/s best ({best:.1f} ms)")
# the sketch shape: (512 x K) x (K x 2000)
m,k,K=512,2000,2**18
u=torch.randn(m,K,dtype=torch.float64,device='cuda'); th=torch.randn(k,K,dtype=torch.float64,device='cuda')
for _ in range(2): y=u@th.T
torch.cuda.synchronize()
best=1e9
for _ in range(5):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); y=u@th.T; e1.record(); torch.cuda.synchronize()
    best=min(best,e0.elapsed_time(e1))
print(f"cuBLAS DGEMM sketch shape {m}x{k}x{K}: {2*m*k*K/best/1e9:.2f} TFLOP/s ({best:.2f} ms)")
