import torch
dev=torch.device("cuda",0)
def timeit(fn, iters=200):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)*1e3/iters
for mb in (7.3, 14.7, 29.5, 59, 118, 472):
    n=int(mb*1e6/4)
    POOL=max(2,int(400/mb))
    outs=[torch.empty(n,device=dev) for _ in range(POOL)]
    ins=[torch.randn(n,device=dev) for _ in range(POOL)]
    tf=timeit(lambda i: outs[i%POOL].fill_(1.0))
    tw=timeit(lambda i: outs[0].fill_(1.0))
    tc=timeit(lambda i: outs[i%POOL].copy_(ins[i%POOL]))
    tr=timeit(lambda i: ins[i%POOL].sum())
    print("%.1f MB: fill cold %.2f us (%.0f GB/s) warm %.2f us (%.0f GB/s); copy %.2f us (%.0f GB/s r+w); sum-read %.2f us (%.0f GB/s)"%(mb,tf,mb*1e6/tf/1e3,tw,mb*1e6/tw/1e3,tc,2*mb*1e6/tc/1e3,tr,mb*1e6/tr/1e3))
