import torch, torch.nn.functional as F, time
torch.backends.cudnn.benchmark=True
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n
for (hw,c) in [(64*64,320),(16*16,1280),(512*512,128),(32*32,1920)]:
    h=int(hw**0.5); x2d=torch.randn(hw,c,device='cuda',requires_grad=True); w=torch.randn(c,device='cuda'); b=torch.randn(c,device='cuda')
    def a():
        y=F.silu(F.group_norm(x2d.t().reshape(1,c,-1),32,w,b,1e-5)); return y.reshape(c,-1).t().contiguous()
    def cl():
        v=x2d.view(1,h,h,c).permute(0,3,1,2); y=F.silu(F.group_norm(v,32,w,b,1e-5)); return y
    y1=a(); y2=cl()
    print(hw,c,'cl out is channels_last:', y2.is_contiguous(memory_format=torch.channels_last), 'maxdiff', float((y2.permute(0,2,3,1).reshape(hw,c)-y1).abs().max()),
          'ms transpose-path %.4f  cl-path %.4f'%(t(a),t(cl)))
    g=torch.randn_like(y1)
    def ab(): y=a(); return torch.autograd.grad(y,x2d,g)
    def clb(): y=cl(); return torch.autograd.grad(y,x2d,g.view(1,h,h,c).permute(0,3,1,2))
    print('   fwd+bwd ms transpose-path %.4f cl-path %.4f'%(t(ab),t(clb)))
# SDPA fp32 timing
for (s,hd,d) in [(4096,8,40),(1024,8,80),(256,8,160)]:
    q=torch.randn(1,hd,s,d,device='cuda',requires_grad=True); k=torch.randn_like(q,requires_grad=True); v=torch.randn_like(q,requires_grad=True)
    f=lambda: F.scaled_dot_product_attention(q,k,v)
    o=f(); g=torch.randn_like(o)
    fb=lambda: torch.autograd.grad(F.scaled_dot_product_attention(q,k,v),(q,k,v),g)
    print('sdpa fp32',s,hd,d,'fwd ms %.4f fwd+bwd %.4f'%(t(f),t(fb)))
    qb,kb,vb=[x.detach().bfloat16().requires_grad_(True) for x in (q,k,v)]
    f2=lambda: F.scaled_dot_product_attention(qb,kb,vb)
    print('sdpa bf16 fwd ms %.4f'%t(f2), 'maxdiff vs fp32', float((f2().float()-o).abs().max()))
