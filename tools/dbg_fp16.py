import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccdm_b200, oracle
from oracle.unet_ref import UnetSpec, make_state_dict
RC64 = UnetSpec(dim=64, dim_mults=(1, 2, 2, 4, 8), in_channels=3, embed_input_dim=128, attn_dim_head=32, attn_heads=4)
dev = torch.device("cuda")
net = ccdm_b200.Unet(dim=64, dim_mults=(1, 2, 2, 4, 8), cond_drop_prob=0.1, precision="fp16")
net.load_state_dict(make_state_dict(RC64, 7)); net = net.to(dev).eval()
B, size = 4, 64
g = torch.Generator().manual_seed(3)
x = torch.randn(B, 3, size, size, generator=g).to(dev)
t = torch.tensor([17, 333, 650, 990], device=dev)
emb = oracle.y2h_sinusoidal(torch.linspace(0.1, 0.9, B, device=dev), 128)
y = net(x, t, emb, cond_drop_prob=0.0)
prog = net.engine().program(B, B, size, size, False)
for name, buf in prog.bufs.items():
    if buf.dtype == torch.bfloat16:
        h = buf.view(torch.float16).float()
        bad = (~torch.isfinite(h)).sum().item()
        print(f"{name:34s} max|x|={h[torch.isfinite(h)].abs().max().item() if bad < h.numel() else float('nan'):10.3f} nonfinite={bad}")
    elif buf.dtype == torch.float32:
        bad = (~torch.isfinite(buf)).sum().item()
        if bad: print(f"{name:34s} fp32 nonfinite={bad}")
for name, pk in net.engine().weights.packs.items():
    t_ = getattr(pk, 'packed', None)
    if t_ is not None and t_.dtype == torch.bfloat16:
        h = t_.view(torch.float16).float(); bad = (~torch.isfinite(h)).sum().item()
        if bad or h.abs().max() > 1000: print("PACK", name, h.abs().max().item(), bad)
