"""One profiled launch each of ccdm_rmsnorm_act and ccdm_block_bwd (C=64, 64x64, B=128) for `ncu --set full
--profile-from-start off`."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccdm_b200 import _lib as L
lib = L.lib()
st = torch.cuda.current_stream().cuda_stream
B, hw, c = 128, 64, 64
rows = B * hw * hw
z = torch.randn(B, hw, hw, c, device="cuda").bfloat16(); dy = torch.randn_like(z); o = torch.empty_like(z)
gain = torch.ones(c, device="cuda"); ss = 0.1 * torch.randn(B, 2 * c, device="cuda"); sums = torch.zeros(3, B, c, device="cuda")
def fwd(): L.check(lib.ccdm_rmsnorm_act(z.data_ptr(), o.data_ptr(), rows, c, hw * hw, gain.data_ptr(), math.sqrt(c), ss.data_ptr(), 2 * c, 0, None, None, L.EPI_SS | L.EPI_SILU, st))
def bwd(): L.check(lib.ccdm_block_bwd(dy.data_ptr(), z.data_ptr(), o.data_ptr(), rows, c, hw * hw, gain.data_ptr(), math.sqrt(c), ss.data_ptr(), 2 * c, 0, sums.data_ptr(), L.EPI_SS | L.EPI_SILU, st))
for _ in range(3): fwd(); bwd()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fwd(); bwd()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
