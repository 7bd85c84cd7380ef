import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/m2-tts_b200/src')
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
for L in (1, 64, 65, 128, 129, 192, 193, 256, 257, 300, 512):
    x = torch.randn(2, L, 96, device="cuda")
    with nat.precision("tf32"):
        a = m.decoder(x).clone()
    b = m.decoder(x).clone()
    torch.cuda.synchronize()
    print(L, float((a-b).abs().max()), flush=True)
for L in (700, 1153, 1500, 3446):
    x = torch.randn(1, L, 96, device="cuda")
    with nat.precision("tf32"):
        a = m.decoder(x).clone()
    b = m.decoder(x).clone()
    torch.cuda.synchronize()
    print(L, float((a-b).abs().max()), flush=True)
for heads, hidden in ((2, 32), (2, 64), (2, 128), (4, 64)):
    kw = dict(STAGE_KWARGS["stage2"]); kw.update(hidden_dim=hidden, num_heads=heads)
    mm = M2TTSModel(**kw).eval().cuda()
    x = torch.randn(2, 333, hidden, device="cuda")
    with nat.precision("tf32"):
        a = mm.decoder(x).clone()
    b = mm.decoder(x).clone()
    torch.cuda.synchronize()
    print("hd", hidden // heads, float((a-b).abs().max()), flush=True)
