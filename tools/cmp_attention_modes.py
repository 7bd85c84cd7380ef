"""Bring-up helper (GPU box): decoder output of the 16-bit split attention (mode 0) against the TF32 split (mode 3)
for sequence lengths around the key-tile boundaries."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/m2-tts_b200/src')
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
lib=nat.lib()
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
for L in (1, 64, 65, 128, 129, 192, 193, 256, 257, 300, 512, 700):
    x = torch.randn(2, L, 96, device="cuda")
    lib.m2tts_set_attention_mode(3); a = m.decoder(x).clone()
    lib.m2tts_set_attention_mode(0); b = m.decoder(x).clone()
    torch.cuda.synchronize()
    d=(a-b).abs()
    print(L, float(d.max()), "rows bad:", (d.amax(dim=(0,2))>1e-3).nonzero().flatten().tolist()[:12])
