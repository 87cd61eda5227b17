import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vfidkr_b200 as V
path, C, s = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B, H, W = 8, 1152, 1984
a = torch.randn(B, C, H // s, W // s, device="cuda"); b = torch.randn_like(a)
V.debug_force_correlation_path(path)
corr = V.Correlation(4, 1, 4, 1, 1, 1)
with torch.no_grad():
    for _ in range(4):
        o = corr(a, b)
torch.cuda.synchronize()
print("ok", float(o.abs().max()))
