import sys; sys.path.insert(0,'/root/repo')
import torch, bench
g = bench.gyroid_device(512, 0, 512, 512, torch.device('cuda',0))
for iso in [0.0, 1e-4, 0.01, 0.1, 0.3001, -1.2001]:
    f = torch.tensor(iso, dtype=torch.float32).item()
    print(iso, int((g == f).sum()))
