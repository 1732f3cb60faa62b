import sys; sys.path.insert(0,'/root/repo')
import torch, bench
g = bench.gyroid_device(512, 0, 512, 512, torch.device('cuda',0))
for iso in bench.ISOS:
    f = torch.tensor(iso, dtype=torch.float32).item()
    m = (g == f)
    print(iso, int(m.sum()), [tuple(int(v) for v in r) for r in m.nonzero()[:6]])
