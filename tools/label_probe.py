"""Developer probe: timing of the label-volume kernel alone (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from platymatch_b200 import device as D
from platymatch_b200.synthetic import make_label_volume
shape, n = (384, 512, 512), 6000
rng = np.random.default_rng(1)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
centers = np.array(shape) / 2.0 + d * (np.array(shape) * 0.42) + rng.normal(0, 4.0, size=(n, 3))
vol = make_label_volume(shape, radius=(3.0, 5.0), seed=2, centers=centers, dtype=np.int32)
dev = torch.from_numpy(vol).cuda()
flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.fill_(1)
    D.label_centroids(dev, 1.0, table_size=n + 1, sync=False)
torch.cuda.synchronize()
print("ok")
