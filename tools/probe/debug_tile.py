import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import neurite_b200 as ne
from oracle import interp as ointerp
rng = np.random.default_rng(0)
vol = rng.standard_normal((1, 20, 40, 64, 1)).astype(np.float32)
flow = rng.uniform(-3, 3, (1, 20, 40, 64, 3)).astype(np.float32)
out = ne.layers.SpatialTransformer()([torch.from_numpy(vol).cuda(), torch.from_numpy(flow).cuda()])
torch.cuda.synchronize()
print('tile kernel ran; equal to oracle:', np.array_equal(out.cpu().numpy(), ointerp.spatial_transformer(vol, flow)))
