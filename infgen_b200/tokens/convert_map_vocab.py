"""Convert the reference map-token vocabulary to a pickle-free .npz (run once, in the build container).

Source: /root/reference/infgen/tokens/map_traj_token5.pkl -> `traj_src` [1024, 11, 2] (5 m polyline tokens sampled at 11
points, token-local frame; reference `infgen/model/infgen.py:202-211`, consumed by `map_decoder.py:79-81`).  Data, not
source; stored as float32 exactly as the reference casts it (`torch.from_numpy(...).to(torch.float)`).
"""
import pickle
import sys
import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/infgen/tokens/map_traj_token5.pkl'
dst = sys.argv[2] if len(sys.argv) > 2 else __file__.replace('convert_map_vocab.py', 'map_traj_token5.npz')
tok = pickle.load(open(src, 'rb'))
traj = np.ascontiguousarray(np.asarray(tok['traj_src'], dtype=np.float64).astype(np.float32))
assert traj.shape == (1024, 11, 2), traj.shape
np.savez_compressed(dst, traj_src=traj)
print('wrote', dst, traj.shape)
