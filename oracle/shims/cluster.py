"""torch_cluster stand-in: `radius` / `radius_graph` with the CUDA-kernel semantics of torch_cluster 1.6.3.

TEST INFRASTRUCTURE ONLY.  Semantics (the contract parity is judged against):
strict `dist^2 < r^2`; only pairs with equal batch id; per y-point keep the first `max_num_neighbors`
x-points in ascending x index; result rows are (y_idx, x_idx), sorted by y then x.
Vectorised per batch segment so the CPU baseline is not strawmanned by a python double loop.
"""
import torch


def radius(x, y, r, batch_x=None, batch_y=None, max_num_neighbors=32, num_workers=1):
    if batch_x is None:
        batch_x = torch.zeros(x.size(0), dtype=torch.long, device=x.device)
    if batch_y is None:
        batch_y = torch.zeros(y.size(0), dtype=torch.long, device=y.device)
    r2 = float(r) * float(r)
    rows, cols = [], []
    all_x = torch.arange(x.size(0), device=x.device)
    all_y = torch.arange(y.size(0), device=y.device)
    for b in torch.unique(batch_y).tolist():
        my = batch_y == b
        mx = batch_x == b
        if not bool(mx.any()):
            continue
        yi, xi = all_y[my], all_x[mx]
        d = y[my][:, None, :] - x[mx][None, :, :]
        within = (d * d).sum(-1) < r2                      # [ny, nx]
        rank = within.cumsum(dim=1)
        keep = within & (rank <= max_num_neighbors)
        nz = keep.nonzero()
        rows.append(yi[nz[:, 0]])
        cols.append(xi[nz[:, 1]])
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long, device=x.device)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, flow='source_to_target', num_workers=1):
    assert flow == 'source_to_target'
    ei = radius(x, x, r, batch, batch, max_num_neighbors if loop else max_num_neighbors + 1)
    row, col = ei[1], ei[0]                               # row = x_idx (source), col = y_idx (target)
    if not loop:
        keep = row != col
        row, col = row[keep], col[keep]
    return torch.stack([row, col], dim=0)
