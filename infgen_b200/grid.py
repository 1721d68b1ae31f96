"""Ego-centric position grid / heading bins (host-side mirror of the reference `Attr_Tokenizer`).

Reference: `infgen/modules/attr_tokenizer.py:24-43` (grid construction), `:77-89` (encode_pos), `:91-99`
(decode_pos), `:101-110` (heading bins).  The table built here is uploaded to the device once per engine; the
per-step `encode_pos` argmin runs in the CUDA advance kernel.
"""
import math
import torch


class PositionGrid:
    """Square lattice of `grid_interval` spacing masked to a disc of `radius`, rows ordered y-descending, x-ascending."""

    def __init__(self, grid_range: float = 150.0, grid_interval: float = 3.0, radius: float = 75.0,
                 angle_interval: float = 3.0):
        num_grid = int(grid_range / grid_interval) + 1
        axis = torch.linspace(0, num_grid - 1, steps=num_grid)
        gx = axis[None, :].expand(num_grid, num_grid)            # x varies along columns
        gy = axis[:, None].expand(num_grid, num_grid)            # y varies along rows
        cells = torch.stack([gx, gy], dim=-1).flip(dims=[0]).reshape(-1, 2)   # y descending
        cells = (cells - num_grid // 2) * grid_interval
        dist = (cells ** 2).sum(-1).sqrt()
        self.square_mask = dist <= radius
        self.cells = cells[self.square_mask].contiguous()        # [G, 2] float32
        self.grid_size = int(self.cells.shape[0])
        self.center_index = self.grid_size // 2
        assert bool(torch.all(self.cells[self.center_index] == 0.0))
        self.heading = math.pi / 2                               # ego looks along +y in the grid frame
        self.angle_interval = angle_interval
        self.angle_size = int(360.0 / angle_interval)
        self.grid_interval = grid_interval
        self.radius = radius

    def encode_pos(self, x: torch.Tensor, y: torch.Tensor, theta_y: torch.Tensor) -> torch.Tensor:
        """Nearest cell of points x[N,2] seen from the ego at y[1,2] with heading theta_y[1] (host version)."""
        rel = x - y
        a = -(theta_y - self.heading)
        c, s = torch.cos(a), torch.sin(a)
        rx = rel[:, 0] * c - rel[:, 1] * s
        ry = rel[:, 0] * s + rel[:, 1] * c
        d = ((torch.stack([rx, ry], -1)[:, None] - self.cells[None]) ** 2).sum(-1).sqrt()
        return torch.argmin(d, dim=-1)
