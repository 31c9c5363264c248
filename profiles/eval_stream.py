"""Streams 4 Mi random 16x16 candidate layouts (compact format, 32 B each = 128 MiB > L2) through kernel (a) on the
engine's stream; used plain and under ncu for profiles/r1_eval_kernel.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import timberborn_support_solver_b200 as T  # noqa: E402

eng = T.Engine(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
n = 4 << 20
r = lambda: torch.randint(0, 1 << 16, (n, 16), dtype=torch.int32, device="cuda")
lay = (r() & r() & r() & r()).to(torch.int16).contiguous()
grid = torch.full((16,), -1, dtype=torch.int16, device="cuda")
out = torch.empty((n, 2), dtype=torch.int32, device="cuda")
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    eng.eval_compact_dev(grid.data_ptr(), 16, 16, lay.data_ptr(), n, out.data_ptr())
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ms = sum(ts[3:]) / len(ts[3:])
print(f"eval 16x16: {n / ms / 1e6:.1f} G layouts/s, {n * 40 / ms / 1e6:.0f} GB/s algorithmic, uncovered mean {out[:, 0].float().mean().item():.2f}")
