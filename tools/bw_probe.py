"""Write-only / copy bandwidth probe (context for the observation kernel's roofline)."""
import torch
n = 1048576 * 36 * 49
a = torch.empty(n, dtype=torch.float32, device='cuda')
b = torch.empty(n, dtype=torch.float32, device='cuda')
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
ms = t(lambda: a.fill_(1.0)); print(f'fill  {n*4/1e9:.2f} GB in {ms:.3f} ms -> {n*4/ms/1e6:.0f} GB/s (write only)')
ms = t(lambda: a.zero_()); print(f'zero  {n*4/1e9:.2f} GB in {ms:.3f} ms -> {n*4/ms/1e6:.0f} GB/s (write only)')
ms = t(lambda: b.copy_(a)); print(f'copy  {2*n*4/1e9:.2f} GB in {ms:.3f} ms -> {2*n*4/ms/1e6:.0f} GB/s (read+write)')
ms = t(lambda: a.sum()); print(f'sum   {n*4/1e9:.2f} GB in {ms:.3f} ms -> {n*4/ms/1e6:.0f} GB/s (read only)')
