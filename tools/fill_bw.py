import torch
def t(fn, reps=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
M, N = 186880, 2304
x = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
ms = t(lambda: x.zero_()); print(f"zero_ {x.numel()*2/1e6:.0f} MB: {ms:.4f} ms = {x.numel()*2/ms/1e9:.2f} TB/s write")
ms = t(lambda: x.fill_(1.5)); print(f"fill_: {ms:.4f} ms = {x.numel()*2/ms/1e9:.2f} TB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy_: {ms:.4f} ms = {2*x.numel()*2/ms/1e9:.2f} TB/s read+write")
ms = t(lambda: x.sum()); print(f"sum (read): {ms:.4f} ms = {x.numel()*2/ms/1e9:.2f} TB/s read")
s = x[:, :576]
ms = t(lambda: torch.relu(s)); print(f"relu strided read 215MB + write: {ms:.4f} ms")
