import torch
N, D, ld = 248428, 78, 160
X = torch.randn(N, ld, device="cuda")
Y = torch.randn(N, D, device="cuda"); Y2 = torch.empty_like(Y)
Z = torch.randn(N, 80, device="cuda")
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
big = torch.empty(64 * 1024 * 1024, device="cuda")
print("dense copy [N,78] us:", t(lambda: Y2.copy_(Y)))
print("interleaved copy X[:,78:156] = X[:,:78] us:", t(lambda: X[:, 78:156].copy_(X[:, :78])))
print("strided read X[:,:78] -> dense us:", t(lambda: Y2.copy_(X[:, :78])))
print("dense -> strided write X[:,78:156] us:", t(lambda: X[:, 78:156].copy_(Y)))
print("aligned interleave X[:,80:160] = X[:,:80] us:", t(lambda: X[:, 80:160].copy_(X[:, :80])))
