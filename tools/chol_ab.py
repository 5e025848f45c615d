import torch, numpy as np
dev=torch.device("cuda:0")
n=2402
rng=np.random.default_rng(0)
B=rng.normal(size=(n,n+8)); A=torch.from_numpy(B@B.T+n*1e-3*np.eye(n)).to(dev)
def timeit(fn,reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
print("lower", timeit(lambda: torch.linalg.cholesky(A)))
print("upper", timeit(lambda: torch.linalg.cholesky(A, upper=True)))
print("cholesky_ex lower", timeit(lambda: torch.linalg.cholesky_ex(A)))
L=torch.linalg.cholesky(A); b=torch.randn(n,1,dtype=torch.float64,device=dev)
print("cholesky_solve", timeit(lambda: torch.cholesky_solve(b,L)))
print("triangular x2", timeit(lambda: torch.linalg.solve_triangular(L.T, torch.linalg.solve_triangular(L,b,upper=False), upper=True)))
