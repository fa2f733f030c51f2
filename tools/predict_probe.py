"""Throughput of the batched predictive mean/variance kernel chain at a GEMM-bound size."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'gp-plus_b200')
import numpy as np, torch
import bench
from gpplus_b200 import _engine as E
for n, m in ((8192, 65536), (2048, 262144), (512, 1048576)):
    X, y = bench.W.c4_workload(n)
    ys = (y - y.min()) / (y.max() - y.min())
    h = bench.W.c4_natural(np.zeros(13))
    eng = E.Engine(xq=X, y=ys, kernel=E.KERNEL_MATERN52, n_noise=1, n_mean=1, device=0)
    eng.factorize(h)
    Xc = np.random.RandomState(0).randn(m, 10)
    Xd = torch.from_numpy(Xc).cuda()
    mu = torch.empty(m, dtype=torch.float64, device="cuda"); var = torch.empty_like(mu)
    eng.predict(Xd, out_mean=mu, out_var=var); torch.cuda.synchronize()
    t0 = time.time(); eng.predict(Xd, out_mean=mu, out_var=var); torch.cuda.synchronize(); dt = time.time() - t0
    t0 = time.time(); mh, vh = eng.predict(Xc); dth = time.time() - t0
    np_ = (n + 127) // 128 * 128
    print("n=%d m=%d: device-resident %.1f ms (%.2f M pred/s, %.1f TFLOP/s on the triangular V product), from host %.1f ms"
          % (n, m, dt * 1e3, m / dt / 1e6, m * float(np_) * np_ / dt / 1e12, dth * 1e3), flush=True)
    assert np.allclose(mh, mu.cpu().numpy()) and np.allclose(vh, var.cpu().numpy())
    eng.close()
