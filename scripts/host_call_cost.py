"""Host-side cost of one call of each hot C-ABI entry point (tiny problem, GPU mostly idle): wall time per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import unscene3d_b200  # noqa
from unscene3d_b200 import engine
from unscene3d_b200.engine import functional as Fn
from unscene3d_b200.synthetic import make_scene

dev = torch.device("cuda")
for nvox in (2000, 200_000):
    s = make_scene(nvox, seed=0, with_masks=False)
    c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
    x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
    cm, key = x0.coordinate_manager, x0.coordinate_map_key
    table = cm.forward_table(key, key, (3, 3, 3))
    cin = cout = 32
    x = torch.randn(s.n, cin, device=dev)
    dy = torch.randn(s.n, cout, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03

    def bench(name, fn, n=200):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"n={nvox:6d} {name:28s} host {(t1 - t) / n * 1e6:8.1f} us/call   with drain {(t2 - t) / n * 1e6:8.1f} us/call", flush=True)

    hi, lo = Fn.bf16_planes(x, True)
    wp = Fn.pack_weights(w, False, False, 3)
    y = torch.empty(s.n, cout, device=dev)
    from unscene3d_b200._lib import lib, check
    st = Fn._stream()
    bench("raw us3d_spconv_gather_mt", lambda: lib.us3d_spconv_gather_mt(hi.data_ptr(), lo.data_ptr(), s.n, table.nbr.data_ptr(), table.n_rows, 27,
                                                       wp.data_ptr(), cin, cout, 3, 0, 0, y.data_ptr(), cout, 0, table.mask.data_ptr(), 0, 0, 0, st))
    bench("Fn.spconv_gather", lambda: Fn.spconv_gather(x, table, w, cin, cout, False, False))
    bench("Fn.pack_weights", lambda: Fn.pack_weights(w, False, False, 3))
    bench("Fn.spconv_wgrad", lambda: Fn.spconv_wgrad(x, table, dy, cin, cout))
    rm, rv = torch.zeros(cin, device=dev), torch.ones(cin, device=dev)
    bench("Fn.bn_batch_stats", lambda: Fn.bn_batch_stats(x, rm, rv, 0.1, 1e-5))
    bench("torch.empty", lambda: torch.empty((s.n, cout), device=dev))
    bench("lib.us3d_relu", lambda: lib.us3d_relu(x.data_ptr(), y.data_ptr(), x.numel(), st))

# ---- alternating configurations (different dynamic shared memory sizes / ksplit paths), as the network issues them
print("alternating kernels:")
s = make_scene(20000, seed=0, with_masks=False)
c4 = torch.from_numpy(np.concatenate([np.zeros((s.n, 1), np.int32), s.coords], 1)).to(dev)
x0 = engine.SparseTensor(torch.zeros(s.n, 1, device=dev), c4)
cm, key = x0.coordinate_manager, x0.coordinate_map_key
k2 = cm.stride(key, (2, 2, 2)); k4 = cm.stride(k2, (2, 2, 2)); k8 = cm.stride(k4, (2, 2, 2))
cfgs = []
for kk, cin, cout in ((key, 32, 32), (k2, 96, 96), (k4, 128, 128), (k8, 256, 256)):
    n = cm.size(kk)
    t = cm.forward_table(kk, kk, (3, 3, 3))
    x = torch.randn(n, cin, device=dev)
    w = torch.randn(27, cin, cout, device=dev) * 0.03
    hi, lo = Fn.bf16_planes(x, True)
    wp = Fn.pack_weights(w, False, False, 3)
    y = torch.empty(n, cout, device=dev)
    cfgs.append((hi, lo, n, t, wp, cin, cout, y, x, w))
st = Fn._stream()


def call(c):
    hi, lo, n, t, wp, cin, cout, y, x, w = c
    lib.us3d_spconv_gather_mt(hi.data_ptr(), lo.data_ptr(), n, t.nbr.data_ptr(), t.n_rows, 27, wp.data_ptr(), cin, cout, 3, 0, 0,
                              y.data_ptr(), cout, 0, t.mask.data_ptr(), 0, 0, 0, st)


for c in cfgs:
    call(c)
torch.cuda.synchronize()
for label, order in (("same config x200", [cfgs[1]] * 200), ("alternating 4 configs", cfgs * 50),
                     ("small map (ksplit) x200", [cfgs[3]] * 200)):
    t = time.perf_counter()
    for c in order:
        call(c)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"  {label:28s} host {(t1 - t) / len(order) * 1e6:8.1f} us/call   with drain {(t2 - t) / len(order) * 1e6:8.1f} us/call", flush=True)
xs = torch.randn(1000, device=dev); ys = torch.empty_like(xs)
for label, fn in (("conv + small us3d kernel", lambda c: (call(c), lib.us3d_relu(xs.data_ptr(), ys.data_ptr(), xs.numel(), st))),
                  ("conv + torch kernel", lambda c: (call(c), ys.add_(1.0))),
                  ("small us3d kernel alone", lambda c: lib.us3d_relu(xs.data_ptr(), ys.data_ptr(), xs.numel(), st)),
                  ("torch kernel alone", lambda c: ys.add_(1.0))):
    fn(cfgs[1]); torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(200):
        fn(cfgs[i % 4])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"  {label:28s} host {(t1 - t) / 200 * 1e6:8.1f} us/iter   with drain {(t2 - t) / 200 * 1e6:8.1f} us/iter", flush=True)
