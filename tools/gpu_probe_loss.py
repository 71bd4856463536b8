"""Quick device timing of the fused loss (not a test; used while tuning)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from uaps_b200 import _lib as L

def run(K, C, B, H, W, flags=0, iters=10):
    dev = torch.device("cuda:0")
    z = [torch.randn(B, C, H, W, device=dev) * 2 for _ in range(K)]
    dz = [torch.empty_like(t) for t in z]
    lib = L.lib()
    ws = torch.zeros(lib.uaps_loss_workspace_bytes(K, C), dtype=torch.uint8, device=dev)
    sums = torch.empty(lib.uaps_loss_sums_count(K, C), dtype=torch.float64, device=dev)
    sc = torch.empty(lib.uaps_loss_scalars_count(K, C), dtype=torch.float32, device=dev)
    go = torch.zeros_like(sc); go[0] = 1.0
    w = L.float_array([1.0 / K] * K)
    zp, dzp = L.ptr_array(z), L.ptr_array(dz)
    st = L.stream_ptr()
    N = B * H * W
    def p1(): L.check(lib.uaps_loss_pass1(zp, K, B, C, H * W, w, None, ws.data_ptr(), sums.data_ptr(), None, None, flags, st), "p1")
    def fin(): L.check(lib.uaps_loss_finalize(sums.data_ptr(), K, C, N, 0.1, 0.1, 0, sc.data_ptr(), st), "fin")
    def p2(): L.check(lib.uaps_loss_pass2(zp, K, B, C, H * W, w, None, sc.data_ptr(), go.data_ptr(), dzp, flags, None, st), "p2")
    for _ in range(3): p1(); fin(); p2()
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("pass1", p1), ("pass2", p2), ("all", lambda: (p1(), fin(), p2()))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / iters
    b1, b2 = 4 * K * C * N, 8 * K * C * N
    print(f"K={K} C={C} B={B} {H}x{W} flags={flags}: pass1 {res['pass1']*1e3:.1f}us {b1/res['pass1']/1e6:.0f} GB/s | "
          f"pass2 {res['pass2']*1e3:.1f}us {b2/res['pass2']/1e6:.0f} GB/s | all {res['all']*1e3:.1f}us "
          f"{(b1+b2)/res['all']/1e6:.0f} GB/s {N/res['all']/1e6:.2f} Gpx/s", flush=True)

if __name__ == "__main__":
    for cfg in [(4,4,64,256,256),(4,4,64,512,512),(4,4,8,200,200),(4,2,32,512,512),(5,2,32,256,512),(2,2,64,512,512),(6,4,32,512,512),(4,4,64,256,256,1)]:
        run(*cfg)
