import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops
lib = _lib.load()
dev = torch.device("cuda:0")

def run(B, Cin, Cout, H, W, k, reps=5):
    x = torch.randn((B, H, W, Cin), device=dev)
    w = ops.prep_weights(torch.randn((Cout, Cin, k, k), device=dev) / (k * Cin ** 0.5), cin_pad=Cin)
    fn = lambda: ops.conv2d_nhwc([x], w, ops.conv_taps(k, k // 2), 1, H, W)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    return ms, fl / ms / 1e9

with torch.no_grad():
    for cfg in [(4, 512, 512, 72, 128, 1), (4, 2048, 512, 72, 128, 1), (4, 4096, 256, 144, 128, 1), (4, 512, 512, 72, 128, 3), (4, 512, 512, 144, 256, 3), (8, 512, 512, 144, 256, 3)]:
        line = f"{cfg}: "
        for cg2 in (0, 1):
            lib.vt_set_option(b"tc_cg2", cg2)
            ms, tf = run(*cfg)
            line += f" cg2={cg2}: {ms:7.3f} ms {tf:6.0f} TF/s |"
        print(line)
