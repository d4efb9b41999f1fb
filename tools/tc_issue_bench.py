import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib
lib = _lib.load()
out = torch.zeros(4, device="cuda")
scratch = torch.zeros(148 * 12288, device="cuda")
names = {2: "converged", 10: "+rotate operands", 26: "+halo-style A desc", 42: "rotate + TMA traffic", 58: "rotate+halo+TMA"}
for N in (64, 128, 256):
    row = []
    for variant in (2, 10, 26, 42, 58):
        out.zero_()
        _lib.check(lib.vt_selftest_tc_gemm(scratch.data_ptr(), None, out.data_ptr(), 0, N, 4000, variant, None))
        torch.cuda.synchronize()
        extra = f" ({out[1].item() / (out[0].item() * 16000):5.1f} B/clk TMA)" if variant & 32 else ""
        row.append(f"{names[variant]}: {out[0].item():6.1f}{extra}")
    print(f"N={N:3d} ideal {N // 2:3d} cyc/MMA | " + " | ".join(row))
