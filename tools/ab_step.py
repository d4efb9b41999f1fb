"""Same-process A/B of library options on the full VToonify-D step (B=4, 576x1024): configurations are interleaved round after
round so that clock / thermal drift hits all of them alike.   python tools/ab_step.py [rounds]   (on the GPU box)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vtoonify_b200 import _lib, ops
from vtoonify_b200.vtoonify import VToonify
from vtoonify_b200.weights import det_inputs, det_state_dict

lib = _lib.load()
dev = torch.device("cuda:0")
BASE = {b"tc_stage_policy": 1, b"tc_halo_pct": 60, b"rsu_epi": 1, b"instnorm_chunks": 296, b"tc_warp_store": int(os.environ.get("VT_TC_WARP_STORE", "1")), b"rsu_bstages": 6, b"tc_m_major": 1, "fuse_stats": True}
CONFIGS = {
    "new (all on)": {},
    "stage_policy 0": {b"tc_stage_policy": 0},
    "halo_pct 50": {b"tc_halo_pct": 50},
    "rsu_epi 0": {b"rsu_epi": 0},
    "instnorm small chunks": {b"instnorm_chunks": 0},
    "warp_store 0": {b"tc_warp_store": 0},
    "rsu_bstages 4": {b"rsu_bstages": 4},
    "tc_m_major 0": {b"tc_m_major": 0},
    "fuse_stats off": {"fuse_stats": False},
    "old (all off)": {b"rsu_bstages": 4, b"tc_m_major": 0, "fuse_stats": False, b"tc_stage_policy": 0, b"tc_halo_pct": 50, b"rsu_epi": 0, b"instnorm_chunks": 0, b"tc_warp_store": 0},
}
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with torch.no_grad():
    m = VToonify(backbone="dualstylegan").eval()
    m.load_state_dict(det_state_dict(m, seed=0)); m.to(dev)
    x, s = det_inputs(4, 576, 1024)
    x, s = x.to(dev), s.to(dev)
    res = {k: [] for k in CONFIGS}
    ref = None
    for r in range(rounds):
        for name, over in CONFIGS.items():
            for k, v in {**BASE, **over}.items():
                if isinstance(k, bytes):
                    lib.vt_set_option(k, v)
                else:
                    ops.set_option(k, v)          # Python-level routing switches
            for _ in range(2):
                y = m(x, s, d_s=0.5).clamp_(-1, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                y = m(x, s, d_s=0.5).clamp_(-1, 1)
            e1.record(); torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 5)
            if r == 0:
                if ref is None:
                    ref = y.clone()
                else:
                    print(f"  {name}: max|diff| vs first config {(y - ref).abs().max().item():.3e}")
    for name, v in res.items():
        print(f"{name:26s} mean {sum(v) / len(v):7.3f} ms   " + " ".join(f"{t:7.3f}" for t in v), flush=True)
