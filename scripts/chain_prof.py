"""Per-role cycle counters of the chain kernel (bring-up aid).  python scripts/chain_prof.py [workload] [envs] [precision]

The wait counters of the producer and MMA warps (prod_wait_empty, mma_wait_x, mma_wait_full) are compiled in only when the
library is built with DPPO_B200_CHAIN_PROF=1 (python -c "from dppo_b200 import build; build.build(force=True)" under that
environment variable); the default build reports 0 for them - the clock reads cost 7 % of the launch."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from dppo_b200 import _lib
from dppo_b200.workloads import get_workload
from tests.helpers import build_model, our_classes

name = sys.argv[1] if len(sys.argv) > 1 else "walker2d"
w = get_workload(name)
E = int(sys.argv[2]) if len(sys.argv) > 2 else w["n_envs"]
prec = sys.argv[3] if len(sys.argv) > 3 else "split3"
ne = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
model = build_model(w, "cuda:0", our_classes())
model.engine_precision = prec
eng = model.engine()
eng.set_launch_shape(ne, cl)
grid = ((E + 15) // 16) * 8
prof = torch.zeros((4096 + grid) * 16, dtype=torch.int64, device="cuda")  # detail rows of a PROF build start at row 4096
lib = _lib.load()
lib.dppo_debug_set_prof.argtypes = [C.c_void_p, C.c_void_p]
lib.dppo_debug_set_prof(eng.ctx, C.c_void_p(prof.data_ptr()))
state = torch.rand(E, 1, w["obs_dim"], device="cuda") * 2 - 1
for _ in range(3):
    eng.sample(state)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.sample(state)
b.record()
torch.cuda.synchronize()
p = prof.view(4096 + grid, 16).cpu()
used = p[:grid][p[:grid][:, 4] > 0]
if "unet" in name and cl == 2:  # track-split pair: even CTAs = main path, odd CTAs = FiLM encoders
    for r, lab in ((0, "main"), (1, "encoders")):
        sub = p[r::2][p[r::2][:, 4] > 0]
        print(f"  -- rank {r} ({lab}), {len(sub)} CTAs")
        for i, n in enumerate(["prod_wait_empty", "prod_total", "mma_wait_x", "mma_wait_full", "mma_total", "epi_wait_layer", "epi_total", "epi_film_wait", "epi_work_warp2"]):
            col = sub[:, i].double()
            print(f"     {n:16s} mean {col.mean() / 1e3:10.1f} kcyc   max {col.max() / 1e3:10.1f} kcyc")
names = ["prod_wait_empty", "prod_total", "mma_wait_x", "mma_wait_full", "mma_total", "epi_wait_layer", "epi_total", "epi_handshake"] + [f"epi_work_warp{i+2}" for i in range(8)]
print(f"{name} E={E} {prec} NE={ne} C={cl}: kernel {a.elapsed_time(b):.3f} ms, {len(used)} CTAs")
for i, n in enumerate(names):
    col = used[:, i].double()
    print(f"  {n:16s} mean {col.mean() / 1e3:10.1f} kcyc   max {col.max() / 1e3:10.1f} kcyc")

# detail rows of a DPPO_B200_CHAIN_PROF=1 build: the MMA warp's operand waits by (layer kind, tile class), early-order waits
n = len(used)
det = p[4096:4096 + n].double()
if det.abs().sum() > 0 and det[:, 4:].abs().sum() == 0:  # pair kernel (leader CTAs): own ring, follower ring, own X, follower X
    lead = det[det[:, 0] + det[:, 1] > 0]
    fol = used[1::2].double()
    print("  follower CTAs: producer wait-empty %.1fk of %.1fk; relay wait-full %.1fk of %.1fk" % (fol[:, 0].mean() / 1e3, fol[:, 1].mean() / 1e3, fol[:, 3].mean() / 1e3, fol[:, 4].mean() / 1e3))
    led = used[0::2].double()
    print("  leader CTAs  : producer wait-empty %.1fk of %.1fk" % (led[:, 0].mean() / 1e3, led[:, 1].mean() / 1e3))
    print("  pair leader waits: own ring %.1fk  follower ring (relay) %.1fk  own X/x0 %.1fk  follower X/x0 %.1fk" % tuple(lead[:, i].mean() / 1e3 for i in range(4)))
elif det.abs().sum() > 0:
    kinds, cls = ["layer0", "l1", "l2", "out"], ["own tile 0 / x0", "own tile 1", "peer tiles"]
    for k in range(4):
        print("  mma_wait_x %-7s" % kinds[k] + "  ".join(f"{cls[c]} {det[:, k * 3 + c].mean() / 1e3:8.1f}k" for c in range(3)))
    print(f"  epi wait tile_done {det[:, 12].mean() / 1e3:8.1f}k   wait early_ok {det[:, 13].mean() / 1e3:8.1f}k")
