"""End-to-end step time (host arrays -> mask on the host) under the upload variants: marker bit packing on/off, z-chunked
upload on/off.  Usage: python tools/e2e_probe.py [size]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy
import torch
from medpy_b200 import synthetic
import medpy_b200.graphcut as gc

size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
vol = synthetic.two_blob_volume((size,) * 3, seed=0)
n = size ** 3
def pin(a):
    t = torch.from_numpy(numpy.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()
keep = [pin(vol["image"]), pin(vol["prob"]), pin(vol["fg"].view(numpy.uint8)), pin(vol["bg"].view(numpy.uint8))]
h_img, h_prob, h_fg, h_bg = (k[1] for k in keep)
h_fg = h_fg.view(numpy.bool_); h_bg = h_bg.view(numpy.bool_)
try:
    print(json.dumps({"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                      "cpu.max": open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None}))
except Exception as exc:
    print("cpu info failed", exc)

def step():
    g = gc.graph_from_voxels(h_fg, h_bg, regional_term=gc.energy_voxel.regional_probability_map, regional_term_args=(h_prob, vol["alpha"]),
                             boundary_term=gc.energy_voxel.boundary_difference_exponential, boundary_term_args=(h_img, vol["sigma"], False))
    t1 = time.perf_counter()
    e = g.maxflow()
    t2 = time.perf_counter()
    m = g.get_mask()
    return e, m, t1, t2

for pack, chunks in ((1, 8), (1, 8), (1, 8), (0, 8)):
    os.environ["MEDPY_GC_PACK_MARKERS"] = str(pack)
    os.environ["MEDPY_GC_CHUNKS"] = str(chunks)
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(16):
        t0 = time.perf_counter()
        e, m, t1, t2 = step()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ts.append((1e3 * (t3 - t0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
    best = min(ts)
    print(json.dumps({"all_totals_ms": [round(t[0], 2) for t in ts]}))
    worst = max(ts)
    print(json.dumps({"worst_step_ms": round(worst[0], 2), "build_call": round(worst[1], 2), "maxflow": round(worst[2], 2), "get_mask": round(worst[3], 2)}))
    print(json.dumps({"pack": pack, "chunks": chunks, "total_ms": best[0], "build_call_ms": best[1], "maxflow_ms": best[2], "get_mask_ms": best[3], "energy": e}))
# raw H2D rate of the same pinned buffers
d = torch.empty(n, dtype=torch.float32, device="cuda")
src = keep[0][0].reshape(-1)
torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"h2d_gbs_pinned_512MB": n * 4 / (t1 - t0) / 1e9}))
hm = torch.empty(n, dtype=torch.uint8).pin_memory(); dm = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter(); hm.copy_(dm, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
print(json.dumps({"d2h_gbs_pinned_128MB": n / (t1 - t0) / 1e9}))
# host-side packing alone (numpy packbits as a yardstick) 
t0 = time.perf_counter(); numpy.packbits(h_fg.view(numpy.uint8).reshape(-1), bitorder="little"); t1 = time.perf_counter()
print(json.dumps({"numpy_packbits_one_plane_ms": 1e3 * (t1 - t0)}))
