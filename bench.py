#!/usr/bin/env python
"""bench.py -- Mvoxels/s to converged min-cut (BASELINE.json metric) on the headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size S]

A "step" is one complete pass of the hot path over one synthetic volume: energy terms (n-link stencil,
regional t-links, markers) -> max-flow to convergence -> mask + energy.  Workload (config.workload): BASELINE
config 3, a 512^3 fp32 two-blob volume with regional_probability_map t-links + boundary_difference_exponential
(sigma = RMS neighbour difference), fg ball / bg shell markers.

  value   : N_voxels * K / t, inputs resident in HBM, t = CUDA events around the K timed steps (max over ranks)
  e2e     : same metric through the public API (medpy_b200.graphcut.graph_from_voxels -> maxflow -> get_mask)
            from pinned HOST buffers, H2D and D2H inside the timed region
  roofline: dominant kernel class (push sweep or relabel sweep), algorithmic bytes / measured launch time
  cpu_baseline: the reference's BK solver (oracle/_ref, real reference sources) or the oracle port, timed on a
            bounded sample (a 256^3 volume of the same generator) on the host, 1 thread (BK is serial)

`--impl reference` times only that CPU arm (rank 0; other ranks exit 0).
N > 1: the volume is z-slab partitioned over the ranks (one process per GPU, NCCL halo exchange of border
heights / pushed flow); same total work => "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mvoxels/sec to min-cut"
UNIT = "Mvoxels/s"
B_ALG = 11  # algorithmic bytes / voxel of config 3: image 4 + prob 4 + fg 1 + bg 1 + mask 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        note = "nvidia-smi -lms 20 from the start of the warm-up steps to the end of the timed steps"
        if not sm:
            # the whole region can be shorter than nvidia-smi's start-up: one query right after it, GPU still hot
            try:
                line = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                      capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
                parts = [x.strip() for x in line.split(",")]
                sm.append(float(parts[1])); mx.append(float(parts[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
                note = "timed region shorter than nvidia-smi's start-up; single query immediately after it"
            except Exception:
                pass
        return {"sm_mhz": float(numpy.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "how": note}


def make_volume(size):
    from medpy_b200 import synthetic
    shape = (size, size, size)
    t0 = time.time()
    vol = synthetic.two_blob_volume(shape, seed=0)
    vol["shape"] = shape
    vol["gen_s"] = time.time() - t0
    return vol


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's BK on the host
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(size, repeats=1):
    """Times the reference's CPU path on a bounded sample: a `size`^3 volume from the same generator, same terms.
    Graph construction uses the oracle's vectorised numpy restatement + a bulk C fill (far faster than the
    reference's per-edge Python loop, ~1 us/edge), so the number flatters the reference."""
    from oracle import energy_terms as et
    from oracle import solvers
    vol = make_volume(size)
    n = int(numpy.prod(vol["shape"]))
    best = None
    for _ in range(repeats):
        t0 = time.time()
        prob = et.build_problem(vol["fg"], vol["bg"], regional=(vol["prob"], vol["alpha"]),
                                boundary=("difference_exponential", vol["image"], vol["sigma"], False))
        t1 = time.time()
        if solvers.have_ref():
            flow, mask, tm = solvers.solve_ref(prob)
            kind = "reference"
            fill, mf, ro = tm["fill_s"], tm["maxflow_s"], tm["readout_s"]
        else:
            flow, mask, tm = solvers.solve_port(prob)
            kind = "port"
            fill, mf, ro = tm["setup_s"], tm["maxflow_s"], 0.0
        total = (t1 - t0) + fill + mf + ro
        r = dict(kind=kind, n=n, terms_s=t1 - t0, fill_s=fill, maxflow_s=mf, readout_s=ro, total_s=total,
                 energy=flow, fg_voxels=int(mask.sum()))
        if best is None or r["total_s"] < best["total_s"]:
            best = r
    return best


def cpu_baseline_obj(r, size):
    return {
        "value": r["n"] / r["total_s"] / 1e6, "unit": UNIT, "cores": 1, "kind": r["kind"],
        "sample": "%d^3 two-blob volume, same generator and terms as the workload (1/%d of its voxels); "
                  "numpy terms %.2fs + C++ graph fill %.2fs + BK maxflow() %.2fs + read-out %.2fs"
                  % (size, max(1, (512 // size) ** 3), r["terms_s"], r["fill_s"], r["maxflow_s"], r["readout_s"]),
        "maxflow_only_value": r["n"] / max(r["maxflow_s"], 1e-9) / 1e6,
        "host_cores_available": os.cpu_count(),
    }


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    size = args.cpu_size
    # warm-up + timed steps, each step one bounded sample
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_run(size)
    t0 = time.time()
    rs = [cpu_reference_run(size) for _ in range(args.steps)]
    dt = time.time() - t0
    n = rs[0]["n"]
    total = sum(r["total_s"] for r in rs)
    value = n * len(rs) / total / 1e6
    best = min(rs, key=lambda r: r["total_s"])
    cb = cpu_baseline_obj(best, size)
    cb["value"] = value
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(rs), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 3 terms (regional_probability_map + boundary_difference_exponential, "
                               "sigma=RMS) on a bounded %d^3 sample of the 512^3 two-blob fp32 volume" % size,
                   "solver": "reference BK (lib/maxflow/src, -O2 -DNDEBUG)" if best["kind"] == "reference" else "oracle BK port",
                   "wall_s": dt},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    size = args.size
    shape = (size, size, size)
    n = int(numpy.prod(shape))
    peak, peak_kind = measured_peak()

    if world > 1:
        from medpy_b200 import distributed as mdist
        result = mdist.bench_slab(shape, args, rank, world, local_rank)   # each rank builds only its own planes
        vol = {"sigma": result["sigma"]}
    else:
        vol = make_volume(size)
        result = bench_single(vol, args, torch)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    out = {
        "metric": METRIC, "value": result["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": result["ms_per_step"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE config 3: %d^3 fp32 two-blob volume, regional_probability_map (alpha 0.1) + "
                               "boundary_difference_exponential (sigma = RMS neighbour difference %.4f), fg balls / bg shell"
                               % (size, vol["sigma"]),
                   "shape": list(shape), "state_dtype": "float64", "parallelism": "zslab%d" % world if world > 1 else "single",
                   "l2": "inputs (%.1f GB) and solver state (%.1f GB) exceed the 126 MB L2; no flush needed"
                         % (n * 10 / 1e9, n * 76 / 1e9),
                   "energy": result.get("energy"), "fg_voxels": result.get("fg_voxels"),
                   "push_sweeps_per_step": result.get("push_sweeps"), "global_relabels_per_step": result.get("global_relabels"),
                   "relabel_sweeps_per_step": result.get("relabel_sweeps"),
                   "hbm_read_roofline_frac": (n * B_ALG / (result["ms_per_step"] * 1e-3)) / (peak * 1e9 * world),
                   "hbm_read_roofline_note": "N*%d B / t_step / (%s peak %.0f GB/s * n_gpus), SURVEY.md §8d" % (B_ALG, peak_kind, peak)},
        "clocks": result["clocks"],
        "e2e": result["e2e"],
        "gpu_launches": result["gpu_launches"],
        "roofline": result["roofline"],
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(args.cpu_size)
        out["cpu_baseline"] = cpu_baseline_obj(r, args.cpu_size)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


def roofline_from_stats(stats_list, n, peak, peak_kind):
    """Roofline of the dominant kernel (by device time in the timed region).  Algorithmic bytes per launch:
      k_boundary  (n-link stencil)   : image 4 B read + six float64 capacities 48 B written      = 52 B/voxel
      k_init_tile (solver state)     : 6 caps + tr read (56 B), excess/sink/label/rmask written (21 B) = 77 B/voxel
      k_push_tile (push pass)        : every voxel's excess + label must be inspected             = 12 B/voxel (dense bound)
      k_relabel_tile (relabel pass)  : residual mask + label                                      =  5 B/voxel (dense bound)
    """
    k = len(stats_list)
    cand = [
        ("k_boundary (n-link stencil, K1)", 52, sum(s["ms_boundary"] for s in stats_list), k),
        ("k_init_tile (solver state init)", 77, sum(s.get("ms_init", 0.0) for s in stats_list), k),
        ("k_push_tile (two-colour push pass)", 12, sum(s["ms_push"] for s in stats_list), sum(s["push_sweeps"] for s in stats_list)),
        ("k_relabel_tile (global relabel)", 5, sum(s["ms_relabel"] for s in stats_list), sum(s["global_relabels"] for s in stats_list)),
    ]
    name, bpv, ms, cnt = max(cand, key=lambda c: c[2])
    cnt = max(cnt, 1)
    # DRAM bytes of that kernel from the committed ncu --set full capture of this workload (profiles/), per launch
    traffic = None
    try:
        if n == 512 ** 3:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_512cubed.json")))["bytes_per_launch"]
            key = {"k_boundary": "k_boundary", "k_init_tile": "k_init_tile", "k_push_tile": "k_push_tile_tma",
                   "k_relabel_tile": "k_bfs_coop"}[name.split(" ")[0]]
            traffic = tj.get(key)
    except Exception:
        traffic = None
    avg = ms / cnt
    achieved = n * bpv / (avg * 1e-3) / 1e9 if avg > 0 else 0.0
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_kind": peak_kind, "launches": cnt, "avg_launch_ms": avg,
            "algorithmic_bytes_per_launch": n * bpv,
            "share_of_step": {c[0].split(" ")[0] + "_ms": c[2] for c in cand} | {
                "solve_ms": sum(s["ms_solve"] for s in stats_list), "terms_ms": sum(s["ms_terms"] for s in stats_list),
                "readout_ms": sum(s["ms_readout"] for s in stats_list)}}


def bench_single(vol, args, torch):
    import medpy_b200.graphcut as gc
    from medpy_b200.graphcut.device import graph_from_device_arrays
    shape = vol["shape"]
    n = int(numpy.prod(shape))
    peak, peak_kind = measured_peak()
    stream = torch.cuda.current_stream()

    # ---- inputs resident in HBM ----
    d_img = torch.from_numpy(vol["image"]).cuda()
    d_prob = torch.from_numpy(vol["prob"]).cuda()
    d_fg = torch.from_numpy(vol["fg"].view(numpy.uint8)).cuda()
    d_bg = torch.from_numpy(vol["bg"].view(numpy.uint8)).cuda()
    d_mask = torch.empty(shape, dtype=torch.uint8, device="cuda")
    graph = None
    stats = []

    def step(record):
        nonlocal graph
        graph = graph_from_device_arrays(d_fg, d_bg, image=d_img, boundary="difference_exponential", sigma=vol["sigma"],
                                         prob=d_prob, alpha=vol["alpha"], graph=graph, stream=stream.cuda_stream)
        e = graph.maxflow()
        graph._nat().get_mask_into(d_mask.data_ptr())
        if record:
            stats.append(graph.stats())
        return e

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    energy = None
    for _ in range(args.steps):
        energy = step(True)
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    fg_vox = int(d_mask.sum().item())
    value = n * args.steps / (ms * 1e-3) / 1e6
    launches = int(sum(s["kernel_launches"] for s in stats))
    roof = roofline_from_stats(stats, n, peak, peak_kind)

    # ---- end to end through the public API from pinned host buffers ----
    def pin(a):
        t = torch.from_numpy(numpy.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = [pin(vol["image"]), pin(vol["prob"]), pin(vol["fg"].view(numpy.uint8)), pin(vol["bg"].view(numpy.uint8))]
    h_img, h_prob, h_fg, h_bg = (k[1] for k in keep)
    h_fg = h_fg.view(numpy.bool_)
    h_bg = h_bg.view(numpy.bool_)
    del graph, d_img, d_prob, d_fg, d_bg, d_mask
    torch.cuda.empty_cache()

    def e2e_step():
        g = gc.graph_from_voxels(h_fg, h_bg, regional_term=gc.energy_voxel.regional_probability_map,
                                 regional_term_args=(h_prob, vol["alpha"]),
                                 boundary_term=gc.energy_voxel.boundary_difference_exponential,
                                 boundary_term_args=(h_img, vol["sigma"], False))
        e = g.maxflow()
        m = g.get_mask()
        return e, m

    # warm-up in the steady-state shape: the previous step's mask is still referenced while the next step runs (as in
    # the timed loop), so the pinned read-back pool already holds the two buffers that pattern needs
    e2e_warm = max(2, min(args.warmup, 3))
    for _ in range(e2e_warm):
        e_e2e, m_e2e = e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_e2e, m_e2e = e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e2e = {"value": n * args.steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(n * 10), "d2h_bytes_per_step": int(n + 8),
           "ms_per_step": 1e3 * dt / args.steps, "api": "medpy_b200.graphcut.graph_from_voxels -> maxflow -> get_mask",
           "energy_matches_resident_run": bool(e_e2e == energy), "timer": "host perf_counter around synchronised steps "
           "(H2D staging and the D2H mask copy are synchronous host calls inside)"}
    return {"value": value, "ms_per_step": ms / args.steps, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roof, "energy": energy, "fg_voxels": fg_vox,
            "push_sweeps": stats[-1]["push_sweeps"], "global_relabels": stats[-1]["global_relabels"],
            "relabel_sweeps": stats[-1]["relabel_sweeps"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=int(os.environ.get("MEDPY_BENCH_SIZE", "512")))
    ap.add_argument("--cpu-size", type=int, default=int(os.environ.get("MEDPY_BENCH_CPU_SIZE", "256")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
