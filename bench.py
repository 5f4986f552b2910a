#!/usr/bin/env python
"""bench.py -- Mrays/s closest-hit on the sibenik stand-in at 4K (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1|c4|c5] [--gather MODE]

Workload (default, all N): config C3 -- sibenik stand-in (75 256 triangles), `render -a 0 -w 3840 -h 2160
-s 16` => 15360 x 8640 = 132 710 400 primary rays per frame on the reference's regular 4x4 sample grid,
image tile-partitioned over the N GPUs (strong scaling).  A "step" = one frame: the traversal kernels on every rank +
RayTracer::resize on the device -> the byte image on rank 0 (N > 1: every rank's resize kernel stores its bytes into
rank 0's memory over NVLink peer memory, `--gather auto`; the NCCL-gather modes are timed beside it in `other_gather`).

  value      rays of the frame / device time of the step (CUDA events on the launching stream, max over
             ranks), scene resident in HBM, L2 flushed between steps.
  e2e        the same frame through the C ABI on HOST buffers, copies inside the timed region.  N = 1: rtx_upload of the
             five reference arrays + rtx_render_download of the float image (tracing and copy pipelined over bands).
             N > 1: every rank rtx_upload + rtx_render_store into ONE page-locked host image all rank processes map, so
             each rank's tiles leave over its own PCIe link.  e2e_u8: device resize + byte download.
  roofline   top level = the level that binds, instruction issue: warp instructions of the dominant kernel (ncu capture of
             THIS build, profiles/traffic.json stamped with the kernel sources' sha256; null when stale or N > 1) over its
             measured time, against SMs x 4 schedulers x the sampled SM clock.  Beside it: the DRAM / L2 / L1 bytes of the
             same capture (hbm_actual, memory_levels) and SURVEY 8d's algorithmic-bytes figure (algorithmic_vs_hbm:
             B = 32 V + 48 T + 48 h + 4 per ray under the reference's exhaustive walk, > 1 because that work is not done).
  cpu_baseline  the reference's own kernel text (oracle/_ref) -- or the C port when it is absent -- on the box's host
             cores, on a fixed row sample of the same frame (every 16th row).
  phases_ms  device time per launch group of a frame; extras (N = 1): configs C4 and C5, an irregular interior, other paths.

--impl reference runs only the CPU arm (rank 0) with the same JSON shape and the same `config`; it loads nothing of this
repo's product (scene from the reference's own mesh.cc + bvh.cc in oracle/_ref).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, nSuperSamples, scene, description)
    "c3": (3840, 2160, 16, "sibenik", "C3 sibenik-standin 3840x2160 s=16 (15360x8640 = 132.7M rays, regular grid)"),
    "c2": (1920, 1080, 4, "sibenik", "C2 sibenik-standin 1920x1080 s=4 (3840x2160 = 8.29M rays)"),
    "c1": (600, 600, 4, "bunny", "C1 bunny 600x600 s=4 (1200x1200 = 1.44M rays)"),
    "c4": (3840, 2160, 1, "bunny_x144", "C4 bunny subdivided x144 (10.16M triangles, 20.3M nodes) 3840x2160 s=1 (8.29M rays)"),
    # C5 is not a frame: 2^28 generated rays against the C2 tree, sharded by contiguous index ranges (bench_c5)
    "c5": (32, 32, 1, "sibenik", "C5 2^28 random-direction rays vs the sibenik stand-in tree (counter-hash generator, seed 1234)"),
}


METRIC = {"c3": "Mrays/s closest-hit (sibenik, 4K)", "c2": "Mrays/s closest-hit (sibenik, 1080p s=4)",
          "c1": "Mrays/s closest-hit (bunny, reference defaults)", "c4": "Mrays/s closest-hit (bunny x144, 4K)"}


CPU_ROW_STEP = 16      # both CPU legs time every 16th row of the frame (rows 8, 24, ...): fixed, so both arms can state it


def make_config(desc, scene_name, ntris, tw, th, world, gather):
    """The `config` object, identical in the GPU arm and in --impl reference (the driver compares the two)."""
    nrows = len(range(CPU_ROW_STEP // 2, th, CPU_ROW_STEP))
    how = {"auto": "every rank's resize kernel stores its bytes into rank 0's memory over NVLink, ordered by a one-element all-reduce "
                   "(one NCCL gather of the bytes + de-interleave where the GPUs cannot map each other's memory)",
           "p2p_u8": "every rank's resize kernel stores its bytes into rank 0's memory over NVLink, ordered by a one-element all-reduce",
           "p2p_float": "every rank's traversal kernel sends finished float tiles into rank 0's image over NVLink, ordered by a one-element all-reduce",
           "u8": "1 NCCL gather of the bytes + de-interleave", "float": "1 NCCL gather of the float tiles + de-interleave"}[gather]
    return {"workload": desc, "scene": scene_name, "triangles": int(ntris), "rays_per_step": int(tw * th),
            "parallelism": "interleaved 32x32 tiles over %d GPU(s), scene replicated, frame assembled on rank 0" % world,
            "step": "trace + RayTracer::resize on the device -> %s image on rank 0%s" % (
                "float" if gather.endswith("float") else "byte", "" if world == 1 else " (N > 1: " + how + ")"),
            "l2": "flushed between timed iterations (256 MiB memset, untimed)",
            "reference_arm_sample": "the CPU arm (--impl reference, cpu_baseline) times every %d-th row of the same %dx%d frame "
                                    "(%d rows = %.2f M rays per pass) on all host cores" % (CPU_ROW_STEP, tw, th, nrows, nrows * tw / 1e6)}


def load_profile(workload):
    """profiles/traffic.json entry of a workload -- only if it was captured on the kernel sources of THIS build
    (tools/make_traffic.py stamps the sha256 of csrc/ + include/rtx_b200.h).  (entry or None, why)"""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        from make_traffic import kernel_source_sha
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            entry = json.load(fh).get(workload)
    except (OSError, ImportError, ValueError) as e:
        return None, "profiles/traffic.json unreadable: %s" % e
    if not entry:
        return None, "no ncu capture of workload %s in profiles/traffic.json" % workload
    have, want = (entry.get("stamp") or {}).get("kernel_src_sha16"), kernel_source_sha()
    if have != want:
        return None, "stale: profiles/traffic.json[%s] was captured on kernel sources %s, this build is %s" % (workload, have, want)
    return entry, "ncu --set full, %s, kernel sources %s" % (entry["stamp"].get("report"), want)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def mesh_of(kind):
    """(vertices, faces) of a workload's mesh.  Generators only: no library of this repo is loaded."""
    from opencl_raytracer_b200 import scenes
    if kind in ("bunny", "bunny_x144"):
        from oracle import pyoracle as po          # only for the staged reference mesh file
        v, f = po.read_mesh_bin(po.staged_bunny_path())
        if kind == "bunny_x144":
            v, f = scenes.subdivided(v, f)
        return v, f, kind
    v, f = scenes.sibenik_standin()
    return v, f, "sibenik_standin"


def build_scene(kind):
    from opencl_raytracer_b200 import scene
    v, f, name = mesh_of(kind)
    return scene.scene_from_mesh(v, f, name=name)


def build_scene_reference(kind):
    """The scene of the reference arm: mesh.cc + bvh.cc of the reference itself (oracle/_ref, compiled from
    /root/reference in the build container), so that arm loads nothing of this repo's product.  Falls back to the
    host builder of this repo only when oracle/_ref is absent (then cpu_baseline.kind says "port")."""
    from oracle import pyoracle as po
    from opencl_raytracer_b200.scene import Scene            # a dataclass; importing it loads no library
    v, f, name = mesh_of(kind)
    if po.ref() is None:
        return build_scene(kind)
    r = po.ref_scene_from_mesh(v, f)
    return Scene(r.faces, r.nodes, r.aabbs, r.vertices, r.normals, r.triangles, r.orig_faces, name=name)


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of one GPU, sampled every 2 ms during the timed region (frames take a few ms)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while True:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if self.stop_flag:          # at least one sample is always taken
                break
            time.sleep(0.002)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=1.0)
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


class CpuArm:
    """The reference CPU path on a bounded row sample of the frame (every CPU_ROW_STEP-th row): the reference's own
    kernel text (oracle/_ref, kind "reference") when it was compiled in the build container, else the C port."""

    def __init__(self, sc, tw, th):
        from oracle import pyoracle as po
        self.po, self.sc, self.tw, self.th = po, sc, tw, th
        self.use_ref = po.ref() is not None
        self.cores = int(po.ref().ref_online_cpus() if self.use_ref else po.port().orc_online_cpus())
        self.rows = (CPU_ROW_STEP // 2, th, CPU_ROW_STEP)
        self.run()                                       # warm-up: page in, start threads

    @property
    def nrays(self):
        return len(range(*self.rows)) * self.tw

    def run(self):
        t = time.perf_counter()
        if self.use_ref:
            self.po.ref_render(self.sc, self.tw, self.th, 1.0, True, rows=self.rows)
        else:
            self.po.render(self.sc, self.tw, self.th, 1.0, True, rows=self.rows, want_ids=False)
        return time.perf_counter() - t

    def run_for(self, budget_s):
        """repeat the sample until budget_s of CPU work is done; (passes, seconds)"""
        n, total = 0, 0.0
        while n < 1 or total < budget_s:
            total += self.run()
            n += 1
        return n, total

    def info(self, mrays, dt, passes=1):
        return {"value": mrays, "unit": "Mrays/s", "cores": self.cores, "kind": "reference" if self.use_ref else "port",
                "sample": "every %d-th row of the %dx%d frame (%d rows, %.2fM rays, %.2f s per pass, %d pass(es)), %d pinned threads"
                          % (self.rows[2], self.tw, self.th, len(range(*self.rows)), self.nrays / 1e6, dt, passes, self.cores)}


def oracle_counters(sc, tw, th):
    """V, T, h of the frame under the reference's exhaustive walk, from a row sample (oracle counting mode)."""
    from oracle import pyoracle as po
    step = max(1, th // 64)
    r = po.render(sc, tw, th, 1.0, True, rows=(step // 2, th, step), want_ids=False, want_counters=True)
    c = r.counters
    return c["V"], c["T"], c["h"]


def extra_c4(local_rank, peak):
    """BASELINE config 4 inside the default line (N = 1): bunny x144 = 10 162 080 triangles (node array > L2) at
    3840x2160, one sample per pixel.  Tree built on the device from the raw mesh; value = best kernel time of 5 frames;
    parity = sampled rows against the oracle + the whole frame against the literal walk on the GPU."""
    from opencl_raytracer_b200 import host
    from oracle import pyoracle as po
    if not os.path.exists(po.staged_bunny_path()):
        return {"unavailable": "oracle/_ref/bunny_mesh.bin not staged"}
    sc = build_scene("bunny_x144")
    rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=1))
    tw, th = rt.totalWidth, rt.totalHeight
    out = {}
    with host.CudaHost(rt, device=local_rank) as h:
        t0 = time.perf_counter()
        h.upload_mesh(sc.vertices, sc.orig_faces, None)
        wall = (time.perf_counter() - t0) * 1e3
        bms, levels = h.build_stats()
        best = 1e9
        for _ in range(5):
            h()
            best = min(best, h.stats()["kernel_ms"])
        h.set_tunable(host.TUNE_RECORD_HITS, 1)          # parity frame: ids + distances recorded (not timed)
        h()
        img = h.download()
        fid, dist = h.download_hits()
    rows = (45, th, 90)
    ys = list(range(*rows))
    ref = po.render(sc, tw, th, 1.0, True, rows=rows, want_counters=True)
    V, T, hf = ref.counters["V"], ref.counters["T"], ref.counters["h"]
    with host.CudaHost(rt, device=local_rank) as h:
        h.set_tunable(host.TUNE_KERNEL, host.KERNEL_EXHAUSTIVE)
        h.set_tunable(host.TUNE_RECORD_HITS, 1)
        h.upload_scene(sc)
        h()
        lit_ms = h.stats()["kernel_ms"]
        fid_x, dist_x = h.download_hits()
        img_x = h.download()
    prof, prof_src = load_profile("c4")
    hbm = None
    if prof:
        tr = (prof.get("dram_bytes_read") or 0) + (prof.get("dram_bytes_write") or 0)
        hbm = {"dram_bytes_per_frame": tr, "GBps": tr / (best * 1e-3) / 1e9, "frac_of_peak": tr / (best * 1e-3) / 1e9 / peak,
               "warps_active_pct": prof.get("warps_active_pct"), "avg_threads_per_instruction": prof.get("avg_threads_per_instruction"),
               "l2_hit_pct": prof.get("l2_hit_pct"), "kernel": prof.get("kernel")}
    B = 32.0 * V + 48.0 * T + 48.0 * hf + 4.0
    out.update({
        "workload": WORKLOADS["c4"][4], "triangles": sc.num_triangles, "rays": tw * th, "kernel_ms": best, "Mrays/s": tw * th / best / 1e3,
        "device_build_ms": bms, "build_levels": levels, "rtx_upload_mesh_wall_ms": wall,
        "parity": {"rows_vs_oracle": len(ys), "rays_vs_oracle": len(ys) * tw,
                   "id_mismatches": int((fid[ys] != ref.face_id[ys]).sum()),
                   "distance_mismatches": int((dist[ys].view(np.uint32) != ref.distance[ys].view(np.uint32)).sum()),
                   "pixel_mismatches": int((img[ys].view(np.uint32) != ref.image[ys].view(np.uint32)).sum()),
                   "whole_frame_vs_literal_walk_on_gpu": {"id_mismatches": int((fid != fid_x).sum()),
                                                          "distance_mismatches": int((dist.view(np.uint32) != dist_x.view(np.uint32)).sum()),
                                                          "pixel_mismatches": int((img.view(np.uint32) != img_x.view(np.uint32)).sum()),
                                                          "literal_walk_ms": lit_ms}},
        "algorithmic_vs_hbm": {"bytes_per_ray": B, "V": V, "T": T, "h": hf, "GBps": B * tw * th / (best * 1e-3) / 1e9,
                               "frac": B * tw * th / (best * 1e-3) / 1e9 / peak},
        "hbm_measured": hbm, "profile": prof_src})
    return out


def extra_irregular(local_rank):
    """The frame of the headline workload on an IRREGULAR scene (scenes.cluttered_interior: 0.98 M triangles from sub-pixel
    to screen-filling): what the frustum front end is worth when tile lists overflow, next to the regular stand-in."""
    from opencl_raytracer_b200 import host, scene, scenes
    from oracle import pyoracle as po
    v, f = scenes.cluttered_interior()
    sc = scene.scene_from_mesh(v, f, name="cluttered_interior")
    rt = host.RayTracer(host.Options(width=3840, height=2160, nSuperSamples=16))
    tw, th = rt.totalWidth, rt.totalHeight
    out = {"scene": "scenes.cluttered_interior(): %d triangles, hundreds of finely tessellated spheres in a room with noise-displaced walls"
                    % sc.num_triangles, "rays": tw * th}
    for name, tun in (("auto", {}), ("per_ray_traversal", {host.TUNE_FRUSTUM: 0})):
        with host.CudaHost(rt, device=local_rank) as h:
            for k, val in tun.items():
                h.set_tunable(k, val)
            h.upload_scene(sc)
            best = 1e9
            for _ in range(3):
                h()
                best = min(best, h.stats()["kernel_ms"])
            out[name] = {"kernel_ms": best, "Mrays/s": tw * th / best / 1e3}
            if name == "auto":
                h.set_tunable(host.TUNE_COUNTERS, 1)
                h()
                st = h.stats()
                out[name].update({"packets_through_the_overflow_launch": st["packet_overflows"],
                                  "share_of_packets": st["packet_overflows"] / (tw * th / 128.0),
                                  "leaf_box_tests_per_ray": st["leafbox_tests"] / (tw * th), "triangle_tests_per_ray": st["tri_tests"] / (tw * th)})
                h.set_tunable(host.TUNE_COUNTERS, 0)
                h()
                img = h.download()
                rows = (135, th, 270)
                ys = list(range(*rows))
                ref = po.render(sc, tw, th, 1.0, True, rows=rows, want_ids=False, want_counters=True)
                out["parity"] = {"rows_vs_oracle": len(ys), "pixel_mismatches": int((img[ys].view(np.uint32) != ref.image[ys].view(np.uint32)).sum())}
                out["exhaustive_walk_V_T"] = [ref.counters["V"], ref.counters["T"]]
    return out


def extra_c5(local_rank, sc, peak):
    """BASELINE config 5 inside the default line (N = 1): 2^28 random rays (counter-hash generator, seed 1234) against the
    stand-in tree; value = best device time of 3 batches; parity = a 2^16 prefix against the oracle + checksums."""
    from opencl_raytracer_b200 import host
    from oracle import pyoracle as po
    rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
    total, n_chk = 1 << 28, 1 << 16
    lo, hi = sc.root_box()
    o, d = po.gen_random_rays(1234, 0, n_chk, lo, hi)
    ref = po.trace_rays(sc, o, d, 100000.0, want_counters=True)
    with host.CudaHost(rt, device=local_rank) as h:
        h.upload_scene(sc)
        _, _, fid, dist = h.trace_random_rays(1234, 0, n_chk, want_arrays=True)
        best, sums = 1e9, None
        h.trace_random_rays(1234, 0, 1 << 24)
        for _ in range(3):
            hits, idsum, _, _ = h.trace_random_rays(1234, 0, total)
            best = min(best, h.stats()["kernel_ms"])
            sums = (hits, idsum)
    V, T, hf = ref.counters["V"], ref.counters["T"], ref.counters["h"]
    B = 32.0 * V + 48.0 * T + 48.0 * hf + 4.0 + 8.0
    prof, prof_src = load_profile("c5")
    issue = None
    if prof:
        issue = {"ncu_issue_active_pct": prof.get("issue_active_pct"), "avg_threads_per_instruction": prof.get("avg_threads_per_instruction"),
                 "l1tex_throughput_pct": prof.get("l1tex_throughput_pct"), "warps_active_pct": prof.get("warps_active_pct"), "kernel": prof.get("kernel")}
    return {"workload": WORKLOADS["c5"][4], "rays": total, "kernel_ms": best, "Mrays/s": total / best / 1e3,
            "parity": {"rays_vs_oracle": n_chk, "id_mismatches": int((fid != ref.face_id).sum()),
                       "distance_mismatches": int((dist.view(np.uint32) != ref.distance.view(np.uint32)).sum()),
                       "hit_count": sums[0], "face_id_sum": sums[1]},
            "algorithmic_vs_hbm": {"bytes_per_ray": B, "V": V, "T": T, "h": hf, "GBps": B * total / (best * 1e-3) / 1e9,
                                   "frac": B * total / (best * 1e-3) / 1e9 / peak,
                                   "note": "tree is L2-resident; see profiles/ for the DRAM traffic (tens of MB per launch)"},
            "profiled": issue, "profile": prof_src, "sweep_over_gpus": "python bench.py --workload c5 --gpus N (profiles/r2_c5_scaling.json)"}


def bench_c5(args, rank, world, local_rank, desc):
    """Config C5: 2^28 arbitrary rays, generated on the device from the counter hash, sharded over the ranks by
    contiguous index ranges; the only exchange is the all-reduce of the (hit count, face-id sum) checksums."""
    import torch
    import torch.distributed as dist
    from opencl_raytracer_b200 import host
    import __graft_entry__ as g
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        g.build(quiet=True)
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    sc = build_scene("sibenik")
    total = 1 << 28
    per = total // world
    first = rank * per
    rt = host.RayTracer(host.Options(width=32, height=32, nSuperSamples=1))
    h = host.CudaHost(rt, device=local_rank)
    h.upload_scene(sc)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sums = torch.zeros(2, dtype=torch.int64, device=dev)
    for _ in range(args.warmup):
        h.trace_random_rays(1234, first, per)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, checks = [], None
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        hits, idsum, _, _ = h.trace_random_rays(1234, first, per)       # blocking; device time in stats (CUDA events on its stream)
        ms = torch.tensor([h.stats()["kernel_ms"]], dtype=torch.float64, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sums[0], sums[1] = hits, idsum
        e0.record()
        if world > 1:
            dist.all_reduce(sums)
        e1.record()
        torch.cuda.synchronize(dev)
        ms += e0.elapsed_time(e1)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        step_ms.append(float(ms.item()))
        checks = (int(sums[0].item()), int(sums[1].item()))
    clocks = sampler.result()
    parity = cpu = roofline = None
    if rank == 0:
        from oracle import pyoracle as po
        lo, hi = sc.root_box()
        n_chk = 1 << 16
        o, d = po.gen_random_rays(1234, 0, n_chk, lo, hi)
        ref = po.trace_rays(sc, o, d, 100000.0, want_counters=True)
        _, _, fid, dist_ = h.trace_random_rays(1234, 0, n_chk, want_arrays=True)
        parity = {"rays_checked": n_chk, "id_mismatches": int((fid != ref.face_id).sum()), "distance_mismatches": int((dist_ != ref.distance).sum()),
                  "hit_count": checks[0], "face_id_sum": checks[1]}
        V, T, hf = ref.counters["V"], ref.counters["T"], ref.counters["h"]
        B = 32.0 * V + 48.0 * T + 48.0 * hf + 4.0 + 8.0
        peak, peak_src = load_peaks()
        kms = float(np.mean(step_ms))
        achieved = B * per / (kms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": 16990208,
                    "peak_source": peak_src, "kernel": "k_trace_persistent (refill kernel)", "kernel_ms": kms, "algorithmic_bytes_per_ray": B,
                    "V": V, "T": T, "h": hf,
                    "note": "tree is L2-resident: DRAM traffic per launch is 17 MB (profiles/r1_c5_refill_ncu.txt); the kernel is bound by the L1 "
                            "data pipe (96 %: 32 different node pairs per warp load)"}
        if world == 1 and not args.no_cpu:
            import time as _t
            n_cpu = 1 << 22
            o, d = po.gen_random_rays(1234, 0, n_cpu, lo, hi)
            t0 = _t.perf_counter()
            po.trace_rays(sc, o, d, 100000.0)
            dt = _t.perf_counter() - t0
            cpu = {"value": n_cpu / dt / 1e6, "unit": "Mrays/s", "cores": int(po.port().orc_online_cpus()), "kind": "port",
                   "sample": "first 2^22 rays of the batch, %.2f s" % dt}
    e2e_host = None
    if rank == 0 and world == 1 and not args.no_e2e:
        # the same kernel fed from HOST arrays (rtx_trace_rays): 32 B/ray up, 8 B/ray down, chunks pipelined over three streams
        from oracle import pyoracle as po
        n_host = 1 << 26
        lo, hi = sc.root_box()
        o, d = po.gen_random_rays(1234, 0, n_host, lo, hi)            # generator only; not timed
        o_p, d_p = torch.from_numpy(o).pin_memory().numpy(), torch.from_numpy(d).pin_memory().numpy()
        del o, d
        f_p = torch.empty(n_host, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        t_p = torch.empty(n_host, dtype=torch.float32).pin_memory().numpy()
        h.trace_rays(o_p[:1 << 22], d_p[:1 << 22])
        import time as _t
        best = 1e9
        for _ in range(3):
            t0 = _t.perf_counter()
            fid_h, dist_h = h.trace_rays(o_p, d_p, out_face_id=f_p, out_distance=t_p)
            best = min(best, _t.perf_counter() - t0)
        _, _, fid_d, _ = h.trace_random_rays(1234, 0, 1 << 20, want_arrays=True)
        e2e_host = {"value": n_host / best / 1e6, "unit": "Mrays/s", "rays": n_host, "ms": best * 1e3, "h2d_bytes_per_step": n_host * 32,
                    "d2h_bytes_per_step": n_host * 8, "matches_device_generated_prefix": bool(np.array_equal(fid_h[:1 << 20], fid_d)),
                    "calls": "rtx_trace_rays(host origins, host directions) -> host face ids + distances; page-locked buffers, chunks of 4 Mi rays pipelined over three streams"}
    h.close()
    if rank == 0:
        t = float(np.sum(step_ms))
        print(json.dumps({
            "metric": "Mrays/s closest-hit (random rays vs sibenik BVH)", "value": total * args.steps / (t * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": sc.name, "triangles": sc.num_triangles, "rays_per_step": total,
                       "parallelism": "contiguous ray-index ranges over %d GPU(s), tree replicated, 1 all-reduce of the checksums" % world,
                       "l2": "flushed between timed iterations (256 MiB memset, untimed)"},
            "clocks": clocks, "e2e": {"value": total * args.steps / (t * 1e-3) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 16,
                                      "calls": "rtx_trace_random_rays: rays generated on the device, only the checksums come back"},
            "e2e_host_rays": e2e_host,
            "gpu_launches": args.steps * world, "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity, "step_ms": step_ms}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather", default="auto", choices=["auto", "u8", "float", "p2p_u8", "p2p_float"],
                    help="what a step ends with on rank 0 and how it gets there (N > 1): u8 = the byte image after RayTracer::resize on "
                         "the device, every rank resizes its own tiles and ONE NCCL gather moves bytes (default); float = ONE NCCL gather "
                         "of the float tiles; p2p_* = no collective, every rank's kernel stores its share into rank 0's image over NVLink "
                         "peer memory (multigpu.TiledRenderer); auto (default) = p2p_u8 when the GPUs can map each other's memory, else u8")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    width, height, nss, scene_kind, desc = WORKLOADS[args.workload]
    if args.workload == "c5" and args.impl != "reference":
        return bench_c5(args, rank, world, local_rank, desc)

    if args.impl == "reference":
        # Nothing of this repo's product on this path: no g.build(), no librtx_*.so.  The scene comes from the
        # reference's own mesh.cc + bvh.cc and the timed loop is the reference's kernel text, both in
        # oracle/_ref/libref_oracle.so (src/render.cc:76-111 is the sequence being reproduced).
        if rank != 0:
            return 0
        sc = build_scene_reference(scene_kind)
        n = int(np.sqrt(nss))
        tw, th = width * n, height * n
        arm = CpuArm(sc, tw, th)
        times = []
        for i in range(args.warmup + args.steps):
            dt = arm.run()
            if i >= args.warmup:
                times.append(dt)
        v = arm.nrays * len(times) / sum(times) / 1e6
        info = arm.info(v, float(np.mean(times)), len(times))
        print(json.dumps({
            "impl": "reference", "metric": METRIC.get(args.workload, METRIC["c3"]), "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": make_config(desc, sc.name, sc.num_triangles, tw, th, max(1, args.gpus), args.gather),
            "rays_timed_per_step": arm.nrays,
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    import torch
    import torch.distributed as dist
    from opencl_raytracer_b200 import host, multigpu
    import __graft_entry__ as g

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        g.build(quiet=True)
    if world > 1:
        dist.barrier()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    sc = build_scene(scene_kind)
    rt = host.RayTracer(host.Options(width=width, height=height, nSuperSamples=nss))
    tw, th = rt.totalWidth, rt.totalHeight
    rays = tw * th

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- value: scene resident, device-timed ----------------
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)                       # kernels, NCCL gather and the events all use this stream
    requested_gather, r = args.gather, None
    if args.gather == "auto":
        args.gather = "u8"
        if world > 1:
            try:
                r = multigpu.TiledRenderer(rt, sc, rank, world, local_rank, gather="p2p_u8")
                args.gather = "p2p_u8"
            except RuntimeError:                          # raised on every rank together (multigpu.TiledRenderer)
                r = None
    if r is None:
        r = multigpu.TiledRenderer(rt, sc, rank, world, local_rank, gather=args.gather)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    for _ in range(args.warmup):
        r.render_frame()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = r.kernel_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    rendezvous = torch.zeros(1, dtype=torch.int32, device=dev)
    for k in range(args.steps):
        flush.zero_()                                   # L2 flush between timed iterations (not timed)
        if world > 1:
            # per-step rendezvous ON THE DEVICE: a one-element all-reduce the launching stream waits for.  All ranks
            # leave it within microseconds of each other, and the host runs ahead enqueueing the frame, so the step's
            # events do not measure the wake-up jitter of N host processes after a host-side barrier.
            dist.all_reduce(rendezvous)
        ev[k][0].record()
        r.render_frame()
        ev[k][1].record()
        torch.cuda.synchronize(dev)
        kernel_ms.append(r.host.stats()["kernel_ms"])
    barrier()
    clocks = sampler.result()
    step_ms = torch.tensor([a.elapsed_time(b) for a, b in ev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(step_ms, op=dist.ReduceOp.MAX)   # max over ranks, per step
    step_ms = step_ms.cpu().numpy()
    total_ms = float(step_ms.sum())
    launches = r.kernel_launches - launches0
    lt = torch.tensor([launches], dtype=torch.int64, device=dev)
    km = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(lt)
        dist.all_reduce(km, op=dist.ReduceOp.MAX)
    value = rays * args.steps / (total_ms * 1e-3) / 1e6
    kernel_ms_mean = float(km.item())

    # per-phase device times (CUDA events around every launch group; separate frames, not the timed ones)
    phase_names = list(host.PHASES) + ["resize", "gather", "deinterleave"]
    r.host.set_tunable(host.TUNE_PHASE_TIMING, 1)
    r.timing = True
    acc = np.zeros(len(phase_names))
    nph = 5
    for _ in range(nph):
        flush.zero_()
        if world > 1:
            dist.all_reduce(rendezvous)
        r.render_frame()
        torch.cuda.synchronize(dev)
        ph = r.phase_ms()
        acc += np.array([ph.get(k, 0.0) for k in phase_names])
    r.host.set_tunable(host.TUNE_PHASE_TIMING, 0)
    r.timing = False
    pt = torch.tensor(acc / nph, dtype=torch.float64, device=dev)
    pt_mean = pt.clone()
    if world > 1:
        dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        dist.all_reduce(pt_mean)
        pt_mean /= world
    phases_ms = {"max_over_ranks": {k: float(v) for k, v in zip(phase_names, pt.cpu().numpy())},
                 "mean_over_ranks": {k: float(v) for k, v in zip(phase_names, pt_mean.cpu().numpy())},
                 "how": "CUDA events around each launch group of %d extra frames (rtx_phase_ms + torch events in multigpu.TiledRenderer); "
                        "'gather' = the NCCL gather (p2p modes: the one-element all-reduce, i.e. rank 0 waiting for the slowest rank), 'resize' includes the peer stores in p2p_u8" % nph}

    # parity spot check inside the bench (rank 0): sampled rows of the gathered frame == oracle
    parity = None
    if rank == 0:
        from oracle import pyoracle as po
        n = rt.n
        ystep = max(1, height // 24)
        ys = list(range(ystep // 3, height, ystep))             # rows of the final image
        ref = np.zeros((th, tw), np.float32)
        for k in range(n):                                       # their n super-sampled rows each
            ref += po.render(sc, tw, th, 1.0, True, rows=(ys[0] * n + k, th, ystep * n), want_ids=False).image
        if args.gather.endswith("u8"):
            got = r.download_u8()
            want = po.resize(ref, width, height, n)
            parity = {"rows_checked": len(ys), "image": "u8 %dx%d" % (width, height),
                      "pixels_differing": int((got[ys] != want[ys]).sum())}
        else:
            got = r.download()
            sel = np.concatenate([np.arange(y * n, y * n + n) for y in ys])
            parity = {"rows_checked": len(sel), "image": "float %dx%d" % (tw, th),
                      "pixels_differing": int((got[sel] != ref[sel]).sum())}
        del got, ref

    # N > 1: the same frame with the other ways of getting it to rank 0, for comparison
    other = None
    if world > 1:
        other = []
        for alt, sync in (("u8", "allreduce"), ("float", "allreduce"), ("p2p_u8", "allreduce"), ("p2p_float", "allreduce"), ("p2p_u8", "flags")):
            if alt == args.gather and sync == "allreduce":
                continue
            try:
                r2 = multigpu.TiledRenderer(rt, sc, rank, world, local_rank, gather=alt, sync=sync)
            except RuntimeError as e:                    # peer memory not available between these GPUs (all ranks agree)
                other.append({"gather": alt, "unavailable": str(e)})
                continue
            for _ in range(3):
                r2.render_frame()
            barrier()
            ts = []
            for _ in range(6):
                flush.zero_()
                dist.all_reduce(rendezvous)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r2.render_frame()
                e1.record()
                torch.cuda.synchronize(dev)
                ts.append(e0.elapsed_time(e1))
            tms = torch.tensor(ts, dtype=torch.float64, device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.mean().item())
            other.append({"gather": alt + (" (ranks ordered by frame counters in peer memory instead of the all-reduce)" if sync == "flags" else ""),
                          "ms_per_step": ms, "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s"})
            barrier()
            r2.close()

    # ---------------- e2e: the reference's five calls on host buffers ----------------
    e2e = e2e_u8 = None
    if not args.no_e2e:
        arrs = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (sc.faces, sc.nodes, sc.aabbs, sc.vertices, sc.normals)]
        np_arrs = [a.numpy() for a in arrs]
        h2d = int(sum(a.numel() * a.element_size() for a in arrs))
        out_f = torch.empty((th, tw), dtype=torch.float32).pin_memory().numpy() if rank == 0 else None
        out_b = torch.empty((height, width), dtype=torch.uint8).pin_memory().numpy() if rank == 0 else None
        # N > 1: download(float*) needs the float tiles gathered, download_u8 the byte tiles -- one renderer per payload
        r_float = r if (world == 1 or args.gather == "float") else multigpu.TiledRenderer(rt, sc, rank, world, local_rank, gather="float")
        r_u8 = r if (world == 1 or args.gather.endswith("u8")) else multigpu.TiledRenderer(rt, sc, rank, world, local_rank, gather="u8")
        hr = r.host

        phase = {}

        def timed(name, fn):
            t0 = time.perf_counter()
            out = fn()
            phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return out

        def render_sync(rr):
            rr.render_frame()
            torch.cuda.synchronize(dev)

        def frame_float():
            timed("upload", lambda: r_float.host.upload(*np_arrs))
            timed("render", lambda: render_sync(r_float))
            if rank == 0:
                timed("download", lambda: r_float.host.download(out_f))

        def frame_pipelined():
            # N = 1: upload, then trace + download as ONE call (rtx_render_download): bands of tile rows, the
            # device->host copy of each finished band overlaps the tracing of the next
            timed("upload", lambda: hr.upload(*np_arrs))
            timed("render+download", lambda: hr.render_download(out_f))

        def frame_u8():
            timed("upload", lambda: r_u8.host.upload(*np_arrs))
            timed("render", lambda: render_sync(r_u8))
            if rank == 0:
                timed("download", lambda: r_u8.host.download_u8(out_b))

        shared = None
        if world > 1:
            try:
                shared = multigpu.SharedHostImage(rt, rank, world)
            except RuntimeError as e:                    # e.g. a small /dev/shm: keep the gather path as the headline
                shared_note = str(e)

        def frame_shared():
            # N > 1: the caller's float image is one page-locked host buffer every rank maps; each rank uploads, traces its tiles
            # and stores them straight into it over its OWN PCIe link (rtx_store_tiles_async) -- no gather, no rank-0 funnel
            st = torch.cuda.current_stream(dev).cuda_stream
            timed("upload", lambda: r_float.host.upload(*np_arrs))

            timed("render+store", lambda: r_float.host.render_store(shared.device_ptr))

        res, phases = [], []
        fns = (frame_float, frame_u8) + ((frame_pipelined,) if world == 1 else ()) + ((frame_shared,) if shared is not None else ())
        for fn in fns:
            for _ in range(2):
                fn()
            barrier()
            phase.clear()
            t0 = time.perf_counter()
            nst = max(3, args.steps // 2)
            for _ in range(nst):
                fn()
                if world > 1:
                    dist.barrier()
            torch.cuda.synchronize(dev)
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            res.append(rays * nst / float(dt.item()) / 1e6)
            phases.append({k: v / nst for k, v in phase.items()})
        e2e = {"value": res[0], "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": rays * 4,
               "calls": "rtx_upload + rtx_render(+gather) + rtx_download(float image), pinned host buffers, wall clock",
               "phase_ms": phases[0]}
        if world > 1 and shared is not None:
            via_gather = e2e
            same = None
            if rank == 0:                                # the two host images must be the same image
                same = bool(np.array_equal(np.asarray(shared.array).view(np.uint32), out_f.view(np.uint32)))
            e2e = {"value": res[2], "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": rays * 4,
                   "calls": "per rank: rtx_upload + rtx_render_store into ONE page-locked host image mapped by every rank process "
                            "(multigpu.SharedHostImage, rtx_host_register): the warp that finishes a tile sends it, so each rank's tiles "
                            "leave over its own PCIe link while the rest is still being traced; wall clock incl. a barrier per frame",
                   "phase_ms": phases[2], "identical_to_gathered_image": same,
                   "via_rank0_gather": {"value": via_gather["value"], "calls": via_gather["calls"], "phase_ms": via_gather["phase_ms"]}}
            shared.close()
        elif world > 1:
            e2e["note"] = "shared host image unavailable (%s): rank 0 downloads the gathered frame" % shared_note
        if world == 1:
            # the headline end-to-end figure: same host buffers, same bytes over PCIe, the library's fused call
            e2e_three_calls = e2e
            e2e = {"value": res[2], "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": rays * 4,
                   "calls": "rtx_upload + rtx_render_download(float image): tracing and device->host copy pipelined over bands of "
                            "tile rows on two streams; pinned host buffers, wall clock",
                   "phase_ms": phases[2],
                   "three_separate_calls": {"value": e2e_three_calls["value"], "calls": e2e_three_calls["calls"],
                                            "phase_ms": e2e_three_calls["phase_ms"]}}
        e2e_u8 = {"value": res[1], "unit": "Mrays/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": width * height,
                  "calls": "rtx_upload + rtx_render(+gather) + rtx_download_u8 (device resize, ray_tracer.cc:3-15 order)",
                  "phase_ms": phases[1]}
        for rr in (r_float, r_u8):
            if rr is not r:
                rr.close()

    # ---------------- extras (N = 1): same frame, other code paths, kernel time only ----------------
    extras = None
    if world == 1:
        extras = {}

        def kernel_ms(h, n=3):
            best = 1e9
            for _ in range(n):
                h()
                best = min(best, h.stats()["kernel_ms"])
            return best

        for name, tun, jit in (("reference_algorithm_on_gpu", {host.TUNE_KERNEL: host.KERNEL_EXHAUSTIVE}, 0),
                               ("per_ray_traversal_no_frustum", {host.TUNE_FRUSTUM: 0}, 0),
                               ("jittered_16spp_seed_0x5EED", {}, 0x5EED)):
            with host.CudaHost(rt, device=local_rank, jitter_seed=jit) as hx:
                for k, v in tun.items():
                    hx.set_tunable(k, v)
                hx.upload_scene(sc)
                ms = kernel_ms(hx)
                extras[name] = {"kernel_ms": ms, "Mrays/s": rays / ms / 1e3}
        # the same frame on the reference's OTHER tree (`-r sah`, bvh.cc:178-236; built by the host builder of this repo, which
        # emits the reference builder's arrays): kernel time + sampled rows against the oracle walking that tree
        try:
            from oracle import pyoracle as po
            from opencl_raytracer_b200 import scene as scene_mod
            t0 = time.perf_counter()
            sc_sah = scene_mod.scene_from_mesh(sc.vertices[:, :3], sc.orig_faces.reshape(-1, 3), sah=True)
            build_s = time.perf_counter() - t0
            with host.CudaHost(rt, device=local_rank) as hx:
                hx.upload_scene(sc_sah)
                ms = kernel_ms(hx)
                got = hx.download()
            rows = list(range(th // 16, th, th // 8))
            bad = 0
            for y in rows:
                want = po.render(sc_sah, tw, th, 1.0, True, rows=(y, y + 1, 1), want_ids=False).image
                bad += int((got[y] != want[y]).sum())
            extras["sah_tree"] = {"kernel_ms": ms, "Mrays/s": rays / ms / 1e3, "host_build_s": build_s,
                                  "parity": {"rows_vs_oracle": len(rows), "pixels_differing": bad},
                                  "what": "the reference's `-r sah` tree of the same mesh (scene_prep.cc::split_sah: the reference builder's arrays, "
                                          "O(n log n) per node instead of O(n^2))"}
            del got
        except Exception as e:                                              # an extra must not cost the headline line
            extras["sah_tree"] = {"unavailable": str(e)}
        with host.CudaHost(rt, device=local_rank) as hx:                  # the tree itself built on the device
            hx.upload_mesh(sc.vertices, sc.orig_faces, sc.normals)
            t0 = time.perf_counter()
            hx.upload_mesh(sc.vertices, sc.orig_faces, sc.normals)
            wall = (time.perf_counter() - t0) * 1e3
            bms, levels = hx.build_stats()
            nodes_d, aabbs_d, tri_d, faces_d = hx.download_tree()
            extras["device_bvh_build"] = {"build_ms": bms, "levels": levels, "rtx_upload_mesh_wall_ms": wall,
                                          "identical_to_host_builder": bool(np.array_equal(nodes_d, sc.nodes) and np.array_equal(tri_d, sc.triangles)
                                                                            and np.array_equal(faces_d, sc.faces)
                                                                            and np.array_equal(aabbs_d.view(np.uint32), np.ascontiguousarray(sc.aabbs, np.float32).reshape(-1, 4).view(np.uint32))),
                                          "what": "rtx_upload_mesh: the reference's longest-axis builder (bvh.cc:59-162) level by level on the device, "
                                                  "then the same validation + flatten as rtx_upload"}
        rt_ao = host.RayTracer(host.Options(width=1920, height=1080, nSuperSamples=4, enableAO=True, aoNumSamples=3, aoMethod=0))
        with host.CudaHost(rt_ao, device=local_rank) as hx:               # the CLI's default options on the C2 frame
            hx.upload_scene(sc)
            ms = kernel_ms(hx)
            hit = float((hx.download() > 0).mean())
            extras["ambient_occlusion_c2_frame"] = {"kernel_ms": ms, "primary_Mrays/s": rt_ao.totalWidth * rt_ao.totalHeight / ms / 1e3,
                                                    "what": "3840x2160 primary rays + uniform AO with 3 rings (28 occlusion rays per hit pixel, "
                                                            "intersect_kernel.cl:214-277); %.1f %% of the pixels are lit" % (100 * hit)}
        peak_hbm = load_peaks()[0]
        extras["c5"] = extra_c5(local_rank, sc, peak_hbm) if args.workload == "c3" else None
        extras["c4"] = extra_c4(local_rank, peak_hbm) if args.workload == "c3" else None
        extras["irregular_interior"] = extra_irregular(local_rank) if args.workload == "c3" else None
        extras["reference_algorithm_on_gpu"]["what"] = ("k_render_exhaustive: the reference kernel's own algorithm (one thread per pixel, "
                                                        "stackless pre-order walk, no culling) compiled for sm_100a")

    # ---------------- roofline + cpu baseline (rank 0, N = 1 only for the CPU leg) ----------------
    roofline = cpu = None
    if rank == 0:
        V, T, hfrac = oracle_counters(sc, tw, th)
        B = 32.0 * V + 48.0 * T + 48.0 * hfrac + 4.0
        peak, peak_src = load_peaks()
        kernel_s = kernel_ms_mean * 1e-3
        k_rays = rays / world
        alg_gbps = B * k_rays / kernel_s / 1e9
        l2_peak = r.host.probe_bandwidth(0, 32 << 20, 20)
        # profiler-only quantities (instruction count, bytes per memory level) come from a capture of THIS build at
        # N = 1, or not at all: a stale or other-N capture is never scaled into the line
        # ... at N > 1 a capture of ONE RANK'S share of the frame (rank 0 of N rendered on one GPU, tools/rank_timing.py
        # under ncu: the very kernel launch a rank of the N-GPU run executes), key "<workload>@N"
        prof, prof_src = load_profile(args.workload if world == 1 else "%s@%d" % (args.workload, world))
        sms = host.device_info(local_rank).sm_count
        ph = phases_ms["mean_over_ranks"]
        render_ph = sum(ph[k] for k in ("tables", "collect_super", "collect", "traversal", "overflow"))
        share = ph["traversal"] / render_ph if render_ph > 0 else 1.0          # dominant kernel's share of the event pair, measured in this run
        issue = hbm_actual = levels = None
        traffic = None
        if prof:
            dom_s = kernel_s * share
            traffic = (prof.get("dram_bytes_read") or 0) + (prof.get("dram_bytes_write") or 0)
            hbm_actual = {"bytes_per_launch": traffic, "GBps": traffic / dom_s / 1e9, "frac_of_peak": traffic / dom_s / 1e9 / peak,
                          "compulsory_bytes": 4 * rays // world, "note": "per rank; compulsory HBM traffic of this workload is the 4 B/ray image write"}
            if prof.get("warp_instructions") and clocks.get("sm_mhz"):
                ach = prof["warp_instructions"] / dom_s / 1e9
                pk = sms * 4 * clocks["sm_mhz"] * 1e6 / 1e9
                issue = {"achieved": ach, "peak": pk, "frac": ach / pk,
                         "warp_instructions_per_launch": prof["warp_instructions"],
                         "thread_instructions_per_ray": (prof["warp_instructions"] * (prof.get("avg_threads_per_instruction") or 32.0)) / (rays / world),
                         "avg_threads_per_instruction": prof.get("avg_threads_per_instruction"),
                         "ncu_issue_active_pct": prof.get("issue_active_pct"), "ncu_warps_active_pct": prof.get("warps_active_pct"),
                         "how": "warp instructions of the dominant kernel%s (ncu capture of this build) / (kernel_ms of this run x its %.1f %% "
                                "share of the event pair, from this run's phases_ms); peak = %d SMs x 4 schedulers x %.0f MHz (median SM clock "
                                "sampled during the timed region)" % (" for one rank's share of the frame" if world > 1 else "", 100 * share, sms, clocks["sm_mhz"])}
            if prof.get("l1_load_bytes") and prof.get("l2_read_bytes_from_l1"):
                levels = {"l1_load_GBps": prof["l1_load_bytes"] / dom_s / 1e9, "l1_hit_pct": prof.get("l1_hit_pct"),
                          "l2_read_GBps": prof["l2_read_bytes_from_l1"] / dom_s / 1e9, "l2_hit_pct": prof.get("l2_hit_pct"),
                          "l2_peak_GBps_probed": l2_peak, "l2_frac_of_probed_peak": prof["l2_read_bytes_from_l1"] / dom_s / 1e9 / l2_peak,
                          "hbm_GBps": traffic / dom_s / 1e9, "hbm_frac_of_peak": traffic / dom_s / 1e9 / peak}
        # top level = the level that binds.  The scene (11 MB) lives in L1/L2, the only compulsory HBM traffic is the
        # image write (2 % of peak), so the bound is warp-instruction issue; SURVEY 8d's algorithmic-bytes figure is kept
        # beside it (algorithmic_vs_hbm) with the reason it exceeds 1.
        roofline = {
            "bound": "issue", "unit": "Gwarp-instr/s",
            "achieved": issue["achieved"] if issue else None, "peak": issue["peak"] if issue else None,
            "frac": issue["frac"] if issue else None,
            "traffic": traffic,
            "kernel": (prof or {}).get("kernel", "k_render_packet<256,4,0,...,2,1,1> (candidate-list kernel)"),
            "kernel_ms": kernel_ms_mean, "dominant_kernel_share_of_kernel_ms": share if prof else None,
            "profile": prof_src, "peak_source": "148 SMs x 4 warp schedulers x 1 instruction / clock at the sampled SM clock",
            "issue": issue, "hbm_actual": hbm_actual, "memory_levels": levels,
            "algorithmic_vs_hbm": {
                "bound": "hbm", "achieved": alg_gbps, "peak": peak, "unit": "GB/s", "frac": alg_gbps / peak, "peak_source": peak_src,
                "algorithmic_bytes_per_ray": B, "V": V, "T": T, "h": hfrac,
                "note": "SURVEY 8d contract figure: B = 32 V + 48 T + 48 h + 4 bytes per ray under the reference's EXHAUSTIVE walk "
                        "(V box tests, T triangle tests per ray, counted by the oracle on a row sample of this frame) x rays / kernel time. "
                        "It exceeds 1 because the kernel does not do that work: the frustum front end leaves ~7 leaf-box and ~4 triangle "
                        "tests per ray and the scene is cache-resident; results are bit-identical (parity_check). Not a roofline fraction."},
            "l2_probe": {"peak_GBps": l2_peak, "source": "rtx_probe_bandwidth: ld.cg float4 sweep of a 32 MiB buffer, this run"}}
        if world == 1 and not args.no_cpu:
            arm = CpuArm(sc, tw, th)
            passes, total_s = arm.run_for(10.0)
            cpu = arm.info(arm.nrays * passes / total_s / 1e6, total_s / passes, passes)
    r.close()

    if rank == 0:
        print(json.dumps({
            "metric": METRIC.get(args.workload, METRIC["c3"]), "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(desc, sc.name, sc.num_triangles, tw, th, world, requested_gather),
            "gather": args.gather,
            "clocks": clocks, "e2e": e2e, "e2e_u8": e2e_u8, "gpu_launches": int(lt.item()),
            "roofline": roofline, "cpu_baseline": cpu, "parity_check": parity, "phases_ms": phases_ms, "other_gather": other, "extras": extras,
            "step_ms": [float(x) for x in step_ms],
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
