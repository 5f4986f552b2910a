"""`python -m opencl_raytracer_b200.render [options] INPUT_MESH OUTPUT_IMAGE` -- the reference's `render` command line
(src/render.cc:16-47: same options, same defaults, same PGM) for hosts without the reference tree: OFF loader, vertex
normals and BVH from include/rtx_scene.h (or the BVH on the device with --device-build), tracing through the C ABI.

  -w/--width, -h/--height, -a/--ambient-occlusion-samples, -d/--ambient-occlusion-max-distance,
  -m/--ambient-occlusion-method [uniform|random], -f/--focal-length, -s/--supersamples, -r/--bvh-strategy [longest]
"""
from __future__ import annotations

import argparse
import sys
import time

from . import host, scene


def parse(argv):
    ap = argparse.ArgumentParser(prog="render", add_help=False,
                                 description="A B200 raytracer that renders triangle meshes in OFF format (reference CLI mirror).")
    ap.add_argument("--help", action="help")
    ap.add_argument("input_mesh")
    ap.add_argument("output_image")
    ap.add_argument("-w", "--width", type=int, default=600)
    ap.add_argument("-h", "--height", type=int, default=600)                      # render.cc:24 re-binds -h to the height
    ap.add_argument("-a", "--ambient-occlusion-samples", type=int, default=3)
    ap.add_argument("-d", "--ambient-occlusion-max-distance", type=float, default=0.2)
    ap.add_argument("-m", "--ambient-occlusion-method", choices=["uniform", "random"], default="uniform")
    ap.add_argument("-f", "--focal-length", type=float, default=1.0)
    ap.add_argument("-s", "--supersamples", type=int, default=4)
    ap.add_argument("-r", "--bvh-strategy", choices=["longest", "sah"], default="longest",
                    help="render.cc:30: cut along the longest axis (default) or the surface-area heuristic of bvh.cc:178-236 "
                         "(same tree as the reference's O(n^2) builder, built in O(n log^2 n))")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--device-build", action="store_true", help="build normals and BVH on the device (rtx_upload_mesh)")
    ap.add_argument("--scene-cache", metavar="DIR", default=None,
                    help="keep the prepared scene (BVH, sorted faces, normals) in DIR, keyed by the mesh file's sha256")
    return ap.parse_args(argv)


def main(argv=None) -> int:
    a = parse(sys.argv[1:] if argv is None else argv)
    opt = host.Options(width=a.width, height=a.height, focalLength=a.focal_length, nSuperSamples=a.supersamples, enableShading=True,
                       enableAO=a.ambient_occlusion_samples != 0, aoMaxDistance=a.ambient_occlusion_max_distance,
                       aoNumSamples=a.ambient_occlusion_samples, aoMethod=0 if a.ambient_occlusion_method == "uniform" else 1,
                       aoAlphaMin=4, aoAlphaMax=90)
    t0 = time.perf_counter()
    sah = a.bvh_strategy == "sah"
    sc = scene.cached_scene_from_off(a.input_mesh, a.scene_cache, sah=sah) if a.scene_cache else scene.scene_from_off(a.input_mesh, sah=sah)
    t1 = time.perf_counter()
    rt = host.RayTracer(opt)
    with host.CudaHost(rt, device=a.device) as h:
        if a.device_build:
            h.upload_mesh(sc.vertices, sc.orig_faces, None)
        else:
            h.upload_scene(sc)
        t2 = time.perf_counter()
        h()
        ms = h.stats()["kernel_ms"]
        u8 = h.download_u8()                                                    # RayTracer::resize on the device, ray_tracer.cc:3-15
    host.write_pgm(a.output_image, u8)                                          # render.cc:130-137
    print("Vertices: %d  Triangles: %d  load+prepare %.1f ms, upload %.1f ms, rendering %.3f ms, total %.1f ms" % (
        sc.vertices.shape[0], sc.num_triangles, (t1 - t0) * 1e3, (t2 - t1) * 1e3, ms, (time.perf_counter() - t0) * 1e3))
    return 0


if __name__ == "__main__":
    sys.exit(main())
