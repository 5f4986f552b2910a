"""Stage a reference mesh for the GPU box.  TEST INFRASTRUCTURE ONLY.

usage: stage_ref_mesh.py <mesh.off> <out.bin>

Loads the OFF file with the REFERENCE's own load_off_mesh (through
oracle/_ref/libref_oracle.so) and writes vertices + faces as raw binary into
oracle/_ref/ (git-ignored; travels with the gpurun snapshot), because
/root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle  # noqa: E402


def main(off_path: str, out_path: str) -> None:
    sc = pyoracle.ref_scene_from_off(off_path)
    verts = sc.vertices[:, :3]
    faces = sc.orig_faces.reshape(-1, 3)
    pyoracle.write_mesh_bin(out_path, verts, faces)
    print("staged %s: %d vertices, %d triangles" % (out_path, verts.shape[0], faces.shape[0]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
