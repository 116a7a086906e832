#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the REFERENCE's own code (oracle/_ref: the reference's
CLOBJloader, CLBVHScene and kernel_bvh.cl compiled verbatim, see oracle/build_ref.py).
Run in the build container (needs /root/reference); the outputs are committed so that the GPU
box, which has no /root/reference, can still check against reference-produced vectors.

  cornell_scene.npz   post-build CLTriangle / CLLinearBVHNode / CLMaterial arrays (meaningful bytes)
  cornell_hits.npz    Intersect() of the 128x128 frame-1 camera rays + 4096 seeded shell rays
  cornell_frames.npz  KernelEntry() 96x64: frame 1 with 1 bounce; frames 1..3 accumulated, 4 bounces,
                      light types 0/1/2
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
import scenes  # noqa: E402


def main():
    tris, nodes, mats = ol.ref_load_scene(scenes.CORNELL, 4)
    # padding bytes are uninitialised in the reference: keep meaningful fields only
    tf = tris.view(np.float32).reshape(-1, 64)
    np.savez_compressed(os.path.join(HERE, "cornell_scene.npz"),
                        tri_floats=tf[:, [0, 1, 2, 4, 5, 8, 9, 10, 20, 21, 22, 24, 25, 28, 29, 30, 40, 41, 42, 44, 45, 48, 49, 50]],
                        tri_mtl=tris.view(np.uint32).reshape(-1, 64)[:, 60],
                        node_bounds=nodes.view(np.float32).reshape(-1, 12)[:, [0, 1, 2, 4, 5, 6]],
                        node_offset=nodes.view(np.uint32).reshape(-1, 12)[:, 8],
                        node_nprims=nodes.view(np.uint16).reshape(-1, 24)[:, 18],
                        node_axis=np.where(nodes.view(np.uint16).reshape(-1, 24)[:, 18] == 0, nodes[:, 38], 255),
                        mat_floats=mats.view(np.float32).reshape(-1, 16)[:, [0, 1, 2, 4, 5, 6, 8, 9, 10, 13, 14]])
    cam = ol.oracle_camera_rays(128, 128, 1)
    shell = scenes.shell_rays(4096, 12.0, seed=5, centre=(0.0, 7.0, 8.0))
    rays = np.concatenate([cam, shell])
    r = ol.ref_closest(tris, nodes, rays)
    np.savez_compressed(os.path.join(HERE, "cornell_hits.npz"), rays=rays.view(np.float32).reshape(-1, 8),
                        hit=r["hit"].astype(np.uint8), tri=r["tri"], t=r["t"], pos=r["pos"], normal=r["normal"])
    W, H = 96, 64
    frames = {}
    img = np.zeros((W * H, 4), dtype=np.float32)
    ol.ref_render(tris, nodes, mats, img, W, H, 1, 1)
    frames["b1_f1"] = img[:, :3].copy()
    for lt in (0, 1, 2):
        img = np.zeros((W * H, 4), dtype=np.float32)
        for fc in (1, 2, 3):
            ol.ref_render(tris, nodes, mats, img, W, H, fc, 4, light_type=lt)
        frames["b4_f123_lt%d" % lt] = img[:, :3].copy()
    img = np.zeros((W * H, 4), dtype=np.float32)
    ol.ref_render(tris, nodes, mats, img, W, H, 0, 9)      # frameCount == 0 branch (kernel_bvh.cl:449-451)
    frames["b9_f0"] = img[:, :3].copy()
    np.savez_compressed(os.path.join(HERE, "cornell_frames.npz"), width=W, height=H, **frames)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
