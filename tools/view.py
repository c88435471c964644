#!/usr/bin/env python
"""Write the camera views (and, with --top-view, the top views) of a few envs as PNG files (the non-interactive part of
play!: the reference blits the transposed image, utils.jl:64-73).
    python tools/view.py --out gpurun_out/views --envs 4 --steps 40 [--format rgb8|xrgb32|gray8]
"""
import argparse
import os
import struct
import sys
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raycastworlds_jl_b200 as rcw  # noqa: E402


def write_png(path, img):
    """img: uint8 [H, W, 3] or [H, W]"""
    h, w = img.shape[:2]
    color = 2 if img.ndim == 3 else 0
    raw = b"".join(b"\x00" + np.ascontiguousarray(img[r]).tobytes() for r in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/views")
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--format", default="rgb8", choices=["rgb8", "xrgb32", "gray8"])
    ap.add_argument("--seed", type=int, default=7)
    ap.add_argument("--top-view", action="store_true", help="also dump update_top_view! (single_room.jl:446-483)")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    env = rcw.BatchedSingleRoom(args.envs, seed=args.seed, obs_format=args.format)
    env.step_random(args.steps)
    obs = env.copy_obs()                      # [E, columns, rows(, 3)]
    st = env.get_state()
    for e in range(args.envs):
        img = obs[e]
        if args.format == "xrgb32":
            img = np.stack([(img >> 16) & 255, (img >> 8) & 255, img & 255], -1).astype(np.uint8)
        img = np.swapaxes(img, 0, 1)          # rows first: the transpose the reference does when blitting
        path = os.path.join(args.out, f"env{e}_{args.format}.png")
        write_png(path, img)
        print(path, "pos", st["pos"][e], "dir", st["dir_au"][e], "goal", st["goal"][e])
    if args.top_view:
        env.render_top_view()
        top = env.copy_top_view()             # uint32 [E, W*pu columns, H*pu rows]
        for e in range(args.envs):
            t = np.swapaxes(top[e], 0, 1)
            img = np.stack([(t >> 16) & 255, (t >> 8) & 255, t & 255], -1).astype(np.uint8)
            path = os.path.join(args.out, f"env{e}_top_view.png")
            write_png(path, img)
            print(path)
    env.close()


if __name__ == "__main__":
    main()
