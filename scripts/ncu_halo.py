#!/usr/bin/env python
"""Tiny driver for an ncu capture of the K-halo kernel: per-step launches (a step only waits for EARLIER steps of
its neighbours, so kernel-serialising profilers cannot deadlock it).  One process; --world 2 --devices 0 1 drives
two GPUs so the capture shows the kernel while its halo crosses NVLink."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hpc_patterns_b200.models.halo import VirtualRing  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--devices", type=int, nargs="*", default=None)
ap.add_argument("--mode", default="pull")
ap.add_argument("--rows", type=int, default=8)
ap.add_argument("--bytes", type=int, default=32 << 20)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
ring = VirtualRing(args.world, args.bytes, args.rows, args.mode, devices=args.devices or [0] * args.world)
ring.step(args.steps)
ring.synchronize()
print("wrong words:", sum(hs.verify_from_init() for hs in ring.ranks))
ring.close()
