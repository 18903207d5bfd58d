#!/usr/bin/env python
"""One process, two GPUs: plain K-p2p put / get launches over NVLink (no flags, nothing to wait for) for ncu."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hpc_patterns_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--bytes", type=int, default=188743680)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
C = hpc_patterns_b200.native()
C.enable_peer_access([0, 1])
torch.cuda.set_device(0)
local = torch.empty(args.bytes, dtype=torch.uint8, device="cuda:0")
local2 = torch.empty(args.bytes, dtype=torch.uint8, device="cuda:0")
remote = torch.empty(args.bytes, dtype=torch.uint8, device="cuda:1")
local.fill_(3)
remote.fill_(5)
torch.cuda.synchronize(0)
torch.cuda.synchronize(1)
st = torch.cuda.current_stream(0).cuda_stream
for _ in range(args.reps):
    for engine in ("tma", "ldst"):
        C.copy(remote.data_ptr(), local.data_ptr(), args.bytes, False, engine, {}, {}, 0, st)    # put
        C.copy(local2.data_ptr(), remote.data_ptr(), args.bytes, True, engine, {}, {}, 0, st)    # get
torch.cuda.synchronize(0)
print("put ok:", bool((remote == 3).all().item()), "get ok:", bool((local2 == 3).all().item()))
