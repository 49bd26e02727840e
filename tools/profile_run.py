#!/usr/bin/env python3
"""Small device-resident scan used under ncu (profiles/): one warm-up pass and one measured pass over --mib MiB."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from gpu_api import scan_buffer  # noqa: E402
from hypergrep_b200 import synth, utils  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--mib", type=int, default=1024)
parser.add_argument("--set", default="c2")
parser.add_argument("--passes", type=int, default=2)
parser.add_argument("--quiet", action="store_true")
parser.add_argument("--patterns", type=int, default=10000, help="pattern count of the c5 set")
parser.add_argument("--distinct-ids", action="store_true", help="one match id per pattern (event mode, general path)")
args = parser.parse_args()
lib = utils._get_hyperscanner_lib()
plants = None
if args.set == "c1":
    patterns = synth.C1_PATTERNS
elif args.set == "c3":
    patterns, plants = synth.c3_patterns()
elif args.set == "lit":   # literals of >= 9 bytes only: the prefilter samples at stride 4
    patterns = [p for p in synth.C2_LITERALS if len(p) >= 9]
else:
    patterns = synth.C2_PATTERNS
host = torch.empty(args.mib << 20, dtype=torch.uint8).pin_memory()
if args.set == "c5":   # configs[4]: long JSON-ish lines, caseless template patterns; the text is a tiled 8 MiB sample
    import time
    import numpy as np
    patterns = synth.c5_patterns(args.patterns)
    sample = np.frombuffer(synth.jsonish_bytes(8 << 20, patterns_to_plant=["session_4242 failed", "code=E31337abcd"]), dtype=np.uint8)
    reps = -(-host.numel() // sample.size)
    host.numpy()[:] = np.tile(sample, reps)[: host.numel()]
    host.numpy()[-1] = 10
    t0 = time.time()
else:
    synth.fill_syslog(host.numpy(), seed=1234, plants=plants, plant_ppm=1000 if plants else 0, lib=lib)
dev = host.cuda()
torch.cuda.synchronize()
flags = [15] * len(patterns) if args.set == "c5" else None   # caseless
for _ in range(args.passes):
    ids = list(range(len(patterns))) if args.distinct_ids else None
    rc, _, st = scan_buffer(lib, dev.data_ptr(), dev.numel(), 1, patterns, flags=flags, ids=ids, collect=False)
    assert rc == 0
    print(f"set={args.set} bytes={st.bytes_scanned} lines={st.lines} matches={st.matches} candidates={st.candidates} "
          f"gpu_ms={st.gpu_ms:.3f} wall_ms={st.wall_ms:.3f} stream_ms={st.stream_kernel_ms:.3f} launches={st.launches} path={st.path} "
          f"GB/s={st.bytes_scanned / st.gpu_ms / 1e6:.1f} stream_GB/s={st.bytes_scanned / st.stream_kernel_ms / 1e6:.1f}")
