#!/usr/bin/env python3
"""A/B runs of the fast-path kernels in ONE process (one ncu launch list covers all of them).

Each variant sets environment switches of the engine and scans the same device-resident text with the same pattern set
plus a variant-specific literal (so that no cached database / gram table of another variant is reused).
    python tools/variants.py --mib 2048 --set c2 --variants new,verify_smem,emit_v1,old
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from gpu_api import scan_buffer  # noqa: E402
from hypergrep_b200 import synth, utils  # noqa: E402

VARIANTS = {
    "new": {},
    "verify_v1": {"GPUGREP_VERIFY": "v1"},
    "verify_v1": {"GPUGREP_VERIFY": "v1"},
    "noreprobe": {"GPUGREP_NO_REPROBE": "1"},
}

parser = argparse.ArgumentParser()
parser.add_argument("--mib", type=int, default=2048)
parser.add_argument("--set", default="c2")
parser.add_argument("--passes", type=int, default=3)
parser.add_argument("--variants", default="new,verify_v1")
args = parser.parse_args()
lib = utils._get_hyperscanner_lib()
plants = None
if args.set == "c1":
    patterns = list(synth.C1_PATTERNS)
elif args.set == "c3":
    patterns, plants = synth.c3_patterns()
elif args.set == "lit":
    patterns = [p for p in synth.C2_LITERALS if len(p) >= 9]
else:
    patterns = list(synth.C2_PATTERNS)
host = torch.empty(args.mib << 20, dtype=torch.uint8).pin_memory()
synth.fill_syslog(host.numpy(), seed=int(os.environ.get("GPUGREP_TEXT_SEED", "1234")), plants=plants, plant_ppm=1000 if plants else 0, lib=lib)
dev = host.cuda()
torch.cuda.synchronize()
SWITCHES = sorted({k for v in VARIANTS.values() for k in v})
for name in args.variants.split(","):
    for key in SWITCHES:
        os.environ.pop(key, None)
    os.environ.update(VARIANTS[name])
    tagged = patterns + ["zq" + name.replace("_", "") + "variantqz"]
    for k in range(args.passes):
        rc, _, st = scan_buffer(lib, dev.data_ptr(), dev.numel(), 1, tagged, collect=False)
        assert rc == 0, rc
    print(f"variant={name:10s} set={args.set} matches={st.matches} candidates={st.candidates} gpu_ms={st.gpu_ms:.3f} stream_ms={st.stream_kernel_ms:.3f} "
          f"launches={st.launches} path={st.path} GB/s={st.bytes_scanned / st.gpu_ms / 1e6:.1f} stream_GB/s={st.bytes_scanned / st.stream_kernel_ms / 1e6:.1f}", flush=True)
