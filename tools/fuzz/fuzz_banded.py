"""Random sizes / world sizes for the band-sharded chain (native C++ driver and the Python pipeline), ranks as threads.

Ad-hoc driver behind the bounded, seeded versions in tests/ (tests/test_multigpu.py); runs on the test emulator
(tests/hostsim), CPU only.  Usage: python tools/fuzz/fuzz_banded.py <seed> <count>
"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import random
from tests import test_multigpu as T
from vulkanpbrt_b200.multigpu import BandPlan
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = random.Random(seed)
done = 0
while done < n:
    world = rng.choice([2, 2, 3, 4])
    W = rng.choice([33, 40, 54, 63, 64, 65, 96, 101, 130])
    H = rng.randint(world * 96, world * 96 + 260)
    taa = rng.random() < 0.7
    frames = rng.choice([10, 17, 19])
    try:
        BandPlan(W, H, world, 32, 12, taa)
    except ValueError as e:
        continue
    cfg = dict(world=world, W=W, H=H, taa=taa, frames=frames)
    which = rng.choice(["native", "native", "python"])
    try:
        if which == "native":
            T.test_native_banded_rank_on_the_emulator(world, taa, W, H, frames)
        else:
            T.test_peer_memory_exchange_protocol_on_the_emulator(world, taa, W, H, frames)
        print("ok", which, cfg, flush=True)
    except BaseException as e:
        print("FAIL", which, cfg, repr(e)[:400], flush=True)
    done += 1
