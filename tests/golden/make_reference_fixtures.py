"""Generates tests/golden/ref_*.npz: inputs and the outputs of the UNMODIFIED reference extension (oracle/_ref/libgnnflow,
built from /root/reference by oracle/build_ref.sh) for the store and the recent sampler.  Needs a GPU (the reference
cannot run without one); run on the GPU box:

    gpurun -- python tests/golden/make_reference_fixtures.py

The fixtures are small (a 6,000-edge stream) and committed, so that the CPU suite can pin the oracle -- and the GPU
suite the CUDA path -- against the reference's real output without the reference being present.
Uniform sampling is not recorded: the reference draws from cuRAND XORWOW state ("parity unpinned")."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_stream  # noqa: E402

BATCH, MINBLK = 700, 5
CASES = {
    "l1": dict(fanouts=[10]),
    "l2": dict(fanouts=[5, 4]),
    "snap3": dict(fanouts=[4, 3], num_snapshots=3, snapshot_time_window=35.5),
    "win": dict(fanouts=[6], snapshot_time_window=100.0),
    "prop": dict(fanouts=[3, 2], num_snapshots=2, snapshot_time_window=77.0, prop_time=True),
}


def stream():
    src, dst, ts, eid = synth_stream(60, 20, 6000, seed=77, t_max=900.0)
    ts = (np.floor(ts * 2) / 2).astype(np.float32)  # ties
    return src, dst, ts, eid


def roots(src, dst, ts):
    rng = np.random.default_rng(5)
    out = []
    for lo in (50, 2500, 5800):
        r = np.concatenate([src[lo:lo + 150], dst[lo:lo + 150], rng.integers(0, 85, 150)]).astype(np.int64)
        t = np.concatenate([ts[lo:lo + 150]] * 3).astype(np.float32)
        out.append((r, t))
    return out


def run(spec, tmp):
    sp, out = os.path.join(tmp, "spec.npz"), os.path.join(tmp, "out.npz")
    np.savez(sp, **spec)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "ref_runner.py"), sp, out])
    return dict(np.load(out))


def main():
    import tempfile
    src, dst, ts, eid = stream()
    base = dict(src=src, dst=dst, ts=ts, eid=eid, batch=BATCH, minblk=MINBLK, adaptive=True)
    with tempfile.TemporaryDirectory() as tmp:
        ref = run(dict(base, mode="store"), tmp)
        np.savez_compressed(os.path.join(HERE, "ref_store.npz"), **{"in_" + k: v for k, v in base.items()}, **ref)
        rb = roots(src, dst, ts)
        for name, case in CASES.items():
            spec = dict(base, mode="sample", fanouts=np.array(case["fanouts"]), policy=0,
                        num_snapshots=case.get("num_snapshots", 1), window=case.get("snapshot_time_window", 0.0),
                        prop_time=case.get("prop_time", False), nroots=len(rb))
            for i, (r, t) in enumerate(rb):
                spec["roots_%d" % i], spec["rts_%d" % i] = r, t
            ref = run(spec, tmp)
            np.savez_compressed(os.path.join(HERE, "ref_sample_%s.npz" % name),
                                **{"in_" + k: v for k, v in spec.items()}, **ref)
            print(name, "ok", len(ref), "arrays")


if __name__ == "__main__":
    main()
