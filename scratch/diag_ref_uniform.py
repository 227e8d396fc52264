"""Diagnosis of the reference's uniform sampler on this platform: runs oracle/_ref's
_TemporalSampler(UNIFORM) in child processes on growing inputs and records return code + stderr of each."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import synth_stream  # noqa: E402

CHILD = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.join(%r, "oracle", "_ref"))
import libgnnflow as ref
spec = np.load(sys.argv[1])
g = ref._DynamicGraph(64 << 20, 1 << 30, ref.MemoryResourceType.CUDA, int(spec["minblk"]), 1024, ref.InsertionPolicy.INSERT, 0, True)
src, dst, ts, eid = spec["src"], spec["dst"], spec["ts"], spec["eid"]
B = int(spec["batch"])
for i in range(0, len(src), B):
    g.add_edges(src[i:i+B], dst[i:i+B], ts[i:i+B], eid[i:i+B])
print("graph built", g.num_edges(), flush=True)
s = ref._TemporalSampler(g, [int(f) for f in spec["fanouts"]], ref.SamplingPolicy.UNIFORM, 1, 0.0, False, 1234)
print("sampler built", flush=True)
r = s.sample(spec["roots"], spec["rts"])
print("sampled", r[0][0].num_src_nodes(), r[0][0].num_dst_nodes(), flush=True)
torch.cuda.synchronize()
print("ok", flush=True)
''' % ROOT


def run(tag, tmp, sanitizer=False, **spec):
    sp = os.path.join(tmp, tag + ".npz")
    np.savez(sp, **spec)
    cmd = [sys.executable, "-c", CHILD, sp]
    if sanitizer:
        cmd = ["compute-sanitizer", "--tool", "memcheck", "--print-limit", "5"] + cmd
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    return dict(tag=tag, rc=r.returncode, stdout=(r.stdout[:6000] if sanitizer else r.stdout[-1500:]), stderr=r.stderr[-3000:])


def main():
    tmp = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
    os.makedirs(tmp, exist_ok=True)
    out = []
    # 1. toy graph, all roots have candidates
    src = np.array([0, 0, 0, 0, 1, 1, 1, 1], dtype=np.int64)
    dst = np.array([2, 3, 4, 5, 2, 3, 4, 5], dtype=np.int64)
    ts = np.arange(8, dtype=np.float32)
    eid = np.arange(8, dtype=np.int64)
    base = dict(src=src, dst=dst, ts=ts, eid=eid, batch=8, minblk=4, fanouts=np.array([2]))
    out.append(run("toy_with_candidates", tmp, **base, roots=np.array([0, 1], dtype=np.int64), rts=np.array([10, 10], dtype=np.float32)))
    # 2. toy graph, a root with edges but none inside the window (a12 zero-candidate UB)
    out.append(run("toy_zero_candidates", tmp, **base, roots=np.array([0, 1], dtype=np.int64), rts=np.array([0, 10], dtype=np.float32)))
    # 3. the test's stream, fan-out 8, 1 200 roots
    s, d, t, e = synth_stream(300, 60, 40000, seed=31, t_max=4000.0)
    t = (np.floor(t * 2) / 2).astype(np.float32)
    lo = 20000
    first_ts = np.full(400, np.inf)
    np.minimum.at(first_ts, s, t)
    keep = first_ts[s[lo:lo + 600]] < t[lo:lo + 600]
    roots = np.concatenate([s[lo:lo + 600][keep], d[lo:lo + 600]]).astype(np.int64)
    rts = np.concatenate([t[lo:lo + 600][keep], t[lo:lo + 600]]).astype(np.float32)
    big = dict(src=s, dst=d, ts=t, eid=e, batch=3000, minblk=7, fanouts=np.array([8]))
    out.append(run("stream_1200_roots", tmp, **big, roots=roots, rts=rts))
    out.append(run("stream_32_roots", tmp, **big, roots=roots[:32], rts=rts[:32]))
    out.append(run("stream_src_roots_only", tmp, **big, roots=roots[:int(keep.sum())], rts=rts[:int(keep.sum())]))
    # dst-only roots (vertices that never were a source: their node_table entry has tail == nullptr)
    out.append(run("stream_dst_roots_only", tmp, **big, roots=roots[int(keep.sum()):], rts=rts[int(keep.sum()):]))
    out.append(run("stream_1200_roots_memcheck", tmp, sanitizer=True, **big, roots=roots, rts=rts))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
