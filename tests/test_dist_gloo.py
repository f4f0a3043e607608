"""N>1 host-side logic on CPU: world_size-2 gloo job. Each rank takes its round-robin share of the
segments (viyadb_b200.dist), computes its partial group table with the oracle, and the partials are
merged with the same commutative rules the NCCL merge uses (sum / min / max / set union) — the result
must equal the single-process result over all segments, and the ncclUniqueId plumbing must deliver
rank 0's 128 bytes to everybody."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch.distributed as dist
import viya_oracle
from helpers import random_table
from viyadb_b200 import dist as vdist

dist.init_process_group("gloo")
rank, world, _ = vdist.world()
conf = {"name": "t", "segment_size": 2000, "dimensions": [{"name": "a"}, {"name": "n", "type": "ushort"}],
        "metrics": [{"name": "count", "type": "count"}, {"name": "s", "type": "long_sum"}, {"name": "mn", "type": "int_min"},
                    {"name": "mx", "type": "int_max"}, {"name": "u", "type": "bitset"}]}
spec = {"a": (1, 7), "n": (0, 30), "count": (1, 2), "s": (-1000, 1000), "mn": (-2**31, 2**31 - 1),
        "mx": (-2**31, 2**31 - 1), "u": ("ids", 40, 2)}
segs, dicts, hidden = random_table(conf, 7, 2000, 77, spec, last_rows=321)   # same seed on every rank
q = {"type": "aggregate", "table": "t", "dimensions": ["a", "n"], "metrics": ["count", "s", "mn", "mx", "u"],
     "filter": {"op": "lt", "column": "n", "value": "20"}}
mine = vdist.segments_for_rank(len(segs), rank, world)
assert mine == list(range(rank, len(segs), world))
uid = vdist.share_unique_id(lambda: bytes(range(128)), rank, dist)
assert uid == bytes(range(128))

# partial of this rank: raw groups, plus the (group, id) pairs of the bitset metric
part = viya_oracle.run_query(conf, [segs[i] for i in mine], dicts, q)
keys = list(zip(*[k.tolist() for k in part["groups"]["keys"]]))
accs = [a.tolist() for a in part["groups"]["accs"]]
pairs = set()
for i in mine:
    seg = segs[i]
    sel = np.nonzero(seg["n"] < 20)[0]
    off, val = seg["u"]
    for r in sel:
        for x in val[off[r]:off[r + 1]]:
            pairs.add((int(seg["a"][r]), int(seg["n"][r]), int(x)))
box = [None] * world
dist.all_gather_object(box, (keys, accs, pairs))
merged, allpairs = {}, set()
for ks, ac, pr in box:
    allpairs |= pr
    for gi, k in enumerate(ks):
        cur = merged.get(k)
        v = [ac[0][gi], ac[1][gi], ac[2][gi], ac[3][gi]]
        merged[k] = v if cur is None else [cur[0] + v[0], cur[1] + v[1], min(cur[2], v[2]), max(cur[3], v[3])]
distinct = {}
for a, n, x in allpairs:
    distinct[(a, n)] = distinct.get((a, n), 0) + 1
full = viya_oracle.run_query(conf, segs, dicts, q)
fk = list(zip(*[k.tolist() for k in full["groups"]["keys"]]))
assert set(fk) == set(merged)
for gi, k in enumerate(fk):
    want = [full["groups"]["accs"][m][gi] for m in range(5)]
    got = merged[k] + [distinct[k]]
    assert [int(x) for x in want] == [int(x) for x in got], (k, want, got)
dist.barrier()
if rank == 0:
    print("GLOO_MERGE_OK", len(fk))
dist.destroy_process_group()
'''


def test_world_size_2_shard_and_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert "GLOO_MERGE_OK" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]
