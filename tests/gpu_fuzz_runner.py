"""Device leg of the differential fuzz, as a standalone process (tests/test_zz_gpu_fuzz.py starts it): every seeded random
query the real reference answered (tests/golden/ref_fuzz*.jsonl), on its own segment bytes, through the C ABI on cuda:0.

  python tests/gpu_fuzz_runner.py aggregate | aggregate_forced_hash | select_search

Prints one JSON line {"which", "total", "ran", "failed": [[test, what], ...]} and exits 1 when anything differs. A process
of its own so that a crash or a poisoned CUDA context stays here."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "oracle"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

import golden_util as G  # noqa: E402


def main(which):
    import viyadb_b200 as vdb
    failed, ran = [], 0
    if which in ("aggregate", "aggregate_forced_hash"):
        import test_gpu_golden as T
        flags = 1 if which == "aggregate_forced_hash" else 0
        recs = [r for r in G.records("ref_fuzz_scenarios.jsonl") if "error" not in r]
        for rec in recs:
            try:
                T.run(vdb, rec, flags=flags)
                ran += 1
            except vdb.VgpuError as e:
                if flags == 1 and e.code == -2:      # key wider than 64 bits under a forced hash table
                    continue
                failed.append([G.rec_id(rec), repr(e)[:200]])
            except AssertionError as e:
                failed.append([G.rec_id(rec), str(e)[:200]])
    else:
        recs = [r for r in G.records("ref_fuzz_select_scenarios.jsonl") if "error" not in r]
        for rec in recs:
            db = vdb.Database({"tables": [rec["table"]]}, device=0)
            try:
                t = db.get_table(rec["table"]["name"])
                t.load_dump(G.seg_path(rec["seg"]))
                out = vdb.MemoryRowOutput()
                try:
                    stats = db.query(rec["query"], out)
                except vdb.VgpuError as e:
                    if e.code != -2:                  # search on a floating-point dimension is outside the device path
                        failed.append([G.rec_id(rec), repr(e)[:200]])
                    continue
                ran += 1
                if out.rows != rec["rows"]:
                    failed.append([G.rec_id(rec), "rows " + str(out.rows)[:90] + " != " + str(rec["rows"])[:90]])
                else:
                    for k, v in rec["stats"].items():
                        if getattr(stats, k) != v:
                            failed.append([G.rec_id(rec), f"{k} {getattr(stats, k)} != {v}"])
            finally:
                db.close()
    print(json.dumps({"which": which, "total": len(recs), "ran": ran, "failed": failed[:20], "nfailed": len(failed)}))
    return 1 if failed or ran == 0 else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "aggregate"))
