"""Loading of the committed golden fixtures (tests/golden/, produced by make_golden.py from the
real reference). Usable on the GPU box: nothing here touches /root/reference."""
import gzip
import json
import os
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
_TMP = {}


def records(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        return []
    return [json.loads(l) for l in open(path)]


def seg_path(name):
    """Decompress a golden segment dump to a temp file once per session."""
    if name not in _TMP:
        with open(os.path.join(GOLDEN, "seg", name), "rb") as f:
            data = gzip.decompress(f.read())
        fd, p = tempfile.mkstemp(suffix=".bin", prefix="vgpu_seg_")
        with os.fdopen(fd, "wb") as f:
            f.write(data)
        _TMP[name] = p
    return _TMP[name]


def rec_id(rec):
    return f"{rec['test']}#{rec.get('seq', '')}"


def is_float_metric_query(rec):
    """float/double SUM/AVG results depend on the order of additions: compare with a tolerance."""
    types = {m["name"]: m["type"] for m in rec["table"].get("metrics", [])}
    return any(t.startswith(("float_", "double_")) and t.endswith(("_sum", "_avg")) for t in types.values())
