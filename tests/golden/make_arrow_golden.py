"""Produces tests/golden/arrow_golden.json: results of the reference's Native (Arrow Acero) plans,
run HERE with Arrow 24.0.0 (pyarrow wheel), on the reference's own large test inputs
(RandomArrayGenerator(42); filter_test.cc:63-78, aggr_test.cc:37-49, take_test.cc:48-72,
join_test.cc:82-121). Inputs come from the oracle's generator, itself pinned to the real
libstdc++/PCG output (generator_golden.json). Only digests are stored.

    python tests/golden/make_arrow_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import oracle  # noqa: E402
from oracle import arrow_native as an  # noqa: E402


def sha(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.uint32).tobytes())
    return h.hexdigest()


out = {"arrow_version": __import__("pyarrow").__version__}

# --- filter / sum: 128 x 65536 full-range, one generator (aggr_test.cc:37-49 shape) ---
g = oracle.RandomArrayGenerator(42)
batches = oracle.make_random_batches(g, 128, 65536)
f = an.FilterNative(batches, use_threads=False)  # single thread keeps batch order deterministic
f.Prepare()
t = f.GetResult()
col = t.column(0).to_numpy()
per_batch = [int((b < (1 << 30)).sum()) for b in batches]
assert col.size == sum(per_batch)
out["filter_128x65536"] = {"rows": int(col.size), "per_batch_first8": per_batch[:8], "sha256": sha(col),
                           "batch0_sha256": sha(col[:per_batch[0]])}
a = an.AggrNative(batches)
a.Prepare()
out["sum_128x65536"] = {"sum": a.Run()}

# --- take: values 128 x 65536, indices 128 x 8192 in [0, 65535] (take_test.cc:48-72) ---
g = oracle.RandomArrayGenerator(42)
vals = oracle.make_random_batches(g, 128, 65536)
idx = oracle.make_random_batches(g, 128, 8192, 0, 65535)
tk = an.TakeNative(vals, idx)
res = tk.Run()
out["take_128x65536_8192"] = {"rows": int(sum(r.size for r in res)), "sha256": sha(*res),
                              "batch0_first8": [int(v) for v in res[0][:8]]}

# --- join LargeTest (join_test.cc:82-121): draw order x batches, then y batches, then fk ---
g = oracle.RandomArrayGenerator(42)
nb, bs = 128, 65536
x = oracle.make_random_batches(g, nb, bs)
pk = oracle.make_index_batches(nb, bs)
y = oracle.make_random_batches(g, nb, bs)
fk = oracle.make_fk_batches(g, bs, nb, bs)
j = an.JoinNative({"fk": fk, "y": y}, {"pk": pk, "x": x})
j.Prepare()
jt = j.Run()
rows = oracle.sort_rows(jt["fk"].to_numpy(), jt["y"].to_numpy(), jt["x"].to_numpy())
out["join_128x65536"] = {"rows": int(jt.num_rows), "columns": jt.schema.names, "sorted_sha256": sha(*rows),
                         "checksum": oracle.triple_checksum(*rows), "first_row": [int(c[0]) for c in rows]}

Path(__file__).with_name("arrow_golden.json").write_text(json.dumps(out, indent=1) + "\n")
print(json.dumps(out, indent=1))
