# dev helper: debug the wide path on one golden case
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bcalm_b200.api import Cdbg
from bcalm_b200 import parity
from oracle import oracle as orc
cases = {c["name"]: c for c in json.load(open("tests/golden/golden.json"))}
for name in sys.argv[1:]:
    c = cases[name]; k = c["k"]
    o = orc.run_any(c["reads"], k, c["abundance_min"])
    with Cdbg(k=k, abundance_min=c["abundance_min"]) as cd:
        cd.push_reads(c["reads"]); st = cd.count(); km, ct = cd.solid_kmers()
        W = km.shape[1] if km.ndim == 2 else 1
        print(name, "inst", st["n_instances"], o.stats["n_instances"], "distinct", st["n_distinct"], o.stats["n_distinct"], "solid", st["n_solid"], o.stats["n_solid"])
        km2 = km.reshape(-1, W)
        order = np.lexsort(tuple(km2[:, i] for i in range(W)))
        ok = np.array_equal(km2[order], o.solid_kmers.reshape(len(o.solid_counts), -1)[:, :W]) if len(order) == len(o.solid_counts) else False
        print("  solid kmers equal:", ok, " counts equal:", np.array_equal(ct[order], o.solid_counts) if ok else None)
        cd.compact(); cd.link(); g = cd.unitigs()
        print("  unitigs", g.n_unitigs, o.n_unitigs, "links", len(g.links), len(o.links), "compactions", g.stats["n_compactions"])
        d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
        print("  parity:", "OK" if not d else d[:300])
