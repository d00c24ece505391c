"""graph_checksum (bcalm_b200/checksum.py, the size-independent parity check of bench.py and the full-scale tests) is invariant exactly
where the parity relation is: the oracle's graph and the REFERENCE's own output (different ids, strands, link order, cycle cut
points) give the same checksum on every golden case, and any change of a sequence, a count or a link changes it."""
import copy

import numpy as np
import pytest

from bcalm_b200 import parity
from bcalm_b200.checksum import graph_checksum, records_to_graph
from conftest import load_golden
from oracle import oracle as orc

CASES = [c for c in load_golden() if c["unitigs_fa"] is not None and c["k"] <= 31]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_and_reference_output_have_the_same_checksum(case):
    k = case["k"]
    o = orc.run(case["reads"], k, case["abundance_min"])
    ref = records_to_graph(parity.parse_unitigs_fa(case["unitigs_fa"]), k)
    a, b = graph_checksum(o), graph_checksum(ref)
    assert a == b
    assert a["bad_overlaps"] == 0


def test_checksum_detects_changes():
    case = [c for c in CASES if c["name"] == "reads_3_k31_a2"][0]
    g = records_to_graph(parity.parse_unitigs_fa(case["unitigs_fa"]), case["k"])
    base = graph_checksum(g)
    h = copy.deepcopy(g); h.kc[3] += 1
    assert graph_checksum(h)["nodes"] != base["nodes"]
    h = copy.deepcopy(g); h.seq = h.seq.copy(); h.seq[40] = ord("A") if h.seq[40] != ord("A") else ord("C")
    assert graph_checksum(h)["nodes"] != base["nodes"]
    h = copy.deepcopy(g); h.links = h.links.copy(); h.links["to_sign"][0] = ord("-") if h.links["to_sign"][0] == ord("+") else ord("+")
    c = graph_checksum(h)
    assert c["links"] != base["links"] and c["bad_overlaps"] > 0
    # chunking does not change the result
    assert graph_checksum(g, chunk_nt=64) == base
