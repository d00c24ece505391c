"""TEST INFRASTRUCTURE (manual): every string over {A,C,G,T} of length k..Lmax (one of each reverse-complement pair) as a single
read, through the emulated CUDA path and through the oracle.

    python tests/emu/exhaustive_check.py K LMAX [PART NPARTS]      # e.g. 4 9 0 7 ... 4 9 6 7 on seven cores (~5 min)

End of round 1: no difference for k = 4, 5, 6 with Lmax = 9."""
import itertools
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import build_emu  # noqa: E402
from bcalm_b200 import api, parity  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def main():
    k, lmax = int(sys.argv[1]), int(sys.argv[2])
    part, nparts = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 1)
    api.LIB_PATH = build_emu.build()
    comp = str.maketrans("ACGT", "TGCA")
    bad = n = 0
    t0 = time.time()
    with api.Cdbg(k=k, abundance_min=1) as c:
        for length in range(k, lmax + 1):
            for idx, tup in enumerate(itertools.product("ACGT", repeat=length)):
                if idx % nparts != part:
                    continue
                s = "".join(tup)
                if s > s[::-1].translate(comp):
                    continue
                g, o = c.build([s]), orc.run([s], k, 1)
                n += 1
                d = parity.diff(parity.graph_to_records(g), parity.graph_to_records(o), k)
                if d != "" or g.stats["n_solid"] != o.stats["n_solid"]:
                    bad += 1
                    print("FAIL", s, d[:200], flush=True)
    print(f"k={k} L<={lmax} part {part}/{nparts}: {n} inputs, {bad} failures, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
