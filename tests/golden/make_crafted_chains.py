"""Writes tests/golden/chains_crafted.json: hand-built chains (tests/_common.craft_chains) whose records
were produced by the reference's own alignChain_edlib through oracle/_ref/libref_shim.so.  They cover what
the front-end-derived chains in chains.json do not: records dropped for having fewer than two anchors,
splits in adjacent gaps, and long error-free stretches.  Run in the build container only
(`make -C oracle ref` first); the JSON is what travels."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _oracle as O  # noqa: E402
from _common import CRAFTED_REF, craft_chains  # noqa: E402
from lordfast_b200 import sim  # noqa: E402


def main():
    assert O.have_ref(), "oracle/_ref missing: run `make -C oracle ref` first"
    ref = sim.make_reference(*CRAFTED_REF)
    idx = O.RefIndex(ref.tobytes())
    out = []
    for q, seeds in craft_chains(ref):
        sam = O.ref_align_chain(idx, seeds, q.tobytes(), 0)
        out.append(dict(read=q.tobytes().decode(), seeds=[list(s) for s in seeds], sam=sam))
    json.dump(out, open(os.path.join(HERE, "chains_crafted.json"), "w"), separators=(",", ":"))
    print(len(out), "chains,", sum(len(c["sam"]) for c in out), "records")


if __name__ == "__main__":
    main()
