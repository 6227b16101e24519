"""Generates tests/golden/graph_ref.json by running the REFERENCE's own pose-graph code
(/root/reference/map_merge_3d/src/graph.cpp, compiled unmodified into oracle/_ref/libgraph_ref.so
by `make -C oracle ref`) on seeded random confidence graphs.  Run in the build container only:
/root/reference does not exist on the GPU box, which is why the outputs are committed.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py  # noqa: E402


def cases():
    rng = np.random.default_rng(2024)
    out = []
    for n_nodes in (2, 3, 4, 5, 8, 12, 32):
        for variant in range(6):
            pairs = [(i, j) for i in range(n_nodes - 1) for j in range(i + 1, n_nodes)]
            if variant % 3 == 1 and len(pairs) > 3:  # drop some pairs (maps without keypoints)
                keep = rng.random(len(pairs)) > 0.3
                pairs = [p for p, k in zip(pairs, keep) if k] or pairs[:1]
            conf = rng.uniform(0.0, 20.0, len(pairs))
            if variant % 3 == 2:  # ties and failed pairs
                conf = np.round(conf / 5.0) * 5.0
                conf[rng.random(len(pairs)) < 0.3] = 5.562684646268003e-309  # 1 / DBL_MAX
            thr = [0.0, 4.0, 10.0][variant % 3] if variant < 3 else [0.0, 7.5, 19.5][variant % 3]
            out.append(dict(st=[list(p) for p in pairs], conf=[float(c) for c in conf], thr=float(thr)))
    return out


def main():
    ref = oracle_py.GraphRef()
    golden = []
    for c in cases():
        inc, te, cen, nn = ref.graph(c["st"], c["conf"], c["thr"])
        golden.append(dict(c, in_component=inc.tolist(), tree_edges=te.tolist(), centers=cen.tolist(), n_nodes=int(nn)))
    with open(os.path.join(os.path.dirname(__file__), "graph_ref.json"), "w") as f:
        json.dump(golden, f)
    print("wrote", len(golden), "cases")


if __name__ == "__main__":
    main()
