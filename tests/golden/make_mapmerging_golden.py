"""Generates tests/golden/mapmerging_ref.json by running the REFERENCE's own code — src/features.cpp, src/matching.cpp,
src/map_merging.cpp and src/graph.cpp compiled unmodified into oracle/_ref/libmapmerging_ref.so (`make -C oracle ref`), with
the PCL classes they drive replaced by stand-ins over the CPU checker's stage functions — on seeded inputs: estimateMapsTransforms and composeMaps results (float bits) and the text that
MapMergingParams::fromCommandLine + operator<< print for a set of command lines.  Run in the build container only:
/root/reference does not exist on the GPU box, which is why the outputs are committed."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_py  # noqa: E402
import mm3d_pkg  # noqa: E402

ESTIMATE_CASES = [
    dict(name="sift_fpfh_matching", params=dict()),
    dict(name="harris_fpfh", params=dict(keypoint_type=1, keypoint_threshold=0.005)),
    dict(name="sac_ia_no_refine", params=dict(estimation_method=1, refine_transform=0, max_iterations=60)),
    dict(name="threshold_disconnects", params=dict(confidence_threshold=1e9)),
    dict(name="rsd", params=dict(descriptor_type=3)),
    dict(name="pfh_default_descriptor", params=dict(descriptor_type=0)),
    dict(name="pfhrgb", params=dict(descriptor_type=1)),
    dict(name="shot", params=dict(descriptor_type=4)),
    dict(name="sc3d", params=dict(descriptor_type=5)),
]
COMMAND_LINES = [
    [],
    ["--resolution", "0.25", "--descriptor_type", "SHOT", "--matching_k", "0", "--refine_transform", "0", "--bogus", "1"],
    ["--keypoint_type", "HARRIS", "--estimation_method", "SAC_IA", "--matching_k", "7", "--refine_transform", "2", "--max_iterations", "12"],
    ["--output_resolution", "0.01", "--confidence_threshold", "2.5", "--transform_epsilon", "1e-3", "--inlier_threshold", "0.3",
     "--max_correspondence_distance", "0.7", "--normal_radius", "0.45", "--descriptor_radius", "1.1", "--outliers_min_neighbours", "20",
     "--keypoint_threshold", "3"],
    ["--keypoint_type", "sift"],
    ["--descriptor_type", "SIFT"],
    ["--resolution"],
]


def inputs():
    mm3d_pkg.load()
    synth = importlib.import_module("map_merge_b200.synth")
    maps, truth = synth.make_maps(5, 3, 9000, 7.0, 5.0, 1, 1, 0.8)
    maps = list(maps) + [np.zeros((0, 4), np.float32)]  # a robot that has not published yet
    return maps, truth


def compose_cases(maps, transforms):
    T = np.array(transforms, np.float32)
    T0 = T.copy(); T0[1] = 0  # a map that could not be placed is skipped
    return [dict(name="all", T=T, res=0.05), dict(name="one_skipped", T=T0, res=0.05), dict(name="coarse", T=T, res=0.3)]


def match_cases():
    rng = np.random.default_rng(77)
    fp = lambda n: (rng.dirichlet(np.ones(11), size=(n, 3)).reshape(n, 33) * 100).astype(np.float32)  # FPFH-like rows
    a = fp(300)
    b = np.concatenate([a[:120] + rng.normal(0, 0.5, (120, 33)).astype(np.float32), fp(200)])
    dup = np.concatenate([a[:50], np.repeat(a[50:51], 20, axis=0)])  # ties: equal distances everywhere
    r2 = rng.uniform(0.1, 0.22, (150, 2)).astype(np.float32)
    return [("fpfh_like", a, b), ("ties", dup, np.concatenate([dup[::-1][:40], a[100:160]])), ("rsd_2d", r2, r2[::-1][:100] + np.float32(0.001))]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32).reshape(-1).tolist()


def main():
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.DEVNULL)
    ref = oracle_py.MapMergingRef()
    maps, _ = inputs()
    out = dict(estimate=[], compose=[], command_lines=[], match=[])
    first = None
    for c in ESTIMATE_CASES:
        T = ref.estimate_maps_transforms(maps, oracle_py.default_params(**c["params"]))
        out["estimate"].append(dict(name=c["name"], params=c["params"], shape=list(T.shape), bits=bits(T)))
        if first is None:
            first = T
    # the trailing empty map has no keypoints, so no pair names it and the reference returns 3 transforms for 4 clouds
    assert first.shape == (3, 4, 4)
    for c in compose_cases(maps[:3], first):
        r = ref.compose_maps(maps[:3], c["T"], c["res"])
        out["compose"].append(dict(name=c["name"], res=c["res"], T_bits=bits(c["T"]), n=int(len(r)), checksum=int(np.bitwise_xor.reduce(r.view(np.uint32).reshape(-1).astype(np.uint64) * np.arange(1, r.size + 1, dtype=np.uint64) % np.uint64(2**61 - 1))),
                                   head_bits=bits(r[:8])))
    # findFeatureCorrespondences (the reference's own reciprocal k-NN cross-match) on seeded descriptor sets
    out["match"] = []
    for name, ds, dt in match_cases():
        for k in (1, 5, 8):
            p, d = ref.match(ds, dt, k)
            out["match"].append(dict(name=name, k=k, pairs=p.reshape(-1).tolist(), dist_bits=bits(d)))
    for argv in COMMAND_LINES:
        out["command_lines"].append(dict(argv=argv, text=ref.params_text(argv)))
    with open(os.path.join(ROOT, "tests", "golden", "mapmerging_ref.json"), "w") as fh:
        json.dump(out, fh, indent=1)
        fh.write("\n")
    print("wrote tests/golden/mapmerging_ref.json:", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
