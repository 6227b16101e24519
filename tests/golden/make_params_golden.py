"""Generates tests/golden/params_ref.json by running oracle/_ref/params_ref: the REFERENCE's own public headers
(/root/reference/map_merge_3d/include/map_merge_3d/{enum,features,matching,map_merging}.h, compiled unmodified by
`make -C oracle ref` against the two stand-in headers in oracle/hdr_stub/).  Run in the build container only:
/root/reference does not exist on the GPU box, which is why the output is committed."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "params_ref")], text=True)
    data = json.loads(out)
    with open(os.path.join(ROOT, "tests", "golden", "params_ref.json"), "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("wrote tests/golden/params_ref.json:", sorted(data))


if __name__ == "__main__":
    main()
