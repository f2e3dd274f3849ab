"""tests/golden/make_demo_golden.py -- writes tests/golden/demo_runme_sha256.json: the SHA-256 of every output file the PURE
reference binary (oracle/_ref/pindel_ref, built from the untouched sources) writes for BASELINE.json configs[0], the
reference's own demo (demo/RUNME: "pindel -i simulated_config.txt -f simulated_reference.fa -o bamtest -c ALL"), with one
thread -- the reference's files depend on its thread count (the order in which an OpenMP loop files reads), so -T 1 is the
reproducible form, as in its own regression script (devtools/run_regressiontests.sh).  Needs /root/reference.
    python tests/golden/make_demo_golden.py"""
import hashlib
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DEMO = "/root/reference/demo"
SUFFIXES = ("D", "SI", "INV", "TD", "LI", "BP", "RP", "INT_final", "CloseEndMapped")


def main():
    tmp = tempfile.mkdtemp(prefix="demo_gold_")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "pindel_ref"), "-i", "simulated_config.txt", "-f", "simulated_reference.fa",
                    "-o", os.path.join(tmp, "bamtest"), "-c", "ALL", "-T", "1"], cwd=DEMO, check=True, capture_output=True)
    out = {}
    for suf in SUFFIXES:
        data = open(os.path.join(tmp, "bamtest_" + suf), "rb").read()
        out[suf] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
    with open(os.path.join(HERE, "demo_runme_sha256.json"), "w") as f:
        json.dump({"command": "pindel -i simulated_config.txt -f simulated_reference.fa -o bamtest -c ALL -T 1", "files": out}, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
