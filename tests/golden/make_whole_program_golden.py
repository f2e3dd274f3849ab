"""tests/golden/make_whole_program_golden.py -- writes tests/golden/reference_whole_program_sha256.json: the SHA-256 of every output
file the PURE reference binary (oracle/_ref/pindel_ref, built from the untouched sources) writes for whole-program commands of
the reference's own that have no usable committed golden:

  demo_runme      BASELINE.json configs[0], demo/RUNME: pindel -i simulated_config.txt -f simulated_reference.fa -c ALL
  colousingbd15   src/update_target_files.sh:10-15: COLO-829 chr20 text input, -b COLO-829.20.BreakDancer.sv -x 5 -B 100 -l -k
                  -u 0.05 -c 20:0-1,500,000  (a BreakDancer file's events as the far end's first windows, far rings up to 131 k)
  colowobd15      src/update_target_files.sh:17-21: the same without -b
                  (test/SmallTest/TargetOutput/ holds files of these two commands from an older Pindel; the current sources do
                  not reproduce them -- SURVEY.md section 4 -- so the reference binary built here is the golden)

all with one thread: the reference's files depend on its thread count (reads are filed from inside OpenMP loops) and -T 1 is
what its own scripts pin.  Needs /root/reference; takes a few minutes.      python tests/golden/make_whole_program_golden.py"""
import hashlib
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SMALL = os.path.join(REF, "test", "SmallTest")
SUFFIXES = ("D", "SI", "INV", "TD", "LI", "BP", "RP", "INT_final", "CloseEndMapped")
COLO15 = ["-p", os.path.join(SMALL, "COLO-829_20-p_ok.txt"), "-f", os.path.join(SMALL, "hs_ref_chr20.fa"), "-x", "5", "-B", "100", "-l", "-k",
          "-u", "0.05", "-c", "20:0-1,500,000"]
COMMANDS = {            # name -> (working directory, arguments without -o / -T)
    "demo_runme": (os.path.join(REF, "demo"), ["-i", "simulated_config.txt", "-f", "simulated_reference.fa", "-c", "ALL"]),
    "colousingbd15": (SMALL, COLO15 + ["-b", os.path.join(SMALL, "COLO-829.20.BreakDancer.sv")]),
    "colowobd15": (SMALL, COLO15),
}


def main():
    tmp = tempfile.mkdtemp(prefix="whole_gold_")
    out = {}
    for name, (cwd, args) in COMMANDS.items():
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "pindel_ref")] + args + ["-o", os.path.join(tmp, name), "-T", "1"], cwd=cwd, check=True,
                       capture_output=True)
        files = {}
        for suf in SUFFIXES:
            data = open(os.path.join(tmp, name + "_" + suf), "rb").read()
            files[suf] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
        out[name] = {"args": [a.replace(REF, "$REF") for a in args] + ["-T", "1"], "files": files}
        print(name, {k: v["bytes"] for k, v in files.items()}, flush=True)
    with open(os.path.join(HERE, "reference_whole_program_sha256.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
