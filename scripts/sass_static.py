"""Static digest of the shipped library's search kernels (no GPU needed): registers, stack frame, shared memory, SASS
instruction count and opcode histogram per template variant, the TMA / mbarrier / vote / redux mnemonics that show what the
kernel is built from, and the unrolled level walk of the headline variant.   python scripts/sass_static.py > profiles/<name>.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pindel_b200", "libpindel_pg.so")


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL)\s*:\s*(\d+)", line))
            cur = None
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(2).strip())
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    print("# Static digest of `pindel_b200/libpindel_pg.so` (cuobjdump; sm_100a)\n")
    print("| kernel | registers | stack frame (B) | static shared (B) | SASS instructions | code (KB) |")
    print("|---|---|---|---|---|---|")
    for f, ins in funcs.items():
        name = demangle(f)
        if "pg_search_kernel" not in name:
            continue
        u = usage.get(f, {})
        print("| `%s` | %s | %s | %s | %d | %.1f |" % (name.split("(")[0].replace("void ", ""), u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), len(ins), len(ins) * 16 / 1024))
    head = [f for f in funcs if "pg_search_kernelILi8ELi7E" in f]
    if head:
        ins = funcs[head[0]]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in ins)
        print("\n## headline variant `pg_search_kernel<8, 7>` (reads <= 256 bp, 7 mismatch levels): opcode histogram\n")
        print(", ".join("%s %d" % kv for kv in ops.most_common(28)))
        marks = ["UBLKCP", "SYNCS", "VOTE", "REDUX", "CREDUX", "SHFL", "FLO", "BREV", "POPC", "VIMNMX", "LDL", "STL", "ATOMG", "RED"]
        print("\nmnemonics: " + ", ".join("%s x%d" % (m, sum(1 for i in ins if re.search(r"(^|\s)%s(\.|\s|$)" % m, re.sub(r"^@!?U?P\d+\s+", "", i)))) for m in marks))
        flo = [k for k, i in enumerate(ins) if i.startswith("FLO.U32 ")]
        if len(flo) >= 4:
            a = flo[1] - 4
            print("\n## the level walk (three of the seven unrolled levels): clear the lowest set bit (VIADD + LOP3), isolate it (LOP3 xor), FLO off the chain, key, top-2 update\n\n```")
            for i in ins[a:a + 30]:
                print("  " + i)
            print("```")


if __name__ == "__main__":
    main()
